// Stages R and U of the CT-MHD step in ONE launch, interleaved by CTA role ("R+U interleaved").
//
// Why: stage R is bound by FP64 issue (60 % pipe, DRAM 27 % busy, the board at its power cap), stage U by memory
// latency (FP64 pipe 18 %, 5 TB/s); run one after the other they cost 27 + 7.5 ms at 512^3.  The fused R+U kernel
// (mhd_ru.cuh) removes the F / EMF round trip but takes stage R's L1 with its shared-memory planes and is slower.
// Here both stages keep their own code (call_face / call_edge, mhd_update_core) and their own access pattern; they
// only share the launch: the linear CTA index walks, slab by slab (the y-chunked order of mhd_cells.cuh) and plane by
// plane, through groups of
//       nR stage-R CTAs of slab-plane (s, g)   followed by   nU stage-U CTAs of slab-plane (s, g - LAG),
// so that the memory-bound update of plane g-LAG runs on the same SMs, at the same time, as the FP64-bound Riemann
// problems of plane g, and reads fluxes / EMFs that are still in L2 (LAG slab-planes of F + EMF are ~12 MB).
//
// Dependencies: the update of cell c reads F / EMF of c, c+1, c+sj, c+sk, ... (planes k and k+1, rows j and j+1).  A
// stage-U CTA therefore owns the rows [j0-1, j1-1) of its slab (one row below the slab's stage-R rows: row j+1 is then
// still inside the slab) and waits until the stage-R CTAs of the slab-planes (s, k), (s, k+1), (s-1, k), (s-1, k+1)
// have all published their completion (a counter per slab-plane: stores, __syncthreads, __threadfence + atomicAdd by
// one thread; the reader polls with ld.acquire.gpu).  Every CTA it waits for has a LOWER linear index, i.e. has been
// dispatched before it (the hardware dispatches the CTAs of a 1-D grid in index order -- the same assumption the
// decoupled look-back scans make), so the wait cannot deadlock; with LAG >= 7 groups (more CTAs than the whole GPU holds)
// between producer and consumer the poll succeeds on the first try.
// Arithmetic: identical to the two-kernel form (same functions, same operand order): results are bitwise equal.
#pragma once
#include "mhd_stage_r.cuh"

namespace ek {

constexpr int RUI_THREADS = 128;
constexpr int RUI_UC = 1; // cells per thread of a stage-U CTA

struct MhdRuiArgs {
  MhdArgs a;           // a.cm.span is a multiple of RUI_THREADS here
  unsigned* flags;     // [nchunk * ksize] stage-R CTAs of the slab-plane that have finished (zeroed before the launch)
  unsigned nR, nU;     // stage-R / stage-U CTAs per group
  unsigned ngroups;    // groups per slab = ksize + lag
  unsigned lag;        // stage U runs this many slab-planes behind stage R
};

// every operand of the update of one cell, in registers (source interface of mhd_update_core, mhd_cells.cuh)
struct MhdUpdRegs {
  static constexpr bool kPreloaded = true;
  double U[11];                     // U^n of the cell; the field on its three upper faces
  double F0[15], Fxi[5], Fyj[5], Fzk[5];
  double E0[3], Ei[2], Ej[2], Ek[2]; // Ei = (Ey, Ez) at c+1, Ej = (Ex, Ez) at c+sj, Ek = (Ex, Ey) at c+sk
  double Eij, Eik, Ejk;              // Ez at c+1+sj, Ey at c+1+sk, Ex at c+sj+sk
  __device__ __forceinline__ double u(int v) const { return U[v]; }
  __device__ __forceinline__ double fx0(int v) const { return F0[v]; }
  __device__ __forceinline__ double fy0(int v) const { return F0[5 + v]; }
  __device__ __forceinline__ double fz0(int v) const { return F0[10 + v]; }
  __device__ __forceinline__ double fxi(int v) const { return Fxi[v]; }
  __device__ __forceinline__ double fyj(int v) const { return Fyj[v]; }
  __device__ __forceinline__ double fzk(int v) const { return Fzk[v]; }
  __device__ __forceinline__ double e0(int c) const { return E0[c]; }
  __device__ __forceinline__ double ei(int c) const { return Ei[c - 1]; }
  __device__ __forceinline__ double ej(int c) const { return Ej[c >> 1]; }
  __device__ __forceinline__ double ek(int c) const { return Ek[c]; }
  __device__ __forceinline__ double eij(int) const { return Eij; }
  __device__ __forceinline__ double eik(int) const { return Eik; }
  __device__ __forceinline__ double ejk(int) const { return Ejk; }
};

#ifndef EK_MHD_RUI_LDCG
#define EK_MHD_RUI_LDCG 1
#endif
__device__ __forceinline__ double rui_ld(const double* p)
{
#if EK_MHD_RUI_LDCG
  return __ldcg(p);
#else
  return *p;
#endif
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p)
{
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int RS, bool INVDT>
__global__ void __launch_bounds__(RUI_THREADS, 4) mhd_rui_kernel(const MhdRuiArgs ra)
{
  const MhdArgs& a = ra.a;
  if (a.ts->finished)
    return;
  const Grid& g = a.g;
  const unsigned G = ra.nR + ra.nU;
  const unsigned grp = blockIdx.x / G, r = blockIdx.x - grp * G;
  const unsigned slab = grp / ra.ngroups, gi = grp - slab * ra.ngroups;
  const unsigned ksize = (unsigned)g.ksize;
  const MhdWork w = mhd_work_split(a.work, g.sv);

  if (r < ra.nR) {
    // ---- stage R of slab-plane (slab, gi): the body of mhd_faces_edges_call_kernel --------------------------------
    if (gi >= ksize)
      return; // the lag groups after the last plane hold stage-U CTAs only
    const unsigned t = (slab * ksize + gi) * a.cm.span + r * RUI_THREADS + threadIdx.x;
    MhdIdx x;
    if (mhd_chunk_cell(g, a.cm, t, x) && mhd_stage_r_active(g, x)) {
      const double* __restrict__ S = w.S;
      const unsigned c = x.c, sj = x.sj, sk = x.sk;
      const double* __restrict__ S000 = S + tile_off<S_TOTAL>(c);
      const double* __restrict__ S100 = S + tile_off<S_TOTAL>(c - 1);
      const double* __restrict__ S010 = S + tile_off<S_TOTAL>(c - sj);
      const double* __restrict__ S001 = S + tile_off<S_TOTAL>(c - sk);
      double* __restrict__ Fc = w.F + tile_off<15>(c);
      double* __restrict__ Ec = w.EMF + tile_off<3>(c);
      call_face<0, RS>(g, S100, S000, Fc);
      call_edge<2>(g, S + tile_off<S_TOTAL>(c - 1 - sj), S010, S100, S000, Ec + 2 * TL);
      call_face<1, RS>(g, S010, S000, Fc);
      call_edge<0>(g, S + tile_off<S_TOTAL>(c - sj - sk), S001, S010, S000, Ec);
      call_face<2, RS>(g, S001, S000, Fc);
      call_edge<1>(g, S + tile_off<S_TOTAL>(c - sk - 1), S100, S001, S000, Ec + TL);
    }
    __syncthreads(); // every store of this CTA has been issued ...
    if (threadIdx.x == 0) {
      __threadfence(); // ... and is visible GPU-wide before the counter moves
      atomicAdd(ra.flags + slab * ksize + gi, 1u);
    }
    return;
  }

  // ---- stage U of slab-plane (slab, k = gi - lag), rows one below the slab's stage-R rows --------------------------
  // A stage-U CTA holds a register slot stage R could use, so it has to be SHORT: one cell per thread, every operand of
  // the update (11 + 30 + 12 values) loaded up front with L2-only loads (they must not evict stage R's lines from L1) ->
  // two memory round trips per CTA: U^n beside the poll of the counters, then the fluxes / EMFs.
  if (gi < ra.lag)
    return;
  const unsigned k = gi - ra.lag;
  const int gw = g.gw;
  if ((int)k < gw || (int)k >= g.ksize - gw)
    return; // no interior cell in this plane
  const unsigned sj = (unsigned)g.isize, sk = (unsigned)g.sk;
  const unsigned yc = a.cm.yc;
  const unsigned j0 = slab * yc;
  const unsigned j1 = (j0 + yc < (unsigned)g.jsize) ? j0 + yc : (unsigned)g.jsize;
  const unsigned ju0 = (j0 > 0u) ? j0 - 1u : 0u; // rows [ju0, j1-1)
  const unsigned c0 = sj * ju0 + sk * k, c1 = sj * (j1 - 1u) + sk * k;
  const unsigned c = (c0 & ~31u) + (r - ra.nR) * RUI_THREADS + threadIdx.x;
  MhdIdx x;
  x.c = c;
  x.sj = sj;
  x.sk = sk;
  x.k = (int)k;
  {
    const unsigned rr = c - k * sk;
    const unsigned j = rr / sj;
    x.j = (int)j;
    x.i = (int)(rr - j * sj);
  }
  const bool mine = c >= c0 && c < c1 && x.i >= gw && x.i < g.isize - gw && x.j >= gw && x.j < g.jsize - gw;
  MhdUpdRegs P;
  const long long sv = g.sv;
  if (mine) {
    const double* __restrict__ Uc = a.Uin + c;
#pragma unroll
    for (int v = 0; v < 8; ++v)
      P.U[v] = rui_ld(Uc + sv * v);
    if constexpr (INVDT) {
      P.U[8] = rui_ld(Uc + sv * IA + 1);
      P.U[9] = rui_ld(Uc + sv * IB + sj);
      P.U[10] = rui_ld(Uc + sv * IC + sk);
    }
  }
  {
    const unsigned nflag = (slab > 0u) ? 4u : 2u; // (s, k), (s, k+1), (s-1, k), (s-1, k+1): polled by four threads at once
    if (threadIdx.x < nflag) {
      const unsigned* f = ra.flags + (slab - (threadIdx.x >> 1)) * ksize + k + (threadIdx.x & 1u);
      while (ld_acquire_u32(f) < ra.nR)
        __nanosleep(64);
    }
  }
  __syncthreads();
  double m = 0.0;
  if (mine) {
    const double* __restrict__ F0 = w.F + tile_off<15>(c);
    const double* __restrict__ Fi = w.F + tile_off<15>(c + 1);
    const double* __restrict__ Fj = w.F + tile_off<15>(c + sj);
    const double* __restrict__ Fk = w.F + tile_off<15>(c + sk);
    const double* __restrict__ E0 = w.EMF + tile_off<3>(c);
    const double* __restrict__ Ei = w.EMF + tile_off<3>(c + 1);
    const double* __restrict__ Ej = w.EMF + tile_off<3>(c + sj);
    const double* __restrict__ Ek = w.EMF + tile_off<3>(c + sk);
#pragma unroll
    for (int v = 0; v < 15; ++v)
      P.F0[v] = rui_ld(F0 + TL * v);
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      P.Fxi[v] = rui_ld(Fi + TL * v);
      P.Fyj[v] = rui_ld(Fj + TL * (5 + v));
      P.Fzk[v] = rui_ld(Fk + TL * (10 + v));
    }
#pragma unroll
    for (int v = 0; v < 3; ++v)
      P.E0[v] = rui_ld(E0 + TL * v);
    P.Ei[0] = rui_ld(Ei + TL * 1);
    P.Ei[1] = rui_ld(Ei + TL * 2);
    P.Ej[0] = rui_ld(Ej + TL * 0);
    P.Ej[1] = rui_ld(Ej + TL * 2);
    P.Ek[0] = rui_ld(Ek + TL * 0);
    P.Ek[1] = rui_ld(Ek + TL * 1);
    if constexpr (INVDT) {
      P.Eij = rui_ld(w.EMF + tile_off<3>(c + 1 + sj) + TL * 2);
      P.Eik = rui_ld(w.EMF + tile_off<3>(c + 1 + sk) + TL * 1);
      P.Ejk = rui_ld(w.EMF + tile_off<3>(c + sj + sk) + TL * 0);
    }
    const double dt = a.ts->dt;
    m = mhd_update_core<INVDT>(g, a.Uin, a.Uout, P, dt / g.dx, dt / g.dy, dt / g.dz, x);
  }
  if constexpr (INVDT) {
    __shared__ double red[RUI_THREADS / 32];
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0)
      red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      double q = red[0];
#pragma unroll
      for (int n = 1; n < RUI_THREADS / 32; ++n)
        q = dmax(q, red[n]);
      if (q > 0.0)
        atomic_max_nonneg(&a.ts->invdt_bits, q);
    }
  }
}

// Stage U as a kernel of its own with the operands preloaded as above (plain, L1-cached loads): every load of a thread is
// in flight at once instead of one batch per conserved variable (in mhd_update_kernel the stores of variable v fence the
// loads of variable v+1: the flux pointers are not provably distinct from the output).  ~110 registers: 2 CTAs per SM.
template <bool INVDT>
__global__ void __launch_bounds__(256, 2) mhd_update_pre_kernel(const MhdArgs a)
{
  if (a.ts->finished)
    return;
  const Grid& g = a.g;
  const MhdWork w = mhd_work_split(a.work, g.sv);
  const int gw = g.gw;
  MhdIdx x;
  double m = 0.0;
  if (mhd_cell_of_thread(a, x) && x.k >= a.upd_klo && x.k < a.upd_khi && x.i >= gw && x.j >= gw && x.k >= gw && x.i < g.isize - gw &&
      x.j < g.jsize - gw && x.k < g.ksize - gw) {
    const unsigned c = x.c, sj = x.sj, sk = x.sk;
    const long long sv = g.sv;
    MhdUpdRegs P;
    const double* __restrict__ Uc = a.Uin + c;
    const double* __restrict__ F0 = w.F + tile_off<15>(c);
    const double* __restrict__ Fi = w.F + tile_off<15>(c + 1);
    const double* __restrict__ Fj = w.F + tile_off<15>(c + sj);
    const double* __restrict__ Fk = w.F + tile_off<15>(c + sk);
    const double* __restrict__ E0 = w.EMF + tile_off<3>(c);
    const double* __restrict__ Ei = w.EMF + tile_off<3>(c + 1);
    const double* __restrict__ Ej = w.EMF + tile_off<3>(c + sj);
    const double* __restrict__ Ek = w.EMF + tile_off<3>(c + sk);
#pragma unroll
    for (int v = 0; v < 8; ++v)
      P.U[v] = Uc[sv * v];
#pragma unroll
    for (int v = 0; v < 15; ++v)
      P.F0[v] = F0[TL * v];
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      P.Fxi[v] = Fi[TL * v];
      P.Fyj[v] = Fj[TL * (5 + v)];
      P.Fzk[v] = Fk[TL * (10 + v)];
    }
#pragma unroll
    for (int v = 0; v < 3; ++v)
      P.E0[v] = E0[TL * v];
    P.Ei[0] = Ei[TL * 1];
    P.Ei[1] = Ei[TL * 2];
    P.Ej[0] = Ej[TL * 0];
    P.Ej[1] = Ej[TL * 2];
    P.Ek[0] = Ek[TL * 0];
    P.Ek[1] = Ek[TL * 1];
    if constexpr (INVDT) {
      P.U[8] = Uc[sv * IA + 1];
      P.U[9] = Uc[sv * IB + sj];
      P.U[10] = Uc[sv * IC + sk];
      P.Eij = w.EMF[tile_off<3>(c + 1 + sj) + TL * 2];
      P.Eik = w.EMF[tile_off<3>(c + 1 + sk) + TL * 1];
      P.Ejk = w.EMF[tile_off<3>(c + sj + sk) + TL * 0];
    }
    const double dt = a.ts->dt;
    m = mhd_update_core<INVDT>(g, a.Uin, a.Uout, P, dt / g.dx, dt / g.dy, dt / g.dz, x);
  }
  if constexpr (INVDT)
    mhd_block_max_invdt(m, a.ts);
}

} // namespace ek
