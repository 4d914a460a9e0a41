// Stages P + E + T of the CT-MHD step, fed by TMA (sm_100a): the conserved state U reaches the SM only through
// cp.async.bulk.tensor box loads that an mbarrier completes; no thread of this kernel issues a global LOAD.
//
// The first fused version (mhd_pet.cuh, kept for grids whose row stride is not a multiple of 16 bytes) issues about 45
// global loads per thread and plane (11 per converted cell, 18 + 9 face-field values for the predictor and the edge
// field); ncu at 512^3: 4.4 long-scoreboard stall cycles per issued instruction at 16 warps per SM, L1 hit rate 20 %
// (the 214 KB of shared memory leave ~30 KB of L1), FP64 pipe 22 %, DRAM 40 %.  Here
//   * one elected thread issues, per plane, nine 4-D box loads (36 x 15 cells of one variable each: the 8 variables of
//     the plane and the field component C of the plane above -- the cell-centred C needs the upper z face) into one of
//     TWO staging buffers: two planes (2 x 39 KB) in flight while the CTA works on the planes before them;
//   * the ring of three planes in shared memory keeps, per cell, r, p, u, v, w and the three FACE fields A, B, C --
//     everything the predictor and the edge field read (cell-centred B = mean of two faces, formed on the fly with
//     the expression cons->prim uses), so their 27 face-field loads are shared-memory reads of data that is there anyway;
//   * a CTA is 32 x 12 cells (384 threads, one per SM) => box 34 x 14.  The first TMA version used 32 x 16 cells = 512
//     threads, which caps the kernel at 128 registers per thread; the predictor wants ~168 (ptxas with the cap lifted),
//     and with 384 threads it gets them: 14.9 -> 11.9 ms at 512^3 (tiles of 16 / 14 / 12 / 10 / 8 rows measured:
//     50.8 / 49.7 / 48.7-49.0 / 49.2 / 49.7 ms per step; two CTAs of 8 rows per SM at 128 registers: 52.7).
// Per plane: E(p+1) from the ring | barrier | predictor of plane p -> S | barrier | staged plane p+2 -> ring slot of
// plane p-1 | barrier | TMA of plane p+4 into the buffer plane p+2 came from.  The per-cell arithmetic is the code of
// mhd_cells.cuh (mhd_cons_to_prim, mhd_elec_core, mhd_trace_core).
#pragma once
#include <cuda.h> // CUtensorMap (type only: the encoder is looked up at run time, ek_api.cu)
#include "mhd_cells.cuh"

namespace ek {

#ifndef EK_PET_TY
#define EK_PET_TY 12 // tile rows (CTA = 32 x EK_PET_TY threads)
#endif
#ifndef EK_PET_NSTAGE
#define EK_PET_NSTAGE 2 // TMA staging buffers (planes in flight ahead of the conversion); 2 needs EK_PET_TY <= 14 to fit 227 KB
#endif
#ifndef EK_PET_MINB
#define EK_PET_MINB 1 // resident CTAs per SM the register budget is cut for
#endif
constexpr int PT_TX = 32, PT_TY = EK_PET_TY, PT_NS = EK_PET_NSTAGE, PT_MINB = EK_PET_MINB;
static_assert(PT_NS == 1 || PT_NS == 2, "one or two staging buffers");
constexpr int PT_BX = PT_TX + 2, PT_BY = PT_TY + 2, PT_BOX = PT_BX * PT_BY;  // tile + 1-cell ring (34 x 14 cells)
constexpr int PT_TILES = (PT_BOX + 31) / 32;                                  // AoSoA tiles per plane of the ring
constexpr int PT_QPLANE = PT_TILES * 8 * 32, PT_EPLANE = PT_TILES * 3 * 32;   // doubles
constexpr int PT_RX = 36, PT_RY = PT_BY + 1;                                  // staged box: columns i0-2 .. i0+33 (the box start must be
                                                                              // 16-byte aligned: an odd start column is an illegal instruction), PT_TY+3 rows
constexpr int PT_RBOX = PT_RX * PT_RY;                                        // doubles per staged variable (one TMA box)
constexpr int PT_RPLANE = (PT_RBOX + 15) / 16 * 16;                           // ... padded: every box lands 128-byte aligned
constexpr unsigned PT_STAGE_BYTES = 9u * PT_RBOX * sizeof(double);            // 8 variables of plane k + C of plane k+1

struct PetTmaSmem {
  double raw[PT_NS][9][PT_RPLANE];  // TMA destinations (128-byte aligned: first member)
  double q[3][PT_QPLANE];    // ring: r, p, u, v, w, A_face, B_face, C_face of planes p-1, p, p+1
  double e[2][PT_EPLANE];    // edge electric field of planes p, p+1
  unsigned long long bar[PT_NS]; // mbarrier of each staging buffer
};

__device__ __forceinline__ unsigned smem_addr_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
  asm volatile(
    "{\n"
    ".reg .pred P1;\n"
    "LAB_WAIT:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
    "@P1 bra DONE;\n"
    "bra LAB_WAIT;\n"
    "DONE:\n"
    "}" ::"r"(smem_addr_u32(bar)), "r"(parity) : "memory");
}
// box load of a 4-D tensor (x, y, z, variable) into shared memory; completion = bytes on the mbarrier
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1, int c2, int c3)
{
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                 smem_addr_u32(dst)),
               "l"((unsigned long long)map), "r"(smem_addr_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// mhd_trace_core's input interface over the shared-memory ring: the cell-centred field components are the mean of the
// two faces of the cell (constoprim_mhd, MHDBaseFunctor3D.h:207-215), formed from the face values the ring holds
struct PetRingIn {
  const double *qm, *q0p, *qp; // ring planes p-1, p, p+1
  const double *e0p, *e1p;     // edge field of planes p, p+1
  unsigned b;                  // box cell of this thread
  __device__ __forceinline__ static double at(const double* plane, unsigned cell, int v) { return plane[tile_off<8>(cell) + TL * v]; }
  // variable v of `cell` in ring plane `pl` (`up` = the plane above it, for C)
  __device__ __forceinline__ static double prim(const double* pl, const double* up, unsigned cell, int v)
  {
    if (v == IA)
      return 0.5 * (at(pl, cell, IA) + at(pl, cell + 1, IA));
    if (v == IB)
      return 0.5 * (at(pl, cell, IB) + at(pl, cell + PT_BX, IB));
    if (v == IC)
      return 0.5 * (at(pl, cell, IC) + at(up, cell, IC));
    return at(pl, cell, v);
  }
  __device__ __forceinline__ double q0(int v) const { return prim(q0p, qp, b, v); }
  __device__ __forceinline__ double qxm(int v) const { return prim(q0p, qp, b - 1, v); }
  __device__ __forceinline__ double qxp(int v) const { return prim(q0p, qp, b + 1, v); }
  __device__ __forceinline__ double qym(int v) const { return prim(q0p, qp, b - PT_BX, v); }
  __device__ __forceinline__ double qyp(int v) const { return prim(q0p, qp, b + PT_BX, v); }
  // the z slopes never ask for C (MHDBaseFunctor3D.h:364-495: the normal component has no slope along its normal)
  __device__ __forceinline__ double qzm(int v) const { return prim(qm, q0p, b, v); }
  __device__ __forceinline__ double qzp(int v) const { return prim(qp, qp, b, v); }
  __device__ __forceinline__ double e0(int c) const { return e0p[tile_off<3>(b) + TL * c]; }
  __device__ __forceinline__ double ei(int c) const { return e0p[tile_off<3>(b + 1) + TL * c]; }
  __device__ __forceinline__ double ej(int c) const { return e0p[tile_off<3>(b + PT_BX) + TL * c]; }
  __device__ __forceinline__ double ek(int c) const { return e1p[tile_off<3>(b) + TL * c]; }
};

__global__ void __launch_bounds__(PT_TX* PT_TY, PT_MINB)
  mhd_pet_tma_kernel(const MhdArgs a, const int zchunk, const __grid_constant__ CUtensorMap umap)
{
  extern __shared__ __align__(1024) unsigned char pet_tma_raw[];
  PetTmaSmem& sm = *reinterpret_cast<PetTmaSmem*>(pet_tma_raw);
  if (a.ts->finished)
    return;
  const Grid& g = a.g;
  const MhdWork w = mhd_work_split(a.work, g.sv);
  const double dt = a.ts->dt;
  const double dtdx = dt / g.dx, dtdy = dt / g.dy, dtdz = dt / g.dz;
  const int t = threadIdx.x;
  const int i0 = 2 + PT_TX * blockIdx.x, j0 = 2 + PT_TY * (int)blockIdx.y; // first cell of the tile (ghosted coordinates)
  const int kbeg = 2 + (int)blockIdx.z * zchunk;
  const int kend = min(kbeg + zchunk, g.ksize - 1); // the predictor runs on planes 2 .. ksize-2
  if (kbeg >= kend)
    return;
  const unsigned sj = (unsigned)g.isize, sk = (unsigned)g.sk;
  constexpr int NT = PT_TX * PT_TY;
  static_assert(PT_BOX > NT && PT_BOX <= 2 * NT, "two passes over the box");

  if (t == 0) {
#pragma unroll
    for (int b = 0; b < PT_NS; ++b)
      mbar_init(&sm.bar[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // The planes kbeg-1 .. kend are staged in order; the n-th of them goes through buffer n % PT_NS (parity of that
  // buffer's mbarrier: (n / PT_NS) & 1).  A plane is issued as soon as the buffer it lands in has been converted.
  int n_issued = 0, n_conv = 0;
  // plane k (8 variables) + C of plane k+1 into a staging buffer; out-of-range parts of a box arrive as zeros
  auto issue_next = [&]() {
    const int k = kbeg - 1 + n_issued;
    if (k > kend)
      return; // never leave a load in flight at exit
    const int b = n_issued % PT_NS;
    ++n_issued;
    if (t == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the buffer was read through the generic proxy
      mbar_arrive_expect_tx(&sm.bar[b], PT_STAGE_BYTES);
#pragma unroll
      for (int v = 0; v < 8; ++v)
        tma_load_4d(&sm.raw[b][v][0], &umap, &sm.bar[b], i0 - 2, j0 - 1, k, v);
      tma_load_4d(&sm.raw[b][8][0], &umap, &sm.bar[b], i0 - 2, j0 - 1, k + 1, IC);
    }
  };
  // staged plane -> ring slot: primitives r, p, u, v, w + the three face fields of every box cell
  auto convert_to = [&](double* __restrict__ qs) {
    const int sb = n_conv % PT_NS;
    mbar_wait(&sm.bar[sb], (unsigned)(n_conv / PT_NS) & 1u);
    ++n_conv;
    const double(*__restrict__ raw)[PT_RPLANE] = sm.raw[sb];
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int b = t + pass * NT;
      if (b < PT_BOX) {
        const int bj = b / PT_BX, bi = b - bj * PT_BX;
        const int o = bj * PT_RX + bi + 1; // staged column 0 is cell i0-2, box column 0 is cell i0-1
        double u[8], q[8];
#pragma unroll
        for (int v = 0; v < 8; ++v)
          u[v] = raw[v][o];
        mhd_cons_to_prim(g, u, raw[IA][o + 1], raw[IB][o + PT_RX], raw[8][o], q);
        double* __restrict__ qc = qs + tile_off<8>((unsigned)b);
#pragma unroll
        for (int v = 0; v < 5; ++v)
          qc[TL * v] = q[v];
        qc[TL * IA] = u[IA];
        qc[TL * IB] = u[IB];
        qc[TL * IC] = u[IC];
      }
    }
  };
  // edge field of plane k from ring planes k-1 (qlo) and k (qhi) -> es
  auto elec_plane = [&](int k, const double* __restrict__ qlo, const double* __restrict__ qhi, double* __restrict__ es) {
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int b = t + pass * NT;
      if (b >= PT_BOX)
        continue;
      const int bj = b / PT_BX, bi = b - bj * PT_BX;
      MhdIdx x;
      x.i = i0 - 1 + bi;
      x.j = j0 - 1 + bj;
      x.k = k;
      if (!(k <= g.ksize - 2 && bi >= 1 && bj >= 1 && mhd_elec_active(g, x)))
        continue;
      const unsigned ub = (unsigned)b;
      MhdElecQ P;
      P.q000 = qhi + tile_off<8>(ub);
      P.q100 = qhi + tile_off<8>(ub - 1);
      P.q010 = qhi + tile_off<8>(ub - PT_BX);
      P.q110 = qhi + tile_off<8>(ub - 1 - PT_BX);
      P.q001 = qlo + tile_off<8>(ub);
      P.q101 = qlo + tile_off<8>(ub - 1);
      P.q011 = qlo + tile_off<8>(ub - PT_BX);
      MhdElecB f;
      f.A0 = P.q000[TL * IA]; f.B0 = P.q000[TL * IB]; f.C0 = P.q000[TL * IC];
      f.Aym = P.q010[TL * IA]; f.Cym = P.q010[TL * IC];
      f.Bxm = P.q100[TL * IB]; f.Cxm = P.q100[TL * IC];
      f.Azm = P.q001[TL * IA]; f.Bzm = P.q001[TL * IB];
      mhd_elec_core(f, P, es + tile_off<3>(ub));
    }
  };

  // ring slot of plane k: k % 3
  // prologue: planes kbeg-1, kbeg, kbeg+1 into the ring, E(kbeg); plane kbeg+2 left in flight
#pragma unroll
  for (int b = 0; b < PT_NS; ++b)
    issue_next();
  convert_to(sm.q[(kbeg - 1) % 3]);
  __syncthreads();
  issue_next();
  convert_to(sm.q[kbeg % 3]);
  __syncthreads();
  issue_next();
  elec_plane(kbeg, sm.q[(kbeg - 1) % 3], sm.q[kbeg % 3], sm.e[kbeg & 1]);
  convert_to(sm.q[(kbeg + 1) % 3]);
  __syncthreads();
  issue_next(); // plane kbeg+2 (+PT_NS-1): converted by the first iterations of the loop

  const int lx = t & (PT_TX - 1), ly = t / PT_TX;
  const unsigned bt = (unsigned)((1 + lx) + PT_BX * (1 + ly)); // this thread's cell inside the box
  for (int p = kbeg; p < kend; ++p) {
    const double* __restrict__ qm = sm.q[(p - 1) % 3];
    const double* __restrict__ q0 = sm.q[p % 3];
    const double* __restrict__ qp = sm.q[(p + 1) % 3];
    elec_plane(p + 1, q0, qp, sm.e[(p + 1) & 1]);
    __syncthreads();
    {
      MhdIdx x;
      x.i = i0 + lx;
      x.j = j0 + ly;
      x.k = p;
      x.sj = sj;
      x.sk = sk;
      x.c = (unsigned)x.i + sj * (unsigned)x.j + sk * (unsigned)p;
      if (x.i <= g.isize - 2 && x.j <= g.jsize - 2 && mhd_trace_active(g, x)) {
        PetRingIn P;
        P.qm = qm;
        P.q0p = q0;
        P.qp = qp;
        P.e0p = sm.e[p & 1];
        P.e1p = sm.e[(p + 1) & 1];
        P.b = bt;
        // the face-centred field values of the predictor (18 global loads before): ring reads
        MhdTraceB fb;
        const unsigned bxm = bt - 1, bxp = bt + 1, bym = bt - PT_BX, byp = bt + PT_BX;
        fb.A0 = PetRingIn::at(q0, bt, IA); fb.Axp = PetRingIn::at(q0, bxp, IA); fb.Aym = PetRingIn::at(q0, bym, IA);
        fb.Ayp = PetRingIn::at(q0, byp, IA); fb.Azm = PetRingIn::at(qm, bt, IA); fb.Azp = PetRingIn::at(qp, bt, IA);
        fb.B0 = PetRingIn::at(q0, bt, IB); fb.Bxm = PetRingIn::at(q0, bxm, IB); fb.Bxp = PetRingIn::at(q0, bxp, IB);
        fb.Byp = PetRingIn::at(q0, byp, IB); fb.Bzm = PetRingIn::at(qm, bt, IB); fb.Bzp = PetRingIn::at(qp, bt, IB);
        fb.C0 = PetRingIn::at(q0, bt, IC); fb.Cxm = PetRingIn::at(q0, bxm, IC); fb.Cxp = PetRingIn::at(q0, bxp, IC);
        fb.Cym = PetRingIn::at(q0, bym, IC); fb.Cyp = PetRingIn::at(q0, byp, IC); fb.Czp = PetRingIn::at(qp, bt, IC);
        mhd_trace_core(g, fb, P, w.S + tile_off<S_TOTAL>(x.c), dtdx, dtdy, dtdz, x);
      }
    }
    if (p + 1 < kend) {
      __syncthreads();                      // every reader of ring plane p-1 is done
      convert_to(sm.q[(p + 2) % 3]);        // staged plane p+2 replaces plane p-1
      __syncthreads();                      // ring plane p+2 visible, its staging buffer free
      issue_next();                         // plane p+2+PT_NS
    }
  }
}

} // namespace ek
