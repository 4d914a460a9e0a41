#!/bin/bash
# The one GPU job script of this repo (runs on the B200 box under gpurun; everything it writes goes to gpurun_out/).
#   bash profiles/gpu.sh TAG STEP [STEP ...]
# STEPs (executed in order; each has its own timeout so that a hang cannot eat the whole call):
#   tests            pytest -m gpu of the whole suite            -> ${TAG}_pytest_gpu.log
#   tests:EXPR       ... restricted to `-k EXPR` / a file path
#   bench[:ARGS]     python bench.py ARGS (default workload)      -> ${TAG}_bench[_ARGS].json
#   launches[:ARGS]  ncu launch list (gpu__time_duration) of bench.py --steps 3        -> ${TAG}_launches.csv
#   ncu:REGEX[:ARGS] ncu --set full of 3 launches of the kernels matching REGEX        -> ${TAG}_REGEX.ncu-rep + _raw.csv
#   metrics:REGEX[:ARGS]  per-kernel time, DRAM bytes, FP64 pipe, instruction counts   -> ${TAG}_metrics.csv
#   multi:N[:ARGS]   tests/test_multigpu.py + torchrun bench on N GPUs                 -> ${TAG}_multigpu.log, ${TAG}_bench_nN.json
#   dropin           the C++ driver on 512^3 (Perf line of bin/euler_b200)            -> ${TAG}_dropin.log
# ARGS use ',' for spaces, e.g. bench:--workload,blast3d,--steps,10
O=gpurun_out; mkdir -p $O
TAG=$1; shift
args() { echo "${1//,/ }"; }
for STEP in "$@"; do
  IFS=: read -r KIND A B <<< "$STEP"
  case $KIND in
    tests)
      WHAT=tests; SEL=""
      if [ -n "$A" ]; then if [[ "$A" == tests/* ]]; then WHAT="$A"; else SEL="-k $A"; fi; fi
      timeout 1500 python -m pytest $WHAT -m gpu -q $SEL > $O/${TAG}_pytest_gpu.log 2>&1; echo "[$STEP] rc=$?"; tail -4 $O/${TAG}_pytest_gpu.log ;;
    bench)
      SUF=$(echo "$A" | tr -c 'A-Za-z0-9\n' '_'); F=$O/${TAG}_bench${SUF:+_$SUF}.json
      timeout 900 python bench.py $(args "$A") > $F 2> ${F%.json}.err; echo "[$STEP] rc=$?"; cut -c1-420 $F; tail -2 ${F%.json}.err ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
        python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary $(args "$A") > $O/${TAG}_launches.log 2>&1; echo "[$STEP] rc=$?"
      python profiles/summarize.py launches $O/${TAG}_launches.csv | tee $O/${TAG}_launches.txt ;;
    metrics)
      M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_fp64.sum,sm__warps_active.avg.pct_of_peak_sustained_active
      timeout 900 ncu --metrics $M --clock-control none -k regex:"$A" -s ${MSKIP:-12} -c ${MCOUNT:-12} --csv --log-file $O/${TAG}_metrics.csv \
        python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary $(args "$B") > $O/${TAG}_metrics.log 2>&1; echo "[$STEP] rc=$?"
      python profiles/summarize.py metrics $O/${TAG}_metrics.csv | tee $O/${TAG}_metrics.txt ;;
    ncu)
      N=$(echo "$A" | tr -c 'A-Za-z0-9\n' '_')
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$A" -s 6 -c 3 -f -o $O/${TAG}_$N \
        python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary $(args "$B") > $O/${TAG}_ncu_$N.log 2>&1; echo "[$STEP] rc=$?"
      ncu -i $O/${TAG}_$N.ncu-rep --page raw --csv > $O/${TAG}_${N}_raw.csv 2>/dev/null
      python profiles/summarize.py raw $O/${TAG}_${N}_raw.csv | tee $O/${TAG}_${N}_summary.txt ;;
    multi)
      timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q > $O/${TAG}_multigpu.log 2>&1; echo "[$STEP] pytest rc=$?"; tail -4 $O/${TAG}_multigpu.log
      F=$O/${TAG}_bench_n$A.json
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $A --master-addr 127.0.0.1 --master-port 29533 \
        bench.py --gpus $A --steps 10 --warmup 3 $(args "$B") > $F 2> ${F%.json}.err; echo "[$STEP] bench rc=$?"; cut -c1-420 $F ;;
    refcuda)   # the reference's own Kokkos-Cuda build on this GPU: hydro blast 512^3 and MHD Orszag-Tang 384^3
      timeout 900 python bench.py --impl reference-cuda --workload blast3d --steps 20 --warmup 3 > $O/${TAG}_refcuda_blast3d.json 2> $O/${TAG}_refcuda_blast3d.err; cut -c1-300 $O/${TAG}_refcuda_blast3d.json
      timeout 900 python bench.py --impl reference-cuda --workload ot3d --steps 10 --warmup 3 > $O/${TAG}_refcuda_ot3d.json 2> $O/${TAG}_refcuda_ot3d.err; cut -c1-300 $O/${TAG}_refcuda_ot3d.json
      timeout 900 python bench.py --workload ot3d --size 384 --steps 10 --no-cpu-baseline > $O/${TAG}_bench_ot3d_384.json 2> $O/${TAG}_bench_ot3d_384.err; cut -c1-300 $O/${TAG}_bench_ot3d_384.json ;;
    dropin)
      python - > $O/${TAG}_dropin.ini <<'PY'
import sys
sys.path.insert(0, ".")
import argparse, bench
ini, _ = bench.workload_ini(argparse.Namespace(workload="ot3d", n=512, shape=None), {"run.nstepmax": 30, "run.tend": 1e6, "run.nlog": 10})
ini.write("/dev/stdout")
PY
      timeout 600 euler_kokkos_b200/bin/euler_b200 $O/${TAG}_dropin.ini > $O/${TAG}_dropin.log 2>&1; echo "[$STEP] rc=$?"; grep -E "time step=|Perf|final time" $O/${TAG}_dropin.log | tail -6 ;;
    *) echo "unknown step $STEP" ;;
  esac
done
