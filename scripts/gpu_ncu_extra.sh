#!/bin/bash
# ncu passes that need a kernel-name filter: the launch list of bench.py restricted to the library's own kernels
# (the input generator alone launches ~1000 torch kernels), and full captures of template instances.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
OURS='regex:tc_|gram128|chol128|apply128|trinv128|colmax128|splitk_reduce|cast_shadow|finish_r12|peer_allreduce|gram32|chol32|apply32'
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 2700 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-e2e > $O/r02_bench_under_ncu.log 2>&1; echo "ncu bench rc=$?"
cap() {  # name demangled-regex count m n
  timeout 240 ncu --set full --clock-control none --kernel-name-base demangled -k "regex:$2" -c $3 -f -o $O/prof_$1 \
      python scripts/gpu_profile_run.py $4 $5 1 > $O/prof_$1.log 2>&1; echo "ncu $1 rc=$?"; tail -1 $O/prof_$1.log | cut -c1-150
}
cap gram16k 'tc_gemm_kernel<\(int\)256, \(bool\)0, \(int\)0>' 5 16384 16384
cap update16k 'tc_update_kernel<\(int\)256' 1 16384 16384
cap update_1m 'tc_update_kernel<\(int\)128' 1 1048576 1024
ls -la $O/*.ncu-rep
