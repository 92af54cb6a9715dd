#!/bin/bash
# Final single-GPU measurement pass of a round: tests, both bench arms, the ncu launch list of bench.py, full
# ncu captures of the dominant kernels, compute-sanitizer runs.  Everything under its own timeout; outputs in
# gpurun_out/ (scripts/make_profiles.py r02 turns them into profiles/).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -rf --no-header > $O/r02_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r02_tests.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err; echo "bench ref rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 > $O/r02_bench_b200.json 2> $O/r02_bench_b200.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-e2e > $O/r02_bench_under_ncu.log 2>&1; echo "ncu bench rc=$?"
cap() {  # name regex count m n
  timeout 240 ncu --set full --clock-control none -k "regex:$2" -c $3 -f -o $O/prof_$1 \
      python scripts/gpu_profile_run.py $4 $5 1 > $O/prof_$1.log 2>&1; echo "ncu $1 rc=$?"
}
cap gram16k 'tc_gemm_kernel<256, *false, *0>|tc_gemm_kernelILi256ELb0ELi0' 5 16384 16384
cap update16k 'tc_update_kernel<256' 1 16384 16384
cap chol16k 'chol128b_kernel' 1 16384 16384
cap gram_f64_16k 'gram128_f64_kernel' 1 16384 16384
cap apply16k 'apply128_kernel' 1 16384 16384
cap gram_i8_1m 'gram128_i8_kernel' 1 1048576 1024
cap apply_tc_1m 'apply128_tc_kernel' 1 1048576 1024
cap gramcast_1m 'tc_gram_cast_kernel' 3 1048576 1024
cap update_1m 'tc_update_kernel<128' 1 1048576 1024
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool python scripts/gpu_sanitize_run.py square panel panel32 host ormqr qdwh > $O/r02_sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; tail -2 $O/r02_sanitize_$tool.log
done
timeout 400 compute-sanitizer --tool memcheck python scripts/gpu_sanitize_run.py tall > $O/r02_sanitize_memcheck_tall.log 2>&1; echo "memcheck tall rc=$?"; tail -2 $O/r02_sanitize_memcheck_tall.log
timeout 400 compute-sanitizer --tool racecheck python scripts/gpu_sanitize_run.py tall > $O/r02_sanitize_racecheck_tall.log 2>&1; echo "racecheck tall rc=$?"; tail -2 $O/r02_sanitize_racecheck_tall.log
