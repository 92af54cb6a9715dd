#!/bin/bash
# Multi-GPU pass: bash scripts/gpu_multi.sh N [tests].  Everything under its own timeout (a hung rank must not
# hold the box).  Outputs in gpurun_out/.
N=$1
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
if [ "$2" = "tests" ]; then
  timeout 500 python -m pytest tests/test_gpu_tsqr.py -m gpu -q -rf --no-header > $O/r02_tests_n$N.log 2>&1; echo "tests rc=$?"; tail -4 $O/r02_tests_n$N.log
fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > $O/r02_bench_n$N.json 2> $O/r02_bench_n$N.err; echo "bench n$N rc=$?"; tail -c 600 $O/r02_bench_n$N.json
LB_PEER_ALLREDUCE=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > $O/r02_bench_n${N}_nccl.json 2> $O/r02_bench_n${N}_nccl.err; echo "bench n$N nccl rc=$?"; tail -c 300 $O/r02_bench_n${N}_nccl.json
