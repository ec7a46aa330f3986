#!/bin/bash
# 8-GPU pass of a finished tree (run under `gpurun --gpus 8`): multi-device GPU tests, then the C3 and full-size C5 bench lines.
set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi_device.py -m gpu -x -q > gpurun_out/f8_tests.log 2>&1
for w in c3 c5; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/f8_bench_$w.json 2> gpurun_out/f8_bench_$w.err
done
