#!/bin/bash
# One-GPU validation + measurement pass of a finished tree (run under gpurun): GPU suite with the product library and with the
# verify / debug-bounds build, bench lines of every config, reference arm, shard emulation, ncu launch list and full capture.
# Everything lands in gpurun_out/ (f1_*); the summaries worth keeping are copied to profiles/ by hand.
set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f1_gpu.log 2>&1
AUDIORT_LIB=$PWD/audio-raytracer_b200/libaudiort_verify.so timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f1_verify.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/f1_bench_c3.json 2> gpurun_out/f1_bench_c3.err
for w in c1 c2 c4; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/f1_bench_$w.json 2> gpurun_out/f1_bench_$w.err; done
python bench.py --workload c5 --rays 524288 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f1_bench_c5s.json 2> gpurun_out/f1_bench_c5s.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f1_bench_ref.json 2> gpurun_out/f1_bench_ref.err
python tools/exp_env.py --workload c3 --frames 4 --shard 0/8 --chunk 16384 > gpurun_out/f1_shard.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/f1_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name 'regex:query_fan|bounce_kernel|perm_loss_binned|fan_match|fan_project|fan_order|perm_bin|permeation_grid|perm_last|echo_stats' -c 16 -f -o gpurun_out/f1_full python tools/prof_frame.py --workload c3 --frames 1 > gpurun_out/f1_full.log 2>&1
ls -la gpurun_out
