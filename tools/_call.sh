timeout 600 python -m pytest tests/test_gpu_binned.py -x -q 2>&1 | tail -15
for w in "c3" "c4" "c5 --rays 524288" "c3 --shard 0/8"; do python tools/prof_frame.py --workload $w --frames 3 2>&1 | tail -1; done
ART_K2_BINNED=0 python tools/prof_frame.py --workload c3 --frames 3 2>&1 | tail -1
