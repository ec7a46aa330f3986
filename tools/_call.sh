mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.log 2>&1; tail -8 gpurun_out/final_pytest.log
