set -x
timeout 500 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "inside" > gpurun_out/r70_pytest.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/r70_pytest.log
