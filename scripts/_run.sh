set -x
python scripts/c1_tc_bench.py 2>&1 | tail -12
python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "conv" 2>&1 | tail -3
