set -x
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r68_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r68_pytest.log
timeout 300 python bench_dense.py --max-tiles 2 --profile-out gpurun_out/r68_dense_prof.json > gpurun_out/r68_dense.log 2>&1; echo rc=$?; grep -o '"value": [0-9.]*' gpurun_out/r68_dense.log | head -1
