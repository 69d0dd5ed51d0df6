set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r62_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r62_pytest.log
for i in 1 2; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r62_bench_$i.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r62_bench_$i.log | tr '\n' ' '; grep -o '"loss": [0-9.]*' gpurun_out/r62_bench_$i.log; done
tail -3 gpurun_out/r62_bench_1.log | cut -c1-200
