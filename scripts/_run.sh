set -x
for n in 1 2 3; do E2_WGRAD_STREAMS=$n python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r58_bench_s$n.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r58_bench_s$n.log | head -1; done
