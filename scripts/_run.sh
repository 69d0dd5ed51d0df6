set -x
for i in 1 2 3; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r60_bench_$i.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r60_bench_$i.log | tr '\n' ' '; grep -o '"clocks": {[^}]*}' gpurun_out/r60_bench_$i.log; done
