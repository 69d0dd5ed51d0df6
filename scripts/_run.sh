set -x
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r64_a.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r64_a.log | tr '\n' ' '; echo prio
E2_MAIN_PRIO=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r64_b.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r64_b.log | tr '\n' ' '; echo noprio
E2_TAIL_LAYERS=3 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r64_c.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r64_c.log | tr '\n' ' '; echo tail3
E2_TAIL_LAYERS=6 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r64_d.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r64_d.log | tr '\n' ' '; echo tail6
