set -x
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r65_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r65_ncu.log 2>&1; echo rc=$?; tail -c 200 gpurun_out/r65_ncu.log
python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/r65_prof_unet3d.json > gpurun_out/r65_bench.log 2>&1; tail -1 gpurun_out/r65_bench.log | cut -c1-1800
