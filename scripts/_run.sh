set -x
python scripts/tc_check.py > gpurun_out/r46_tc_check.log 2>&1; tail -8 gpurun_out/r46_tc_check.log
python -m pytest tests -m gpu -x -q > gpurun_out/r46_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r46_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/r46_prof_unet3d.json > gpurun_out/r46_bench.log 2>&1; tail -c 300 gpurun_out/r46_bench.log
python scripts/prof_layer.py 32 114 130 130 64 dgrad 3
ncu --set full --clock-control none --import-source on -k regex:k_conv_zstack -s 2 -c 1 -o gpurun_out/r46_conv1_dgrad python scripts/prof_layer.py 32 114 130 130 64 dgrad 3 > gpurun_out/r46_ncu_a.log 2>&1
