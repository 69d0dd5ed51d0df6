set -x
E2_ZS_INFO=1 python scripts/tc_check.py > gpurun_out/r47_tc_check.log 2>&1; grep -c OK gpurun_out/r47_tc_check.log; grep -E "FAIL|Error|error|ALL" gpurun_out/r47_tc_check.log | head
python -m pytest tests -m gpu -x -q > gpurun_out/r47_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r47_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/r47_prof_unet3d.json > gpurun_out/r47_bench.log 2>&1; tail -c 300 gpurun_out/r47_bench.log
E2_ZS_NARROW=1 E2_ZS_NOSPLIT=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/r47_prof_unet3d_old.json > gpurun_out/r47_bench_old.log 2>&1; tail -c 300 gpurun_out/r47_bench_old.log
