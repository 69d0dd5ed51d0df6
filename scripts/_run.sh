set -x
python scripts/tc_check.py > gpurun_out/r53_tc_check.log 2>&1; grep -c OK gpurun_out/r53_tc_check.log; grep -E "FAIL|Error|error|ALL" gpurun_out/r53_tc_check.log | head; grep upconv gpurun_out/r53_tc_check.log | cut -c1-60,110-200
python -m pytest tests -m gpu -x -q > gpurun_out/r53_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r53_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/r53_prof_unet3d.json > gpurun_out/r53_bench.log 2>&1; tail -c 400 gpurun_out/r53_bench.log
E2_PACK_OVERLAP=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r53_bench_nooverlap.log 2>&1; tail -c 400 gpurun_out/r53_bench_nooverlap.log
