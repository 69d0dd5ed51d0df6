set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r51_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r51_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/r51_prof_unet3d.json > gpurun_out/r51_bench.log 2>&1; tail -c 600 gpurun_out/r51_bench.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
