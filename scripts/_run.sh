set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r56_dp2.log 2>&1; echo rc=$?; grep -o '{"metric.*' gpurun_out/r56_dp2.log | cut -c1-400; grep -o '"e2e".*' gpurun_out/r56_dp2.log | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r56_ref2.log 2>&1; echo rc=$?; tail -c 300 gpurun_out/r56_ref2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench_dense.py --max-tiles 6 > gpurun_out/r56_dense2.log 2>&1; echo rc=$?; grep -o '{"metric.*' gpurun_out/r56_dense2.log | cut -c1-500
