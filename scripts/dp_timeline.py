#!/usr/bin/env python
"""CUDA-event timeline of ONE eager data-parallel training step (no nsys in this image): when each gradient bucket's
all-reduce starts / ends on the comm stream relative to the start of the backward pass, and when the compute streams are
done -- i.e. which collective is exposed at the end of the step and by how much.

    torchrun --nproc-per-node N scripts/dp_timeline.py [out.json]
"""
import contextlib, io, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elektronn2_b200 import examples, parallel, neuromancer as nm
from elektronn2_b200.config import config as e2cfg

rank, world, local = parallel.init_from_env()
torch.cuda.set_device(local)
dist = torch.distributed
e2cfg.use_cuda_graph = False
np.random.seed(2)
with contextlib.redirect_stdout(io.StringIO()):
    model = examples.unet3d()
dp = parallel.DataParallel(model)
nm.optimiser.Optimiser.setlr(5e-4), nm.optimiser.Optimiser.setwd(0.5e-4), nm.optimiser.Optimiser.setmom(0.9)
ish = [1 if s is None else s for s in model.input_node.shape.shape]
tsh = [1 if s is None else s for s in model.target_node.shape.shape]
x = np.random.RandomState(1000 + rank).rand(*ish).astype(np.float32)
t = np.random.RandomState(1001 + rank).randint(0, 2, tsh).astype(np.float32)
plan = model._train_plan(1)
dp.broadcast_parameters(plan.store)
opt = model.optimisers['Adam']
plan.feed({model.input_node: x, model.target_node: t})
for _ in range(3):
    plan.train_step(opt)
torch.cuda.synchronize()
dist.barrier()
# one traced step: forward, then the backward body with events around it
opt.dev_sync(plan.store)
plan._train_body_fwd(opt)
torch.cuda.synchronize()
dist.barrier()
dp.trace = []
b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
b0.record()
plan._train_body_bwd(opt, dp)
b1.record()
torch.cuda.synchronize()
rows = [dict(floats=[int(s), int(e)], mbytes=(e - s) * 4 / 1e6, start_us=b0.elapsed_time(t0) * 1e3, end_us=b0.elapsed_time(t1) * 1e3)
        for s, e, t0, t1 in dp.trace]
out = dict(world=world, rank=rank, backward_total_us=b0.elapsed_time(b1) * 1e3, buckets=rows,
           last_collective_end_us=max(r['end_us'] for r in rows) if rows else None,
           note='eager step (no CUDA graph), Adam + re-pack inside the timed region; times relative to the first backward launch')
allr = [None] * world
dist.all_gather_object(allr, out)
if rank == 0:
    txt = json.dumps(dict(ranks=allr), indent=1)
    if len(sys.argv) > 1:
        os.makedirs(os.path.dirname(os.path.abspath(sys.argv[1])), exist_ok=True)
        open(sys.argv[1], 'w').write(txt)
    r0 = allr[0]
    print('backward+update %.0f us; buckets:' % r0['backward_total_us'])
    for r in r0['buckets']:
        print('  [%9d, %9d) %6.2f MB  start %7.0f us  end %7.0f us  (%.0f us)' % (r['floats'][0], r['floats'][1], r['mbytes'], r['start_us'], r['end_us'], r['end_us'] - r['start_us']))
plan.release_graphs()
dist.barrier()
torch.cuda.synchronize()
dist.destroy_process_group()
