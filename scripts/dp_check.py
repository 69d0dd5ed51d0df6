#!/usr/bin/env python
"""2-rank data-parallel check on GPUs: after 3 fused training steps (graph-captured NCCL buckets, early optimiser
update) the parameters must be identical on both ranks and equal to a single-process run that averages the two
ranks' gradients by hand (unfused path, E2_FUSE_OPT=0 semantics via execute() + manual all-reduce + opt.step())."""
import contextlib, io, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elektronn2_b200 import examples, parallel, neuromancer as nm

rank, world, local = parallel.init_from_env()
torch.cuda.set_device(local)
dist = torch.distributed


def build():
    nm.model_manager.reset() if hasattr(nm.model_manager, 'reset') else None
    np.random.seed(2)
    with contextlib.redirect_stdout(io.StringIO()):
        m = examples.unet3d_litelite()
    nm.optimiser.Optimiser.setlr(1e-3)
    nm.optimiser.Optimiser.setwd(0.5e-4)
    nm.optimiser.Optimiser.setmom(0.9)
    return m


def data(m, r):
    ish = [1 if s is None else s for s in m.input_node.shape.shape]
    tsh = [1 if s is None else s for s in m.target_node.shape.shape]
    return (np.random.RandomState(1000 + r).rand(*ish).astype(np.float32),
            np.random.RandomState(2000 + r).randint(0, 2, tsh).astype(np.float32))


# (a) fused data-parallel path
m = build()
dp = parallel.DataParallel(m)
x, t = data(m, rank)
plan = m._train_plan(1)
dp.broadcast_parameters(plan.store)
for _ in range(3):
    m.trainingstep(x, t, optimiser='Adam')
torch.cuda.synchronize()
pa = plan.store.P.clone()
# ranks agree
other = pa.clone()
dist.broadcast(other, src=0)
same = bool((other == pa).all())
# (b) reference: same model, no DataParallel hooks, gradients averaged by hand, plain optimiser step
m2 = build()
plan2 = m2._train_plan(1)
dist.broadcast(plan2.store.P, src=0)
plan2.store.version += 1
opt2 = m2.optimisers['Adam']
for _ in range(3):
    plan2.feed({m2.input_node: x, m2.target_node: t})
    plan2.execute()
    dist.all_reduce(plan2.store.G)
    plan2.store.G.mul_(1.0 / world)
    opt2.step(plan2.store)
torch.cuda.synchronize()
pb = plan2.store.P
err = float((pa - pb).abs().max() / pb.abs().max())
print('rank %d: ranks identical %s, fused-DP vs manual max rel diff %.3e, graph %s' % (rank, same, err, bool(plan._opt_graphs)), flush=True)
ok = same and err < 2e-5
dist.barrier()
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0 if ok else 1)      # no interpreter teardown: NCCL communicators referenced by CUDA graphs can block at exit
