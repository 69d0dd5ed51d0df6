#!/usr/bin/env python
"""Secondary benchmark: tiled dense prediction (BASELINE.json configs 4 and 5).

    python bench_dense.py [--volume 512 512 512] [--patch 54 400 400] [--uint8-out] [--max-tiles N]
    python -m torch.distributed.run --nproc-per-node N ... bench_dense.py ...   (tiles sharded over ranks)

``examples/neuro3d.py`` re-built with max-fragment-pooling (``override_mfp_to_active``, model.py:623-729)
predicts a synthetic uint8 EM volume tile by tile through ``Node.predict_dense`` (node_basic.py:860-1012),
the call a user of the reference makes: every tile crosses PCIe as uint8, the probabilities come back as
float32 (or uint8 = trunc(p*255) with --uint8-out).  Metric: predicted voxels / wall second, the reference's own
convention (node_basic.py:1003-1007), H2D/D2H and host-side tile assembly included.  With N ranks the tile
list is cut into contiguous blocks (dense.shard_tiles); there is no collective on the data path, the time
is the max over ranks.  ``bench.py`` stays the headline benchmark; this prints one JSON line of its own.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--volume', type=int, nargs=3, default=[512, 512, 512])
    ap.add_argument('--patch', type=int, nargs=3, default=[54, 400, 400])
    ap.add_argument('--uint8-out', action='store_true')
    ap.add_argument('--max-tiles', type=int, default=0, help='bound the run: only the first N tiles of each rank')
    ap.add_argument('--compute', default=None, choices=['tf32', 'f32'])
    ap.add_argument('--profile-out', default=None, help='per-launch CUDA-event table of one tile (JSON)')
    args = ap.parse_args()

    import torch
    from elektronn2_b200 import examples, parallel, neuromancer as nm
    from elektronn2_b200.config import config as e2cfg
    from elektronn2_b200.neuromancer import dense
    if args.compute:
        e2cfg.compute = args.compute
    rank, world, local = parallel.init_from_env()
    torch.cuda.set_device(local)
    dist = torch.distributed
    np.random.seed(2)
    with contextlib.redirect_stdout(io.StringIO()):
        base = examples.neuro3d()
        model = nm.rebuild_model(base, override_mfp_to_active=True, imposed_patch_size=tuple(args.patch))
    node = model.prediction_node
    patch = [int(s) for s in model.input_node.shape.spatial_shape]
    # RandomState(3).randint(0, 256, (1, 512, 512, 512), uint8); generated per z-slab
    vol = np.empty([1] + list(args.volume), dtype=np.uint8)
    rs = np.random.RandomState(3)
    for z in range(0, args.volume[0], 64):
        vol[0, z:z + 64] = rs.randint(0, 256, vol[0, z:z + 64].shape, dtype=np.uint8)
    tile_sh, prob_sh, pred_sh, n_tiles = dense.tile_geometry(node, vol.shape[1:])
    tiles = dense.tile_list(n_tiles)
    per = (len(tiles) + world - 1) // world
    lo, hi = rank * per, min(len(tiles), (rank + 1) * per)
    if args.max_tiles:
        hi = min(hi, lo + args.max_tiles)
    out = np.zeros([node.shape['f']] + list(pred_sh), dtype=np.uint8 if args.uint8_out else np.float32)
    # warm-up: one tile (plan build, CUDA-graph capture), not timed
    dense.predict_dense(node, vol, as_uint8=args.uint8_out, tile_range=(lo, lo + 1), out=out)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, st = dense.predict_dense(node, vol, as_uint8=args.uint8_out, tile_range=(lo, hi), out=out, return_stats=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    done = hi - lo
    if world > 1:
        tt = torch.tensor([dt, float(done)], device='cuda', dtype=torch.float64)
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        dt, done = float(mx[0]), int(tt[1])
    if rank != 0:
        sys.stdout.flush()
        os._exit(0)          # no interpreter teardown with live NCCL communicators (see bench.py:_finish)
    if args.profile_out:
        plan = next(iter(node._plans.values()))
        prof = plan.profile(repeats=3)
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
        json.dump(dict(tile_ms=sum(p[4] for p in prof),
                       launches=[dict(label=p[0], kind=p[1], flops=p[2], bytes=p[3], ms=p[4]) for p in prof]),
                  open(args.profile_out, 'w'), indent=1)
    vox_per_tile = float(np.prod(prob_sh))
    n_vox = min(done * vox_per_tile, float(np.prod(pred_sh)) if done == len(tiles) else done * vox_per_tile)
    flops_per_vox = 0.0
    for n in model.nodes.values():
        w = getattr(n, 'w', None)
        if w is not None and hasattr(w, 'shape'):
            flops_per_vox += 2.0 * float(np.prod(w.shape))
    line = dict(metric='predict_dense voxels/sec', value=n_vox / dt, unit='voxels/s', n_gpus=world, seconds=dt,
                tiles=done, tiles_total=len(tiles), higher_is_better=True, scaling='strong', dtype=e2cfg.compute,
                data='synthetic',
                config=dict(workload='examples/neuro3d.py + MFP predict_dense, uint8 volume %s, patch %s -> %s voxels per tile'
                                     % (list(args.volume), patch, [int(v) for v in prob_sh]),
                            parallelism='tiles sharded over %d rank(s), no collective' % world,
                            output='uint8' if args.uint8_out else 'float32'),
                e2e=dict(value=n_vox / dt, unit='voxels/s', h2d_bytes=st['h2d_bytes'], d2h_bytes=st['d2h_bytes']),
                tflops_asymptotic=n_vox * flops_per_vox / dt / 1e12)
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        os._exit(0)
    return 0


if __name__ == '__main__':
    sys.exit(main())
