#!/usr/bin/env python
"""Benchmark of the ELEKTRONN2 volumetric-CNN hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--workload unet3d] [--impl reference]

Metric (BASELINE.json): 3-D U-Net training voxels/s (+ conv3d TFLOP/s in ``roofline``).
Workload: ``examples/unet3d.py`` (configs[2]; it is the configuration the metric is quoted
on and it fits one GPU), one (1,1,116,132,132) float32 patch per GPU, data-parallel over
N GPUs ("weak" scaling: global batch == N).  Voxels/s follows the reference's own
convention: prod(input batch shape) / step time (training/trainer.py:283-284).

One JSON line is printed by rank 0.  ``value`` times K steps with inputs resident in
HBM; ``e2e`` times K calls of ``Model.trainingstep(x_host, t_host)`` -- the call a user
of the reference makes -- including the pinned H2D copies and the D2H loss read.
``--impl reference`` times the Theano-equivalent CPU restatement (oracle/theano_cpu.py)
on the host cores: Theano itself cannot be installed here.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    'unet3d': dict(patch=(116, 132, 132)),
    'unet3d_litelite': dict(patch=(22, 140, 140)),
    'neuro3d_lite': dict(patch=(11, 155, 155)),
    'neuro3d': dict(patch=(23, 185, 185)),
}


# BASELINE.json configs 4 and 5: tiled dense prediction of a uint8 EM volume with examples/neuro3d.py re-built with MFP
DENSE = {
    'predict_dense_512': dict(volume=(512, 512, 512), patch=(54, 400, 400), seed=3, uint8_out=False,
                              cpu_patch=(23, 185, 185)),
    'predict_dense_2048': dict(volume=(2048, 2048, 2048), patch=(54, 400, 400), seed=4, uint8_out=True,
                               cpu_patch=(23, 185, 185)),
}


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], bf16=d['bf16_tflops'], bf16_sustained=d.get('bf16_tflops_sustained'),
                    source='measured')
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source='fallback')


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region.  ONE long-running ``nvidia-smi -lms`` child
    (a Python thread forking a query every 200 ms stalled the timed host loop through the GIL); rows are kept by
    their timestamp, so only samples taken between ``mark_begin`` and ``stop`` count."""
    Q = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.p, self.t0 = index, None, None

    def start(self):
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '50'], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def mark_begin(self):
        import datetime
        self.t0 = datetime.datetime.now()

    def stop(self):
        import datetime
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        t1 = datetime.datetime.now()
        time.sleep(0.12)                 # let the sample that covers the end of the region be printed
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ''
        rows, inside = [], []
        for line in out.strip().splitlines():
            f = [v.strip() for v in line.split(',')]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], '%Y/%m/%d %H:%M:%S.%f')
                float(f[1]), float(f[2]), float(f[3])
            except Exception:
                continue
            rows.append(f[1:])
            if self.t0 is not None and self.t0 - datetime.timedelta(milliseconds=60) <= ts <= t1 + datetime.timedelta(milliseconds=60):
                inside.append(f[1:])
        use = inside or rows[-5:]
        if not use:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        sm = sorted(float(r[0]) for r in use)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith('active') for r in use)]
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=float(use[0][1]), reasons=reasons,
                    power_w_max=max(float(r[2]) for r in use), samples=len(use), samples_in_timed_region=len(inside))


def synthetic_batch(model, seed):
    ish = [1 if s is None else s for s in model.input_node.shape.shape]
    tsh = [1 if s is None else s for s in model.target_node.shape.shape]
    x = np.random.RandomState(seed).rand(*ish).astype(np.float32)          # EM-like [0,1) float32
    t = np.random.RandomState(seed + 1).randint(0, 2, tsh).astype(np.float32)
    return x, t


def cpu_reference(workload, steps, warmup, budget_s=240.0):
    """Theano-equivalent CPU restatement on all host cores.  ALWAYS the patch named in ``config.workload``: when the
    projected time exceeds ``budget_s`` the number of timed steps is cut (and reported), never the patch."""
    import torch
    from oracle import nets as onets, theano_cpu
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    builder = onets.BUILDERS[workload]
    patch = WORKLOADS[workload]['patch']
    t0 = time.perf_counter()
    cal = theano_cpu.time_training(builder, patch, steps=1, warmup=0)       # first step (includes one-off costs)
    first = time.perf_counter() - t0
    per = cal['seconds_per_step']
    left = budget_s - first
    n_warm = max(0, min(warmup - 1, int(left * 0.2 / per)))
    n_steps = max(1, min(steps, int((left - n_warm * per) / per)))
    if n_steps == 1 and n_warm == 0 and left < per:
        r = cal                                                              # slow host: the calibration step is the sample
    else:
        r = theano_cpu.time_training(builder, patch, steps=n_steps, warmup=n_warm)
    return dict(value=r['voxels_per_s'], unit='voxels/s', cores=r['cores'], kind='port',
                sample='%d timed fwd+bwd+Adam step(s) (after %d warm-up) of %s on the full (1,1,%d,%d,%d) patch, '
                       'conv3d2d/pool_2d decomposition on torch-CPU float32 (Theano-equivalent restatement, not Theano)'
                       % ((n_steps, n_warm + 1, workload) + tuple(patch)),
                ms_per_step=r['seconds_per_step'] * 1e3, patch=list(patch), steps_timed=n_steps)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='unet3d', choices=sorted(WORKLOADS) + sorted(DENSE))
    ap.add_argument('--compute', default=None, choices=['tf32', 'f32'])
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-solo-check', action='store_true', help='N > 1: skip the per-GPU single-process timing')
    ap.add_argument('--profile-out', default=None, help='write the per-launch timing table (JSON) here')
    args = ap.parse_args()
    if args.workload in DENSE:
        return dense_main(args)
    W = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    K = max(args.steps, 1)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    config = dict(workload='examples/%s.py training step (fwd + bwd + Adam), patch (1,1,%d,%d,%d) per GPU'
                           % ((args.workload,) + WORKLOADS[args.workload]['patch']),
                  global_batch=world, parallelism='dp%d' % world,
                  l2_policy='activations of one step (~2.5 GB for unet3d) exceed the 126 MB L2')

    if args.impl == 'reference':
        if rank != 0:
            return 0
        r = cpu_reference(args.workload, K, args.warmup)
        line = dict(metric='3D U-Net train voxels/sec', value=r['value'], unit='voxels/s', n_gpus=world, steps=K,
                    warmup=args.warmup, ms_per_step=r['ms_per_step'], higher_is_better=True, scaling='weak',
                    vs_baseline=None, dtype='f32', data='synthetic', config=config, impl='reference',
                    steps_timed=r['steps_timed'],
                    cpu_baseline=dict(value=r['value'], unit='voxels/s', cores=r['cores'], kind=r['kind'],
                                      sample=r['sample']),
                    e2e=dict(value=r['value'], unit='voxels/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return 0

    import torch
    import elektronn2_b200
    from elektronn2_b200 import examples, parallel, neuromancer as nm
    from elektronn2_b200.config import config as e2cfg
    if args.compute:
        e2cfg.compute = args.compute
    if args.no_graph:
        e2cfg.use_cuda_graph = False
    rank, world, local = parallel.init_from_env()
    torch.cuda.set_device(local)
    dist = torch.distributed
    sampler = ClockSampler(local)
    sampler.start()                                     # running well before the timed region starts
    np.random.seed(2)                                   # identical initial weights on every rank
    with contextlib.redirect_stdout(io.StringIO()):
        model = examples.BUILDERS[args.workload]()
    dp = None
    solo = None
    if world > 1 and not args.no_solo_check:
        # Every rank first times the SAME step on its own GPU without data parallelism (no collective, no peer to wait
        # for): synchronous data parallelism runs at the pace of the slowest GPU of the box, so the spread of these
        # numbers is the part of the N-GPU step that no overlap scheme can recover.
        if os.environ.get('E2_SOLO_AFTER_NCCL_INIT'):      # experiment: is an initialised (idle) communicator felt?
            dist.all_reduce(torch.zeros(1 << 20, device='cuda'))
            torch.cuda.synchronize()
        nm.model_manager.reset()
        np.random.seed(2)
        with contextlib.redirect_stdout(io.StringIO()):
            m0 = examples.BUILDERS[args.workload]()
        nm.optimiser.Optimiser.setlr(5e-4), nm.optimiser.Optimiser.setwd(0.5e-4), nm.optimiser.Optimiser.setmom(0.9)
        x0, t0_ = synthetic_batch(m0, 1000 + rank)
        p0 = m0._train_plan(1)
        o0 = m0.optimisers['Adam']
        p0.feed({m0.input_node: x0, m0.target_node: t0_})
        for _ in range(W):
            p0.train_step(o0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(K):
            p0.train_step(o0)
        b.record()
        torch.cuda.synchronize()
        mine = torch.tensor([a.elapsed_time(b) / K], device='cuda', dtype=torch.float64)
        allms = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allms, mine)
        solo = [round(float(v), 4) for v in allms]
        p0.release_graphs()
        del p0, m0, o0
        nm.model_manager.reset()
        np.random.seed(2)
        with contextlib.redirect_stdout(io.StringIO()):
            model = examples.BUILDERS[args.workload]()
    if world > 1:
        dp = parallel.DataParallel(model)
    nm.optimiser.Optimiser.setlr(5e-4)                  # examples/unet3d.py optimiser_params
    nm.optimiser.Optimiser.setwd(0.5e-4)
    nm.optimiser.Optimiser.setmom(0.9)
    x, t = synthetic_batch(model, 1000 + rank)
    plan = model._train_plan(1)
    if dp is not None:
        dp.broadcast_parameters(plan.store)
    opt = model.optimisers['Adam']
    # forward + backward launches (counted by the library in an eager pass) + the optimiser / re-pack launches of the step
    launches_per_step = plan.launches_per_step(with_pack=False) + plan.update_launches(opt)
    n_vox = float(np.prod(x.shape)) * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        plan.train_step(opt)                            # fwd + bwd (+ all-reduce) + Adam + weight re-pack

    # ---- value: inputs resident in HBM ------------------------------------------------
    plan.feed({model.input_node: x, model.target_node: t})
    for _ in range(W):
        device_step()
    barrier()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        device_step()
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    # ---- e2e: the public API call with host buffers (page-locked, as the contract says) --
    x = torch.from_numpy(x).pin_memory().numpy()
    t = torch.from_numpy(t).pin_memory().numpy()
    for _ in range(2):
        model.trainingstep(x, t, optimiser='Adam')
    barrier()
    import gc
    gc.collect()
    gc.disable()                                        # no collector pauses inside the timed host loop
    t0 = time.perf_counter()
    for _ in range(K):
        loss, _, _ = model.trainingstep(x, t, optimiser='Adam')
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    gc.enable()
    clocks = sampler.stop()
    if world > 1:
        tt = torch.tensor([dev_ms, e2e_ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(tt[0]), float(tt[1])
    if rank != 0:
        _finish(dist, world, model)
        return 0

    # ---- roofline of the dominant kernel family (eager, per-launch CUDA events) ---------
    peaks = load_peaks()
    prof = plan.profile(repeats=3, opt=opt)
    step_ms = sum(p[4] for p in prof)
    fam = {}
    for label, kind, flops, nbytes, ms in prof:
        key = label.split(':')[0]
        f = fam.setdefault(key, dict(ms=0.0, flops=0.0, bytes=0.0, n=0, tensor=0))
        f['ms'] += ms
        f['flops'] += flops if kind == 'tensor' else 0.0
        f['bytes'] += nbytes
        f['n'] += 1
        f['tensor'] += kind == 'tensor'
    top = max(fam, key=lambda k: fam[k]['ms'])
    tf = fam[top]
    tensor_peak = peaks['bf16'] / 2.0 if e2cfg.compute == 'tf32' else None   # TF32 = half the bf16 rate
    if tf['tensor'] and tensor_peak:
        ach = tf['flops'] / (tf['ms'] * 1e-3) / 1e12
        roof = dict(bound='tensor', kernel=top, achieved=ach, peak=tensor_peak, unit='TFLOP/s', frac=ach / tensor_peak,
                    traffic=None, launches=tf['n'], avg_launch_ms=tf['ms'] / tf['n'], share_of_step=tf['ms'] / step_ms,
                    peak_source='%s cuBLAS bf16 %.1f TFLOP/s (burst) / 2 for TF32' % (peaks['source'], peaks['bf16']))
    else:
        ach = tf['bytes'] / (tf['ms'] * 1e-3) / 1e9
        roof = dict(bound='hbm', kernel=top, achieved=ach, peak=peaks['hbm'], unit='GB/s', frac=ach / peaks['hbm'],
                    traffic=None, launches=tf['n'], avg_launch_ms=tf['ms'] / tf['n'], share_of_step=tf['ms'] / step_ms,
                    peak_source='%s copy bandwidth' % peaks['source'])
    # DRAM traffic of that family from the committed ncu pass (scripts/ncu_traffic.py), per launch like `achieved`
    tpath = os.path.join(ROOT, 'profiles', 'traffic_%s.json' % args.workload)
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        tf_ = tj.get('families', {}).get(top)
        n_launch = tf['n']
        if not tf_ and top in ('conv_fwd', 'conv_dgrad') and 'conv_fwd+dgrad' in tj.get('families', {}):
            # forward and dgrad run the same kernels (k_conv_zstack_tc / tap kernel): ncu cannot tell them apart, so the
            # per-launch traffic is that of the combined family
            tf_ = tj['families']['conv_fwd+dgrad']
            n_launch = tf_['launches_per_step']
            roof['traffic_note'] = 'conv_fwd and conv_dgrad share kernels: DRAM bytes per launch of the combined family'
        if tf_:
            roof['traffic'] = tf_['dram_bytes_per_step'] / n_launch
            roof['traffic_source'] = 'profiles/traffic_%s.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, %s)' % (
                args.workload, tj.get('source', ''))
            roof['algorithmic_bytes'] = tf['bytes'] / tf['n']
    all_flops = sum(p[2] for p in prof)
    roof['step_conv_tflops'] = all_flops / (dev_ms / K * 1e-3) / 1e12
    if args.profile_out:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
        json.dump(dict(step_ms_eager=step_ms, families=fam,
                       launches=[dict(label=p[0], kind=p[1], flops=p[2], bytes=p[3], ms=p[4]) for p in prof]),
                  open(args.profile_out, 'w'), indent=1)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference(args.workload, 1, 0, budget_s=30.0)
        cpu = dict(value=r['value'], unit='voxels/s', cores=r['cores'], kind=r['kind'], sample=r['sample'])

    in_bytes = x.nbytes + t.nbytes
    line = dict(metric='3D U-Net train voxels/sec', value=n_vox * K / (dev_ms * 1e-3), unit='voxels/s', n_gpus=world,
                steps=K, warmup=W, ms_per_step=dev_ms / K, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype=e2cfg.compute, data='synthetic', config=config, roofline=roof, cpu_baseline=cpu,
                e2e=dict(value=n_vox * K / (e2e_ms * 1e-3), unit='voxels/s', h2d_bytes_per_step=in_bytes,
                         d2h_bytes_per_step=16, ms_per_step=e2e_ms / K, host_buffers='page-locked numpy arrays'),
                gpu_launches=launches_per_step * K, clocks=clocks, loss=float(loss),
                **(dict(solo_ms_per_step_by_rank=solo, allreduce='copy engines over symmetric peer memory (E2_DP_CE=1)'
                        if getattr(dp, 'use_ce', False) else 'NCCL') if world > 1 else {}),
                cuda_graph=bool(plan._opt_graphs) or plan._graph is not None)
    print(json.dumps(line))
    _finish(dist, world, model)
    return 0


def cpu_dense_reference(spec, steps, budget_s=240.0):
    """The reference's CPU way to predict one dense tile (no MFP: prod(strides) shifted passes, node_basic.py:832-856)
    on all host cores; a step = one tile of the (23,185,185) training patch (10x84x84 output voxels)."""
    import torch
    from oracle import nets as onets, theano_cpu
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    t0 = time.perf_counter()
    r = theano_cpu.time_dense_tile(onets.neuro3d, spec['cpu_patch'], (2, 4, 4))
    first = time.perf_counter() - t0
    n = max(0, min(steps - 1, int((budget_s - first) / max(r['seconds'], 1e-9))))
    secs = [r['seconds']]
    for i in range(n):
        secs.append(theano_cpu.time_dense_tile(onets.neuro3d, spec['cpu_patch'], (2, 4, 4), seed=i + 1)['seconds'])
    vox = float(np.prod(r['out_spatial']))
    per = float(np.mean(secs))
    return dict(value=vox / per, unit='voxels/s', cores=r['cores'], kind='port', ms_per_step=per * 1e3, steps_timed=len(secs),
                sample='%d tile(s) of %s output voxels each, %d shifted forward passes of examples/neuro3d.py at patch %s per '
                       'tile (the reference\'s non-MFP dense path, node_basic.py:832-856), conv3d2d/pool_2d decomposition on '
                       'torch-CPU float32 (Theano-equivalent restatement, not Theano)'
                       % (len(secs), r['out_spatial'], r['passes'], list(spec['cpu_patch'])))


def dense_main(args):
    """``--workload predict_dense_512 | predict_dense_2048``: a step = one (54,400,400) tile through the MFP graph."""
    spec = DENSE[args.workload]
    K = max(args.steps, 1)
    W = max(args.warmup, 3)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    config = dict(workload='predict_dense, examples/neuro3d.py + MFP (override_mfp_to_active), uint8 volume %s, tile patch '
                           '(1,1,%d,%d,%d), %s output' % ((list(spec['volume']),) + tuple(spec['patch']) +
                                                          ('uint8' if spec['uint8_out'] else 'float32',)),
                  parallelism='tiles sharded over %d rank(s) in contiguous blocks, no collective' % world,
                  l2_policy='activations of one tile (~6 GB) exceed the 126 MB L2')
    if args.impl == 'reference':
        if rank != 0:
            return 0
        r = cpu_dense_reference(spec, K)
        line = dict(metric='predict_dense voxels/sec', value=r['value'], unit='voxels/s', n_gpus=world, steps=K,
                    warmup=args.warmup, ms_per_step=r['ms_per_step'], higher_is_better=True, scaling='weak',
                    vs_baseline=None, dtype='f32', data='synthetic', config=config, impl='reference',
                    steps_timed=r['steps_timed'],
                    cpu_baseline=dict(value=r['value'], unit='voxels/s', cores=r['cores'], kind=r['kind'], sample=r['sample']),
                    e2e=dict(value=r['value'], unit='voxels/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return 0

    import ctypes as C
    import torch
    from elektronn2_b200 import _lib, examples, parallel, neuromancer as nm
    from elektronn2_b200.config import config as e2cfg
    from elektronn2_b200.neuromancer import dense
    if args.compute:
        e2cfg.compute = args.compute
    rank, world, local = parallel.init_from_env()
    torch.cuda.set_device(local)
    dist = torch.distributed
    sampler = ClockSampler(local)
    sampler.start()
    np.random.seed(2)
    with contextlib.redirect_stdout(io.StringIO()):
        base = examples.neuro3d()
        model = nm.rebuild_model(base, override_mfp_to_active=True, imposed_patch_size=tuple(spec['patch']))
    node = model.prediction_node
    vol_sp = spec['volume']
    tile_sh, prob_sh, pred_sh, n_tiles = dense.tile_geometry(node, vol_sp)
    tiles = dense.tile_list(n_tiles)
    per = (len(tiles) + world - 1) // world
    lo = rank * per
    hi = min(len(tiles), lo + per, lo + K)          # bounded sample: the first K tiles of this rank's contiguous block
    mine = tiles[lo:hi]
    # the volume: zero pages are never touched; only the z-slab(s) this rank's tiles read get synthetic uint8 data
    vol = np.zeros((1,) + tuple(vol_sp), dtype=np.uint8)
    for z_t in sorted(set(t[0] for t in mine)):
        z0, z1 = z_t * int(prob_sh[0]), min(vol_sp[0], z_t * int(prob_sh[0]) + int(tile_sh[0]))
        vol[0, z0:z1] = np.random.RandomState(spec['seed'] * 1000 + z_t).randint(0, 256, vol[0, z0:z1].shape, dtype=np.uint8)
    as_u8 = spec['uint8_out']
    # output buffer: only the rows this rank writes are touched
    out = np.zeros([node.shape['f']] + [int(v) for v in pred_sh], dtype=np.uint8 if as_u8 else np.float32)

    def placed_voxels(ts):
        n = 0
        for t in ts:
            n += int(np.prod([min(int(prob_sh[i]), int(pred_sh[i]) - t[i] * int(prob_sh[i])) for i in range(3)]))
        return n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up through the public API (plan build, CUDA-graph capture) ---------------------------------------
    dense.predict_dense(node, vol, as_uint8=as_u8, tile_range=(lo, lo + 1), out=out)
    runner = dense._TileRunner(node, as_u8, True)
    h = runner.h
    n_in = runner.dev_u8.numel()

    def device_tile():
        s = h.stream()
        h.call('e2_u8_to_f32', _lib.ptr(runner.dev_u8), runner.t_in.ptr(), n_in, C.c_float(255.0), s)
        runner.plan.execute()
        h.call('e2_ndhwc_to_ncdhw', C.byref(runner.out.desc), runner.out.ptr(), _lib.ptr(runner.out_ncdhw), s)
        if as_u8:
            h.call('e2_f32_to_u8', _lib.ptr(runner.out_ncdhw), _lib.ptr(runner.out_u8), runner.out_ncdhw.numel(),
                   C.c_float(255.0), s)

    # ---- value: the tile already resident in HBM (full tile of random uint8) ------------------------------------
    runner.dev_u8.copy_(torch.randint(0, 256, runner.in_shape, dtype=torch.uint8))
    before = h.launches
    device_tile()                                   # the plan replays its CUDA graph: only the boundary kernels count here
    torch.cuda.synchronize()
    edge = h.launches - before
    before = h.launches
    runner.plan.pack()
    packs = h.launches - before
    launches_per_tile = edge + runner.plan.launches_per_step() - packs     # eager pass of the launch list, counted by the library
    for _ in range(W):
        device_tile()
    barrier()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        device_tile()
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    # ---- e2e: Node.predict_dense on the host volume (uint8 in over PCIe, probabilities back, host-side assembly) ----
    barrier()
    t0 = time.perf_counter()
    _, st = dense.predict_dense(node, vol, as_uint8=as_u8, tile_range=(lo, hi), out=out, return_stats=True)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    vox_tile = float(np.prod(prob_sh))
    dev_vox, e2e_vox, n_done = vox_tile * K, float(placed_voxels(mine)), float(len(mine))
    if world > 1:
        tt = torch.tensor([dev_ms, e2e_s], device='cuda', dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(tt[0]), float(tt[1])
        ss = torch.tensor([dev_vox, e2e_vox, n_done], device='cuda', dtype=torch.float64)
        dist.all_reduce(ss, op=dist.ReduceOp.SUM)
        dev_vox, e2e_vox, n_done = float(ss[0]), float(ss[1]), float(ss[2])
    if rank != 0:
        _finish(dist, world, model)
        return 0
    peaks = load_peaks()
    prof = runner.plan.profile(repeats=3)
    tile_ms = sum(p[4] for p in prof)
    fam = {}
    for label, kind, flops, nbytes, ms in prof:
        f = fam.setdefault(label.split(':')[0], dict(ms=0.0, flops=0.0, bytes=0.0, n=0))
        f['ms'] += ms
        f['flops'] += flops if kind == 'tensor' else 0.0
        f['bytes'] += nbytes
        f['n'] += 1
    top = max(fam, key=lambda k: fam[k]['ms'])
    tf = fam[top]
    tensor_peak = peaks['bf16'] / 2.0
    ach = tf['flops'] / (tf['ms'] * 1e-3) / 1e12
    roof = dict(bound='tensor', kernel=top, achieved=ach, peak=tensor_peak, unit='TFLOP/s', frac=ach / tensor_peak, traffic=None,
                launches=tf['n'], avg_launch_ms=tf['ms'] / tf['n'], share_of_step=tf['ms'] / tile_ms,
                peak_source='%s cuBLAS bf16 %.1f TFLOP/s (burst) / 2 for TF32' % (peaks['source'], peaks['bf16']))
    if 'mfp_fwd' in fam:
        m = fam['mfp_fwd']
        gbs = m['bytes'] / (m['ms'] * 1e-3) / 1e9
        roof['mfp_fwd'] = dict(bound='hbm', achieved=gbs, peak=peaks['hbm'], unit='GB/s', frac=gbs / peaks['hbm'],
                               launches=m['n'], ms=m['ms'], share_of_step=m['ms'] / tile_ms)
    if args.profile_out:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
        json.dump(dict(tile_ms_eager=tile_ms, families=fam,
                       launches=[dict(label=p[0], kind=p[1], flops=p[2], bytes=p[3], ms=p[4]) for p in prof]),
                  open(args.profile_out, 'w'), indent=1)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_dense_reference(spec, 1, budget_s=30.0)
        cpu = dict(value=r['value'], unit='voxels/s', cores=r['cores'], kind=r['kind'], sample=r['sample'])
    tiles_e2e = int(n_done)
    line = dict(metric='predict_dense voxels/sec', value=dev_vox / (dev_ms * 1e-3), unit='voxels/s', n_gpus=world, steps=K,
                warmup=W, ms_per_step=dev_ms / K, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype=e2cfg.compute, data='synthetic', config=config, roofline=roof, cpu_baseline=cpu,
                e2e=dict(value=e2e_vox / e2e_s, unit='voxels/s', h2d_bytes_per_step=st['h2d_bytes'] / max(len(mine), 1),
                         d2h_bytes_per_step=st['d2h_bytes'] / max(len(mine), 1), seconds=e2e_s, tiles=tiles_e2e,
                         tiles_total=len(tiles), tiles_per_rank=len(mine),
                         api='Node.predict_dense(raw_img uint8 host array) -> host array'),
                gpu_launches=int(launches_per_tile * K), clocks=clocks, tile_output_voxels=[int(v) for v in prob_sh],
                cuda_graph=runner.plan._graph is not None)
    print(json.dumps(line))
    _finish(dist, world, model)
    return 0


def _finish(dist, world, model=None):
    """Orderly teardown, so that interpreter exit hooks run: drop the CUDA graphs (they reference the NCCL
    communicators and the streams), drain the device, destroy the process group, return normally."""
    import gc
    import torch
    sys.stdout.flush()
    sys.stderr.flush()
    if model is not None:
        plans = list(getattr(model, '_train_plans', {}).values()) + list(getattr(model, '_ext_plans', {}).values())
        for n in getattr(model, 'nodes', {}).values():
            plans += list(getattr(n, '_plans', {}).values())
        for plan in plans:
            plan.release_graphs()
    gc.collect()
    torch.cuda.synchronize()
    if world > 1 and dist.is_initialized():
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == '__main__':
    sys.exit(main())
