"""Multi-GPU execution: one process per GPU, torch.distributed (NCCL over NVLink /
NVSwitch) for the plumbing.

The reference has no multi-GPU code at all; both modes below are new
capabilities whose single-GPU numerics are what the parity tests pin.

* Training shards the batch: every rank holds the full model and its own patch
  (reference batch_size is 1, so global batch == world size).  Gradients live in ONE
  flat buffer laid out in reverse graph order, so the backward pass completes it front
  to back; it is all-reduced in buckets on a side stream, each bucket launched as soon
  as the wgrad kernels that fill it have been enqueued, overlapping with the remaining
  dgrad/wgrad work.  The mean over ranks is folded into the loss gradient
  (grad_scale = 1/world), so the collective is a plain sum.
  Caveat: MultinoulliNLL normalises by the *local* labelled-voxel count
  (loss.py:342-345); the mean of shard gradients equals the global-batch gradient when
  every shard has the same labelled count (true for dense labels).
* Dense inference shards the tile list (neuromancer/dense.shard_tiles): tiles are
  independent given their halos, which are re-read from the host volume -> no
  collective on the data path.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if os.environ.get('E2_NCCL_MAX_CTAS'):
            # every NCCL channel is a CTA that occupies an SM for the duration of a collective; the persistent conv
            # kernels size their grids to the SM count, so SMs taken by NCCL cost them a second wave
            os.environ.setdefault('NCCL_MAX_CTAS', os.environ['E2_NCCL_MAX_CTAS'])
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def bucket_ranges(entries, total, bucket_floats):
    """Split [0,total) of the flat gradient buffer into contiguous buckets of roughly
    ``bucket_floats`` that end on parameter boundaries.  ``entries``: (key, param, offset,
    size) in buffer order.  Returns [(start, end, last_entry_index)]."""
    out, start = [], 0
    for i, (_, _, off, size) in enumerate(entries):
        end = off + (size + 3) // 4 * 4
        last = i == len(entries) - 1
        if end - start >= bucket_floats or last:
            out.append((start, total if last else end, i))
            start = end
    return out


class _EventWork(object):
    """``Work``-like handle of a collective that is just stream-ordered work on the comm stream."""

    def __init__(self, event):
        self.event = event

    def wait(self):
        torch.cuda.current_stream().wait_event(self.event)


class DataParallel(object):
    """Gradient all-reduce for a ``neuromancer.Model``."""

    def __init__(self, model, bucket_mb=1024.0, overlap=True):
        self.model = model
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        bucket_mb = float(os.environ.get('E2_DP_BUCKET_MB', bucket_mb))
        self.bucket_floats = int(bucket_mb * 1024 * 1024 / 4)
        self._nocomm = os.environ.get('E2_DP_NOCOMM') == '1'      # experiments only: bookkeeping without the collectives
        # Measured on 2 and 8 B200: every additional bucket costs the step ~0.03 ms whatever carries it
        # (NCCL or the copy engines), and the backward pass is long enough behind the split offset to hide one large
        # collective -- so by default there is ONE bucket for everything below the fused optimiser's split offset and
        # one bucket per layer for the E2_DP_TAIL layers nearest the input (their gradients are produced last).
        self.tail_entries = int(os.environ.get('E2_DP_TAIL', '2'))
        self.overlap = overlap
        self.comm_stream = None
        self.producer_streams = []            # set by the executor: the streams its wgrad kernels run on
        self.graph_ok = os.environ.get('E2_DP_GRAPH', '1') != '0'   # capture the step incl. the collectives
        model.data_parallel = self
        self.bytes_reduced = 0
        self.trace = None                     # list of (start, end, start_event, end_event) per bucket when tracing (eager steps only)
        # E2_DP_CE=1: all-reduce over NVSwitch peer memory with the COPY ENGINES (reduce-scatter by pulling, local sum,
        # all-gather by pulling; see _ce_allreduce) instead of NCCL's SM-resident kernels
        self.use_ce = os.environ.get('E2_DP_CE', '0') == '1' and self.world > 1
        self._hdl = self._stage = None

    def broadcast_parameters(self, store):
        """Rank 0's weights everywhere (identical init is also given by the shared seed)."""
        if self.world > 1:
            dist.broadcast(store.P, src=0)
            store.version += 1

    def grad_scale(self):
        return 1.0 / self.world

    # -- copy-engine all-reduce over symmetric (peer-mapped) memory --------------------------------------------
    def alloc_gradient_buffer(self, numel, device):
        """The flat gradient buffer.  With E2_DP_CE it is allocated as symmetric memory and exchanged with the peers
        (torch.distributed._symmetric_memory: one VMM allocation per rank, mapped into every rank's address space over
        NVLink), so that a rank can read any peer's gradients with a plain device-to-device copy."""
        if not self.use_ce:
            return torch.zeros(numel, dtype=torch.float32, device=device)
        import torch.distributed._symmetric_memory as symm_mem
        g = symm_mem.empty(numel, dtype=torch.float32, device=device)
        g.zero_()
        self._hdl = symm_mem.rendezvous(g, dist.group.WORLD.group_name)
        self._numel = numel
        return g

    def _ce_allreduce(self, G, s, e):
        """Sum G[s:e] over the ranks, on the current (comm) stream, without occupying SMs for the transfers.

        rank r owns slice r of the bucket: it PULLS that slice from every peer's buffer (copy engines, NVLink 5 through
        the NVSwitch: every peer at full bandwidth), adds the world-1 staged copies to its own (one small kernel; each
        element is summed by exactly one rank in a fixed order, so all ranks end up bit-identical), then pulls the
        other ranks' finished slices.  Three peer barriers (signal pads in peer memory) order the phases:
        gradients ready -> slices summed -> everybody has read everything (the buffer may be overwritten)."""
        hdl, world, rank = self._hdl, self.world, self.rank
        n = e - s
        per = ((n + world - 1) // world + 3) // 4 * 4
        lo = [min(e, s + r * per) for r in range(world)]
        hi = [min(e, s + (r + 1) * per) for r in range(world)]
        if self._stage is None or self._stage.shape[1] < per:
            self._stage = torch.empty((world - 1, per), dtype=torch.float32, device=G.device)
        mine = hi[rank] - lo[rank]
        hdl.barrier(channel=0, timeout_ms=20000)
        if mine > 0:
            for step in range(1, world):
                peer = (rank + step) % world
                src = hdl.get_buffer(peer, (mine,), torch.float32, lo[rank])
                self._stage[step - 1, :mine].copy_(src, non_blocking=True)
            G[lo[rank]:hi[rank]].add_(self._stage[:world - 1, :mine].sum(dim=0))
        hdl.barrier(channel=1, timeout_ms=20000)
        for step in range(1, world):
            peer = (rank + step) % world
            cnt = hi[peer] - lo[peer]
            if cnt > 0:
                G[lo[peer]:hi[peer]].copy_(hdl.get_buffer(peer, (cnt,), torch.float32, lo[peer]), non_blocking=True)
        hdl.barrier(channel=2, timeout_ms=20000)

    def allreduce_gradients(self, store):
        """Simple form: one collective over the whole flat buffer on the current stream.
        Used when the step did not go through ``begin_step`` / ``finish_step``."""
        if self.world > 1 and not self._step_reduced:
            if self.use_ce:
                self._ce_allreduce(store.G, 0, store.total)
            else:
                dist.all_reduce(store.G, op=dist.ReduceOp.SUM)
            self.bytes_reduced += store.G.numel() * 4
        self._step_reduced = False

    # -- overlapped, bucketed form (driven by executor.Plan.execute) ------------------
    _step_reduced = False

    def begin_step(self, store, split=None):
        """``split``: an offset that must be a bucket boundary (the fused optimiser updates [0, split) as soon as
        those buckets are reduced)."""
        if getattr(self, '_buckets', None) is None or getattr(self, '_split', None) != split:
            self._split = split
            # weight region in buckets; the layers nearest the input -- whose gradients the backward pass produces
            # LAST -- get a bucket each, so that when the final wgrad kernel ends only its own few KB (merged with the
            # bias region, which is contiguous with it) are still to be reduced: the exposed tail of the step is one
            # latency-bound collective instead of a multi-MB one
            w_entries = [e for e in store.entries if e[2] < store.n_reg]
            self._buckets = [(s, e) for s, e, _ in bucket_ranges(w_entries, store.n_reg, self.bucket_floats)]
            cuts = set()
            if split:
                cuts.add(int(split))
            for e in w_entries[-self.tail_entries:] if self.tail_entries else []:
                cuts.add(int(e[2]))
            for c in sorted(cuts):
                self._buckets = [r for (s, e) in self._buckets for r in ([(s, c), (c, e)] if s < c < e else [(s, e)])]
            if store.total > store.n_reg:
                if self._buckets and self.tail_entries:
                    s_last, _ = self._buckets.pop()
                    self._buckets.append((s_last, store.total))
                else:
                    self._buckets.append((store.n_reg, store.total))
            if torch.cuda.is_available() and store.G.is_cuda:
                self.comm_stream = torch.cuda.Stream(device=store.G.device)
        self._next, self._works = 0, []

    def _launch_bucket(self, store, s, e):
        if self._nocomm:
            return
        view = store.G[s:e]
        if self.comm_stream is not None and self.overlap:
            ev = torch.cuda.Event()
            ev.record()                       # everything enqueued so far on the current stream ...
            evs = [ev]
            for st in self.producer_streams:  # ... and on every stream that runs wgrad kernels
                e2 = torch.cuda.Event()
                e2.record(st)
                evs.append(e2)
            with torch.cuda.stream(self.comm_stream):
                for e2 in evs:
                    self.comm_stream.wait_event(e2)
                if self.trace is not None:
                    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    t0.record(self.comm_stream)
                if self.use_ce:
                    self._ce_allreduce(store.G, s, e)
                    done = torch.cuda.Event()
                    done.record(self.comm_stream)
                    self._works.append(_EventWork(done))
                else:
                    self._works.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, async_op=True))
                if self.trace is not None:
                    self._works[-1].wait()        # stream-ordered on the comm stream
                    t1.record(self.comm_stream)
                    self.trace.append((s, e, t0, t1))
        else:
            self._works.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, async_op=True))
        self.bytes_reduced += (e - s) * 4

    def on_gradients_ready(self, end_offset):
        """Called after a wgrad launch: gradients in [0, end_offset) of the weight region
        are (stream-ordered) complete -> launch every bucket that lies inside."""
        store = self.model._store
        while self._next < len(self._buckets) and self._buckets[self._next][1] <= min(end_offset, store.n_reg) \
                and self._buckets[self._next][0] < store.n_reg:
            s, e = self._buckets[self._next]
            self._launch_bucket(store, s, e)
            self._next += 1

    def wait_launched(self):
        """Make the current stream wait for every collective launched so far in this step."""
        for w in self._works:
            w.wait()

    def finish_step(self, store):
        while self._next < len(self._buckets):
            s, e = self._buckets[self._next]
            self._launch_bucket(store, s, e)
            self._next += 1
        for w in self._works:
            w.wait()                          # compute stream waits for the collectives
        if self.comm_stream is not None and self.overlap:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self._step_reduced = True
