"""Tiled dense prediction (reference: Node._predict_densetile / Node.predict_dense,
neuromancer/node_basic.py:805-1012).

Same tile geometry, padding, scaling and output conventions as the reference:
  * integer input is scaled by 1/255 (node_basic.py:904-910);
  * tile input = patch + strides - 1, useful output = out * strides (:939-940);
  * end tiles are zero-padded on the right and cropped after (:971-988);
  * as_uint8: trunc(prob * 255) (:990-996);
  * strides != 1 without MFP: prod(strides) forward calls on shifted crops,
    interleaved into the output (:832-856); all strides == 1 (U-Net, or MFP +
    FragmentsToDense): one call per tile (:825-829).
What differs is where the work happens: a tile crosses PCIe once as uint8 (or float32
if that is what the caller holds), is scaled on the device, and the probabilities come
back already converted.  ``tile_range`` shards the tile list for multi-GPU inference:
tiles are independent given their halos, which are re-read from the host volume, so
there is no collective on the data path.
"""
import ctypes as C
import logging
import time

import numpy as np
import torch

from .. import _lib

logger = logging.getLogger('elektronn2log')


def tile_geometry(node, raw_spatial):
    """(tile_sh, prob_sh, pred_sh, n_tiles) exactly as node_basic.py:938-954."""
    offset = node.shape.offsets
    if np.any(np.less(offset, 0)):
        raise ValueError("Cannot predict dense because the CNN contains UpConvs which cause unknown FOVs. "
                         "If you use UpConvs you should not need predict dense anyway!")
    ps = np.array(node.input_nodes[0].shape.spatial_shape)
    strides = np.array(node.shape.strides)
    out_sh = np.array(node.shape.spatial_shape)
    tile_sh = ps + strides - 1
    prob_sh = out_sh * strides
    pred_sh = np.array([raw_spatial[i] - 2 * offset[i] for i in range(3)])
    n_tiles = [int(np.ceil(float(pred_sh[i]) / prob_sh[i])) for i in range(3)]
    return tile_sh, prob_sh, pred_sh, n_tiles


def tile_list(n_tiles):
    """Tile indices in the reference's loop order (z outer, then x, then y)."""
    return [(z, x, y) for z in range(n_tiles[0]) for x in range(n_tiles[1]) for y in range(n_tiles[2])]


def shard_tiles(tiles, rank, world):
    """Contiguous block of the tile list for ``rank`` (keeps host reads sequential)."""
    per = (len(tiles) + world - 1) // world
    return tiles[rank * per:(rank + 1) * per]


class _TileRunner(object):
    """Owns the plan and the staging buffers for one (node, patch) pair."""

    def __init__(self, node, as_uint8, int_input):
        from .executor import Plan
        self.node = node
        inp = node.input_nodes[0]
        b = inp.shape['b'] or 1
        if int(b) != 1:
            # the reference feeds raw_img[None] (node_basic.py:825): one tile per call
            raise ValueError("predict_dense needs an input node with batch size 1 or None (got %d)" % int(b))
        self.plan = node._plans.get(int(b)) or Plan(node.model, [node], b)
        node._plans[int(b)] = self.plan
        self.h = self.plan.h
        self.t_in, self.pinned_f32, self.staging = self.plan.inputs[inp]
        d = self.t_in.desc
        if d.c != 1 and int_input:
            raise NotImplementedError("integer multi-channel input")
        self.in_shape = (d.n, d.c, d.z, d.x, d.y)
        self.int_input = int_input
        if int_input:
            self.dev_u8 = torch.empty(self.in_shape, dtype=torch.uint8, device=self.plan.device)
        out = self.plan.val[node]
        od = out.desc
        self.out = out
        self.out_ncdhw = torch.empty((od.n, od.c, od.z, od.x, od.y), dtype=torch.float32, device=self.plan.device)
        self.as_uint8 = as_uint8
        if as_uint8:
            self.out_u8 = torch.empty_like(self.out_ncdhw, dtype=torch.uint8)
        self.h2d_bytes = self.d2h_bytes = 0

    def forward(self, patch):
        """patch: (ch,z,x,y) host array (uint8 if int_input else float32 already scaled).
        Returns the (n_lab, zo, xo, yo) host array (float32, or uint8 = trunc(p*255))."""
        self.submit(patch)
        return self.collect()

    # Two-deep software pipeline for the one-call-per-tile case: ``submit`` enqueues H2D -> network -> conversion ->
    # D2H for a tile and returns; ``collect`` waits for the OLDEST submitted tile.  With two slots of pinned input /
    # output staging the host prepares tile i+1 and assembles tile i-1 while the GPU computes tile i.
    SLOTS = 2

    def _slots(self):
        if getattr(self, '_slot_bufs', None) is None:
            self._slot_bufs = []
            for _ in range(self.SLOTS):
                pin_in = (torch.empty(self.in_shape, dtype=torch.uint8) if self.int_input
                          else torch.empty(self.in_shape, dtype=torch.float32)).pin_memory()
                host_out = torch.empty(self.out_ncdhw.shape, dtype=torch.uint8 if self.as_uint8 else torch.float32).pin_memory()
                self._slot_bufs.append(dict(pin_in=pin_in, host_out=host_out, done=torch.cuda.Event(), busy=False))
            self._queue, self._next_slot = [], 0
        return self._slot_bufs

    def submit(self, patch):
        h, s = self.h, self.h.stream()
        slots = self._slots()
        slot = slots[self._next_slot]
        if slot['busy']:
            raise RuntimeError("predict_dense pipeline: collect() the oldest tile before submitting a third one")
        self._next_slot = (self._next_slot + 1) % self.SLOTS
        if self.int_input:
            slot['pin_in'][0].copy_(torch.from_numpy(np.ascontiguousarray(patch)))
            self.dev_u8.copy_(slot['pin_in'], non_blocking=True)
            h.call('e2_u8_to_f32', _lib.ptr(self.dev_u8), self.t_in.ptr(), self.dev_u8.numel(), C.c_float(255.0), s)
            self.h2d_bytes += self.dev_u8.numel()
        else:
            slot['pin_in'][0].copy_(torch.from_numpy(np.ascontiguousarray(patch, dtype=np.float32)))
            dst = self.staging if self.staging is not None else \
                self.t_in.buf[self.t_in.offset:self.t_in.offset + slot['pin_in'].numel()].view(self.in_shape)
            dst.copy_(slot['pin_in'], non_blocking=True)
            self.h2d_bytes += slot['pin_in'].numel() * 4
        self.plan.execute()
        h.call('e2_ndhwc_to_ncdhw', C.byref(self.out.desc), self.out.ptr(), _lib.ptr(self.out_ncdhw), s)
        if self.as_uint8:
            h.call('e2_f32_to_u8', _lib.ptr(self.out_ncdhw), _lib.ptr(self.out_u8), self.out_ncdhw.numel(),
                   C.c_float(255.0), s)
            slot['host_out'].copy_(self.out_u8, non_blocking=True)
        else:
            slot['host_out'].copy_(self.out_ncdhw, non_blocking=True)
        slot['done'].record()
        slot['busy'] = True
        self._queue.append(slot)
        self.d2h_bytes += slot['host_out'].numel() * slot['host_out'].element_size()

    def collect(self):
        slot = self._queue.pop(0)
        slot['done'].synchronize()
        slot['busy'] = False
        return slot['host_out'].numpy()[0]


def predict_dense(node, raw_img, as_uint8=False, pad_raw=False, tile_range=None, out=None, return_stats=False):
    """See module docstring.  ``raw_img``: (ch, z, x, y).  Returns (n_lab, z', x', y')."""
    if node.shape.ndim != 3:
        raise NotImplementedError("predict_dense on the B200 path handles 3-D nets")
    raw_img = np.asarray(raw_img)
    if raw_img.ndim != 4:
        raise ValueError("raw_img must be (ch, z, x, y)")
    int_input = np.issubdtype(raw_img.dtype, np.integer)
    if int_input and raw_img.dtype != np.uint8:
        raw_img = raw_img.astype(np.float32) / 255   # reference semantics for other int types
        int_input = False
    elif not int_input:
        raw_img = raw_img.astype(np.float32, copy=False)
    offset = node.shape.offsets
    t_start = time.time()
    if pad_raw:
        raw_img = np.pad(raw_img, [(0, 0)] + [(o, o) for o in offset], mode='symmetric')
    tile_sh, prob_sh, pred_sh, n_tiles = tile_geometry(node, raw_img.shape[1:])
    if np.any(pred_sh <= 0):
        raise ValueError("raw image %s is smaller than the field of view %s" % (raw_img.shape[1:], node.shape.fov))
    logger.info("Predicting img %s in %i Blocks: (%i, %i, %i)"
                % (raw_img.shape, int(np.prod(n_tiles)), n_tiles[0], n_tiles[1], n_tiles[2]))   # node_basic.py:958-959
    n_lab = node.shape['f']
    strides = [int(s) for s in node.shape.strides]
    patch = [int(s) for s in node.input_nodes[0].shape.spatial_shape]
    dtype = np.uint8 if as_uint8 else np.float32
    predictions = out if out is not None else np.zeros([n_lab] + list(pred_sh), dtype=dtype)
    runner = _TileRunner(node, as_uint8, int_input)
    tiles = tile_list(n_tiles)
    if tile_range is not None:
        tiles = tiles[tile_range[0]:tile_range[1]]
    one_call = all(s == 1 for s in strides)
    prob = np.zeros([n_lab] + list(prob_sh), dtype=dtype)
    def place(lo, right, end_tile, p):
        if end_tile:
            p = p[:, :prob_sh[0] - right[0], :prob_sh[1] - right[1], :prob_sh[2] - right[2]]
        predictions[:, lo[0]:lo[0] + prob_sh[0], lo[1]:lo[1] + prob_sh[1], lo[2]:lo[2] + prob_sh[2]] = p

    pending = None
    for (z_t, x_t, y_t) in tiles:
        lo = [z_t * prob_sh[0], x_t * prob_sh[1], y_t * prob_sh[2]]
        raw_tile = raw_img[:, lo[0]:lo[0] + tile_sh[0], lo[1]:lo[1] + tile_sh[1], lo[2]:lo[2] + tile_sh[2]]
        right = np.subtract(tile_sh, raw_tile.shape[1:])
        end_tile = bool(np.any(right > 0))
        if end_tile:
            raw_tile = np.pad(raw_tile, [(0, 0)] + [(0, int(r)) for r in right], mode='constant')
        if one_call:
            # the GPU works on this tile while the host assembles the previous one and cuts the next
            runner.submit(raw_tile)
            if pending is not None:
                place(*pending, runner.collect())
            pending = (lo, right, end_tile)
        else:
            for x_off in range(strides[1]):
                for y_off in range(strides[2]):
                    for z_off in range(strides[0]):
                        cut = raw_tile[:, z_off:z_off + patch[0], x_off:x_off + patch[1], y_off:y_off + patch[2]]
                        prob[:, z_off::strides[0], x_off::strides[1], y_off::strides[2]] = runner.forward(cut)
            place(lo, right, end_tile, prob)
    if pending is not None:
        place(*pending, runner.collect())
    dt = time.time() - t_start
    n_vox = float(np.prod(pred_sh))
    logger.info(" Inference speed: %.3f MB or MPix /s, time %.3f s (%d of %d blocks)"
                % (n_vox / 1e6 / dt, dt, len(tiles), int(np.prod(n_tiles))))                    # node_basic.py:1003-1007
    if return_stats:
        return predictions, dict(seconds=dt, tiles=len(tiles), n_tiles=n_tiles, h2d_bytes=runner.h2d_bytes,
                                 d2h_bytes=runner.d2h_bytes, launches=runner.h.launches)
    return predictions
