"""Tiled dense prediction (TEST INFRASTRUCTURE).

Restates Node._predict_densetile (node_basic.py:805-858) and
Node.predict_dense (node_basic.py:860-1012) for 3-D nets, parameterised by a
``forward(patch (1,ch,z,x,y)) -> (1,n_lab,zo,xo,yo)`` callable.
"""
import numpy as np


def tile_grid(raw_spatial, patch, out_spatial, strides, offsets):
    """Tile geometry: node_basic.py:938-954."""
    tile_sh = np.add(patch, strides) - 1
    prob_sh = np.multiply(out_spatial, strides)
    pred_sh = np.array([raw_spatial[i] - 2 * offsets[i] for i in range(3)])
    n_tiles = [int(np.ceil(float(pred_sh[i]) / prob_sh[i])) for i in range(3)]
    return tile_sh, prob_sh, pred_sh, n_tiles


def predict_densetile(forward, raw_tile, out_arr, patch, strides):
    """node_basic.py:822-858."""
    if np.all(np.equal(strides, 1)):
        out_arr[:] = forward(raw_tile[None])[0]
        return out_arr
    for x_off in range(strides[1]):
        for y_off in range(strides[2]):
            for z_off in range(strides[0]):
                cut = raw_tile[None, :, z_off:z_off + patch[0], x_off:x_off + patch[1],
                               y_off:y_off + patch[2]]
                out_arr[:, z_off::strides[0], x_off::strides[1], y_off::strides[2]] = forward(cut)[0]
    return out_arr


def predict_dense(forward, raw_img, patch, out_spatial, strides, offsets, n_lab,
                  as_uint8=False, pad_raw=False):
    """node_basic.py:898-1012 (3-D case).  ``raw_img`` is (ch,z,x,y)."""
    m = 255 if np.issubdtype(raw_img.dtype, np.integer) else 1   # :904-908
    raw_img = raw_img.astype(np.float32) / m                      # :910
    if pad_raw:
        raw_img = np.pad(raw_img, [(0, 0)] + [(o, o) for o in offsets], mode='symmetric')
    raw_sh = raw_img.shape[1:]
    tile_sh, prob_sh, pred_sh, (zt, xt, yt) = tile_grid(raw_sh, patch, out_spatial, strides, offsets)
    prob_arr = np.zeros([n_lab] + list(prob_sh), np.float32)
    predictions = np.zeros([n_lab] + list(pred_sh), np.uint8 if as_uint8 else np.float32)
    for z_t in range(zt):
        for x_t in range(xt):
            for y_t in range(yt):
                raw_tile = raw_img[:, z_t * prob_sh[0]:z_t * prob_sh[0] + tile_sh[0],
                                   x_t * prob_sh[1]:x_t * prob_sh[1] + tile_sh[1],
                                   y_t * prob_sh[2]:y_t * prob_sh[2] + tile_sh[2]]
                end_tile = not np.all(np.equal(raw_tile.shape[1:], tile_sh))
                if end_tile:                                        # :972-977 zero right-pad
                    right = np.subtract(tile_sh, raw_tile.shape[1:])
                    raw_tile = np.pad(raw_tile, [(0, 0)] + [(0, int(r)) for r in right], mode='constant')
                prob = predict_densetile(forward, raw_tile, prob_arr, patch, strides)
                if end_tile:                                        # :987-988
                    prob = prob[:, :prob_sh[0] - right[0], :prob_sh[1] - right[1], :prob_sh[2] - right[2]]
                if as_uint8:
                    prob = prob * 255                               # :990-991 (then C truncation)
                predictions[:, z_t * prob_sh[0]:(z_t + 1) * prob_sh[0],
                            x_t * prob_sh[1]:(x_t + 1) * prob_sh[1],
                            y_t * prob_sh[2]:(y_t + 1) * prob_sh[2]] = prob
    return predictions
