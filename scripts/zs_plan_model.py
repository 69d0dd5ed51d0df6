#!/usr/bin/env python
"""Offline cycle model of the z-stack conv kernel's tile plans (CPU only).

plan_zstack (csrc/e2_conv_zstack_tc.cu) ranks plans by MMA issue cycles alone.  The measurements of round 2
(profiles/r2_ncu_conv1_dgrad.txt, r2_zs_trace_conv1_tz.txt, r2_mma_bench_2cta.txt) say a stage is bound by the
shared-memory port: operand fetch of every MMA + the TMA fills of planes and weights + the epilogue's staging traffic, with
a 48-cycle floor per instruction.  This script enumerates the plans the kernel could run for a layer under that model and
prints where it disagrees with the library's pick (read through e2_debug_zstack_plan), i.e. the experiments worth a GPU run
(E2_ZS_BN / E2_ZS_TZ force a plan).

    stage  = max(sum_q nk * max(48, n_q / 2), sum_q nk * (32 + n_q / 4) + fill / 128 + epilogue share) + 250
    fill   = kz * BN * 128 (weights of the stage) + NP * plane bytes / (kx * ky) (planes, once per channel block)
Calibration: conv1 fwd 2343 measured / 2326 model cycles per stage; conv1 dgrad 2138 / 2001; dense conv5 0.49 / 0.58 of peak.
"""
import ctypes as C
import math
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

TX, TY, EPI_WARPS, SMS = 16, 8, 8, 148
KS_OPTS = [1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 24]


def plans(K, N, On, Oz, Ox, Oy, kz, kx, ky, gate=False, may_split=True):
    S, T9 = kz, kx * ky
    XH, YP = TX + kx - 1, TY + ky - 1
    plane_bytes = XH * YP * 128
    plane_stride = (plane_bytes + 1023) // 1024 * 1024
    CBn = (K + 31) // 32
    nk_avg = ((K + 7) // 8) / CBn
    ntx, nty = -(-Ox // TX), -(-Oy // TY)
    out = []
    for b in range(min(256, (N + 15) // 16 * 16), 15, -16):
        ntn = -(-N // b)
        if ntn > 1 and b % 32:
            continue
        w_bytes = S * b * 128
        epi_bytes = EPI_WARPS * 4096 + (1 if ntn == 1 else EPI_WARPS) * b * 4
        budget = 227 * 1024 - 2048 - epi_bytes
        for tz in range(8, 0, -1):
            if min(tz, S) * b > 256 or 2 * tz * b > 512:
                continue
            np_ = tz + S - 1
            if np_ > 11 or np_ * plane_stride + 2 * w_bytes > budget:
                continue
            rest = budget - np_ * plane_stride
            wsl = min(8, rest // w_bytes)
            dbl = 0
            if wsl >= 3 and rest - 3 * w_bytes >= np_ * plane_stride:
                dbl, wsl = 1, min(8, (rest - np_ * plane_stride) // w_bytes)
            dur = port = 0.0
            for q in range(np_):
                nblk = min(q, tz - 1) - max(0, q - (S - 1)) + 1
                n = nblk * b
                dur += max(48.0, n / 2.0)
                port += 32.0 + n / 4.0
            fill = w_bytes + np_ * plane_bytes / T9
            # epilogue: st.shared + TMA-store read of every output chunk (+ TMA write and ld.shared of the gate tile)
            epi = tz * 128 * b * 4 * (4 if gate else 2) / 128.0
            ntz = -(-Oz // tz)
            for ks in KS_OPTS:
                if ks > 1 and (not may_split or ks > CBn):
                    break
                cb_per = -(-CBn // ks)
                if -(-CBn // cb_per) != ks:
                    continue
                stages = cb_per * T9
                stage = max(dur * nk_avg, port * nk_avg + fill / 128.0 + (epi / stages if ks == 1 else 0.0)) + 250.0
                stage = max(stage, 1200.0 / max(1, wsl - 1))
                tiles = On * ntz * ntx * nty * ntn * ks
                waves = -(-tiles // SMS)
                cost = waves * (stages * stage + (0 if dbl else 800) + 200) + tz * (b / 32.0) / 2 * 1200
                if ks > 1:
                    c_bytes = On * Oz * Ox * Oy * ((N + 3) // 4 * 4) * 4.0
                    cost += 8000 + (ks + 1.0) * c_bytes / 2500.0
                useful = 2.0 * On * Oz * Ox * Oy * N * K * S * T9
                out.append(dict(bn=b, tz=tz, ks=ks, ntn=ntn, wsl=wsl, dbl=dbl, cost=cost,
                                frac=useful / (cost * SMS * 2 * 128 * 8 * 2)))   # of 2048 MAC/clk/SM
    return out


def library_pick(K, N, Oz, Ox, Oy, kz, kx, ky, may_split=True):
    from elektronn2_b200 import _lib
    out = (C.c_int * 8)()
    if not _lib.lib.e2_debug_zstack_plan(SMS, K, N, Oz, Ox, Oy, kz, kx, ky, 1 if may_split else 0, out):
        return None
    return dict(zip('bn tz ks cb_per wsl nslot units ntn'.split(), out))


UNET3D = [  # name, c_in, c_out, output extent of the forward conv
    ('conv1', 32, 64, (112, 128, 128)), ('conv2', 64, 64, (54, 62, 62)), ('conv3', 64, 128, (52, 60, 60)),
    ('conv4', 128, 128, (24, 28, 28)), ('conv5', 128, 256, (22, 26, 26)), ('conv8', 768, 256, (12, 16, 16)),
    ('conv9', 256, 256, (10, 14, 14)), ('conv10', 384, 128, (18, 26, 26)), ('conv11', 128, 128, (16, 24, 24)),
    ('conv12', 192, 64, (30, 46, 46)), ('conv13', 64, 64, (28, 44, 44)),
]


def main():
    rows = []
    for name, ci, co, (z, x, y) in UNET3D:
        rows.append((name + ' fwd', ci, co, 1, z, x, y, False))
        rows.append((name + ' dgrad', co, ci, 1, z + 2, x + 2, y + 2, True))
    for label, K, N, On, z, x, y, gate in rows:
        cand = plans(K, N, On, z, x, y, 3, 3, 3, gate=gate)
        if not cand:
            continue
        best = min(cand, key=lambda c: c['cost'])
        lib = library_pick(K, N, z, x, y, 3, 3, 3)
        mine = [c for c in cand if lib and (c['bn'], c['tz'], c['ks']) == (lib['bn'], lib['tz'], lib['ks'])]
        cur = mine[0] if mine else None
        gain = (cur['cost'] / best['cost'] - 1.0) * 100 if cur else float('nan')
        print('%-13s K %3d N %3d  library bn %3d tz %d ks %2d -> model %.2f of peak | model best bn %3d tz %d ks %2d (%.2f)  %+5.1f %%'
              % (label, K, N, lib['bn'] if lib else 0, lib['tz'] if lib else 0, lib['ks'] if lib else 0,
                 cur['frac'] if cur else float('nan'), best['bn'], best['tz'], best['ks'], best['frac'], gain))


if __name__ == '__main__':
    main()
