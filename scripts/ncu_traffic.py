#!/usr/bin/env python
"""Per-family DRAM traffic of one training step from an ncu launch list.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --csv --log-file launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph
    python scripts/ncu_traffic.py launches.csv profiles/traffic_unet3d.json

Kernels are mapped to the op families bench.py reports (conv_fwd / conv_dgrad share kernels, so they are
reported together as conv_fwd+dgrad).  Steps are counted by the k_softmax_nll_fwd launches in the list.
"""
import csv
import json
import sys

FAMILY = [
    ('k_wgrad_halo_tc', 'conv_wgrad'), ('k_wgrad_halo_reduce', 'conv_wgrad'), ('k_c1_wgrad', 'conv_wgrad'),
    ('k_wgrad_zs', 'conv_wgrad'),
    ('k_reduce_gemm', 'conv_wgrad'),
    ('k_wgrad_tc_reduce', 'upconv_wgrad'), ('k_wgrad_tc', 'upconv_wgrad'), ('k_bias_grad', 'upconv_wgrad'),
    ('k_conv_zstack_tc', 'conv_fwd+dgrad'), ('k_zstack_reduce', 'conv_fwd+dgrad'),
    ('k_c1_fwd', 'conv_fwd+dgrad'), ('k_gather_gemm_tc', 'conv/upconv tap kernel'), ('k_gather_gemm_reduce', 'conv/upconv tap kernel'),
    ('k_gather_gemm', 'conv_fwd+dgrad'),
    ('k_maxpool_fwd', 'pool_fwd'), ('k_maxpool_bwd', 'pool_bwd'), ('k_crop_fwd', 'crop_concat_fwd'), ('k_crop_bwd', 'crop_concat_bwd'),
    ('k_mfp_fwd', 'mfp_fwd'), ('k_mfp_bwd', 'mfp_bwd'),
    ('k_softmax_nll', 'loss'), ('k_adam_pack', 'adam+pack'), ('k_adam', 'adam'), ('k_pack', 'pack'), ('k_repack', 'pack'),
]


def family(name):
    for key, fam in FAMILY:
        if key in name:
            return fam
    return 'other'


def main(src, dst):
    rows = list(csv.reader(open(src, errors='replace')))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    H = rows[hdr]
    ki, mi, vi, ii = H.index('Kernel Name'), H.index('Metric Name'), H.index('Metric Value'), H.index('ID')
    ui = H.index('Metric Unit')
    launches = {}
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        d = launches.setdefault(r[ii], dict(name=r[ki]))
        v = float(r[vi].replace(',', ''))
        unit = r[ui].lower()
        scale = {'byte': 1.0, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'ns': 1.0, 'us': 1e3, 'ms': 1e6}.get(unit, 1.0)
        d[r[mi]] = v * scale
    steps = sum(1 for d in launches.values() if 'k_softmax_nll_fwd' in d['name'])
    fam = {}
    for d in launches.values():
        f = fam.setdefault(family(d['name']), dict(launches=0, dram_read_bytes=0.0, dram_write_bytes=0.0, time_ns=0.0))
        f['launches'] += 1
        f['dram_read_bytes'] += d.get('dram__bytes_read.sum', 0.0)
        f['dram_write_bytes'] += d.get('dram__bytes_write.sum', 0.0)
        f['time_ns'] += d.get('gpu__time_duration.sum', 0.0)
    out = dict(source=src, steps=steps, note='per training step; ncu per-launch values are cold-cache and serialised', families={})
    tot = sum(f['time_ns'] for f in fam.values())
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]['time_ns']):
        out['families'][k] = dict(launches_per_step=f['launches'] / max(steps, 1),
                                  dram_bytes_per_step=(f['dram_read_bytes'] + f['dram_write_bytes']) / max(steps, 1),
                                  dram_read_bytes_per_step=f['dram_read_bytes'] / max(steps, 1),
                                  dram_write_bytes_per_step=f['dram_write_bytes'] / max(steps, 1),
                                  time_us_per_step=f['time_ns'] / 1e3 / max(steps, 1), share_of_step=f['time_ns'] / tot)
    json.dump(out, open(dst, 'w'), indent=1)
    for k, v in out['families'].items():
        print('%-26s %6.1f launches  %8.1f MB  %8.1f us  %5.1f %%' % (k, v['launches_per_step'], v['dram_bytes_per_step'] / 1e6,
                                                                    v['time_us_per_step'], 100 * v['share_of_step']))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
