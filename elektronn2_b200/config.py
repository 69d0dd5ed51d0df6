"""Backend configuration.

The reference selects its conv/pool backend with module-level flags in
elektronn2/config.py:79-81 (use_manual_cudnn_conv, use_manual_cudnn_conv_not_w1,
use_manual_cudnn_pool).  The B200 backend has one switch of the same kind:
``compute`` picks the arithmetic of the conv GEMMs.
"""


class _Config(object):
    backend = 'b200'
    # 'tf32': tcgen05.mma kind::tf32 with fp32 accumulation (rel. error ~3e-4 at K~7k);
    # 'f32' : CUDA-core FFMA, exact fp32 products (parity / debugging).
    compute = 'tf32'
    use_cuda_graph = True
    # pooling backward tie rule: 'first' (cuDNN-like single winner) or 'all' (Theano CPU)
    pool_tie_mode = 'first'
    # Model.time_per_step / loss_smooth average over this many steps (elektronn2/config.py:77)
    time_per_step_smoothing_length = 50


config = _Config()
