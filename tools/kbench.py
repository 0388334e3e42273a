"""Per-kernel CUDA-event timings at the BASELINE configs[1] layer shapes (N = 5120 breaths, bf16), L2-cold:
every call works on one of ROT buffer sets (> 126 MB apart in total), so inputs come from HBM like inside a step.

    python tools/kbench.py [bn] [conv] [wgrad] [stem]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepards_b200 import _lib  # noqa: E402

DEV = "cuda"
N, GROUP = 5120, 20
SHAPES = [(64, 56), (128, 28), (256, 14), (512, 7)]
ROT = 6
BF = torch.bfloat16


def timeit(fn, reps=18):
    """us per call; the calls are captured in a CUDA graph so that the host launch cost (ctypes) does not count."""
    for i in range(ROT):
        fn(i)
    torch.cuda.synchronize()
    if os.environ.get("KBENCH_NO_GRAPH"):
        return 1e-9
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


def st():
    return torch.cuda.current_stream().cuda_stream


def bn():
    G = N // GROUP
    for c, l in SHAPES:
        rows = GROUP * l
        xs = [torch.randn(N, l, c, device=DEV).to(BF) for _ in range(ROT)]
        gs = [torch.randn(N, l, c, device=DEV).to(BF) for _ in range(ROT)]
        outs = [torch.empty(N, l, c, device=DEV, dtype=BF) for _ in range(ROT)]
        out2 = [torch.empty(N, l, c, device=DEV, dtype=BF) for _ in range(ROT)]
        gamma, beta = torch.ones(c, device=DEV), torch.zeros(c, device=DEV)
        mean, rstd = torch.empty(G, c, device=DEV), torch.empty(G, c, device=DEV)
        dgp, dbp = torch.empty(G, c, device=DEV), torch.empty(G, c, device=DEV)
        mb = N * l * c * 2 / 1e6

        def fwd(i, res=False):
            k = i % ROT
            _lib.call("dards_gbn_fwd", xs[k].data_ptr(), outs[k].data_ptr(), gs[k].data_ptr() if res else None,
                      gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(), rstd.data_ptr(), G, rows, c, c, c, c, 1e-5, 1,
                      _lib.BF16, st())

        def bwd(i, mode=1, dres=False):
            k = i % ROT
            _lib.call("dards_gbn_bwd", gs[k].data_ptr(), xs[k].data_ptr(), outs[k].data_ptr() if mode == 2 else None,
                      gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(), rstd.data_ptr(), out2[k].data_ptr(), 0,
                      gs[k].data_ptr() if dres else None, dgp.data_ptr(), dbp.data_ptr(),
                      G, rows, c, c, c, c, c, c, mode, _lib.BF16, st())

        t = timeit(lambda i: fwd(i, False))
        print("gbn_fwd  C=%3d L=%2d plain      : %6.1f us  %5.0f GB/s (r+w = %.0f MB)" % (c, l, t, 2 * mb / t * 1e3, 2 * mb))
        t = timeit(lambda i: fwd(i, True))
        print("gbn_fwd  C=%3d L=%2d residual   : %6.1f us  %5.0f GB/s (2r+w)" % (c, l, t, 3 * mb / t * 1e3))
        t = timeit(lambda i: bwd(i, 1, False))
        print("gbn_bwd  C=%3d L=%2d mode1      : %6.1f us  %5.0f GB/s (2r+w)" % (c, l, t, 3 * mb / t * 1e3))
        t = timeit(lambda i: bwd(i, 2, True))
        print("gbn_bwd  C=%3d L=%2d mode2+dres : %6.1f us  %5.0f GB/s (3r+2w)" % (c, l, t, 5 * mb / t * 1e3), flush=True)


def conv():
    from deepards_b200 import kernels as K
    for c, l in SHAPES:
        xs = [torch.randn(N, l, c, device=DEV).to(BF) for _ in range(ROT)]
        outs = [torch.empty(N, l, c, device=DEV, dtype=BF) for _ in range(ROT)]
        w = torch.randn(c, c, 3, device=DEV) * 0.05
        kio, koi = K.pack_conv_weight(w, BF)
        fl = 2.0 * N * l * c * c * 3

        def f(i):
            k = i % ROT
            _lib.call("dards_conv1d_fwd", xs[k].data_ptr(), koi.data_ptr(), outs[k].data_ptr(), None, N, l, l, c, c, c, c, 0,
                      3, 1, 1, _lib.BF16, 1, st())

        def d(i):
            k = i % ROT
            _lib.call("dards_conv1d_dgrad", xs[k].data_ptr(), kio.data_ptr(), outs[k].data_ptr(), None, N, l, l, c, c, c, c,
                      0, 3, 1, 1, _lib.BF16, 1, st())

        nb = _lib.fn("dards_conv1d_wgrad_workspace_bytes")(N, l, c, c, 3, 1)
        ws = torch.empty(max(nb // 4, 1), device=DEV)
        dw = torch.empty(c, c, 3, device=DEV)

        def wg(i):
            k = i % ROT
            _lib.call("dards_conv1d_wgrad", xs[k].data_ptr(), outs[k].data_ptr(), dw.data_ptr(), 0, ws.data_ptr(),
                      ws.numel() * 4, N, l, l, c, c, c, c, 3, 1, 1, _lib.BF16, 1, st())

        dwt = torch.zeros(3, c, c, device=DEV)

        def wga(i):  # accumulate mode (what the training step runs): L2 reduce-adds into the tap-major buffer, no reduce kernel
            k = i % ROT
            _lib.call("dards_conv1d_wgrad_accum", xs[k].data_ptr(), outs[k].data_ptr(), dwt.data_ptr(), N, l, l, c, c, c, c,
                      3, 1, 1, _lib.BF16, st())

        for name, fn in (("fwd", f), ("dgrad", d), ("wgrad", wg), ("wgacc", wga)):
            t = timeit(fn)
            print("conv %-5s C=%3d L=%2d k3 s1: %6.1f us  %6.1f TFLOP/s" % (name, c, l, t, fl / t * 1e-6), flush=True)


def stem():
    G = N // GROUP
    c0 = 64
    xs = [torch.randn(N, 224, device=DEV) for _ in range(ROT)]
    outs = [torch.empty(N, 56, c0, device=DEV, dtype=BF) for _ in range(ROT)]
    w = torch.randn(c0, 1, 7, device=DEV) * 0.3
    gamma, beta = torch.ones(c0, device=DEV), torch.zeros(c0, device=DEV)
    mean, rstd = torch.empty(G, c0, device=DEV), torch.empty(G, c0, device=DEV)
    dwp, dgp, dbp = torch.empty(G, c0 * 7, device=DEV), torch.empty(G, c0, device=DEV), torch.empty(G, c0, device=DEV)

    def f(i):
        k = i % ROT
        _lib.call("dards_stem_fwd", xs[k].data_ptr(), w.data_ptr(), gamma.data_ptr(), beta.data_ptr(), outs[k].data_ptr(),
                  mean.data_ptr(), rstd.data_ptr(), G, GROUP, c0, c0, 1e-5, 0, None, 0, _lib.BF16, st())

    def b(i):
        k = i % ROT
        _lib.call("dards_stem_bwd", outs[k].data_ptr(), xs[k].data_ptr(), w.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                  mean.data_ptr(), rstd.data_ptr(), dwp.data_ptr(), dgp.data_ptr(), dbp.data_ptr(), G, GROUP, c0, c0, 0,
                  None, 0, _lib.BF16, st())

    print("stem fwd: %6.1f us" % timeit(f))
    print("stem bwd: %6.1f us" % timeit(b), flush=True)


def convbn():
    """conv -> BatchNorm (+ residual) -> ReLU: two kernels (conv, gbn_fwd) against the fused path (one kernel where a group
    fits on chip, else conv with statistics partials + the streaming normalisation)."""
    from deepards_b200 import kernels as K
    G = N // GROUP
    for c, l in SHAPES:
        rows = GROUP * l
        xs = [torch.randn(N, l, c, device=DEV).to(BF) for _ in range(ROT)]
        rs = [torch.randn(N, l, c, device=DEV).to(BF) for _ in range(ROT)]
        ys = [torch.empty(N, l, c, device=DEV, dtype=BF) for _ in range(ROT)]
        outs = [torch.empty(N, l, c, device=DEV, dtype=BF) for _ in range(ROT)]
        w = torch.randn(c, c, 3, device=DEV) * 0.05
        kio, koi = K.pack_conv_weight(w, BF)
        gamma, beta = torch.ones(c, device=DEV), torch.zeros(c, device=DEV)
        mean, rstd = torch.empty(G, c, device=DEV), torch.empty(G, c, device=DEV)
        shape = (N, GROUP, l, l, c, c, 3, 1, 1)
        mode = _lib.fn("dards_conv1d_bn_mode")(*shape, _lib.BF16)
        entries = _lib.fn("dards_conv1d_bn_part_entries")(*shape) if mode == 1 else 0
        part = torch.empty(max(G * entries * 3 * c, 1), device=DEV)
        fl = 2.0 * N * l * c * c * 3

        def two(i, res):
            k = i % ROT
            _lib.call("dards_conv1d_fwd", xs[k].data_ptr(), koi.data_ptr(), ys[k].data_ptr(), None, N, l, l, c, c, c, c, 0,
                      3, 1, 1, _lib.BF16, 1, st())
            _lib.call("dards_gbn_fwd", ys[k].data_ptr(), outs[k].data_ptr(), rs[k].data_ptr() if res else None,
                      gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(), rstd.data_ptr(), G, rows, c, c, c, c, 1e-5,
                      1 | _lib.HINT_LAST_USE, _lib.BF16, st())

        def fused(i, res):
            k = i % ROT
            rp = rs[k].data_ptr() if res else None
            _lib.call("dards_conv1d_bn_fwd", xs[k].data_ptr(), koi.data_ptr(), ys[k].data_ptr(), outs[k].data_ptr(),
                      rp if mode == 2 else None, gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                      part.data_ptr(), N, GROUP, l, l, c, c, c, c, c, c, 3, 1, 1, 1e-5, 1, _lib.BF16, st())
            if mode == 1:
                _lib.call("dards_gbn_apply_fwd", ys[k].data_ptr(), outs[k].data_ptr(), rp, gamma.data_ptr(), beta.data_ptr(),
                          part.data_ptr(), entries, mean.data_ptr(), rstd.data_ptr(), None, None, None, None, 0, None, None,
                          G, rows, c, c, c, c, 0, 1e-5, 1, _lib.BF16, st())

        def conv_only(i):
            k = i % ROT
            _lib.call("dards_conv1d_bn_fwd", xs[k].data_ptr(), koi.data_ptr(), ys[k].data_ptr(), outs[k].data_ptr(),
                      None, gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                      part.data_ptr(), N, GROUP, l, l, c, c, c, c, c, c, 3, 1, 1, 1e-5, 1, _lib.BF16, st())

        for res in (False, True):
            t2 = timeit(lambda i: two(i, res))
            tf = timeit(lambda i: fused(i, res))
            print("conv+bn%s C=%3d L=%2d: conv, gbn_fwd %6.1f us | fused (mode %d) %6.1f us  %6.1f TFLOP/s" %
                  ("+res" if res else "    ", c, l, t2, mode, tf, fl / tf * 1e-6), flush=True)
        if mode == 1:
            print("          conv with statistics partials alone: %6.1f us" % timeit(conv_only), flush=True)


def copy():
    """the practical floor: torch's elementwise copy / read-only reduction of one 36.7 MB activation tensor, same rotation"""
    for mb, shape in ((36.7, (5120, 56, 64)), (147, (4 * 5120, 56, 64))):
        xs = [torch.randn(*shape, device=DEV).to(BF) for _ in range(ROT)]
        outs = [torch.empty_like(x) for x in xs]
        t = timeit(lambda i: outs[i % ROT].copy_(xs[i % ROT]))
        print("torch copy   %6.1f MB: %6.1f us  %5.0f GB/s (r+w)" % (mb, t, 2 * mb / t * 1e3))
        t = timeit(lambda i: outs[i % ROT].copy_(outs[i % ROT]) if False else torch.add(xs[i % ROT], 1, out=outs[i % ROT]))
        print("torch add    %6.1f MB: %6.1f us  %5.0f GB/s (r+w)" % (mb, t, 2 * mb / t * 1e3))
        acc = torch.zeros(shape[-1], device=DEV, dtype=torch.float32)
        t = timeit(lambda i: torch.sum(xs[i % ROT], dim=(0, 1), dtype=torch.float32, out=acc))
        print("torch sum    %6.1f MB: %6.1f us  %5.0f GB/s (read only)" % (mb, t, mb / t * 1e3), flush=True)


def convprof():
    """two small-layer conv launches (fwd) and one wgrad for `ncu --set full -k regex:tc_`"""
    global SHAPES, ROT
    SHAPES, ROT = [(64, 56)], 2
    conv()


def bnprof():
    """a handful of BN launches for `ncu --set full -k regex:gbn`"""
    global SHAPES, ROT
    SHAPES, ROT = [(64, 56), (512, 7)], 2
    bn()


if __name__ == "__main__":
    what = sys.argv[1:] or ["bn", "conv", "stem"]
    torch.cuda.set_device(0)
    _lib.load()
    if os.environ.get("KBENCH_CONV3"):
        _lib.call("dards_tc_debug_set", 5, int(os.environ["KBENCH_CONV3"]))  # opt in to the single-load 3-tap kernel
    if os.environ.get("KBENCH_WGRAD_FUSE"):
        _lib.call("dards_tc_debug_set", 7, int(os.environ["KBENCH_WGRAD_FUSE"]))  # 0: three MMAs per k-step at C=64
    if os.environ.get("KBENCH_TILE_BALANCE"):
        _lib.call("dards_tc_debug_set", 8, int(os.environ["KBENCH_TILE_BALANCE"]))  # 1: wave-balanced tile width
    if os.environ.get("KBENCH_SHARED"):
        _lib.call("dards_tc_debug_set", 10, int(os.environ["KBENCH_SHARED"]))  # 0: wide k3/s1 layers on the per-tap kernel; 2: all on the shared-weight loop
    if os.environ.get("KBENCH_CB_BSTAGES"):
        _lib.call("dards_tc_debug_set", 12, int(os.environ["KBENCH_CB_BSTAGES"]))
    if os.environ.get("KBENCH_SHAPES"):
        SHAPES = [tuple(int(v) for v in sh.split("x")) for sh in os.environ["KBENCH_SHAPES"].split(",")]
    if os.environ.get("KBENCH_CB_SHARE"):
        _lib.call("dards_tc_debug_set", 14, int(os.environ["KBENCH_CB_SHARE"]))
    if os.environ.get("KBENCH_CB_WIDE"):
        _lib.call("dards_tc_debug_set", 13, int(os.environ["KBENCH_CB_WIDE"]))
    if os.environ.get("KBENCH_CB_PERTAP"):
        _lib.call("dards_tc_debug_set", 11, int(os.environ["KBENCH_CB_PERTAP"]))
    if os.environ.get("KBENCH_PAIR"):
        _lib.call("dards_tc_debug_set", 17, int(os.environ["KBENCH_PAIR"]))   # 0: wide layers on the single-CTA kernel instead of the cta_group::2 one
    if os.environ.get("KBENCH_PAIR_STAGES"):
        _lib.call("dards_tc_debug_set", 18, int(os.environ["KBENCH_PAIR_STAGES"]))
    if os.environ.get("KBENCH_WGRAD_PAIR"):
        _lib.call("dards_tc_debug_set", 19, int(os.environ["KBENCH_WGRAD_PAIR"]))   # 0: C >= 256 weight gradients on single CTAs instead of CTA pairs
    if os.environ.get("KBENCH_STAGES"):
        _lib.call("dards_tc_debug_set", 6, int(os.environ["KBENCH_STAGES"]))
    for wname in what:
        {"bn": bn, "conv": conv, "stem": stem, "bnprof": bnprof, "convprof": convprof, "copy": copy, "convbn": convbn}[wname]()
