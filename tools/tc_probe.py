"""Probe of the tcgen05 wgrad descriptor variants on the GPU (prints, never asserts).
mode 0: descriptor base_offset = (start >> 7) & 7 for row-shifted starts; mode 1: base_offset = 0."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepards_b200 import _lib, kernels as K  # noqa: E402
from tests.helpers import rel_err  # noqa: E402

CASES = [(40, 64, 64, 56, 3, 1, 1), (40, 64, 128, 56, 3, 2, 1), (40, 64, 128, 56, 1, 2, 0), (20, 128, 128, 28, 3, 1, 1),
         (33, 512, 512, 7, 3, 1, 1), (20, 96, 128, 28, 1, 1, 0), (20, 128, 32, 14, 3, 1, 1), (256, 64, 64, 56, 3, 1, 1)]


def cl(x):
    return x.permute(0, 2, 1).contiguous()


for mode in (0, 1):
    _lib.call("dards_tc_debug_set", 3, mode)
    for case in CASES:
        n, cin, cout, l, k, s, p = case
        g = torch.Generator().manual_seed(7)
        x = torch.randn(n, cin, l, generator=g).cuda().bfloat16()
        lo = (l + 2 * p - k) // s + 1
        dy = torch.randn(n, cout, lo, generator=g).cuda().bfloat16()
        ref = K.conv1d_wgrad(cl(x), cl(dy), k, s, p, impl=0)
        try:
            got = K.conv1d_wgrad(cl(x), cl(dy), k, s, p, impl=1)
            torch.cuda.synchronize()
            per_tap = [rel_err(got[:, :, t], ref[:, :, t]) for t in range(k)]
            print("mode %d case %s: err %.3e per-tap %s" % (mode, case, rel_err(got, ref), ["%.2e" % e for e in per_tap]), flush=True)
        except Exception as e:  # noqa: BLE001
            print("mode %d case %s: EXCEPTION %s" % (mode, case, str(e)[:200]), flush=True)
            sys.exit(1)
