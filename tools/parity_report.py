"""Parity report on the GPU box: B200 backend vs CPU oracle (fp32) vs CPU oracle in fp64 ("exact").

For every parameter-gradient tensor prints
    e_ref  = err(oracle fp32, oracle fp64)   -- the reference arithmetic's own rounding sensitivity
    e_b200 = err(B200 fp32,  oracle fp64)
    d      = err(B200 fp32,  oracle fp32)
(err = max|a-b| / max|b|), plus the bf16 path's logits error and per-tensor gradient cosine.
Usage: python tools/parity_report.py [resnet18|densenet18] [B]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import deepards_b200 as D  # noqa: E402
from oracle import cnn_linear_oracle as O  # noqa: E402
from tests.helpers import cosine, rel_err  # noqa: E402


def run(backbone, B, seed=21):
    sd = O.cnn_linear_state(backbone, seed=seed, bn_perturb=0.1)
    x = O.synthetic_breaths(B, seed=100)
    t = O.synthetic_targets(B, seed=100)
    o32, l32, g32 = O.forward_backward(sd, x, t)
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    o64, l64, g64 = O.forward_backward(sd64, x.double(), t.double())
    res = {}
    for prec in ("fp32", "bf16"):
        bb = D.resnet18() if backbone == "resnet18" else D.densenet18(drop_rate=0.0)
        net = D.CNNLinearNetwork(bb, 20, 0)
        net.load_state_dict(sd)
        net = net.cuda()
        net.precision = prec
        out = net(x.cuda(), None)
        loss = F.binary_cross_entropy_with_logits(out, t.cuda())
        loss.backward()
        res[prec] = (out.detach().cpu(), float(loss), {n: p.grad.cpu() for n, p in net.named_parameters() if p.grad is not None})
    print("== %s B=%d  loss: fp64 %.7f  ref32 %.7f  b200-fp32 %.7f  b200-bf16 %.7f" %
          (backbone, B, float(l64), float(l32), res["fp32"][1], res["bf16"][1]))
    print("logits: e_ref %.2e  e_b200 %.2e  d %.2e | bf16 err %.2e" %
          (rel_err(o32, o64), rel_err(res["fp32"][0], o64), rel_err(res["fp32"][0], o32), rel_err(res["bf16"][0], o64)))
    print("%-58s %9s %9s %9s | %8s %9s" % ("gradient", "e_ref", "e_b200", "d", "bf16cos", "bf16err"))
    for k in g64:
        g = res["fp32"][2][k]
        gb = res["bf16"][2][k]
        print("%-58s %9.2e %9.2e %9.2e | %8.4f %9.2e" % (k[-58:], rel_err(g32[k], g64[k]), rel_err(g, g64[k]),
                                                       rel_err(g, g32[k]), cosine(gb, g64[k]), rel_err(gb, g64[k])))


if __name__ == "__main__":
    bbs = [sys.argv[1]] if len(sys.argv) > 1 else ["resnet18", "densenet18"]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    for bb in bbs:
        run(bb, B)
