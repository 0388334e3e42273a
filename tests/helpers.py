"""Shared helpers for the parity tests (CPU and GPU)."""
import os

import numpy as np
import torch

from oracle import cnn_linear_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SAMPLE_STRIDE = 61

# name -> (oracle state kwargs, forward kwargs, per_breath)
CASES = {
    "resnet18_p64_B2_randn": (dict(backbone="resnet18", seed=1, bn_perturb=0.1), {}, False),
    "resnet18_p16_B3_synth": (dict(backbone="resnet18", seed=2, bn_perturb=0.1, initial_planes=16), {}, False),
    "resnet18_p16_B2_avgpool": (dict(backbone="resnet18", seed=3, bn_perturb=0.1, initial_planes=16),
                                dict(first_pool_type="avg"), False),
    "densenet18_B2_real": (dict(backbone="densenet18", seed=4, bn_perturb=0.1), {}, False),
    "densenet18_B3_synth": (dict(backbone="densenet18", seed=5, bn_perturb=0.1), {}, False),
    "resnet18_p16_B2_perbreath": (dict(backbone="resnet18", seed=6, bn_perturb=0.1, initial_planes=16, per_breath=True),
                                  {}, True),
}


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    skw, fkw, per_breath = CASES[name]
    sd = O.cnn_linear_state(**skw)
    return z, sd, fkw, per_breath


def rel_err(a, b):
    """max |a-b| / max |b|  -- the tolerance definition of SURVEY.md section 8c."""
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    denom = float(b.abs().max())
    if denom == 0.0:
        return float((a - b).abs().max())
    return float((a - b).abs().max()) / denom


def cosine(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def check_grads_against_golden(z, grads, tol, what=""):
    """grads: {state_dict name: tensor}.  Compares with full / sampled golden gradients."""
    worst = 0.0
    n = 0
    for key in z.files:
        if key.startswith("grad/"):
            name = key[5:]
            e = rel_err(grads[name].detach().cpu(), z[key])
        elif key.startswith("gradsample/"):
            name = key[11:]
            g = grads[name].detach().cpu().reshape(-1)
            ref = z[key]
            # relative to the max of the WHOLE tensor is not available; the sample max is a lower bound
            e = rel_err(g[::SAMPLE_STRIDE], ref)
            st = z["gradstat/" + name]
            l2 = float(g.double().norm())
            assert abs(l2 - st[1]) <= max(tol, 1e-5) * st[1] + 1e-12, (what, name, "l2", l2, st[1])
        elif key.startswith("nograd/"):
            name = key[7:]
            assert grads.get(name) is None or float(torch.as_tensor(grads[name]).abs().max()) == 0.0, (what, name)
            continue
        else:
            continue
        assert e <= tol, "%s grad %s rel err %.3e > %.1e" % (what, name, e, tol)
        worst = max(worst, e)
        n += 1
    assert n > 0
    return worst
