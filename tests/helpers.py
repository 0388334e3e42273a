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
    # all 20 real sequences of the reference's tests/test_dataset.pkl (oracle/make_golden_real20.py)
    "densenet18_B20_real_all": (dict(backbone="densenet18", seed=7, bn_perturb=0.1), {}, False),
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


def golden_grad_errors(z, grads):
    """{param name: max|mine-golden| / max|golden|} over the full / sampled golden gradients."""
    out = {}
    for key in z.files:
        if key.startswith("grad/"):
            name = key[5:]
            out[name] = rel_err(grads[name].detach().cpu(), z[key])
        elif key.startswith("gradsample/"):
            name = key[11:]
            out[name] = rel_err(grads[name].detach().cpu().reshape(-1)[::SAMPLE_STRIDE], z[key])
        elif key.startswith("nograd/"):
            name = key[7:]
            g = grads.get(name)
            assert g is None or float(torch.as_tensor(g).abs().max()) == 0.0, name
    return out


def reference_sensitivity(sd, x, t, n_perturb=4, **kw):
    """How far the REFERENCE arithmetic's own fp32 gradients are from the exact (fp64) ones on this step.

    Runs the oracle in fp64 once and in fp32 on the nominal input plus `n_perturb` copies of it scaled by
    (1 + s * 2^-21) -- a few ulps, i.e. nothing but rounding noise.  Each run takes a different set of
    near-tie ReLU / max-pool decisions; the worst max-abs-relative error over runs and tensors is the step's
    inherent fp32 sensitivity.  Returns (grads fp32 nominal, grads fp64, sensitivity)."""
    _, _, g64 = oracle_fp64(sd, x, t, **kw)
    _, _, g32 = O.forward_backward({k: v.clone() for k, v in sd.items()}, x, t, **kw)
    sens = max(rel_err(g32[k], g64[k]) for k in g64)
    for s_ in [1, -1, 2, -2][:n_perturb]:
        _, _, gp = O.forward_backward({k: v.clone() for k, v in sd.items()}, x * (1.0 + s_ * 2.0 ** -21), t, **kw)
        sens = max(sens, max(rel_err(gp[k], g64[k]) for k in g64))
    return g32, g64, sens


def conditioned_grad_check(mine, ref32, ref64, tol, what="", sensitivity=None):
    """The fp32 gate of the parity tests.

    A gradient tensor passes if it is within `tol` (max-abs relative) of the fp32 reference result, OR -- where
    the step itself is ill-conditioned -- if it is as close to the exact (fp64) gradient as the reference's own
    fp32 arithmetic gets on that step (factor 3 on `sensitivity`, see reference_sensitivity; hard ceiling 0.2).
    Why the second clause exists: with >10^7 ReLU / max-pool decisions per step, dozens of pre-activations sit
    within fp32 rounding of zero, and one flipped decision in a late layer moves the gradient of every layer
    below it by 1e-3..5e-2 of its max.  The reference's own fp32-vs-fp64 difference shows exactly that
    ("Parity"), so NO independent fp32 implementation -- including the reference on another CPU --
    can be held to 1e-4 on those tensors.  Logits and loss are continuous in rounding noise and stay at 1e-4.
    Returns (worst strict error, number of tensors that needed the conditioned clause)."""
    worst, conditioned, bad = 0.0, 0, []
    if sensitivity is None:
        sensitivity = max(rel_err(ref32[k], g64) for k, g64 in ref64.items())
    allow = min(0.2, max(tol, 3.0 * sensitivity))
    for k, g64 in ref64.items():
        m = mine[k].detach().cpu()
        d = rel_err(m, ref32[k])
        if d <= tol:
            worst = max(worst, d)
            continue
        e_mine = rel_err(m, g64)
        if e_mine <= allow:
            conditioned += 1
        else:
            bad.append("%s: |b200-ref32| %.2e, |b200-fp64| %.2e > allowance %.2e (reference fp32 sensitivity %.2e)" %
                       (k, d, e_mine, allow, sensitivity))
    assert not bad, "%s gradients out of tolerance:\n  %s" % (what, "\n  ".join(bad))
    return worst, conditioned


def oracle_fp64(sd, x, t, **kw):
    sd64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    return O.forward_backward(sd64, x.double(), t.double(), **kw)


STEM_PARAMS = ("breath_block.conv1.weight", "breath_block.bn1.weight", "breath_block.bn1.bias",
               "breath_block.features.conv0.weight", "breath_block.features.norm0.weight",
               "breath_block.features.norm0.bias")


def pinned_decision_check(plan, sd, x, t, mine, tol, what="", per_breath=False, **kw):
    """fp32 gradient gate that is immune to ReLU decision flips.

    A training step takes >10^6 discrete decisions (is this unit active?).  Two correct fp32 implementations
    disagree on the handful of units whose pre-activation lies within rounding noise of zero, and ONE such
    disagreement in a late layer moves whole gradient tensors by 1e-3..5e-2 of their max -- so comparing raw
    gradients at 1e-4 tests luck, not correctness.  This check separates the two questions:

      1. decisions: every ReLU mask of the B200 forward (read from the plan's activation buffers) equals the
         oracle's, except at units that are provably near-ties (|z| < 1e-4 of the site's mean |z| in the oracle);
      2. arithmetic: with the oracle's decisions pinned to the B200 masks (oracle.DecisionHooks), every parameter
         gradient agrees within `tol` (max-abs relative) -- the north_star's 1e-4, strictly.

    The stem's ReLU+max-pool decisions are recomputed inside the fused stem kernel and cannot be read back; the
    three stem parameters are therefore held to `tol` when no stem unit is a near-tie and to 5e-2 otherwise.
    Returns (number of flipped decisions, worst pinned error)."""
    b, g = x.shape[0], x.shape[1]
    masks = {}
    for site, buf in plan.sites.items():
        n, l, c = buf.shape
        masks[site] = (buf.float() > 0).permute(0, 2, 1).reshape(b, g, c, l).cpu()
    rec = {}
    O.forward_backward({k: v.clone() for k, v in sd.items()}, x, t, per_breath=per_breath,
                       hooks=O.DecisionHooks(record=rec), **kw)
    flips = 0
    for site, m in masks.items():
        z = torch.stack([rec[site][i] for i in range(b)])
        dis = (z > 0) != m
        k = int(dis.sum())
        if k:
            flips += k
            margin = float(z[dis].abs().max() / z.abs().mean())
            assert margin < 1e-4, "%s: %d decisions at %s differ and are NOT near-ties (margin %.2e)" % (what, k, site, margin)
    _, _, g_pin = O.forward_backward({k: v.clone() for k, v in sd.items()}, x, t, per_breath=per_breath,
                                     hooks=O.DecisionHooks(masks=masks), **kw)
    z0 = torch.stack([rec["relu0"][i] for i in range(b)])
    stem_near_tie = float(z0.abs().min() / z0.abs().mean()) < 1e-5
    worst, bad = 0.0, []
    for k, gp in g_pin.items():
        e = rel_err(mine[k].detach().cpu(), gp)
        limit = tol
        if k in STEM_PARAMS and e > tol and stem_near_tie:
            limit = 5e-2
        if e > limit:
            bad.append("%s: %.2e > %.1e" % (k, e, limit))
        elif limit == tol:
            worst = max(worst, e)
    assert not bad, "%s: gradients differ with decisions pinned (flips=%d):\n  %s" % (what, flips, "\n  ".join(bad))
    return flips, worst


def plan_of(net, n_breaths, precision="fp32"):
    plans = [p for p in net.__dict__.get("_dards_plans", {}).values() if p.N == n_breaths and p.precision == precision]
    assert plans, "no plan was built"
    return max(plans, key=lambda p: p.fwd_serial)
