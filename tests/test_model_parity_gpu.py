"""GPU: whole-network parity of the B200 backend.

  * golden vectors recorded from the reference's own modules (tests/golden, oracle/make_golden.py):
    logits, loss, every parameter gradient, BatchNorm running statistics, GradCAM tensors -- fp32, <= 1e-4 rel
    (relative to the max-abs of each tensor: the tolerance BASELINE.json's north_star states)
  * the CPU oracle on fresh seeded inputs at config-1 size (B = 16)
  * the bf16 / tcgen05 path against the oracle with the separately stated looser tolerance
  * size-independent properties at config-2 size (B = 256): shard exactness and gradient additivity
"""
import io

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import cnn_linear_oracle as O  # noqa: E402
from tests.helpers import (CASES, cosine, golden_grad_errors, load_case, pinned_decision_check, plan_of,  # noqa: E402
                           rel_err)

FP32_TOL = 1e-4          # north_star: logits and gradients within 1e-4 relative error in fp32
# bf16 storage + tcgen05 path (the "separately stated, looser tolerance" of the north_star).  bf16 rounding (2^-9)
# of every stored activation flips ~0.3 % of the ReLU decisions, so gradients are compared by direction:
BF16_LOGIT_TOL = 1e-1    # logits, max-abs relative (measured 2e-2 .. 5e-2)
BF16_LOSS_TOL = 2e-2     # absolute, loss ~0.7
BF16_GRAD_COS = 0.80     # every parameter-gradient tensor with >= 64 elements (measured min 0.84)
BF16_GRAD_COS_MEAN = 0.93  # mean over those tensors (measured 0.95 .. 0.97)


def build(name_or_kw, sd, precision="fp32", per_breath=False, fkw=None):
    import deepards_b200 as D
    kw = dict(name_or_kw)
    backbone = kw.pop("backbone")
    kw.pop("seed", None)
    kw.pop("bn_perturb", None)
    kw.pop("per_breath", None)
    fkw = fkw or {}
    if backbone == "resnet18":
        bb = D.resnet18(first_pool_type=fkw.get("first_pool_type", "max"), **kw)
    else:
        bb = D.densenet18(**kw)
        for m in bb.modules():
            if hasattr(m, "drop_rate"):
                m.drop_rate = 0.0  # parity needs a deterministic step (SURVEY.md point 4)
    net = D.CNNSingleBreathLinearNetwork(bb) if per_breath else D.CNNLinearNetwork(bb, 20, 0)
    net.load_state_dict(sd, strict=True)
    net = net.cuda()
    net.precision = precision
    net.train()
    return net


def step(net, x, t):
    net.zero_grad()
    out = net(x.cuda(), None)
    loss = F.binary_cross_entropy_with_logits(out, t.cuda())
    loss.backward()
    grads = {n: p.grad for n, p in net.named_parameters()}
    return out.detach().cpu(), float(loss), grads


@pytest.mark.parametrize("name", sorted(CASES))
def test_fp32_matches_reference_golden(name):
    z, sd, fkw, per_breath = load_case(name)
    net = build(CASES[name][0], sd, "fp32", per_breath, fkw)
    out, loss, grads = step(net, torch.from_numpy(z["x"]), torch.from_numpy(z["target"]))
    assert rel_err(out, z["logits"]) <= FP32_TOL, rel_err(out, z["logits"])
    assert abs(loss - float(z["loss"])) <= FP32_TOL
    errs = golden_grad_errors(z, grads)
    assert len(errs) > 10
    if max(errs.values()) > FP32_TOL:
        # some tensor is off by more than 1e-4 from the recorded fp32 reference run: that is only acceptable if a
        # ReLU decision flipped at a near-tie -- prove it, and hold the arithmetic to 1e-4 with decisions pinned
        x, t = torch.from_numpy(z["x"]), torch.from_numpy(z["target"])
        flips, worst = pinned_decision_check(plan_of(net, x.shape[0] * x.shape[1]), sd, x, t, grads, FP32_TOL, name,
                                             per_breath=per_breath, **fkw)
        assert flips > 0, "gradients differ from the golden run although every ReLU decision agrees"
    else:
        flips, worst = 0, max(errs.values())
    print("%s: worst grad rel err %.2e (decisions pinned: %s, flipped near-tie decisions: %d)" %
          (name, worst, flips > 0, flips))
    sd_after = net.state_dict()
    for key in z.files:
        if key.startswith("buf/"):
            got = sd_after[key[4:]].cpu()
            if key.endswith("num_batches_tracked"):
                assert int(got) == int(z[key]), key
            else:
                assert rel_err(got, z[key]) <= FP32_TOL, key


@pytest.mark.parametrize("backbone", ["resnet18", "densenet18"])
def test_fp32_matches_oracle_config1(backbone):
    """BASELINE.json configs[0]: batch 16 x 20 x 1 x 224, fp32 forward+backward, against the CPU oracle."""
    skw = dict(backbone=backbone, seed=21, bn_perturb=0.1)
    sd = O.cnn_linear_state(**skw)
    x = O.synthetic_breaths(16, seed=100)
    t = O.synthetic_targets(16, seed=100)
    ref_out, ref_loss, ref_grads = O.forward_backward(sd, x, t)
    net = build(skw, sd, "fp32")
    out, loss, grads = step(net, x, t)
    assert rel_err(out, ref_out) <= FP32_TOL
    assert abs(loss - float(ref_loss)) <= FP32_TOL
    flips, worst = pinned_decision_check(plan_of(net, 16 * 20), sd, x, t, grads, FP32_TOL, backbone)
    raw = max(rel_err(grads[k].cpu(), g) for k, g in ref_grads.items())
    print("%s config1: %d near-tie decisions flipped; worst grad rel err %.2e with decisions pinned (%.2e raw)" %
          (backbone, flips, worst, raw))
    for k, g in grads.items():
        if k not in ref_grads:
            assert g is None, k  # conv1_alt / conv2 / bn2 never receive a gradient


@pytest.mark.parametrize("backbone", ["resnet18", "densenet18"])
def test_bf16_tensor_core_path_close_to_oracle(backbone):
    skw = dict(backbone=backbone, seed=22, bn_perturb=0.1)
    sd = O.cnn_linear_state(**skw)
    x = O.synthetic_breaths(8, seed=101)
    t = O.synthetic_targets(8, seed=101)
    ref_out, ref_loss, ref_grads = O.forward_backward(sd, x, t)
    net = build(skw, sd, "bf16")
    out, loss, grads = step(net, x, t)
    assert rel_err(out, ref_out) <= BF16_LOGIT_TOL, rel_err(out, ref_out)
    assert abs(loss - float(ref_loss)) <= BF16_LOSS_TOL
    cs = []
    for k, g in ref_grads.items():
        if g.numel() >= 64:
            c = cosine(grads[k].cpu(), g)
            cs.append(c)
            assert c >= BF16_GRAD_COS, (k, c)
    assert sum(cs) / len(cs) >= BF16_GRAD_COS_MEAN, sum(cs) / len(cs)
    print("%s bf16: logits rel err %.2e, grad cosine min %.4f mean %.4f" % (backbone, rel_err(out, ref_out), min(cs),
                                                                          sum(cs) / len(cs)))


def test_bf16_simt_and_tcgen05_agree(monkeypatch):
    """Same bf16 storage, CUDA-core convs vs tcgen05 convs: network outputs must agree to bf16 noise."""
    skw = dict(backbone="resnet18", seed=23, bn_perturb=0.1)
    sd = O.cnn_linear_state(**skw)
    x = O.synthetic_breaths(4, seed=102)
    t = O.synthetic_targets(4, seed=102)
    monkeypatch.setenv("DEEPARDS_B200_CONV_IMPL", "simt")
    o1, l1, g1 = step(build(skw, sd, "bf16"), x, t)
    monkeypatch.delenv("DEEPARDS_B200_CONV_IMPL")
    o2, l2, g2 = step(build(skw, sd, "bf16"), x, t)
    assert rel_err(o2, o1) < 3e-2
    for k in g1:
        if g1[k] is not None and g1[k].numel() >= 64:
            # same storage precision, different summation order + ReLU flips; B = 4 sequences only, so single flips
            # show: measured minimum 0.965 (layer1 BN biases), typical > 0.99
            assert cosine(g2[k], g1[k]) > 0.95, k


def test_gradcam_tensors_match_reference():
    """gradcam.py:40-65, 83-99 through the B200 `features` module: A, dA and the model output."""
    z, sd, _, _ = load_case("densenet18_B2_real")
    net = build(CASES["densenet18_B2_real"][0], sd, "fp32")
    x = torch.from_numpy(z["x"][0]).cuda()
    a = net.breath_block.features(x)
    saved = {}
    a.register_hook(lambda g: saved.__setitem__("dA", g))
    y = net.breath_block.avgpool(F.relu(a)).view(-1)
    mo = net.linear_final(y).unsqueeze(0)
    net.zero_grad()
    mo[0, int(z["cam/target"])].backward()
    assert rel_err(a.detach().cpu(), z["cam/A"]) <= FP32_TOL
    assert rel_err(saved["dA"].cpu(), z["cam/dA"]) <= FP32_TOL
    assert rel_err(mo.detach().cpu(), z["cam/out"]) <= FP32_TOL
    # the full backward ran through the conv stack: parameters have gradients
    assert net.breath_block.features.conv0.weight.grad is not None
    # forward_no_pool (ProtoPNet hook, densenet.py:191-193)
    assert rel_err(net.breath_block.forward_no_pool(x).detach().cpu(), np.maximum(z["cam/A"], 0)) <= FP32_TOL


def test_backbone_standalone_and_group_is_batch():
    sd = O.cnn_linear_state("resnet18", seed=24, initial_planes=16, bn_perturb=0.1)
    net = build(dict(backbone="resnet18", initial_planes=16), sd, "fp32")
    x = O.synthetic_breaths(2, seed=3).reshape(40, 1, 224)
    ref = O.backbone_forward(sd, x)  # one BatchNorm group of 40 breaths
    got = net.breath_block(x.cuda())
    assert got.shape == (40, 128)
    assert rel_err(got.detach().cpu(), ref) <= FP32_TOL


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_shard_exactness_and_gradient_additivity_at_full_size(precision):
    """config 2 size (B = 256): every sequence is independent of the others (BatchNorm per sequence), so
    out(x)[i:j] == out(x[i:j]) bit for bit, and the batch gradient is the mean of the shard gradients --
    which is what makes data-parallel sharding exact (SURVEY.md point 3)."""
    skw = dict(backbone="resnet18", seed=25, bn_perturb=0.1)
    sd = O.cnn_linear_state(**skw)
    net = build(skw, sd, precision)
    B, S = 256, 64
    x = O.synthetic_breaths(B, seed=103)
    t = O.synthetic_targets(B, seed=103)
    out, loss, grads = step(net, x, t)
    full = {k: g.clone() for k, g in grads.items() if g is not None}
    acc = {k: torch.zeros_like(g) for k, g in full.items()}
    for i in range(0, B, S):
        o, l, g = step(net, x[i:i + S], t[i:i + S])
        assert torch.equal(o, out[i:i + S]), "sequences are not independent"
        for k in acc:
            acc[k] += g[k] * (S / B)
    tol = 2e-4 if precision == "fp32" else 3e-2
    for k in acc:
        assert rel_err(acc[k], full[k]) <= tol, (k, rel_err(acc[k], full[k]))


def test_ragged_and_tiny_batches():
    """B = 1 and an odd B; the reference trims odd batches to even but must accept any B >= 1 at test time."""
    skw = dict(backbone="densenet18", seed=26, bn_perturb=0.1)
    sd = O.cnn_linear_state(**skw)
    net = build(skw, sd, "fp32")
    for b in (1, 5):
        x = O.synthetic_breaths(b, seed=104 + b)
        with torch.no_grad():
            got = net(x.cuda(), None).cpu()
        assert got.shape == (b, 2)
        assert rel_err(got, O.cnn_linear_forward(sd, x)) <= FP32_TOL


def test_interface_errors_and_pickling():
    import deepards_b200 as D
    sd = O.cnn_linear_state("resnet18", seed=27, initial_planes=16)
    net = build(dict(backbone="resnet18", initial_planes=16), sd, "fp32")
    with pytest.raises(Exception, match="sequence length of 224"):
        net(torch.zeros(1, 20, 1, 200, device="cuda"), None)
    with pytest.raises(NotImplementedError):
        net(torch.zeros(1, 20, 1, 224, device="cuda", requires_grad=True), None)
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        D.CNNLinearNetwork(D.resnet18(initial_planes=16), 20, 0)(torch.zeros(1, 20, 1, 224), None)
    x = O.synthetic_breaths(2, seed=1).cuda()
    with torch.no_grad():
        before = net(x, None)
    buf = io.BytesIO()
    torch.save(net, buf)  # train_ards_detector.py:364 pickles the whole module
    buf.seek(0)
    net2 = torch.load(buf, weights_only=False)
    with torch.no_grad():
        after = net2(x, None)
    assert torch.equal(before, after)
    # backward twice without a forward in between must fail loudly, not reuse stale activations
    out = net(x, None)
    out.sum().backward()
    with pytest.raises(RuntimeError):
        out.sum().backward()


def test_dropout_training_mode_is_stochastic_but_eval_like_when_off():
    import deepards_b200 as D
    sd = O.cnn_linear_state("densenet18", seed=28)
    net = D.CNNLinearNetwork(D.densenet18(), 20, 0)
    net.load_state_dict(sd)
    net = net.cuda().train()
    x = O.synthetic_breaths(2, seed=2).cuda()
    with torch.no_grad():
        a, b = net(x, None), net(x, None)
        assert not torch.equal(a, b)  # drop_rate 0.2 active in train mode (densenet.py:37-39)
        net.eval()
        c, d = net(x, None), net(x, None)
        assert torch.equal(c, d)
        assert rel_err(c.cpu(), O.cnn_linear_forward(sd, x.cpu())) <= FP32_TOL  # BN still uses batch stats
