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
BF16_LOGIT_TOL = 8e-2    # logits, max-abs relative (measured 2.0e-2 ResNet-18, 5.0e-2 DenseNet-18)
BF16_LOSS_TOL = 5e-3     # absolute, loss ~0.7 (measured < 1e-3)
BF16_GRAD_COS = 0.82     # every parameter-gradient tensor with >= 64 elements (measured min 0.907 / 0.898)
BF16_GRAD_COS_MEAN = 0.93  # mean over those tensors (measured 0.955 / 0.951)


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
    net.precision = "fp32"
    x = O.synthetic_breaths(2, seed=2).cuda()
    with torch.no_grad():
        a, b = net(x, None), net(x, None)
        assert not torch.equal(a, b)  # drop_rate 0.2 active in train mode (densenet.py:37-39)
        net.eval()
        c, d = net(x, None), net(x, None)
        assert torch.equal(c, d)
        assert rel_err(c.cpu(), O.cnn_linear_forward(sd, x.cpu())) <= FP32_TOL  # BN still uses batch stats


# ---------------------------------------------------------------------------------------------------------------
# bf16 path where the headline lives: the benchmarked batch size, and a training trajectory
# ---------------------------------------------------------------------------------------------------------------
# Gates = about twice the values measured on the B200 (tests/golden/parity_report_r02.txt).  At B = 256 a parameter
# gradient is a sum over 256 sequences, so the bf16 rounding noise averages out and the directions agree better than at
# B = 8.
# measured (profiles/r02_parity_report.txt): ResNet-18 logits 1.9e-2, |dloss| 1e-4, cosine min 0.896 (stem conv) / mean 0.951
# / all gradients as one vector 0.994; DenseNet-18 3.0e-2, 2e-5, 0.560 (stem conv: 448 numbers whose 256-sequence mean nearly
# cancels with random targets, so what is left is mostly rounding; 0.87 with another seed) / 0.938 / 0.753.  The fp32
# reference itself is only good to 5e-3 on these gradients (oracle fp32 vs fp64, tools/parity_report.py), and the
# trajectory test below shows that 200 optimizer steps follow the fp32 run to 1e-2 in the loss.
BF16_B256 = {
    "resnet18": dict(logit=4e-2, loss=1e-3, cos_min=0.80, cos_mean=0.90, cos_all=0.93),
    "densenet18": dict(logit=6e-2, loss=1e-3, cos_min=0.40, cos_mean=0.88, cos_all=0.60),
}


@pytest.mark.parametrize("backbone", ["resnet18", "densenet18"])
def test_bf16_step_at_the_benchmarked_batch_size_vs_oracle(backbone):
    """BASELINE.json configs[1] / [2] size: one bf16 training step on 256 x 20 x 1 x 224 against the fp32 CPU oracle
    (logits, loss, every parameter gradient by direction).  wgrad's split-K count and the tile schedules depend on the
    batch, so this is the configuration bench.py times."""
    B = 256
    skw = dict(backbone=backbone, seed=23, bn_perturb=0.1)
    sd = O.cnn_linear_state(**skw)
    x = O.synthetic_breaths(B, seed=105)
    t = O.synthetic_targets(B, seed=105)
    torch.set_num_threads(max(1, (torch.get_num_threads())))
    ref_out, ref_loss, ref_grads = O.forward_backward(sd, x, t)
    net = build(skw, sd, "bf16")
    out, loss, grads = step(net, x, t)
    g = BF16_B256[backbone]
    e_logit = rel_err(out, ref_out)
    cs = {k: cosine(grads[k].cpu(), v) for k, v in ref_grads.items() if v.numel() >= 64}
    worst = min(cs, key=cs.get)
    mean = sum(cs.values()) / len(cs)
    keys = [k for k in ref_grads]
    c_all = cosine(torch.cat([grads[k].cpu().reshape(-1) for k in keys]), torch.cat([ref_grads[k].reshape(-1) for k in keys]))
    print("%s bf16 B=256: logits rel err %.3e, |dloss| %.3e, grad cosine min %.4f (%s) mean %.4f over %d tensors, all "
          "gradients as one vector %.4f" % (backbone, e_logit, abs(loss - float(ref_loss)), cs[worst], worst, mean, len(cs), c_all))
    assert c_all >= g["cos_all"], c_all
    assert e_logit <= g["logit"], e_logit
    assert abs(loss - float(ref_loss)) <= g["loss"]
    assert cs[worst] >= g["cos_min"], (worst, cs[worst])
    assert mean >= g["cos_mean"], mean


@pytest.mark.parametrize("backbone", ["resnet18", "densenet18"])
def test_bf16_training_trajectory_tracks_the_fp32_oracle(backbone):
    """200 steps of the reference's optimizer (clamp 0.01 + SGD lr 1e-3 momentum 0.9 nesterov wd 1e-4,
    train_ards_detector.py:416-422, 474-476) on the same 4 rotating batches of 8 sequences: the bf16 B200 trainer against
    torch.optim.SGD on the fp32 CPU oracle.  The headline number is a bf16 TRAINING rate, so the claim to check is that the
    trajectory -- not just one step -- follows fp32: per-step loss within a band, the smoothed curves within a tighter
    one, and both runs must actually learn."""
    from deepards_b200.data_parallel import DataParallelTrainer
    steps, B = 200, 8
    kw = dict(initial_planes=16) if backbone == "resnet18" else {}
    sd0 = O.cnn_linear_state(backbone, seed=41, bn_perturb=0.1, **kw)
    batches = [(O.synthetic_breaths(B, seed=600 + i), O.synthetic_targets(B, seed=600 + i)) for i in range(4)]
    # learnable signal: the class decides the sign of a small offset on every breath of the sequence
    batches = [(x + 0.5 * (t[:, :1] * 2 - 1).view(B, 1, 1, 1), t) for x, t in batches]
    ref = {k: v.clone() for k, v in sd0.items()}
    names = [k for k, v in ref.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))]
    leaves = {k: ref[k].clone().requires_grad_(True) for k in names}
    opt = torch.optim.SGD([leaves[k] for k in names], lr=1e-3, momentum=0.9, weight_decay=1e-4, nesterov=True)
    ref_losses = []
    for i in range(steps):
        x, t = batches[i % 4]
        cur = dict(ref)
        cur.update({k: leaves[k].detach() for k in names})
        _, loss, grads = O.forward_backward(cur, x, t, clip_val=0.01, running_update=backbone == "resnet18")
        ref_losses.append(float(loss))
        opt.zero_grad()
        for k in names:
            leaves[k].grad = grads[k].clone() if k in grads else None
        opt.step()
    import deepards_b200 as D
    bb = D.resnet18(**kw) if backbone == "resnet18" else D.densenet18(drop_rate=0.0)
    net = D.CNNLinearNetwork(bb, 20, 0)
    net.load_state_dict({k: v.clone() for k, v in sd0.items()})
    net = net.cuda().train()
    net.precision = "bf16"
    tr = DataParallelTrainer(net, lr=1e-3, optimizer="sgd", weight_decay=1e-4, clip_val=0.01, use_graph=True)
    dev = [(x.cuda(), t.cuda()) for x, t in batches]
    losses = [float(tr.train_step(*dev[i % 4])) for i in range(steps)]
    a, b = torch.tensor(losses), torch.tensor(ref_losses)
    smooth = lambda v: v.unfold(0, 20, 1).mean(1)  # noqa: E731
    d_step = float((a - b).abs().max())
    d_smooth = float((smooth(a) - smooth(b)).abs().max())
    print("%s bf16 trajectory: loss %.4f -> %.4f (fp32 oracle %.4f -> %.4f); max |dloss| per step %.4f, smoothed %.4f" %
          (backbone, float(a[:8].mean()), float(a[-8:].mean()), float(b[:8].mean()), float(b[-8:].mean()), d_step, d_smooth))
    assert float(b[-8:].mean()) < float(b[:8].mean()) - 0.02, "the fp32 oracle run did not learn: the test has no signal"
    assert float(a[-8:].mean()) < float(a[:8].mean()) - 0.02
    # measured: ResNet-18 0.0054 / 0.0025, DenseNet-18 0.0122 / 0.0064 (profiles/r02_parity_report.txt)
    assert d_step <= 0.025, d_step
    assert d_smooth <= 0.013, d_smooth


def test_reference_clamp_hooks_on_every_parameter_incl_the_unused_ones():
    """train_ards_detector.py:474-476 registers `lambda grad: torch.clamp(grad, -clip, clip)` on EVERY parameter.  ResNet's
    conv1_alt / conv2 / bn2 never take part in forward(): like in the reference they are not in the autograd graph, their
    hook is never called (a None gradient would make clamp raise) and their .grad stays None; the optimizer skips them."""
    import deepards_b200 as D
    torch.manual_seed(0)
    net = D.CNNLinearNetwork(D.resnet18(initial_planes=16), 20, 0).cuda().train()
    net.precision = "fp32"
    calls = []
    for n, p in net.named_parameters():
        p.register_hook(lambda g, n=n: (calls.append(n), torch.clamp(g, -0.01, 0.01))[1])
    opt = torch.optim.SGD(net.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4, nesterov=True)
    x, t = O.synthetic_breaths(3, seed=9).cuda(), O.synthetic_targets(3, seed=9).cuda()
    for _ in range(2):
        opt.zero_grad()
        F.binary_cross_entropy_with_logits(net(x, None), t).backward()
        opt.step()
    unused = [n for n, p in net.named_parameters() if p.grad is None]
    assert sorted(unused) == sorted(n for n, _ in net.named_parameters()
                                    if n.startswith(("breath_block.conv1_alt.", "breath_block.conv2.", "breath_block.bn2.")))
    assert not set(unused) & set(calls)
    assert all(float(p.grad.abs().max()) <= 0.01 + 1e-9 for p in net.parameters() if p.grad is not None)


def test_dropout_masks_do_not_depend_on_the_sharding():
    """DenseNet in train() mode (drop_rate 0.2 active, densenet.py:33-39): the masks are keyed by the global sequence
    index, so running sequences [2, 4) of a batch as their own call with sequence offset 2 reproduces their logits inside
    the whole batch bit for bit -- the single-GPU statement of "a 2-rank step equals the 1-rank step with dropout on"
    (tests/test_dp_multi_gpu.py checks the real thing on two devices)."""
    import deepards_b200 as D
    from deepards_b200 import engine
    sd = O.cnn_linear_state("densenet18", seed=51, bn_perturb=0.1)
    x = O.synthetic_breaths(4, seed=77).cuda()

    def run(xs, first):
        torch.manual_seed(0)
        net = D.CNNLinearNetwork(D.densenet18(), 20, 0)
        net.load_state_dict(sd)
        net = net.cuda().train()
        net.precision = "bf16"
        from deepards_b200.torch_cnn_linear_network import _drop_key
        plan = engine.get_plan(net, net.breath_block, net.linear_final, xs.shape[0] * 20, 20, "bf16", "cnn_linear",
                               dropout=_drop_key(net.breath_block), update_running=True)
        assert plan.dropout
        plan.set_sequence_offset(first)
        plan.load_input(xs)
        plan.run_forward()
        plan.mark_no_backward()
        torch.cuda.synchronize()
        return plan.logits.clone()

    whole = run(x, 0)
    assert torch.equal(run(x[2:], 2), whole[2:])
    assert torch.equal(run(x[:2], 0), whole[:2])
    assert not torch.equal(run(x[2:], 0), whole[2:])      # without the offset the shard would reuse the masks of [0, 2)


def test_dropout_masks_of_a_ragged_split_over_several_steps():
    """The 3 + 2 split of tests/test_dp_multi_gpu.py, forward only, fp32, five consecutive steps of the same plans (the
    step counter is part of the Philox key): each shard's logits equal its rows of the whole batch bit for bit at every
    step, and the masks change from step to step."""
    import deepards_b200 as D
    from deepards_b200 import engine
    from deepards_b200.torch_cnn_linear_network import _drop_key
    sd = O.cnn_linear_state("densenet18", seed=41, bn_perturb=0.1)
    x = O.synthetic_breaths(5, seed=500).cuda()

    def plan_for(n_seq, first):
        net = D.CNNLinearNetwork(D.densenet18(drop_rate=0.2), 20, 0)
        net.load_state_dict(sd)
        net = net.cuda().train()
        net.precision = "fp32"
        plan = engine.get_plan(net, net.breath_block, net.linear_final, n_seq * 20, 20, "fp32", "cnn_linear",
                               dropout=_drop_key(net.breath_block), update_running=True)
        assert plan.dropout
        plan.set_sequence_offset(first)
        return net, plan

    (n_w, whole), (n_a, a), (n_b, b) = plan_for(5, 0), plan_for(3, 0), plan_for(2, 3)
    prev = None
    for step in range(5):
        outs = []
        for plan, xs in ((whole, x), (a, x[:3]), (b, x[3:])):
            plan.load_input(xs)
            plan.run_forward()
            plan.mark_no_backward()
            torch.cuda.synchronize()
            outs.append(plan.logits.clone())
        assert torch.equal(outs[1], outs[0][:3]) and torch.equal(outs[2], outs[0][3:]), step
        assert prev is None or not torch.equal(prev, outs[0])
        prev = outs[0]


def test_dropout_backward_masks_of_a_ragged_split_add_up():
    """Backward of the same 3 + 2 split: with the loss sum(w * logits), the parameter gradient of the whole batch is the
    sum of the two shards' gradients (each shard regenerates its masks from the global sequence index in the backward
    pass too) -- three consecutive steps, fp32."""
    import deepards_b200 as D
    from deepards_b200 import engine
    from deepards_b200.torch_cnn_linear_network import _drop_key
    sd = O.cnn_linear_state("densenet18", seed=41, bn_perturb=0.1)
    x = O.synthetic_breaths(5, seed=501).cuda()
    w = torch.randn(5, 2, device="cuda")
    keep = []

    def plan_for(n_seq, first):
        net = D.CNNLinearNetwork(D.densenet18(drop_rate=0.2), 20, 0)
        net.load_state_dict(sd)
        net = net.cuda().train()
        net.precision = "fp32"
        keep.append(net)
        plan = engine.get_plan(net, net.breath_block, net.linear_final, n_seq * 20, 20, "fp32", "cnn_linear",
                               dropout=_drop_key(net.breath_block), update_running=True)
        plan.set_sequence_offset(first)
        return plan

    plans = [(plan_for(5, 0), x, w), (plan_for(3, 0), x[:3], w[:3]), (plan_for(2, 3), x[3:], w[3:])]
    for step in range(3):
        grads = []
        for plan, xs, ws in plans:
            plan.load_input(xs)
            plan.run_forward()
            plan.dlogits.copy_(ws.reshape(plan.dlogits.shape))
            plan.run_backward()
            torch.cuda.synchronize()
            grads.append(plan.grad_flat.clone())
        err = float((grads[1] + grads[2] - grads[0]).abs().max() / grads[0].abs().max())
        assert err <= 1e-5, (step, err)
