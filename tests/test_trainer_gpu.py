"""GPU: the training step as a whole (DataParallelTrainer: forward + BCEWithLogits + backward + clamp + optimizer) and
the large-batch inference sweep (BASELINE.json configs[3]) through size-independent properties."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import cnn_linear_oracle as O  # noqa: E402
from tests.helpers import rel_err  # noqa: E402


def _net(backbone, sd, precision, **kw):
    import deepards_b200 as D
    bb = D.resnet18(**kw) if backbone == "resnet18" else D.densenet18(drop_rate=0.0)
    net = D.CNNLinearNetwork(bb, 20, 0)
    net.load_state_dict(sd)
    net = net.cuda().train()
    net.precision = precision
    return net


@pytest.mark.parametrize("backbone,kw", [("resnet18", dict(initial_planes=16)), ("densenet18", {})])
def test_trainer_steps_match_torch_sgd_on_the_oracle(backbone, kw):
    """Three steps of clamp(+-0.01) + SGD(lr 1e-3, momentum 0.9, nesterov, wd 1e-4) (train_ards_detector.py:416-422,
    474-476): the fused flat-buffer update after the B200 backward == torch.optim.SGD on the oracle's clamped gradients.
    The eager and the CUDA-graph replay of the step are bit-identical."""
    from deepards_b200.data_parallel import DataParallelTrainer
    sd0 = O.cnn_linear_state(backbone, seed=31, bn_perturb=0.1, **kw)
    batches = [(O.synthetic_breaths(4, seed=200 + i), O.synthetic_targets(4, seed=200 + i)) for i in range(4)]
    # ---- oracle + torch.optim.SGD on CPU ----
    ref = {k: v.clone() for k, v in sd0.items()}
    names = [k for k, v in ref.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))]
    leaves = {k: ref[k].clone().requires_grad_(True) for k in names}
    opt = torch.optim.SGD([leaves[k] for k in names], lr=1e-3, momentum=0.9, weight_decay=1e-4, nesterov=True)
    ref_losses = []
    for x, t in batches:
        cur = dict(ref)
        cur.update({k: leaves[k].detach() for k in names})
        _, loss, grads = O.forward_backward(cur, x, t, clip_val=0.01, running_update=backbone == "resnet18")
        ref_losses.append(float(loss))
        opt.zero_grad()
        for k in names:
            leaves[k].grad = grads[k].clone() if k in grads else None
        opt.step()
    # ---- B200: eager, then graph ----
    results = {}
    for use_graph in (False, True):
        net = _net(backbone, {k: v.clone() for k, v in sd0.items()}, "fp32", **kw)
        tr = DataParallelTrainer(net, lr=1e-3, optimizer="sgd", weight_decay=1e-4, clip_val=0.01, use_graph=use_graph)
        losses = []
        for x, t in batches:
            losses.append(float(tr.train_step(x.cuda(), t.cuda())))
        torch.cuda.synchronize()
        results[use_graph] = (losses, {k: v.detach().cpu().clone() for k, v in net.state_dict().items()})
    losses, params = results[False]
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 1e-4, (losses, ref_losses)
    # integration tolerance: a ReLU / max-pool decision that flips at a near-tie moves single gradient elements by
    # about the clamp value, i.e. a parameter by ~lr * clip per step (the update arithmetic itself is held to 1e-6 by
    # test_fused_optimizers_match_torch)
    for k in names:
        assert rel_err(params[k], leaves[k].detach()) < 1e-3, k
    g_losses, g_params = results[True]
    assert g_losses == losses
    for k in params:
        assert torch.equal(params[k], g_params[k]), k


@pytest.mark.parametrize("backbone", ["resnet18", "densenet18"])
def test_inference_sweep_padded_single_breath_classifier(backbone):
    """configs[3]: CNNSingleBreathLinearNetwork on padded breaths (zeros after the breath, dataset.py:1233-1237), forward
    only, 16 384 breaths in one call.  Sequences are independent, so any slice of a big call equals the small call bit
    for bit; the small call is held to the fp32 oracle."""
    import deepards_b200 as D
    sd = O.cnn_linear_state(backbone, seed=32, bn_perturb=0.1, per_breath=True)
    bb = D.resnet18() if backbone == "resnet18" else D.densenet18(drop_rate=0.0)
    net = D.CNNSingleBreathLinearNetwork(bb)
    net.load_state_dict(sd)
    # the reference never calls eval() on this path (train_ards_detector.py:448): BatchNorm uses batch statistics at test
    # time; DenseNet's eval() only switches dropout off
    net = net.cuda().eval() if backbone == "densenet18" else net.cuda().train()
    n_seq = 16384 // 20 + 1
    g = torch.Generator().manual_seed(5)
    x = O.synthetic_breaths(n_seq, seed=300)
    lens = torch.randint(60, 200, (n_seq, 20), generator=g)
    x = x * (torch.arange(224).view(1, 1, 1, 224) < lens.view(n_seq, 20, 1, 1))  # zero padding after the breath
    for precision in ("fp32", "bf16"):
        net.precision = precision
        with torch.no_grad():
            big = net(x.cuda(), None)
            small = net(x[100:104].cuda(), None)
        assert big.shape == (n_seq, 20, 2)
        assert torch.equal(big[100:104], small)
        if precision == "fp32":
            ref = O.cnn_linear_forward(sd, x[100:104], per_breath=True)
            assert rel_err(small.cpu(), ref) <= 1e-4


def test_changing_the_learning_rate_after_graph_capture_takes_effect():
    """lr is an argument of the fused optimizer kernel, i.e. a constant of the captured CUDA graph; changing
    `trainer.lr` (decay, warm-up) must re-capture instead of being silently ignored.  Graph replay == eager, bit for bit."""
    from deepards_b200.data_parallel import DataParallelTrainer
    sd0 = O.cnn_linear_state("densenet18", seed=33, bn_perturb=0.1)
    x, t = O.synthetic_breaths(4, seed=210).cuda(), O.synthetic_targets(4, seed=210).cuda()
    out = {}
    for use_graph in (False, True):
        # fp32 plans are bit-reproducible (the bf16 path's accumulate-mode weight gradients are not)
        net = _net("densenet18", {k: v.clone() for k, v in sd0.items()}, "fp32")
        tr = DataParallelTrainer(net, lr=1e-3, optimizer="sgd", weight_decay=1e-4, clip_val=0.01, use_graph=use_graph)
        for i in range(8):
            if i == 5:
                tr.lr = 1e-2        # after the graph has been captured (third call)
            tr.train_step(x, t)
        torch.cuda.synchronize()
        out[use_graph] = tr.param_flat.clone()
    assert torch.equal(out[False], out[True])
