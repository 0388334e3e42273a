"""CPU: the oracle restatement agrees with the golden vectors recorded from the reference's own
modules (oracle/make_golden.py), and the three structural invariants of SURVEY.md hold."""
import numpy as np
import pytest
import torch

from oracle import cnn_linear_oracle as O
from tests.helpers import CASES, check_grads_against_golden, load_case, rel_err

TOL = 2e-5  # same library, same machine class: only summation-order noise


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name):
    z, sd, fkw, per_breath = load_case(name)
    wsum = np.array([float(v.double().abs().sum()) for v in sd.values() if v.is_floating_point()])
    np.testing.assert_allclose(wsum, z["wsum"], rtol=1e-12)  # seeded weights did not drift
    x = torch.from_numpy(z["x"])
    t = torch.from_numpy(z["target"])
    fkw = dict(fkw)
    if "resnet" in name:
        fkw["running_update"] = True
    out, loss, grads = O.forward_backward(sd, x, t, per_breath=per_breath, **fkw)
    assert rel_err(out, z["logits"]) <= TOL
    assert abs(float(loss) - float(z["loss"])) <= TOL
    check_grads_against_golden(z, grads, TOL, name)
    for key in z.files:
        if key.startswith("buf/"):
            got = sd[key[4:]]
            if key.endswith("num_batches_tracked"):
                assert int(got) == int(z[key])
            else:
                assert rel_err(got, z[key]) <= TOL, key


def test_gradcam_oracle_matches_reference():
    z, sd, _, _ = load_case("densenet18_B2_real")
    a, da, out = O.gradcam_forward_backward(sd, torch.from_numpy(z["x"][0]), int(z["cam/target"]))
    assert rel_err(a, z["cam/A"]) <= TOL
    assert rel_err(da, z["cam/dA"]) <= TOL
    assert rel_err(out, z["cam/out"]) <= TOL


def test_bn_is_per_sequence_not_per_batch():
    """SURVEY.md point 1: feeding all B*20 breaths through the backbone at once changes the result."""
    sd = O.cnn_linear_state("resnet18", seed=11, initial_planes=16)
    x = O.synthetic_breaths(3, seed=3)
    loop = torch.cat([O.backbone_forward(sd, x[i]) for i in range(3)])
    flat = O.backbone_forward(sd, x.reshape(60, 1, 224))
    assert (loop - flat).abs().max() > 1e-3


def test_batch_sharding_is_exact():
    """SURVEY.md point 3: model(x[:2]) == model(x)[:2] -- data parallel over B needs no SyncBN."""
    sd = O.cnn_linear_state("densenet18", seed=12)
    x = O.synthetic_breaths(4, seed=4)
    full = O.cnn_linear_forward(sd, x)
    assert torch.equal(O.cnn_linear_forward(sd, x[:2]), full[:2])
    assert torch.equal(O.cnn_linear_forward(sd, x[2:]), full[2:])


def test_running_stats_closed_form():
    """SURVEY.md hard part 6: rm = 0.9^B rm0 + 0.1 sum_i 0.9^(B-1-i) mean_i, unbiased variance."""
    sd = O.cnn_linear_state("resnet18", seed=13, initial_planes=16)
    x = O.synthetic_breaths(3, seed=5)
    sd2 = {k: v.clone() for k, v in sd.items()}
    O.cnn_linear_forward(sd2, x, running_update=True)
    w = sd["breath_block.conv1.weight"]
    means, uvars = [], []
    for i in range(3):
        y = torch.nn.functional.conv1d(x[i], w, stride=2, padding=3)
        means.append(y.mean(dim=(0, 2)))
        uvars.append(y.var(dim=(0, 2), unbiased=True))
    rm = torch.zeros(16)
    rv = torch.ones(16)
    for m, v in zip(means, uvars):
        rm = 0.9 * rm + 0.1 * m
        rv = 0.9 * rv + 0.1 * v
    assert rel_err(sd2["breath_block.bn1.running_mean"], rm) < 1e-5
    assert rel_err(sd2["breath_block.bn1.running_var"], rv) < 1e-5
    assert int(sd2["breath_block.bn1.num_batches_tracked"]) == 3
    assert int(sd2["breath_block.bn2.num_batches_tracked"]) == 0  # never-used stem bn2


def test_grouped_bn_helper_equals_loop():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(60, 8, 14, generator=g)
    w = torch.randn(8, generator=g)
    b = torch.randn(8, generator=g)
    ref = torch.cat([torch.nn.functional.batch_norm(x[i * 20:(i + 1) * 20], None, None, w, b, True, 0.0, 1e-5)
                     for i in range(3)])
    assert rel_err(O.grouped_batchnorm(x, w, b, 20), ref) < 1e-6


def test_sequence_length_guard():
    sd = O.cnn_linear_state("resnet18", seed=1, initial_planes=16)
    with pytest.raises(Exception, match="sequence length of 224"):
        O.cnn_linear_forward(sd, torch.zeros(1, 20, 1, 200))
