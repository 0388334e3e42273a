"""CPU: the oracle's GradCAM maps, window scaling and sibling heads against vectors recorded from the UNMODIFIED
reference (oracle/make_golden_extra.py: deepards/gradcam.py, dataset.py and models/torch_cnn_linear_network.py run
in the build container)."""
import os

import numpy as np
import pytest
import torch

from oracle import cnn_linear_oracle as O
from tests.helpers import GOLDEN, rel_err


def _z(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def test_oracle_gradcam_equals_reference_gradcam_py():
    z = _z("gradcam_densenet18")
    sd = O.cnn_linear_state("densenet18", seed=4, bn_perturb=0.1)
    for i in range(z["x"].shape[0]):
        xi = torch.from_numpy(z["x"][i])
        for tn, t in (("none", None), ("t0", 0), ("t1", 1)):
            a, da, out = O.gradcam_forward_backward(sd, xi, t)
            assert np.array_equal(a.numpy(), z["A/%d" % i]) and np.array_equal(da.numpy(), z["dA/%d/%s" % (i, tn)])
            assert np.array_equal(out.numpy(), z["out/%d" % i])
            read, _, _ = O.gradcam_read_cam(sd, xi, t)
            assert np.array_equal(read, z["read/%d/%s" % (i, tn)])          # uint8, bit-exact
            seq, raw, _ = O.gradcam_seq_cam(sd, xi, t)
            assert np.array_equal(seq, z["seq/%d/%s" % (i, tn)])
            assert np.array_equal(np.maximum(0, raw), z["unnorm/%d/%s" % (i, tn)])


def test_gradient_of_the_norm5_output_has_a_closed_form():
    """What dards_gradcam evaluates instead of a backward pass: dA = (A > 0) * W[target] / 7, bit-exact."""
    z = _z("gradcam_densenet18")
    sd = O.cnn_linear_state("densenet18", seed=4, bn_perturb=0.1)
    w = sd["linear_final.weight"].numpy()
    for i in range(z["x"].shape[0]):
        a = z["A/%d" % i]
        for tn, t in (("t0", 0), ("t1", 1)):
            pred = np.where(a > 0, (w[t].reshape(20, 128, 1) / np.float32(7)).astype(np.float32), np.float32(0))
            assert np.array_equal(pred, z["dA/%d/%s" % (i, tn)])


def test_oracle_window_scaling_is_bit_exact():
    z = _z("scaling_real")
    mu, std = float(z["mu"]), float(z["std"])
    assert np.array_equal(O.scale_windows(z["raw"], mu, std).numpy(), z["scaled"])
    got = O.scale_windows(z["padded_raw"], mu, std, padded=True).numpy()
    assert np.array_equal(got, z["padded_scaled"])
    assert np.all(got[z["padded_raw"] == 0] == 0)      # the zero padding stays zero (dataset.py:1375-1377)


def test_cam_resize_known_properties():
    """cv2's 8-bit INTER_LINEAR restated (parity unpinned: OpenCV is absent here); properties any linear resize with
    half-pixel centres has: constants stay constant, ends are clamped, the map is monotone for monotone input."""
    assert np.all(O.cam_resize_linear_u8(np.full(7, 93, np.uint8)) == 93)
    ramp = np.array([0, 40, 80, 120, 160, 200, 255], np.uint8)
    r = O.cam_resize_linear_u8(ramp)
    assert r.shape == (224,) and r[0] == 0 and r[-1] == 255 and np.all(np.diff(r.astype(int)) >= 0)
    assert np.all(r[:16] == 0) and np.all(r[-16:] == 255)          # first / last half source pixel are clamped
    assert abs(int(r[16 + 32 * 3]) - 120) <= 1                      # source pixel centres map to themselves


@pytest.mark.parametrize("kind", ["to_mean", "compr_to_rf", "double_linear", "regressor", "lstm"])
def test_oracle_sibling_heads_equal_reference(kind):
    z = _z("sibling_heads")
    sd = O.cnn_linear_state("resnet18", seed=8, bn_perturb=0.1, initial_planes=16, per_breath=True)
    if kind == "double_linear":
        sd["linear_intermediate.weight"], sd["linear_intermediate.bias"] = sd["linear_final.weight"], sd["linear_final.bias"]
    for k in z.files:
        if k.startswith(kind + "/sd/"):
            sd[k[len(kind) + 4:]] = torch.from_numpy(z[k])
    x = torch.from_numpy(z["x"])
    if kind == "regressor":
        out = O.regressor_forward(sd, x.reshape(40, 1, 224))
    else:
        out = O.cnn_linear_forward(sd, x, head=kind)
    assert rel_err(out, z[kind + "/logits"]) < 1e-6


def test_patient_vote_oracle_matches_the_references_own_loop():
    """tests/golden/patient_votes.npz was produced by DeepARDSResults.perform_patient_predictions itself
    (oracle/make_golden_heads2.py): arbitrary patient ids in interleaved order, a unanimous patient, a zero-vote patient and
    an exact tie (np.argmax -> class 0)."""
    import numpy as np
    z = _z("patient_votes")
    got = O.patient_vote_table(z["patient"], z["y"], z["pred"])
    assert got.shape == z["table"].shape == (9, 14)
    assert np.array_equal(got, z["table"])
    tie = got[got[:, 6] == got[:, 11]]            # OTHER_votes == ARDS_votes
    assert len(tie) >= 1 and (tie[:, 12] == 0.0).all() and (tie[:, 13] == 0.5).all()
