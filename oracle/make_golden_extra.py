"""Generate tests/golden/{gradcam_densenet18,scaling_real,sibling_heads}.npz from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Runs only in the build container, where /root/reference is mounted:

    python oracle/make_golden_extra.py

GradCAM: imports deepards/gradcam.py itself.  Its module-level imports that are absent from this image (cv2,
matplotlib, ventmap, imblearn, algorithms, mock, prettytable -- none of them touched by the classes used here) are
satisfied by empty stub modules, and `Tensor.cuda()` (gradcam.py:106, no GPU in this container) is the identity.
The reference's `MaxMinNormCam.generate_read_cam`, `.generate_cam` and `UnNormalizedCam.generate_cam`
(gradcam.py:125-154, 195-205) then run on the reference's own DenseNet-18 `CNNLinearNetwork` with the oracle's seeded
state_dict, on two REAL sequences of deepards/tests/test_dataset.pkl and one synthetic sequence.

Sibling heads: the reference's CNNLinearToMean / CNNLinearComprToRF / CNNDoubleLinearNetwork
(torch_cnn_linear_network.py:7-46, 70-89) and CNNRegressor (torch_cnn_bm_regressor.py) over its ResNet-18 (16 planes):
logits, loss and a sample of the parameter gradients of one BCE / MSE step.

Scaling: `(data - mu) / std` and the padded-breath rule through the reference's own `_get_padding_mask`
(dataset.py:1375-1379, 1406-1409) on raw float64 windows of the same pickle.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

STUBS = ("cv2", "matplotlib", "ventmap", "imblearn", "algorithms", "mock", "prettytable")


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {})


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in STUBS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        return _Stub(spec.name)

    def exec_module(self, m):
        pass


sys.meta_path.insert(0, _Finder())
torch.Tensor.cuda = lambda self, *a, **k: self  # no GPU here; gradcam.py:106 calls one_hot.cuda()

from oracle import cnn_linear_oracle as O  # noqa: E402
from oracle.make_golden import real_sequences  # noqa: E402  (also imports the reference model modules)
from deepards import gradcam as G  # noqa: E402
from deepards import dataset as DS  # noqa: E402
from deepards.models.densenet import densenet18 as ref_densenet18  # noqa: E402
from deepards.models.torch_cnn_linear_network import CNNLinearNetwork as RefCNNLinearNetwork  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    xr, _, mu, std = real_sequences(2)
    xs = O.synthetic_breaths(1, seed=321).numpy()
    x = np.concatenate([xr, xs]).astype(np.float32)          # (3, 20, 1, 224)
    # a sequence of ONE breath repeated 20 times: what get_camout_for_breath feeds generate_cam (patient_gradcam.py:213-218)
    rep = np.repeat(x[2:3, 7:8], 20, axis=1)
    x = np.concatenate([x, rep])                             # (4, 20, 1, 224)

    sd = O.cnn_linear_state("densenet18", seed=4, bn_perturb=0.1)
    dn = ref_densenet18()
    for m in dn.modules():
        if hasattr(m, "drop_rate"):
            m.drop_rate = 0.0                                # the reference leaves dropout on (no eval(), gradcam.py:76-78)
    model = RefCNNLinearNetwork(dn, 20, 0)
    model.load_state_dict(sd, strict=True)

    rec = {"x": x, "wsum": np.array([float(v.double().abs().sum()) for v in sd.values() if v.is_floating_point()])}
    for i in range(x.shape[0]):
        xi = torch.from_numpy(x[i])
        for tname, target in (("none", None), ("t0", 0), ("t1", 1)):
            cam = G.MaxMinNormCam(model)
            read, mo = cam.generate_read_cam(xi, target)
            rec["read/%d/%s" % (i, tname)] = read.astype(np.uint8)
            rec["out/%d" % i] = mo.detach().numpy()
            seq, _ = G.MaxMinNormCam(model).generate_cam(xi, target)
            rec["seq/%d/%s" % (i, tname)] = seq.astype(np.uint8)
            un, _ = G.UnNormalizedCam(model).generate_cam(xi, target)
            rec["unnorm/%d/%s" % (i, tname)] = un.astype(np.float32)
            a, da, _ = G.GradCam(model).generate_one_hot_grad_and_output(xi, target)
            rec["A/%d" % i] = a
            rec["dA/%d/%s" % (i, tname)] = da
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "gradcam_densenet18.npz")
    np.savez_compressed(path, **rec)
    print("%s  %.1f KB  out[0]=%s" % (path, os.path.getsize(path) / 1024, rec["out/0"].ravel()))

    # ---- input scaling --------------------------------------------------------------------------------------
    import pickle
    from oracle.make_golden import real_sequences as _rs  # noqa: F401

    class _Any(object):
        def __init__(self, *a, **k):
            pass

        def __setstate__(self, st):
            if isinstance(st, dict):
                self.__dict__.update(st)

    class U(pickle.Unpickler):
        def find_class(self, module, name):
            try:
                return super().find_class(module, name)
            except Exception:
                return type(name, (_Any,), {})

    with open("/root/reference/deepards/tests/test_dataset.pkl", "rb") as f:
        d = U(f, encoding="latin1").load()
    mu, std = d.scaling_factors[None]
    raw = np.stack([d.all_sequences[i][1] for i in range(4)]).astype(np.float64)      # (4, 20, 1, 224)
    scaled = torch.from_numpy((raw - mu) / std).float().numpy()                       # dataset.py:1379 + .float()
    # padded-breath rule on zero-padded copies (dataset.py:1233-1237), through the reference's own mask function
    padded_raw = raw.copy()
    lens = np.random.RandomState(3).randint(60, 201, size=(4, 20))
    for i in range(4):
        for j in range(20):
            padded_raw[i, j, 0, lens[i, j]:] = 0.0
    mask = DS.ARDSRawDataset._get_padding_mask(None, padded_raw, mu)                  # dataset.py:1406-1409
    padded_scaled = torch.from_numpy((padded_raw - mask) / std).float().numpy()       # dataset.py:1375-1377
    path = os.path.join(OUT, "scaling_real.npz")
    np.savez_compressed(path, raw=raw, scaled=scaled, padded_raw=padded_raw, padded_scaled=padded_scaled,
                        mu=np.float64(mu), std=np.float64(std))
    print("%s  %.1f KB  mu=%r std=%r" % (path, os.path.getsize(path) / 1024, mu, std))


def sibling_heads():
    from deepards.models.resnet import resnet18 as ref_resnet18
    from deepards.models import torch_cnn_linear_network as RN
    from deepards.models.torch_cnn_bm_regressor import CNNRegressor as RefCNNRegressor
    from deepards.models.torch_cnn_lstm_combo import CNNLSTMNetwork as RefCNNLSTMNetwork
    rec = {}
    x = O.synthetic_breaths(2, seed=55)
    t = O.synthetic_targets(2, seed=55)
    base = O.cnn_linear_state("resnet18", seed=8, bn_perturb=0.1, initial_planes=16, per_breath=True)
    gen = torch.Generator().manual_seed(99)
    for kind, make in (("to_mean", lambda bb: RN.CNNLinearToMean(bb)),
                       ("compr_to_rf", lambda bb: RN.CNNLinearComprToRF(bb)),
                       ("double_linear", lambda bb: RN.CNNDoubleLinearNetwork(bb, 20, 0)),
                       ("regressor", lambda bb: RefCNNRegressor(bb, 3)),
                       ("lstm", lambda bb: RefCNNLSTMNetwork(bb, 0, False, 32))):
        model = make(ref_resnet18(initial_planes=16))
        sd = dict(base)
        if kind == "double_linear":
            sd["linear_intermediate.weight"] = base["linear_final.weight"]
            sd["linear_intermediate.bias"] = base["linear_final.bias"]
            sd["linear_final.weight"] = torch.randn(2, 40, generator=gen) * 0.1
            sd["linear_final.bias"] = torch.randn(2, generator=gen) * 0.1
        if kind == "regressor":
            sd["linear_final.weight"] = torch.randn(3, 128, generator=gen) * 0.1
            sd["linear_final.bias"] = torch.randn(3, generator=gen) * 0.1
        if kind == "lstm":
            sd["linear_final.weight"] = torch.randn(2, 32, generator=gen) * 0.1
            sd["linear_final.bias"] = torch.randn(2, generator=gen) * 0.1
            for k, shape in (("lstm.weight_ih_l0", (128, 128)), ("lstm.weight_hh_l0", (128, 32)),
                             ("lstm.bias_ih_l0", (128,)), ("lstm.bias_hh_l0", (128,))):
                sd[k] = torch.randn(*shape, generator=gen) * 0.1
        model.load_state_dict(sd, strict=True)
        model.train()
        model.zero_grad()
        if kind == "lstm":
            # train_ards_detector.py feeds NaN metadata and no initial state for cnn_lstm
            out, _ = model(x, torch.tensor(float("nan")), None)
            loss = torch.nn.BCEWithLogitsLoss()(out, t.unsqueeze(1).repeat(1, 20, 1))
        elif kind == "regressor":
            xin = x.reshape(40, 1, 224)
            tgt = torch.randn(40, 3, generator=gen)
            out = model(xin, None)
            loss = torch.nn.MSELoss()(out, tgt)
            rec[kind + "/target"] = tgt.numpy()
        else:
            out = model(x, None)
            loss = torch.nn.BCEWithLogitsLoss()(out, t)
        loss.backward()
        rec[kind + "/logits"] = out.detach().numpy()
        rec[kind + "/loss"] = np.array(loss.item(), np.float64)
        rec[kind + "/keys"] = np.array(list(model.state_dict().keys()))
        for name, p in model.named_parameters():
            if p.grad is None:
                continue
            g = p.grad.detach().reshape(-1).numpy()
            if name.startswith(("linear", "lstm")) or name in ("breath_block.conv1.weight", "breath_block.bn1.weight"):
                rec[kind + "/grad/" + name] = g.reshape(p.shape).copy()
            elif name.endswith("conv2.weight"):
                rec[kind + "/gradsample/" + name] = g[::61].copy()
                rec[kind + "/gradmax/" + name] = np.array(np.abs(g).max())
        for k, v in sd.items():
            if k.startswith(("linear", "lstm")):
                rec[kind + "/sd/" + k] = v.numpy()
    rec["x"] = x.numpy()
    rec["target"] = t.numpy()
    path = os.path.join(OUT, "sibling_heads.npz")
    np.savez_compressed(path, **rec)
    print("%s  %.1f KB  to_mean logits[0]=%s" % (path, os.path.getsize(path) / 1024, rec["to_mean/logits"][0]))


if __name__ == "__main__":
    main()
    sibling_heads()
