"""CPU oracle for the deepards cnn_linear hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``deepards_b200/`` may import this
module; it is used by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker and
as the timed CPU baseline -- never as the product path.

What it is: a functional, fp32, torch-CPU restatement of the reference's module
graph for this path.  The reference has no arithmetic of its own -- every
operation is a ``torch.nn`` library call -- so the
restatement is "the same library calls in the same order on an explicit
``state_dict``", written without any ``nn.Module`` so that nothing of the
reference's class structure is needed at run time (``/root/reference`` does not
exist on the GPU box).

Pinned against: the reference's own modules imported from ``/root/reference``
(``oracle/make_golden.py`` generates ``tests/golden/*.npz`` from them; the CPU
test-suite checks this oracle against those files).  The reference's test-suite
holds no golden vector for this path, so these generated
vectors are the pin.

Reference call sites restated here
  * ResNet stem / stages / pooling ........ deepards/models/resnet.py:141-163
  * BasicBlock ............................ deepards/models/resnet.py:24-40
  * downsample (conv1x1 stride s + BN) .... deepards/models/resnet.py:125-131
  * DenseNet features / head pooling ...... deepards/models/densenet.py:117-150, 179-193
  * _DenseLayer ........................... deepards/models/densenet.py:18-43
  * _Transition ........................... deepards/models/densenet.py:68-80
  * CNNLinearNetwork.forward .............. deepards/models/torch_cnn_linear_network.py:104-113
  * CNNSingleBreathLinearNetwork.forward .. deepards/models/torch_cnn_linear_network.py:57-67
  * BCEWithLogits loss .................... deepards/train_ards_detector.py:530, 929-930
  * gradient clamp hook ................... deepards/train_ards_detector.py:474-476
  * GradCAM forward/backward .............. deepards/gradcam.py:40-65, 83-107
  * GradCAM maps (read / sequence) ........ deepards/gradcam.py:125-162, 195-205
  * window scaling (input contract) ....... deepards/dataset.py:1375-1379, 1406-1409
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
SEQ_LEN = 224


# --------------------------------------------------------------------------
# parameter construction (names are the reference's state_dict keys)
# --------------------------------------------------------------------------
def _conv_w(gen, cout, cin, k):
    # He-normal with fan = k * cout (resnet.py:115-118, densenet.py:156-159)
    std = math.sqrt(2.0 / (k * cout))
    return torch.randn(cout, cin, k, generator=gen, dtype=torch.float32) * std


def _bn(sd, name, c, running, gen=None, perturb=0.0):
    w = torch.ones(c)
    b = torch.zeros(c)
    if perturb and gen is not None:
        # parity tests perturb gamma/beta so that a wrong affine is visible
        w = w + perturb * torch.randn(c, generator=gen)
        b = b + perturb * torch.randn(c, generator=gen)
    sd[name + ".weight"] = w
    sd[name + ".bias"] = b
    if running:
        sd[name + ".running_mean"] = torch.zeros(c)
        sd[name + ".running_var"] = torch.ones(c)
        sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def _linear(sd, name, fin, fout, gen):
    bound = 1.0 / math.sqrt(fin)
    sd[name + ".weight"] = (torch.rand(fout, fin, generator=gen) * 2 - 1) * bound
    sd[name + ".bias"] = (torch.rand(fout, generator=gen) * 2 - 1) * bound


def resnet_state(seed: int = 0, layers: Sequence[int] = (2, 2, 2, 2), initial_planes: int = 64,
                 bn_perturb: float = 0.0, prefix: str = "") -> "OrderedDict[str, torch.Tensor]":
    """state_dict of the reference's 1-D BasicBlock ResNet (resnet.py:83-139), incl.
    the never-used conv1_alt/conv2/bn2 (resnet.py:90-96).  Deterministic in `seed`;
    the values are this oracle's own draw with the reference's init distribution."""
    gen = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    p = initial_planes
    sd["conv1.weight"] = _conv_w(gen, p, 1, 7)
    sd["conv1_alt.weight"] = _conv_w(gen, p, 1, 3)
    _bn(sd, "bn1", p, True, gen, bn_perturb)
    sd["conv2.weight"] = _conv_w(gen, p, p, 7)
    _bn(sd, "bn2", p, True, gen, bn_perturb)
    inpl = p
    for li, nb in enumerate(layers, 1):
        planes = p * 2 ** (li - 1)
        for bi in range(nb):
            stride = 2 if (li > 1 and bi == 0) else 1
            pre = "layer%d.%d." % (li, bi)
            sd[pre + "conv1.weight"] = _conv_w(gen, planes, inpl, 3)
            _bn(sd, pre + "bn1", planes, True, gen, bn_perturb)
            sd[pre + "conv2.weight"] = _conv_w(gen, planes, planes, 3)
            _bn(sd, pre + "bn2", planes, True, gen, bn_perturb)
            if stride != 1 or inpl != planes:
                sd[pre + "downsample.0.weight"] = _conv_w(gen, planes, inpl, 1)
                _bn(sd, pre + "downsample.1", planes, True, gen, bn_perturb)
            inpl = planes
    if prefix:
        sd = OrderedDict((prefix + k, v) for k, v in sd.items())
    return sd


def densenet_state(seed: int = 0, block_config: Sequence[int] = (2, 2, 2, 2), growth_rate: int = 32,
                   num_init_features: int = 64, bn_size: int = 4, in_chans: int = 1,
                   bn_perturb: float = 0.0, prefix: str = "") -> "OrderedDict[str, torch.Tensor]":
    """state_dict of the reference's 1-D DenseNet (densenet.py:96-167): no BN buffers
    (track_running_stats=False, densenet.py:107)."""
    gen = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["features.conv0.weight"] = _conv_w(gen, num_init_features, in_chans, 7)
    _bn(sd, "features.norm0", num_init_features, False, gen, bn_perturb)
    nf = num_init_features
    for i, nl in enumerate(block_config, 1):
        for j in range(1, nl + 1):
            pre = "features.denseblock%d.denselayer%d." % (i, j)
            cin = nf + (j - 1) * growth_rate
            _bn(sd, pre + "norm1", cin, False, gen, bn_perturb)
            sd[pre + "conv1.weight"] = _conv_w(gen, bn_size * growth_rate, cin, 1)
            _bn(sd, pre + "norm2", bn_size * growth_rate, False, gen, bn_perturb)
            sd[pre + "conv2.weight"] = _conv_w(gen, growth_rate, bn_size * growth_rate, 3)
        nf = nf + nl * growth_rate
        if i != len(block_config):
            pre = "features.transition%d." % i
            _bn(sd, pre + "norm", nf, False, gen, bn_perturb)
            sd[pre + "conv.weight"] = _conv_w(gen, nf // 2, nf, 1)
            nf = nf // 2
    _bn(sd, "features.norm5", nf, False, gen, bn_perturb)
    if prefix:
        sd = OrderedDict((prefix + k, v) for k, v in sd.items())
    return sd


def backbone_out_filters(sd: Dict[str, torch.Tensor], prefix: str = "breath_block.") -> int:
    if prefix + "features.norm5.weight" in sd:
        return sd[prefix + "features.norm5.weight"].numel()
    last = [k for k in sd if k.startswith(prefix + "layer4.") and k.endswith("bn2.weight")]
    return sd[sorted(last)[-1]].numel()


def cnn_linear_state(backbone: str = "resnet18", seed: int = 0, sub_batch: int = 20, per_breath: bool = False,
                     bn_perturb: float = 0.0, **kw) -> "OrderedDict[str, torch.Tensor]":
    """state_dict of CNNLinearNetwork(backbone(), sub_batch, 0) (torch_cnn_linear_network.py:97-102)
    or, with per_breath, CNNSingleBreathLinearNetwork (torch_cnn_linear_network.py:50-55)."""
    if backbone.startswith("resnet"):
        sd = resnet_state(seed, bn_perturb=bn_perturb, prefix="breath_block.", **kw)
    elif backbone.startswith("densenet"):
        sd = densenet_state(seed, bn_perturb=bn_perturb, prefix="breath_block.", **kw)
    else:
        raise ValueError(backbone)
    f = backbone_out_filters(sd)
    gen = torch.Generator().manual_seed(seed + 7919)
    _linear(sd, "linear_final", f if per_breath else f * sub_batch, 2, gen)
    return sd


# --------------------------------------------------------------------------
# forward graphs
# --------------------------------------------------------------------------
def _batchnorm(x, sd, name, running_update: bool):
    """nn.BatchNorm1d in training mode on ONE sub-batch: biased batch variance for the
    normalisation, eps 1e-5; running buffers (if the layer has them and running_update
    is set) get momentum 0.1 with the unbiased variance -- exactly F.batch_norm."""
    rm = sd.get(name + ".running_mean") if running_update else None
    rv = sd.get(name + ".running_var") if running_update else None
    y = F.batch_norm(x, rm, rv, sd[name + ".weight"], sd[name + ".bias"], True, BN_MOMENTUM, BN_EPS)
    if rm is not None and (name + ".num_batches_tracked") in sd:
        sd[name + ".num_batches_tracked"].add_(1)
    return y


class DecisionHooks(object):
    """Test instrumentation for the ReLU decisions (see tests/helpers.py, "pinned decisions").

    record: {site: {sequence index: pre-activation tensor}} filled during the forward when not None.
    masks : {site: bool tensor (B, N, C, L)}; where given, relu(z) is replaced by z * mask, i.e. the decision
            "is this unit active" is taken from outside (from the implementation under test) instead of from
            the sign of this run's own rounding noise.  Sites: "relu0" (stem, before the pool), "layerI.J.relu1",
            "layerI.J.relu2", "denseblockI.denselayerJ.relu1|relu2", "transitionI.relu", "relu5"."""

    def __init__(self, record=None, masks=None):
        self.record, self.masks, self.seq = record, masks or {}, 0


def _act(z, site, hooks):
    if hooks is not None:
        if hooks.record is not None:
            hooks.record.setdefault(site, {})[hooks.seq] = z.detach()
        m = hooks.masks.get(site)
        if m is not None:
            return z * m[hooks.seq].to(z.dtype)
    return F.relu(z)


def resnet_forward(sd, x, prefix: str = "", first_pool_type: str = "max", double_conv_first: bool = False,
                   running_update: bool = False, hooks=None):
    """(N, 1, 224) -> (N, 8*initial_planes); BN statistics over the whole N (resnet.py:141-163)."""
    g = lambda k: sd[prefix + k]
    bn = lambda t, name: _batchnorm(t, _Prefixed(sd, prefix), name, running_update)
    if not double_conv_first:
        y = F.conv1d(x, g("conv1.weight"), stride=2, padding=3)
        y = bn(y, "bn1")
    else:
        y = F.conv1d(x, g("conv1_alt.weight"), stride=1, padding=1)
        y = bn(y, "bn1")
        y = F.conv1d(y, g("conv2.weight"), stride=2, padding=3)
        y = bn(y, "bn2")
    y = _act(y, "relu0", hooks)
    if first_pool_type == "max":
        y = F.max_pool1d(y, 3, 2, 1)
    else:
        y = F.avg_pool1d(y, 3, 2, 1)
    li = 1
    while (prefix + "layer%d.0.conv1.weight" % li) in sd:
        bi = 0
        while (prefix + "layer%d.%d.conv1.weight" % (li, bi)) in sd:
            pre = "layer%d.%d." % (li, bi)
            stride = 2 if (li > 1 and bi == 0) else 1
            out = F.conv1d(y, g(pre + "conv1.weight"), stride=stride, padding=1)
            out = _act(bn(out, pre + "bn1"), pre + "relu1", hooks)
            out = F.conv1d(out, g(pre + "conv2.weight"), stride=1, padding=1)
            out = bn(out, pre + "bn2")
            if (prefix + pre + "downsample.0.weight") in sd:
                res = F.conv1d(y, g(pre + "downsample.0.weight"), stride=stride)
                res = bn(res, pre + "downsample.1")
            else:
                res = y
            y = _act(out + res, pre + "relu2", hooks)
            bi += 1
        li += 1
    y = F.avg_pool1d(y, 7, 1)
    return y.reshape(y.shape[0], -1)


class _Prefixed(dict):
    """dict view that prepends a prefix on lookup (writes go through to the base tensors)."""

    def __init__(self, base, prefix):
        super().__init__()
        self._b, self._p = base, prefix

    def __getitem__(self, k):
        return self._b[self._p + k]

    def __contains__(self, k):
        return (self._p + k) in self._b

    def get(self, k, default=None):
        return self._b.get(self._p + k, default)


def densenet_features(sd, x, prefix: str = "", drop_rate: float = 0.0, training: bool = False, hooks=None):
    """(N, C0, 224) -> (N, 128, 7): `DenseNet.features` (densenet.py:117-150).  BN always uses
    batch statistics (track_running_stats=False).  Dropout only if drop_rate>0 and training."""
    g = lambda k: sd[prefix + k]
    bn = lambda t, name: F.batch_norm(t, None, None, g(name + ".weight"), g(name + ".bias"), True, 0.0, BN_EPS)
    y = F.conv1d(x, g("features.conv0.weight"), stride=2, padding=3)
    y = _act(bn(y, "features.norm0"), "relu0", hooks)
    y = F.max_pool1d(y, 3, 2, 1)
    i = 1
    while (prefix + "features.denseblock%d.denselayer1.conv1.weight" % i) in sd:
        j = 1
        while (prefix + "features.denseblock%d.denselayer%d.conv1.weight" % (i, j)) in sd:
            pre = "features.denseblock%d.denselayer%d." % (i, j)
            t = _act(bn(y, pre + "norm1"), pre[9:] + "relu1", hooks)
            t = F.conv1d(t, g(pre + "conv1.weight"))
            t = _act(bn(t, pre + "norm2"), pre[9:] + "relu2", hooks)
            t = F.conv1d(t, g(pre + "conv2.weight"), padding=1)
            if drop_rate > 0:
                t = F.dropout(t, p=drop_rate, training=training)
            y = torch.cat([y, t], 1)
            j += 1
        pre = "features.transition%d." % i
        if (prefix + pre + "conv.weight") in sd:
            t = _act(bn(y, pre + "norm"), pre[9:] + "relu", hooks)
            t = F.conv1d(t, g(pre + "conv.weight"))
            y = F.avg_pool1d(t, 2, 2)
        i += 1
    return bn(y, "features.norm5")


def densenet_forward(sd, x, prefix: str = "", **kw):
    """densenet.py:179-189: relu(features) -> AvgPool1d(7) -> flatten."""
    f = densenet_features(sd, x, prefix, **kw)
    y = F.avg_pool1d(_act(f, "relu5", kw.get("hooks")), 7, 1)
    return y.reshape(f.shape[0], -1)


def backbone_forward(sd, x, prefix: str = "breath_block.", **kw):
    if (prefix + "features.conv0.weight") in sd:
        kw.pop("running_update", None)
        return densenet_forward(sd, x, prefix, **kw)
    return resnet_forward(sd, x, prefix, **kw)


def cnn_linear_forward(sd, x, per_breath: bool = False, head: Optional[str] = None, **kw):
    """The reference's per-sequence loop (torch_cnn_linear_network.py:104-113): each
    x[i] of shape (20, C, 224) is one BatchNorm sub-batch.  Returns (B, 2), or
    (B, 20, 2) for the per-breath head (torch_cnn_linear_network.py:57-67).
    head: None / 'cnn_linear' / 'per_breath', or a sibling head on the same loop --
    'to_mean' (:7-25), 'compr_to_rf' (:28-46), 'double_linear' (:70-89), 'lstm' (torch_cnn_lstm_combo.py:28-50,
    zero initial state, NaN metadata; returns the (B, S, 2) per-step outputs)."""
    if x.shape[-1] != SEQ_LEN:
        raise Exception("input breaths must have sequence length of 224")
    head = head or ("per_breath" if per_breath else "cnn_linear")
    w, b = sd["linear_final.weight"], sd["linear_final.bias"]
    rows = []
    for i in range(x.shape[0]):
        if kw.get("hooks") is not None:
            kw["hooks"].seq = i
        feat = backbone_forward(sd, x[i], **kw)
        if head == "per_breath":
            rows.append(F.linear(feat, w, b).unsqueeze(0))
        elif head == "cnn_linear":
            rows.append(F.linear(feat.reshape(-1), w, b).unsqueeze(0))
        elif head == "double_linear":
            mid = F.linear(feat, sd["linear_intermediate.weight"], sd["linear_intermediate.bias"])
            rows.append(F.linear(mid.reshape(-1).unsqueeze(0), w, b))
        elif head in ("to_mean", "compr_to_rf", "lstm"):
            rows.append(feat.unsqueeze(0))
        else:
            raise ValueError(head)
    out = torch.cat(rows, 0)
    if head == "to_mean":
        return F.linear(torch.mean(out, dim=1), w, b)
    if head == "compr_to_rf":
        return F.linear(torch.median(out, dim=1)[0], w, b)
    if head == "lstm":
        return F.linear(lstm_forward(sd, out), w, b)
    return out


def lstm_forward(sd, x, prefix: str = "lstm."):
    """nn.LSTM(num_layers=1, batch_first=True) from a zero state (torch_cnn_lstm_combo.py:19, 46), restated from the
    torch.nn.LSTM documentation: gates (i, f, g, o) = W_ih x_t + b_ih + W_hh h_{t-1} + b_hh."""
    w_ih, w_hh = sd[prefix + "weight_ih_l0"], sd[prefix + "weight_hh_l0"]
    b_ih, b_hh = sd[prefix + "bias_ih_l0"], sd[prefix + "bias_hh_l0"]
    hid = w_hh.shape[1]
    h = x.new_zeros(x.shape[0], hid)
    c = x.new_zeros(x.shape[0], hid)
    outs = []
    for t in range(x.shape[1]):
        gates = F.linear(x[:, t], w_ih, b_ih) + F.linear(h, w_hh, b_hh)
        i, f, g, o = gates.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h.unsqueeze(1))
    return torch.cat(outs, 1)


def regressor_forward(sd, x, **kw):
    """CNNRegressor.forward, torch_cnn_bm_regressor.py:14-19: the backbone on a FLAT batch x:(N, 1, 224)
    (BatchNorm over all N breaths), then Linear."""
    if x.shape[-1] != SEQ_LEN:
        raise Exception("input breaths must have sequence length of 224")
    return F.linear(backbone_forward(sd, x, **kw).squeeze(), sd["linear_final.weight"], sd["linear_final.bias"])


def bce_with_logits(outputs, target):
    """torch.nn.BCEWithLogitsLoss() (mean over all B*2 elements), train_ards_detector.py:530."""
    return F.binary_cross_entropy_with_logits(outputs, target)


def forward_backward(sd, x, target, clip_val: Optional[float] = None, per_breath: bool = False, **kw):
    """One training step's forward + backward (train_ards_detector.py:153, 162-163) on leaf copies
    of the floating-point parameters.  Returns (logits, loss, {name: grad}).  Parameters that
    take no part in the graph (conv1_alt, conv2, bn2) get no entry, like `.grad is None`."""
    leaves = OrderedDict()
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            leaves[k] = v.detach().clone().requires_grad_(True)
        else:
            leaves[k] = v
    out = cnn_linear_forward(leaves, x, per_breath=per_breath, **kw)
    loss = bce_with_logits(out, target)
    names = [k for k, v in leaves.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    gd = OrderedDict()
    for k, gr in zip(names, grads):
        if gr is None:
            continue
        if clip_val is not None:
            gr = gr.clamp(-clip_val, clip_val)  # the register_hook clamp, train_ards_detector.py:474-476
        gd[k] = gr
    return out.detach(), loss.detach(), gd


def gradcam_forward_backward(sd, x, target: Optional[int] = None):
    """gradcam.py:40-65, 83-99 for ONE sequence x:(20, 1, 224): returns A=(20,128,7) features,
    dA (gradient of the chosen logit wrt A) and the (1,2) model output.  DenseNet only."""
    prefix = "breath_block."
    a = densenet_features(sd, x, prefix).detach().requires_grad_(True)
    # the reference differentiates through `features` as well (parameters require grad); dA does
    # not depend on that part of the graph, so the oracle cuts it here.
    y = F.avg_pool1d(F.relu(a), 7, 1).reshape(-1)
    out = F.linear(y, sd["linear_final.weight"], sd["linear_final.bias"]).unsqueeze(0)
    if target is None:
        target = int(out.argmax())
    (da,) = torch.autograd.grad(out[0, target], a)
    return a.detach(), da, out.detach()


def cam_normalize(cam):
    """MaxMinNormCam.normalize, gradcam.py:156-161: ReLU, min-max to [0,1], truncate to uint8.
    numpy float32 arithmetic like the reference (cam arrays are float32 there)."""
    import numpy as np
    cam = np.maximum(np.asarray(cam, dtype=np.float32), 0)
    with np.errstate(invalid="ignore", divide="ignore"):
        cam = (cam - np.min(cam)) / (np.max(cam) - np.min(cam))
        return np.uint8(cam * 255)


def gradcam_read_cam(sd, x, target: Optional[int] = None):
    """MaxMinNormCam.generate_read_cam, gradcam.py:125-136: one CAM row per breath of the sequence.
    Returns (cam uint8 (20,7), raw float32 cam (20,7) before normalisation, model output (1,2))."""
    import numpy as np
    a, da, out = gradcam_forward_backward(sd, x, target)
    conv_output, grad = a.numpy(), da.numpy()
    weights = np.mean(grad, axis=(2,))
    raw = np.zeros((conv_output.shape[0], conv_output.shape[2]), dtype=np.float32)
    for i in range(conv_output.shape[0]):
        for j in range(conv_output.shape[1]):
            raw[i] += weights[i, j] * conv_output[i, j, :]
    cam = np.stack([cam_normalize(raw[i]) for i in range(raw.shape[0])])
    return cam, raw, out


def gradcam_seq_cam(sd, x, target: Optional[int] = None, normalize: bool = True):
    """MaxMinNormCam.generate_cam (gradcam.py:138-154) / UnNormalizedCam.generate_cam (:195-205): ONE CAM row for
    the whole sequence (channel weights and activations averaged over the breaths).  Returns (cam, raw, out)."""
    import numpy as np
    a, da, out = gradcam_forward_backward(sd, x, target)
    conv_output, grad = a.numpy(), da.numpy()
    weights = np.mean(grad, axis=(0, 2))
    conv_output = np.mean(conv_output, axis=0)
    raw = np.zeros(conv_output.shape[1:], dtype=np.float32)
    for i, w in enumerate(weights):
        raw += w * conv_output[i, :]
    return (cam_normalize(raw) if normalize else raw), raw, out


def cam_resize_linear_u8(cam, out_len: int = SEQ_LEN):
    """cv2.resize(cam, (1, out_len)) of a uint8 column (patient_gradcam.py:217, 227-229), i.e. OpenCV's default
    INTER_LINEAR for 8-bit images: half-pixel centres, coefficients quantised to 11 bits, the vertical pass computing
    ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2 on rows pre-scaled by 2^11.
    PARITY UNPINNED: OpenCV (opencv-python, an un-vendored third-party dependency, environment-py3.yml) is absent
    from this image, so this restates the published algorithm (modules/imgproc/src/resize.cpp, 4.x) without a vector
    from the library itself."""
    import numpy as np
    cam = np.asarray(cam, dtype=np.uint8).ravel()
    n = cam.shape[0]
    scale = n / float(out_len)
    out = np.zeros(out_len, dtype=np.uint8)
    for d in range(out_len):
        fy = (d + 0.5) * scale - 0.5
        sy = int(math.floor(fy))
        fy -= sy
        y0, y1 = min(max(sy, 0), n - 1), min(max(sy + 1, 0), n - 1)
        b0 = int(np.rint(np.float32((1.0 - fy) * 2048)))
        b1 = int(np.rint(np.float32(fy * 2048)))
        s0, s1 = int(cam[y0]) * 2048, int(cam[y1]) * 2048
        out[d] = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2
    return out


def scale_windows(data, mu: float, std: float, padded: bool = False):
    """ARDSRawDataset.__getitem__ scaling, dataset.py:1375-1379 (+ `_get_padding_mask`, :1406-1409) followed by the
    trainer's `.float()` (train_ards_detector.py:150-151): float64 arithmetic, one rounding to float32.
    padded=True is the padded_breath_by_breath rule: mu is subtracted only where the raw sample is non-zero."""
    import numpy as np
    data = np.asarray(data, dtype=np.float64)
    if padded:
        mask = np.zeros(data.shape)
        np.put(mask, np.where(data.ravel() != 0)[0], v=mu)
        data = (data - mask) / std
    else:
        data = (data - mu) / std
    return torch.from_numpy(data).float()


# --------------------------------------------------------------------------
# batched formulation (same numbers, one pass over all B*20 breaths) -- used by the
# tests to show "grouped BN with group = 20" == the reference loop
# --------------------------------------------------------------------------
def grouped_batchnorm(x, weight, bias, group: int):
    """BatchNorm over groups of `group` consecutive rows of x:(N, C, L)."""
    n, c, l = x.shape
    xg = x.reshape(n // group, group, c, l)
    mean = xg.mean(dim=(1, 3), keepdim=True)
    var = xg.var(dim=(1, 3), unbiased=False, keepdim=True)
    y = (xg - mean) / torch.sqrt(var + BN_EPS)
    y = y * weight.view(1, 1, c, 1) + bias.view(1, 1, c, 1)
    return y.reshape(n, c, l)


# --------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------
DATASET_MU = 2.056
DATASET_STD = 28.08


def synthetic_breaths(n_seq: int, seed: int = 1234, sub_batch: int = 20) -> torch.Tensor:
    """(n_seq, sub_batch, 1, 224) fp32 ventilator-flow-like windows, z-scored with the
    dataset constants stored in the reference's tests/test_dataset.pkl."""
    gen = torch.Generator().manual_seed(seed)
    n = n_seq * sub_batch
    ni = torch.randint(40, 81, (n, 1), generator=gen).float()
    peak = 30.0 + 40.0 * torch.rand(n, 1, generator=gen)
    t = torch.arange(SEQ_LEN, dtype=torch.float32).view(1, -1)
    insp = peak * torch.sin(math.pi * t / ni)
    exp_ = -0.6 * peak * torch.exp(-(t - ni) / 25.0)
    flow = torch.where(t < ni, insp, exp_) + 1.5 * torch.randn(n, SEQ_LEN, generator=gen)
    flow = (flow - DATASET_MU) / DATASET_STD
    return flow.view(n_seq, sub_batch, 1, SEQ_LEN).contiguous()


def synthetic_targets(n_seq: int, seed: int = 1234) -> torch.Tensor:
    gen = torch.Generator().manual_seed(seed + 1)
    cls = (torch.rand(n_seq, generator=gen) < 0.5).long()
    return F.one_hot(cls, 2).float()


def patient_vote_table(patient, y_true, y_pred, n_classes=2):
    """CPU restatement of the per-patient loop of DeepARDSResults.perform_patient_predictions (deepards/metrics.py:572-600,
    helpers get_tps/get_fps/get_tns/get_fns :29-62): patients in order of first appearance (`y_test.patient.unique()`); per
    class n the counts over the patient's windows of (actual == n & pred == n), (actual != n & pred == n),
    (actual != n & pred != n), (actual == n & pred != n) and the votes (pred == n); prediction = np.argmax(votes),
    pred_frac = votes[1] / sum(votes).  Returns an (n_patients, 2 + 5 * n_classes + 2) float64 table with the reference's
    column order: patient, patho, [tps, fps, tns, fns, votes] per class, prediction, pred_frac."""
    import numpy as np
    patient, y_true, y_pred = (np.asarray(a) for a in (patient, y_true, y_pred))
    seen, rows = set(), []
    for pt in patient:
        if pt in seen:
            continue
        seen.add(pt)
        m = patient == pt
        actual, pred = y_true[m], y_pred[m]
        row = [float(pt), float(actual[0])]                      # pt_rows.y.unique()[0]
        votes = []
        for n in range(n_classes):
            row += [float(np.sum((actual == n) & (pred == n))), float(np.sum((actual != n) & (pred == n))),
                    float(np.sum((actual != n) & (pred != n))), float(np.sum((actual == n) & (pred != n))),
                    float(np.sum(pred == n))]
            votes.append(row[-1])
        row += [float(np.argmax(votes)), votes[1] / sum(votes)]
        rows.append(row)
    return np.array(rows, dtype=np.float64)
