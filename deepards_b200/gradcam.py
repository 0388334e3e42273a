"""GradCAM on the B200 backend -- drop-in for the model-facing classes of deepards/gradcam.py.

`GradCam`, `MaxMinNormCam`, `UnNormalizedCam` and `FracTotalNormCam` keep the reference's constructor and method
names (gradcam.py:68-205): `generate_one_hot_grad_and_output(input, target)`, `generate_cam(input, target=None)`,
`generate_read_cam(input, target)`, with `input` one sequence (20, 1, 224) and numpy results of the same shapes
and dtypes.  `generate_read_cams` / `generate_cams` are the batched forms for a whole recording (BASELINE config 5):
(B, 20, 1, 224) in, all maps of all sequences out of ONE forward plan + ONE `dards_gradcam` launch.

Where the reference runs forward + a one-hot backward through the whole DenseNet for every call and reduces A / dA
with numpy on the host (gradcam.py:83-99, 125-154), this runs the `features` forward plan only: the gradient of a
logit w.r.t. the norm5 output is (A > 0) * W[target] / 7 in closed form (ReLU -> AvgPool1d(7) -> Linear), which the
kernel evaluates together with the reductions, the min-max normalisation and the 7 -> 224 linear resize of
patient_gradcam.py:217, 227-229.  Requires a `CNNLinearNetwork` over a DenseNet (`breath_block.features`), like the
reference (gradcam.py:45).
"""
import numpy as np
import torch

from . import _lib, autograd as _ag, engine


class CamMaps(object):
    """Device results of one `dards_gradcam` launch over B sequences (all torch tensors on the GPU)."""
    __slots__ = ("logits", "target", "read_raw", "read_u8", "seq_raw", "seq_u8", "read_resized", "seq_resized",
                 "conv_output", "gradients")


def _features_plan(model, n_breaths, group):
    bb = model.breath_block
    feats = getattr(bb, "features", None)
    if feats is None or not hasattr(feats, "_drop_key"):
        raise TypeError("GradCAM needs a deepards_b200 DenseNet breath_block (`.features`), like gradcam.py:45")
    training = bb.training
    return engine.get_plan(feats, feats, None, n_breaths, group, _ag.module_precision(model), "features",
                           dropout=feats._drop_key() if training else (), update_running=False)


def compute_maps(model, x, target=None, resized_len=0, want_tensors=False):
    """x: (B, group, 1, 224) or one sequence (group, 1, 224) on the model's device.  target: None (predicted class of
    each sequence), an int, or an int tensor/sequence of B class indices.  Returns CamMaps (device tensors)."""
    if x.shape[-1] != 224:
        raise Exception('input breaths must have sequence length of 224')
    if x.dim() == 3:
        x = x.unsqueeze(0)
    if x.dim() != 4 or x.shape[2] != 1:
        raise NotImplementedError("GradCAM takes (B, breaths, 1, 224) or (breaths, 1, 224) inputs, got %s" % (tuple(x.shape),))
    lin = model.linear_final
    b, group = x.shape[0], x.shape[1]
    plan = _features_plan(model, b * group, group)
    if x.device != plan.device:
        raise RuntimeError("input is on %s but the network is on %s" % (x.device, plan.device))
    f = plan.out_features
    if lin.in_features != group * f:
        raise RuntimeError("linear_final expects %d features, the sequence provides %d" % (lin.in_features, group * f))
    n_out = lin.out_features
    with torch.no_grad():
        plan.load_input(x if x.dtype == torch.float32 else x.float())
        plan.run_forward()
        plan.mark_no_backward()
    dev, L = plan.device, plan.feat_map.shape[1]
    m = CamMaps()
    m.logits = torch.empty((b, n_out), dtype=torch.float32, device=dev)
    m.target = torch.empty((b,), dtype=torch.int32, device=dev)
    m.read_raw = torch.empty((b, group, L), dtype=torch.float32, device=dev)
    m.read_u8 = torch.empty((b, group, L), dtype=torch.uint8, device=dev)
    m.seq_raw = torch.empty((b, L), dtype=torch.float32, device=dev)
    m.seq_u8 = torch.empty((b, L), dtype=torch.uint8, device=dev)
    m.read_resized = torch.empty((b, group, resized_len), dtype=torch.uint8, device=dev) if resized_len else None
    m.seq_resized = torch.empty((b, resized_len), dtype=torch.uint8, device=dev) if resized_len else None
    m.conv_output = torch.empty((b * group, f, L), dtype=torch.float32, device=dev) if want_tensors else None
    m.gradients = torch.empty((b * group, f, L), dtype=torch.float32, device=dev) if want_tensors else None
    tdev, tfix = None, -1
    if target is not None:
        if isinstance(target, (int, np.integer)):
            tfix = int(target)
        else:
            tdev = torch.as_tensor(target, dtype=torch.int32).to(dev).contiguous()
            if tdev.numel() != b:
                raise ValueError("need one target per sequence (%d), got %d" % (b, tdev.numel()))

    def ptr(t):
        return t.data_ptr() if t is not None else None

    d = _lib.GradcamDesc(a=plan.feat_map.data_ptr(), w=lin.weight.data_ptr(), bias=lin.bias.data_ptr(),
                         target_dev=ptr(tdev), logits=ptr(m.logits), target_used=ptr(m.target), read_raw=ptr(m.read_raw),
                         read_u8=ptr(m.read_u8), seq_raw=ptr(m.seq_raw), seq_u8=ptr(m.seq_u8),
                         read_resized=ptr(m.read_resized), seq_resized=ptr(m.seq_resized), conv_out=ptr(m.conv_output),
                         grad_out=ptr(m.gradients), a_stride=f, target=tfix, n_groups=b, group=group, l=L, f=f,
                         n_out=n_out, resized_len=int(resized_len), dtype=plan.dt, reserved=0)
    _lib.call("dards_gradcam", d, torch.cuda.current_stream(dev).cuda_stream)
    return m


def recording_maps(model, x, targets=None, resized_len=224, group=None):
    """BASELINE config 5: the read maps of a whole recording, x (B, 20, 1, 224), sharded over the ranks of `group`
    (one process per GPU, model replicated).  Sequences are independent, so every rank runs its contiguous shard through
    one forward plan + one `dards_gradcam` launch and the uint8 maps / logits are gathered in order -- no collective on
    the data path.  Returns (maps (B, 20, resized_len) uint8, logits (B, 2)) on every rank."""
    from .data_parallel import shard_bounds, sharded_map
    import torch.distributed as dist
    tsr = None
    if targets is not None and not isinstance(targets, (int, np.integer)):
        tsr = torch.as_tensor(targets, dtype=torch.int32)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    lo = shard_bounds(x.shape[0], dist.get_world_size(group), dist.get_rank(group))[0] if multi else 0

    def fn(shard):
        if shard.shape[0] == 0:
            dev = next(model.parameters()).device
            return (torch.empty((0, x.shape[1], resized_len or 7), dtype=torch.uint8, device=dev),
                    torch.empty((0, model.linear_final.out_features), dtype=torch.float32, device=dev))
        t = targets if tsr is None else tsr[lo:lo + shard.shape[0]]
        m = compute_maps(model, shard, t, resized_len=resized_len)
        return (m.read_resized if resized_len else m.read_u8), m.logits

    return sharded_map(fn, x, group)


class GradCam(object):
    """Produces class activation maps (gradcam.py:68-107)."""

    def __init__(self, model):
        self.model = model

    def generate_one_hot_grad_and_output(self, input, target):
        """-> (conv_output (20,F,7) ndarray, guided_gradients (20,F,7) ndarray, model_output (1,2) tensor)."""
        m = compute_maps(self.model, input, target, want_tensors=True)
        return m.conv_output.cpu().numpy(), m.gradients.cpu().numpy(), m.logits[:1]


class MaxMinNormCam(GradCam):
    """Maps normalised by their own min / max (gradcam.py:110-162)."""

    def generate_read_cam(self, input, target):
        m = compute_maps(self.model, input, target)
        return m.read_u8[0].cpu().numpy(), m.logits[:1]

    def generate_cam(self, input, target=None):
        m = compute_maps(self.model, input, target)
        return m.seq_u8[0].cpu().numpy(), m.logits[:1]

    # ---- batched forms (one launch for a whole recording) ----
    def generate_read_cams(self, inputs, targets=None, resized_len=0):
        """inputs (B,20,1,224) -> uint8 maps (B,20,7) [or (B,20,resized_len)], logits (B,2); device tensors."""
        m = compute_maps(self.model, inputs, targets, resized_len=resized_len)
        return (m.read_resized if resized_len else m.read_u8), m.logits

    def generate_cams(self, inputs, targets=None, resized_len=0):
        m = compute_maps(self.model, inputs, targets, resized_len=resized_len)
        return (m.seq_resized if resized_len else m.seq_u8), m.logits


class UnNormalizedCam(GradCam):
    """ReLU of the raw sequence map (gradcam.py:195-205)."""

    def generate_cam(self, input, target=None):
        m = compute_maps(self.model, input, target)
        return torch.clamp_min(m.seq_raw[0], 0).cpu().numpy(), m.logits[:1]


class FracTotalNormCam(GradCam):
    """Target map as a fraction of target + other-class map (gradcam.py:165-192); two kernel launches."""

    def generate_cam(self, input, target):
        raise NotImplementedError('Havent done this yet')

    def generate_read_cam(self, input, target):
        mt = compute_maps(self.model, input, target)
        mo = compute_maps(self.model, input, (target + 1) % 2)
        t = torch.clamp_min(mt.read_raw[0], 0)
        o = torch.clamp_min(mo.read_raw[0], 0)
        cam = torch.nan_to_num(t / (o + t) * 255, nan=0.0).to(torch.uint8)
        return cam.cpu().numpy(), mt.logits[:1]
