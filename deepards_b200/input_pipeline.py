"""The input contract of the path on the device.

`ARDSRawDataset.__getitem__` scales every window on the host in float64 numpy, `(data - mu) / std`
(deepards/dataset.py:1375-1379; padded_breath_by_breath subtracts mu only from the non-zero samples, :1406-1409), the
DataLoader collates, and the trainer converts with `.float()` and copies to the GPU (train_ards_detector.py:150-152).
`WindowScaler` takes the STORED windows instead (float64 as the dataset holds them, or float32), already on the device
or in pinned host memory, and produces the float32 network input with one kernel -- same float64 arithmetic and the
same single rounding, so the result is bit-identical to the reference's.
"""
import torch

from . import _lib


class WindowScaler(object):
    def __init__(self, mu, std, padded=False):
        if float(std) == 0.0:
            raise ValueError("std must not be 0")
        self.mu, self.std, self.padded = float(mu), float(std), bool(padded)

    @classmethod
    def for_dataset_type(cls, mu, std, dataset_type):
        """padded rule iff 'padded_breath_by_breath' is in the dataset type (dataset.py:1375)."""
        return cls(mu, std, padded='padded_breath_by_breath' in dataset_type)

    def __call__(self, raw, device=None, out=None):
        """raw: float64 / float32 tensor of windows (any shape).  A host tensor is copied to `device` first
        (non-blocking when pinned).  Returns a float32 tensor of the same shape on the device."""
        if raw.dtype not in (torch.float64, torch.float32):
            raise TypeError("raw windows must be float64 or float32, got %s" % raw.dtype)
        if raw.device.type != "cuda":
            if device is None:
                raise RuntimeError("deepards_b200 runs on CUDA devices only: pass device= for host windows "
                                   "(there is no CPU fallback)")
            raw = raw.to(device, non_blocking=True)
        raw = raw.contiguous()
        if out is None:
            out = torch.empty(raw.shape, dtype=torch.float32, device=raw.device)
        elif out.dtype != torch.float32 or out.numel() != raw.numel() or out.device != raw.device or not out.is_contiguous():
            raise ValueError("out must be a contiguous float32 tensor of the same size on the same device")
        _lib.call("dards_scale_windows", raw.data_ptr(), 1 if raw.dtype == torch.float64 else 0, out.data_ptr(),
                  raw.numel(), self.mu, self.std, 1 if self.padded else 0,
                  torch.cuda.current_stream(raw.device).cuda_stream)
        return out
