"""Synthetic workload generator for the benchmarks: ventilator-flow-like breath windows and one-hot
targets.  Pure torch on the CPU; bench.py's product arm and the tools take their batches from here (the CPU checker used by the
tests keeps its own copy of the recipe; tests/test_boundary_cpu.py holds the two bit-identical).

Breath model: inspiratory length ni ~ U{40..80} samples, peak flow p ~ U(30, 70); flow = p sin(pi t / ni) for t < ni, then
-0.6 p exp(-(t - ni) / 25) up to 224 samples, plus N(0, 1.5) noise; z-scored with the constants of the reference's
deepards/tests/test_dataset.pkl (mu = 2.056, std = 28.08), i.e. what ARDSRawDataset.__getitem__ hands to the network
(deepards/dataset.py:1375-1379)."""
import math

import torch
import torch.nn.functional as F

SEQ_LEN = 224
DATASET_MU = 2.056
DATASET_STD = 28.08


def synthetic_breaths(n_seq, seed=1234, sub_batch=20):
    """(n_seq, sub_batch, 1, 224) fp32 windows."""
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


def synthetic_targets(n_seq, seed=1234):
    """(n_seq, 2) one-hot of Bernoulli(0.5), the target layout of train_ards_detector.py:929-930."""
    gen = torch.Generator().manual_seed(seed + 1)
    cls = (torch.rand(n_seq, generator=gen) < 0.5).long()
    return F.one_hot(cls, 2).float()
