"""Is the 2-rank-vs-1-rank divergence of the DenseNet-with-dropout trainer test a bug or the dynamics of the step?
(1) eager vs CUDA-graph replay of the same 5 steps with dropout on: must be bit-identical;
(2) the same 5 steps from initial weights perturbed by 1e-7 (relative), with and without dropout: how fast do two fp32
    trajectories part when nothing but rounding-sized noise distinguishes them?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import deepards_b200 as D  # noqa: E402
from deepards_b200.data_parallel import DataParallelTrainer  # noqa: E402
from oracle import cnn_linear_oracle as O  # noqa: E402  (tools/: measurement script, not product code)


def run(drop, use_graph, perturb=0.0, lr=1e-2):
    sd = O.cnn_linear_state("densenet18", seed=41, bn_perturb=0.1)
    if perturb:
        g = torch.Generator().manual_seed(7)
        sd = {k: (v * (1 + perturb * torch.randn(v.shape, generator=g)) if v.is_floating_point() and "running" not in k else v)
              for k, v in sd.items()}
    net = D.CNNLinearNetwork(D.densenet18(drop_rate=drop), 20, 0)
    net.load_state_dict(sd)
    net = net.cuda().train()
    net.precision = "fp32"
    batches = [(O.synthetic_breaths(5, seed=500 + i).cuda(), O.synthetic_targets(5, seed=500 + i).cuda()) for i in range(4)]
    tr = DataParallelTrainer(net, lr=lr, clip_val=0.01, use_graph=use_graph)
    losses = [float(tr.train_step(*batches[i % 4])) for i in range(5)]
    return losses, tr.param_flat.clone()


for drop in (0.2, 0.0):
    l_e, p_e = run(drop, False)
    l_g, p_g = run(drop, True)
    print("drop %.1f: eager == graph: losses %s, parameters %s" % (drop, l_e == l_g, bool(torch.equal(p_e, p_g))))
    l_p, p_p = run(drop, False, perturb=1e-7)
    print("drop %.1f: initial weights perturbed by 1e-7: loss |diff| per step %s, parameters rel diff %.2e" %
          (drop, " ".join("%.1e" % abs(a - b) for a, b in zip(l_e, l_p)), float((p_e - p_p).abs().max() / p_e.abs().max())))
