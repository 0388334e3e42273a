"""Golden vector from ALL 20 real sequences of the reference's deepards/tests/test_dataset.pkl (build container only:
imports the unmodified reference from /root/reference).  DenseNet-18 (the reference's default backbone,
deepards/defaults.yml:18), drop_rate 0 so the step is deterministic; records logits, loss and every parameter gradient
(large tensors sampled, as in make_golden.py).

    python oracle/make_golden_real20.py      ->  tests/golden/densenet18_B20_real_all.npz
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import make_golden as G  # noqa: E402  (reference imports, real_sequences, run_case)
from oracle import cnn_linear_oracle as O  # noqa: E402


def main():
    import torch
    torch.manual_seed(0)
    torch.set_num_threads(8)
    xr, tr, mu, std = G.real_sequences(20)
    assert xr.shape == (20, 20, 1, 224), xr.shape
    sd = O.cnn_linear_state("densenet18", seed=7, bn_perturb=0.1)
    dn = G.ref_densenet18()
    for m in dn.modules():
        if hasattr(m, "drop_rate"):
            m.drop_rate = 0.0
    G.run_case("densenet18_B20_real_all", G.RefCNNLinearNetwork(dn, 20, 0), sd, xr, tr,
               extra={"scaling": __import__("numpy").array([mu, std])})


if __name__ == "__main__":
    main()
