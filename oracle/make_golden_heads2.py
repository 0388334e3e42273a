"""Golden vectors for the last two sibling components of SURVEY.md 8f-4 (build container only; imports the unmodified
reference from /root/reference):

  * CNNTransformerNetwork (deepards/models/cnn_transformer.py:8-44 over deepards/models/transformer.py).  The reference
    file is Python 2: `xrange` and an integer `/` in MultiHeadAttention.__init__ (`head_size = hidden_size / num_heads`).
    It is run here under a two-line compatibility shim that restores exactly those Python-2 semantics (builtins.xrange =
    range; head_size floor-divided after construction) -- nothing else is touched.  Dropout (p = 0.2 in every block) is set
    to 0 for a deterministic step, as for the other heads.
  * patient-vote aggregation: the per-patient loop of DeepARDSResults.perform_patient_predictions
    (deepards/metrics.py:572-600) run by the reference's own method on synthetic window predictions; third-party
    imports of metrics.py that are absent here (matplotlib, prettytable, dtwco, ...) are stubbed, they are not used by
    that loop.

    python oracle/make_golden_heads2.py   ->   tests/golden/transformer_head.npz, tests/golden/patient_votes.npz
"""
import builtins
import os
import sys
import types
from unittest import mock as umock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from oracle import cnn_linear_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def transformer_head():
    builtins.xrange = range                                   # Python-2 shim (1/2)
    from deepards.models.resnet import resnet18 as ref_resnet18
    from deepards.models.cnn_transformer import CNNTransformerNetwork as Ref
    import deepards.models.transformer as T
    orig_init = T.MultiHeadAttention.__init__

    def init(self, input_size, hidden_size, num_heads):       # Python-2 shim (2/2): int / int is floor division
        orig_init(self, input_size, hidden_size, num_heads)
        self.head_size = hidden_size // num_heads

    T.MultiHeadAttention.__init__ = init
    torch.manual_seed(7)
    model = Ref(ref_resnet18(initial_planes=16), 0, False, 64, 2)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    base = O.cnn_linear_state("resnet18", seed=9, bn_perturb=0.1, initial_planes=16, per_breath=True)
    sd = model.state_dict()
    for k, v in base.items():
        if k.startswith("breath_block."):
            sd[k] = v
    gen = torch.Generator().manual_seed(123)
    for k in list(sd):
        if k.startswith(("transformer.", "linear_final.")):
            sd[k] = (torch.randn(sd[k].shape, generator=gen) * (0.1 if k.endswith("weight") and sd[k].dim() == 2 else 0.05) +
                     (1.0 if ("norm.weight" in k) else 0.0))
    model.load_state_dict(sd, strict=True)
    model.train()
    x = O.synthetic_breaths(2, seed=56)
    t = O.synthetic_targets(2, seed=56)
    out = model(x, torch.tensor(float("nan")))
    loss = torch.nn.BCEWithLogitsLoss()(out, t.unsqueeze(1).repeat(1, 20, 1))
    loss.backward()
    rec = {"x": x.numpy(), "target": t.numpy(), "logits": out.detach().numpy(), "loss": np.array(loss.item(), np.float64),
           "keys": np.array(list(model.state_dict().keys()))}
    for name, p in model.named_parameters():
        if p.grad is None:
            continue
        if name.startswith(("transformer.", "linear_final.")) or name in ("breath_block.conv1.weight", "breath_block.bn1.weight"):
            rec["grad/" + name] = p.grad.detach().numpy().copy()
    for k, v in sd.items():
        if k.startswith(("transformer.", "linear_final.")):
            rec["sd/" + k] = v.numpy()
    path = os.path.join(OUT, "transformer_head.npz")
    np.savez_compressed(path, **rec)
    print("%s  %.1f KB  logits[0,0]=%s loss=%.6f" % (path, os.path.getsize(path) / 1024, rec["logits"][0, 0], loss.item()))


def patient_votes():
    stubs = {"matplotlib": {}, "matplotlib.pyplot": {}, "prettytable": {"PrettyTable": object}, "seaborn": {},
             "mock": {"Mock": umock.Mock}, "dtwco": {}, "dtwco.warping": {}, "dtwco.warping.core": {"dtw": None},
             "ventmap": {}, "ventmap.raw_utils": {}, "imblearn": {}, "imblearn.under_sampling": {}, "algorithms": {},
             "algorithms.breath_meta": {}}
    for m, attrs in stubs.items():
        try:
            __import__(m)
        except Exception:
            mod = types.ModuleType(m)
            for k, v in attrs.items():
                setattr(mod, k, v)
            sys.modules[m] = mod
    import pandas as pd
    import deepards.metrics as M
    rs = np.random.RandomState(11)
    # 9 patients with arbitrary ids, in an interleaved order; 5..60 windows each; one patient with every vote for ARDS,
    # one with none, one with an exact tie
    ids = [413, 27, 1999, 8, 356, 77, 1201, 5, 640]
    rows = []
    for i, pt in enumerate(ids):
        y = int(rs.rand() < 0.5)
        n = int(rs.randint(5, 61))
        if i == 2:
            pred = np.ones(n, int)
        elif i == 3:
            pred = np.zeros(n, int)
        elif i == 4:
            n = 10
            pred = np.array([0, 1] * 5)
        else:
            pred = (rs.rand(n) < (0.7 if y else 0.3)).astype(int)
        rows += [(pt, y, int(p)) for p in pred]
    order = rs.permutation(len(rows))
    rows = [rows[j] for j in order]
    y_test = pd.DataFrame({"patient": [r[0] for r in rows], "y": [r[1] for r in rows]})
    predictions = pd.Series([r[2] for r in rows])
    res = object.__new__(M.DeepARDSResults)
    res.pathos = {0: 'OTHER', 1: 'ARDS'}
    cols = ["patient", "patho"]
    for n, patho in res.pathos.items():
        cols.extend(["{}_tps".format(patho), "{}_fps".format(patho), "{}_tns".format(patho), "{}_fns".format(patho),
                     "{}_votes".format(patho)])
    cols += ["prediction", 'pred_frac', 'epoch_num', 'fold_num']
    res.results = pd.DataFrame([], columns=cols)
    try:
        res.perform_patient_predictions(y_test, predictions, 0, 3)
    except Exception as e:      # the aggregate reporting behind the per-patient loop needs the stubbed libraries
        print("after the per-patient loop:", type(e).__name__, str(e)[:80])
    table = res.results
    assert len(table) == len(ids), len(table)
    path = os.path.join(OUT, "patient_votes.npz")
    np.savez_compressed(path, patient=y_test.patient.values.astype(np.int64), y=y_test.y.values.astype(np.int64),
                        pred=predictions.values.astype(np.int64), columns=np.array(cols[:-2]),
                        table=table[cols[:-2]].values.astype(np.float64))
    print("%s  %.1f KB\n%s" % (path, os.path.getsize(path) / 1024, table[cols[:-2]].head(4)))


if __name__ == "__main__":
    transformer_head()
    patient_votes()
