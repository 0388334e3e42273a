"""Patient-level vote aggregation on the device.

The reference classifies windows (sequences) and then votes per patient: the per-patient loop of
`DeepARDSResults.perform_patient_predictions` (deepards/metrics.py:572-600) counts, for every class n, the true/false
positives/negatives of the patient's windows and the number of windows predicted n, takes the majority class as the
patient's prediction (np.argmax: ties go to class 0) and the fraction of ARDS votes as `pred_frac`.  With pandas that is
five boolean-indexing passes per patient and class on the host; here it is one histogram over (patient, true class,
predicted class) on the device, so the window predictions never have to leave the GPU one by one.

    table = patient_vote_table(patient_ids, y_true, y_pred)         # all int64 tensors of one entry per window

returns a dict of per-patient tensors in the reference's row order (patients by first appearance) and column names:
patient, patho, {OTHER,ARDS}_{tps,fps,tns,fns,votes}, prediction, pred_frac."""
import torch

PATHOS = {0: "OTHER", 1: "ARDS"}


def patient_vote_table(patient, y_true, y_pred, n_classes=2):
    if not (patient.shape == y_true.shape == y_pred.shape) or patient.dim() != 1:
        raise ValueError("patient, y_true and y_pred must be 1-D tensors of the same length")
    if patient.numel() == 0:
        raise ValueError("no windows")
    dev = patient.device
    patient, y_true, y_pred = patient.long(), y_true.long(), y_pred.long()
    if int(y_true.min()) < 0 or int(y_true.max()) >= n_classes or int(y_pred.min()) < 0 or int(y_pred.max()) >= n_classes:
        raise ValueError("class labels must be in [0, %d)" % n_classes)
    ids, inv = torch.unique(patient, sorted=True, return_inverse=True)
    # the reference walks y_test.patient.unique(): order of first appearance
    first = torch.full((ids.numel(),), patient.numel(), dtype=torch.long, device=dev)
    first.scatter_reduce_(0, inv, torch.arange(patient.numel(), device=dev), reduce="amin")
    order = torch.argsort(first)
    rank = torch.empty_like(order)
    rank[order] = torch.arange(order.numel(), device=dev)
    p = rank[inv]                                                     # dense patient index in row order
    n_pt = ids.numel()
    hist = torch.bincount((p * n_classes + y_true) * n_classes + y_pred, minlength=n_pt * n_classes * n_classes)
    hist = hist.view(n_pt, n_classes, n_classes)                      # [patient][true][predicted]
    total = hist.sum(dim=(1, 2))
    out = {"patient": ids[order]}
    # `pt_rows.y.unique()[0]`: the label of the patient's first window
    out["patho"] = y_true[first[order]]
    votes = hist.sum(dim=1)                                           # [patient][predicted]
    for n in range(n_classes):
        name = PATHOS.get(n, str(n))
        tps = hist[:, n, n]
        fps = votes[:, n] - tps
        fns = hist[:, n, :].sum(dim=1) - tps
        out[name + "_tps"], out[name + "_fps"] = tps, fps
        out[name + "_tns"], out[name + "_fns"] = total - tps - fps - fns, fns
        out[name + "_votes"] = votes[:, n]
    out["prediction"] = torch.argmax(votes, dim=1)                    # first maximum, like np.argmax
    out["pred_frac"] = votes[:, 1].double() / total.double()
    return out
