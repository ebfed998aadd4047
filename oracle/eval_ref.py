"""TEST INFRASTRUCTURE - CPU oracle of the reference's validation loop (ocr/train/crnn.py:142-240, `evaluation`).

The training script cannot be imported (it opens LMDB data sets and starts training at import time), so its
`evaluation` function is restated here over the oracle's forward pass (oracle/ocr_ref.py), with the reference's own
third-party arithmetic: `torch.nn.CTCLoss(zero_infinity=True)` (crnn.py:119) on `preds.log_softmax(2).permute(1, 0, 2)`
(:190), `torch.nn.CrossEntropyLoss(ignore_index=0)` (:121, :203-208), `F.softmax` / `cumprod` for the confidence
(:215-235).  The label converters' `encode` (tools/recog_utils.py:24-30, :84-96) and the `Averager` (:122-141) are
restated as well and pinned against the LIVE classes by tests/golden/ref_eval.npz (oracle/make_golden_eval.py).
`ctc_loss_alpha` restates the CTC forward recursion itself in numpy float64 (the blueprint of ctc_loss_kernel) and is
pinned against torch.nn.CTCLoss in tests/test_eval_oracle.py.

Where the reference's attention branch cannot run as written - `net(..., trainning=False)` raises TypeError (:198) and
`AttnLabelConverter.encode` returns after the first label - the intended behaviour is restated (greedy decode with the
B = 1 semantics of CRNN.getPreds; every row encoded).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import ocr_ref

ALPHABET = "0123456789abcdefghijklmnopqrstuvwxyz"


def ctc_encode(labels):
    """CTCLabelConverter.encode (recog_utils.py:24-30): blank = 0, characters from 1."""
    d = {c: i + 1 for i, c in enumerate(ALPHABET)}
    return (torch.IntTensor([d[c] for c in "".join(labels)]), torch.IntTensor([len(s) for s in labels]))


def attn_encode(labels, batch_max_len=25):
    """AttnLabelConverter.encode (recog_utils.py:84-96) as intended: [GO] = 0, [s] = 1, every row filled."""
    chars = ["[GO]", "[s]"] + list(ALPHABET)
    d = {c: i for i, c in enumerate(chars)}
    out = torch.zeros(len(labels), batch_max_len + 2, dtype=torch.long)
    for i, t in enumerate(labels):
        ids = [d[c] for c in t] + [d["[s]"]]
        out[i, 1:1 + len(ids)] = torch.LongTensor(ids)
    return out, torch.IntTensor([len(s) + 1 for s in labels])


def ctc_loss_alpha(logits, target):
    """-log p(target | logits) by the CTC forward recursion, numpy float64.  logits [T, C]; target: class indices > 0.
    inf when no alignment exists."""
    x = np.asarray(logits, np.float64)
    lp = x - x.max(1, keepdims=True)
    lp = lp - np.log(np.exp(lp).sum(1, keepdims=True))
    T = lp.shape[0]
    L = len(target)
    ext = [0] * (2 * L + 1)
    ext[1::2] = [int(t) for t in target]
    S = len(ext)
    a = np.full(S, -np.inf)
    a[0] = lp[0, 0]
    if S > 1:
        a[1] = lp[0, ext[1]]
    for t in range(1, T):
        b = np.full(S, -np.inf)
        for s in range(S):
            terms = [a[s]]
            if s > 0:
                terms.append(a[s - 1])
            if s > 1 and ext[s] != 0 and ext[s] != ext[s - 2]:
                terms.append(a[s - 2])
            m = max(terms)
            if m > -np.inf:
                b[s] = m + np.log(sum(np.exp(v - m) for v in terms)) + lp[t, ext[s]]
        a = b
    tail = [a[S - 1]] + ([a[S - 2]] if S > 1 else [])
    m = max(tail)
    if m == -np.inf:
        return np.inf
    return -(m + np.log(sum(np.exp(v - m) for v in tail)))


def batch_losses(preds, labels, head="CTC", batch_max_len=25):
    """The loss half of evaluation() (crnn.py:186-208) on given preds [B, 26, C] fp32 torch.
    Returns (cost scalar, per-crop unreduced losses, correct flags, prediction strings, label strings)."""
    B = preds.shape[0]
    if head == "CTC":
        text, length = ctc_encode(labels)
        sizes = torch.IntTensor([preds.size(1)] * B)
        lsm = preds.log_softmax(2).permute(1, 0, 2)
        cost = torch.nn.CTCLoss(zero_infinity=True)(lsm, text, sizes, length)
        per = torch.nn.CTCLoss(zero_infinity=True, reduction="none")(lsm, text, sizes, length)
        _, idx = preds.max(2)
        pred_s = [ocr_ref.ctc_decode(idx[b].numpy()) for b in range(B)]
        gt_s = list(labels)
        correct = [int(p == g) for p, g in zip(pred_s, gt_s)]
        return float(cost), per.numpy(), np.array(correct, np.int32), pred_s, gt_s
    text, length = attn_encode(labels, batch_max_len)
    p = preds[:, :text.shape[1] - 1, :]
    target = text[:, 1:]
    cost = torch.nn.CrossEntropyLoss(ignore_index=0)(p.contiguous().view(-1, p.shape[-1]), target.contiguous().view(-1))
    per = torch.nn.CrossEntropyLoss(ignore_index=0, reduction="none")(
        p.contiguous().view(-1, p.shape[-1]), target.contiguous().view(-1)).view(B, -1).sum(1)
    chars = ["[GO]", "[s]"] + list(ALPHABET)
    _, idx = p.max(2)
    pred_s = ["".join(chars[int(i)] for i in idx[b]) for b in range(B)]
    gt_s = ["".join(chars[int(i)] for i in target[b]) for b in range(B)]
    correct = []
    for gt, pred in zip(gt_s, pred_s):            # crnn.py:222-230, str.find semantics kept (-1 cuts the last character)
        gt = gt[:gt.find("[s]")]
        pred = pred[:pred.find("[s]")]
        correct.append(int(pred == gt))
    return float(cost), per.numpy(), np.array(correct, np.int32), pred_s, gt_s


def evaluation(crnn_sd, val_batches, head="CTC", batch_max_len=25):
    """evaluation() (crnn.py:142-240) over (crops, labels) batches: crops are uint8 gray arrays resized like the
    validation data set does (ResizeNormalize((100, 32)), tools/dataset.py:37-47).  Returns (valid_loss, accuracy,
    per-batch details)."""
    total, n_sum, correct, n = 0.0, 0, 0, 0
    details = []
    with torch.no_grad():
        for crops, labels in val_batches:
            preds = torch.cat([ocr_ref.crnn_forward(crnn_sd, ocr_ref.crop_to_tensor(c)[1], head=head) for c in crops], 0)
            cost, per, ok, pred_s, gt_s = batch_losses(preds, labels, head, batch_max_len)
            probs = F.softmax(preds, dim=2).max(dim=2)[0]
            total += cost                        # Averager.add of a 0-d tensor: sum += v, n_count += 1
            n_sum += 1
            correct += int(ok.sum())
            n += len(crops)
            details.append(dict(cost=cost, loss=per, correct=ok, preds=pred_s, labels=gt_s,
                                ids=preds.max(2)[1].numpy(),
                                conf=probs.cumprod(dim=1)[:, -1].numpy()))
    return (total / n_sum if n_sum else 0), correct / float(n) * 100, details
