"""CPU: the evaluation oracle (oracle/eval_ref.py, restating ocr/train/crnn.py:142-240) against the goldens recorded
from the LIVE reference's label converters / Averager and torch's losses (oracle/make_golden_eval.py), the numpy CTC
forward recursion against torch.nn.CTCLoss, and the host-side mirror's converters against the same goldens."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_eval.npz")


def _gold():
    return np.load(GOLD)


def test_encoders_match_live_reference():
    from oracle import eval_ref
    from lightly_ocr_b200 import hostops
    g = _gold()
    labels = [str(s) for s in g["labels"]]
    text, length = eval_ref.ctc_encode(labels)
    assert np.array_equal(text.numpy(), g["ctc_text"]) and np.array_equal(length.numpy(), g["ctc_length"])
    atext, alen = eval_ref.attn_encode(labels)
    assert np.array_equal(atext.numpy(), g["attn_text"]) and np.array_equal(alen.numpy(), g["attn_length"])
    # the product's host-side converters (lightly_ocr_b200/hostops.py) encode identically
    t2, l2 = hostops.CTCLabelConverter(hostops.ALPHABET).encode(labels)
    assert np.array_equal(t2, g["ctc_text"]) and np.array_equal(l2, g["ctc_length"])
    a2, al2 = hostops.AttnLabelConverter(hostops.ALPHABET).encode(labels)
    assert np.array_equal(a2, g["attn_text"]) and np.array_equal(al2, g["attn_length"])


def test_batch_losses_match_goldens():
    from oracle import eval_ref
    g = _gold()
    labels = [str(s) for s in g["labels"]]
    cost, per, ok, pred_s, _ = eval_ref.batch_losses(torch.from_numpy(g["ctc_logits"]), labels, "CTC")
    assert abs(cost - float(g["ctc_cost"])) < 1e-6
    assert np.allclose(per, g["ctc_loss"], rtol=1e-6, atol=1e-6)
    assert pred_s == [str(s) for s in g["ctc_decoded"]]
    assert np.array_equal(ok, g["ctc_correct"]) and 0 < ok.sum() < len(ok)
    cost, per, ok, pred_s, _ = eval_ref.batch_losses(torch.from_numpy(g["attn_logits"]), labels, "Attention")
    assert abs(cost - float(g["attn_cost"])) < 1e-6
    assert pred_s == [str(s) for s in g["attn_decoded"]]
    assert np.array_equal(ok, g["attn_correct"]) and 0 < ok.sum() < len(ok)


def test_ctc_alpha_recursion_matches_torch():
    """The numpy restatement of the CTC forward recursion (blueprint of ctc_loss_kernel) against torch.nn.CTCLoss:
    goldens, an infeasible target (inf -> 0 under zero_infinity) and an empty one."""
    from oracle import eval_ref
    g = _gold()
    off = 0
    n_inf = 0
    for i, L in enumerate(g["ctc_length"]):
        tg = g["ctc_text"][off:off + L]
        off += L
        v = eval_ref.ctc_loss_alpha(g["ctc_logits"][i], tg)
        if np.isinf(v):
            n_inf += 1
            assert g["ctc_loss"][i] == 0.0
        else:
            assert abs(v - g["ctc_loss"][i]) <= 2e-5 * max(1.0, abs(v)), (i, v, g["ctc_loss"][i])
    assert n_inf == 1


def test_averager_matches_live_reference():
    from lightly_ocr_b200.evaluate import Averager
    av = Averager()
    for v in (0.5, 1.25, 3.0):
        av.add(v)
    assert abs(av.val() - float(_gold()["averager"])) < 1e-7
