"""TEST INFRASTRUCTURE - generates tests/golden/ref_eval.npz with the LIVE reference's label converters and Averager
(/root/reference/ocr/tools/recog_utils.py; this container only) and torch's loss functions called exactly as
ocr/train/crnn.py:186-208 calls them, on seeded random logits.  Usage: python -m oracle.make_golden_eval
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALPHABET = "0123456789abcdefghijklmnopqrstuvwxyz"


def cases(seed=0, n=48):
    """Seeded (logits [n, 26, C], labels) for both heads: peaked logits that mostly spell the label (so losses are
    finite and some predictions are correct), repeated characters, empty and over-long labels."""
    rng = np.random.default_rng(seed)
    labels = []
    for i in range(n):
        L = int(rng.integers(0, 14)) if i % 7 else int(rng.integers(20, 26))
        s = "".join(ALPHABET[int(rng.integers(0, 36))] for _ in range(L))
        if i % 5 == 0 and L >= 2:
            s = s[:1] + s[:1] + s[2:]                      # a doubled character needs a blank in between
        labels.append(s)
    labels[3] = ""
    labels[4] = "aabbccddeeffgghhiijjkkllm"                      # 25 symbols, 12 doubles: no 26-step alignment exists
    return rng, labels


def main():
    sys.path.insert(0, ROOT)
    from oracle import ref_env
    scratch = "/tmp/locr_ref_scratch"
    os.makedirs(scratch, exist_ok=True)
    dst = ref_env.stage("CTC", {"x": torch.zeros(1)}, {"x": torch.zeros(1)}, scratch)
    out = {}
    with ref_env.imported(dst):
        from tools import recog_utils as ru
        rng, labels = cases()
        n = len(labels)
        out["labels"] = np.array(labels)
        # ---- CTC: logits that follow the label with blanks between, plus noise
        conv = ru.CTCLabelConverter(ALPHABET)
        text, length = conv.encode(labels, batch_max_len=25)
        out["ctc_text"], out["ctc_length"] = text.numpy(), length.numpy()
        logits = rng.normal(0, 1.0, (n, 26, 37)).astype(np.float32)
        for i, s in enumerate(labels):
            path = []
            for j, c in enumerate(s):
                if len(s) <= 12 or (j > 0 and s[j - 1] == c):
                    path.append(0)
                path.append(conv.dict[c])
            path = (path + [0] * 26)[:26]
            for t, k in enumerate(path):
                logits[i, t, k] += 5.0 if i % 3 else 1.0
        preds = torch.from_numpy(logits)
        sizes = torch.IntTensor([26] * n)
        lsm = preds.log_softmax(2).permute(1, 0, 2)
        out["ctc_logits"] = logits
        out["ctc_cost"] = torch.nn.CTCLoss(zero_infinity=True)(lsm, text, sizes, length).numpy()
        out["ctc_loss"] = torch.nn.CTCLoss(zero_infinity=True, reduction="none")(lsm, text, sizes, length).numpy()
        _, idx = preds.max(2)
        # the live decode overwrites its `text` argument (recog_utils.py:43): valid for one sequence per call
        dec = [conv.decode(idx[i].data, torch.IntTensor([26]))[0] for i in range(n)]
        out["ctc_decoded"] = np.array(dec)
        out["ctc_correct"] = np.array([int(p == g) for p, g in zip(dec, labels)], np.int32)
        # ---- Attention: the live encode fills row 0 only (it returns inside its loop): recorded per single label
        aconv = ru.AttnLabelConverter(ALPHABET)
        rows, lens = [], []
        for s in labels:
            t, l = aconv.encode([s], batch_max_len=25)
            rows.append(t.cpu().numpy()[0])
            lens.append(int(l.cpu().numpy()[0]))
        atext = torch.from_numpy(np.stack(rows))
        out["attn_text"], out["attn_length"] = atext.numpy(), np.array(lens, np.int32)
        alog = rng.normal(0, 1.0, (n, 26, 38)).astype(np.float32)
        for i in range(n):
            for t in range(26):
                k = int(atext[i, t + 1])
                if i % 4 == 1 and t == len(labels[i]):
                    continue                                   # no [s] bump: some predictions never stop
                alog[i, t, k if (k or t <= len(labels[i])) else int(rng.integers(0, 38))] += 4.0 if i % 3 else 0.5
        apreds = torch.from_numpy(alog)
        p = apreds[:, :atext.shape[1] - 1, :]
        target = atext[:, 1:]
        ce = torch.nn.CrossEntropyLoss(ignore_index=0)
        out["attn_logits"] = alog
        out["attn_cost"] = ce(p.contiguous().view(-1, p.shape[-1]), target.contiguous().view(-1)).numpy()
        _, aidx = p.max(2)
        # the live decode overwrites its `text` argument here too (recog_utils.py:117): one row per call
        apred = [aconv.decode(aidx[i:i + 1], torch.IntTensor(lens[i:i + 1]))[0] for i in range(n)]
        agt = [aconv.decode(target[i:i + 1], torch.IntTensor(lens[i:i + 1]))[0] for i in range(n)]
        ok = []
        for gt, pred in zip(agt, apred):                       # crnn.py:222-230
            gt = gt[:gt.find("[s]")]
            pred = pred[:pred.find("[s]")]
            ok.append(int(pred == gt))
        out["attn_decoded"] = np.array(apred)
        out["attn_correct"] = np.array(ok, np.int32)
        # ---- Averager over 0-d costs
        av = ru.Averager()
        for v in (0.5, 1.25, 3.0):
            av.add(torch.tensor(v))
        out["averager"] = np.array(float(av.val()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_eval.npz"), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()})
    print("ctc correct %d / %d, attn correct %d / %d, ctc cost %.4f attn cost %.4f, zero losses %d" % (
        out["ctc_correct"].sum(), n, out["attn_correct"].sum(), n, out["ctc_cost"], out["attn_cost"],
        int((out["ctc_loss"] == 0).sum())))


if __name__ == "__main__":
    main()
