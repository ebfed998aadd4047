"""Validation loop of the reference's training script (ocr/train/crnn.py:142-240, `evaluation`) on the B200 engine:
SURVEY 8f row 4, the inference-side half of the training adjacency.

The reference feeds batches of a validation set through the recogniser, computes the validation loss
(`torch.nn.CTCLoss(zero_infinity=True)` over `preds.log_softmax(2)`, or `CrossEntropyLoss(ignore_index=0)` over the
attention decoder's steps), decodes greedily and counts exact matches.  Here one `locr_evaluate` call per batch does
all of that on the GPU: recognition, the CTC forward (alpha) recursion or the cross entropy straight from the logits
in HBM, the collapsed arg-max path compared with the label; only per-crop scalars come back.

Same return tuple as the reference.  Differences, all because the reference's attention branch cannot run as written:
`net(img, preds_text, trainning=False)` (crnn.py:198) raises TypeError (the keyword is `training`) and
`AttnLabelConverter.encode` returns after the first label (recog_utils.py:96) - the intended behaviour is implemented
(greedy decode, every row encoded).  The backward pass / optimiser step of `train_batch` (crnn.py:243-268) is training
and out of scope: a CUDA library behind a C ABI has no autograd graph.
"""
import time

import numpy as np

from .hostops import ALPHABET, AttnLabelConverter, CTCLabelConverter


class Averager:
    """recog_utils.py:122-141: mean of the batch costs."""

    def __init__(self):
        self.reset()

    def add(self, v):
        v = np.asarray(v, np.float32)
        self.n_count += v.size
        self.sum += float(v.sum())

    def reset(self):
        self.n_count = 0
        self.sum = 0.0

    def val(self):
        return self.sum / float(self.n_count) if self.n_count != 0 else 0


def evaluation(engine, val_batches, config=None):
    """engine: bridge.Pipeline (CTC or Attention head); val_batches: iterable of (crops, labels) - crops a list of uint8
    gray (HxW) or BGR (HxWx3) arrays, labels a list of str over the alphabet; config: dict with `batch_max_len`
    (default 25) and optionally `max_iter`.

    Returns (valid_loss, accuracy, preds_, confidence_, label, infer_, len_data) like the reference: preds_,
    confidence_ and label are those of the LAST batch (the reference overwrites them per batch)."""
    config = config or {}
    bml = int(config.get("batch_max_len", 25))
    attn = engine.head not in (0, "CTC")      # bridge.HEAD_CTC
    converter = AttnLabelConverter(ALPHABET) if attn else CTCLabelConverter(ALPHABET)
    num_correct, len_data, infer_ = 0, 0, 0.0
    avg_loss = Averager()
    preds_, confidence_, label = [], [], []
    for i, (crops, labels) in enumerate(val_batches):
        if "max_iter" in config and i >= int(config["max_iter"]):
            break
        len_data += len(crops)
        loss_text, len_loss = converter.encode(list(labels), batch_max_len=bml)
        start_ = time.time()
        o = engine.evaluate(crops, loss_text, len_loss)
        infer_ += time.time() - start_
        avg_loss.add(o["cost"])
        num_correct += int(o["correct"].sum())
        if attn:
            # evaluation() prints / returns the raw 26-token strings and the labels re-decoded from the targets
            preds_ = converter.decode(o["ids"], len_loss)
            label = converter.decode(loss_text[:, 1:], len_loss)
        else:
            preds_ = o["text"]
            label = list(labels)
        confidence_ = list(o["conf"])
    accuracy = num_correct / float(len_data) * 100 if len_data else 0.0
    return avg_loss.val(), accuracy, preds_, confidence_, label, infer_, len_data
