"""TEST INFRASTRUCTURE - synthetic checkpoints (re-exported from lightly_ocr_b200/synth/weights.py) plus the one-off
calibration that needs the oracle's forward pass (`build_calibrations`, run by oracle/make_golden.py in the authoring
container; its results are committed as lightly_ocr_b200/synth/calib_*.npz)."""
from collections import OrderedDict

import numpy as np
import torch

from lightly_ocr_b200.synth import specs
from lightly_ocr_b200.synth import weights as _w
from lightly_ocr_b200.synth.weights import (calibrate_craft, craft_calibrated, craft_state_dict,  # noqa: F401
                                            crnn_calibrated, crnn_state_dict, has_fp32_checkpoint, _pca_head,
                                            _plant_ink_path)


def build_calibrations():
    """Measure BatchNorm statistics on sample inputs (as a few training-mode batches would), then fit the score head
    (CRAFT) / the prediction head (CRNN, rows along the principal directions of the contextual features so that the
    argmax depends on the input).  Run once in the authoring container by oracle/make_golden.py; results committed."""
    import os
    from . import ocr_ref, receipts
    _HERE = os.path.dirname(os.path.abspath(_w.__file__))
    img0 = receipts.receipt(0)
    with torch.no_grad():
        # ---------------- CRAFT
        base = craft_state_dict(0, ink=False)
        x, _, _ = ocr_ref.craft_preproc(np.ascontiguousarray(img0[:640, :480]))
        ocr_ref.BN_CALIBRATE = True
        ocr_ref.craft_forward(base, x)
        ocr_ref.BN_CALIBRATE = False
        bn_keys = [k for k in base if k.endswith("running_mean") or k.endswith("running_var")]
        xfull, _, _ = ocr_ref.craft_preproc(img0)
        for ink in (False, True):
            sd = OrderedDict((k, v.clone()) for k, v in base.items())
            if ink:
                _plant_ink_path(sd)
            taps = {}
            ocr_ref.craft_forward(sd, xfull, taps)
            calibrate_craft(sd, taps["h16"][0], ink=ink)
            out = {k: base[k].numpy() for k in bn_keys}
            out["conv_cls.8.weight"] = sd["conv_cls.8.weight"].numpy()
            out["conv_cls.8.bias"] = sd["conv_cls.8.bias"].numpy()
            np.savez_compressed(os.path.join(_HERE, "calib_craft_ink%d.npz" % int(ink)), **out)
        # ---------------- CRNN
        samples = receipts.crops(96, seed=7)
        xb = torch.cat([ocr_ref.crop_to_tensor(g)[1] for g in samples], 0)
        for head in ("CTC", "Attention"):
            sd = crnn_state_dict(1, head=head)
            ocr_ref.BN_CALIBRATE = True
            ocr_ref.crnn_forward(sd, xb, "CTC" if head == "CTC" else "skip")
            ocr_ref.BN_CALIBRATE = False
            bn_keys = [k for k in sd if k.endswith("running_mean") or k.endswith("running_var")]
            taps = {}
            ocr_ref.crnn_forward(sd, xb, "CTC" if head == "CTC" else "skip", taps)
            out = {k: sd[k].numpy() for k in bn_keys}
            ncls = 37 if head == "CTC" else 38
            if head == "CTC":
                feats = taps["contextual"].reshape(-1, specs.HIDDEN).double()
                W, bvec = _pca_head(feats, ncls, gain=3.0)
                bvec[0] += 2.0   # blank a little more likely, strings of plausible length
                out["Prediction.weight"] = W.float().numpy()
                out["Prediction.bias"] = bvec.float().numpy()
            else:
                hs = ocr_ref.attention_decode(sd, taps["contextual"], ncls, return_hidden=True)[1]
                W, bvec = _pca_head(hs.reshape(-1, specs.HIDDEN).double(), ncls, gain=3.0)
                # the two special tokens take the lowest-variance directions (the start state is an outlier along
                # the leading ones and would otherwise always decode [s] first)
                W, bvec = torch.roll(W, 2, 0), torch.roll(bvec, 2, 0).clone()
                bvec[1] += 5.0   # [s]: appears at varied positions for a part of the crops
                bvec[0] -= 6.0   # [GO] is never emitted by a trained decoder
                out["Prediction.generator.weight"] = W.float().numpy()
                out["Prediction.generator.bias"] = bvec.float().numpy()
            np.savez_compressed(os.path.join(_HERE, "calib_crnn_%s.npz" % head.lower()), **out)


