"""The drop-in `net.CRAFT` / `net.CRNN` classes against outputs recorded from the LIVE reference (tests/golden/*.npz,
made by oracle/make_golden.py from /root/reference/ocr): the secondary API surface of ocr/net.py that pipeline.py does
not exercise - `preproc`, `self.net(x)`, `getCoords`, `getPreds` - plus the post-processing on the reference's own
score maps (bit-exact) and the synthetic-map goldens of `getDetBoxes` (bit-exact)."""
import contextlib
import importlib
import io
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _setup(tmp_path, head):
    from lightly_ocr_b200.synth import weights
    d = tmp_path / ("ocr_" + head)
    (d / "save_models").mkdir(parents=True)
    torch.save(weights.craft_calibrated(0, ink=True), str(d / "save_models" / "CRAFT.pth"))
    torch.save(weights.crnn_calibrated(1, head), str(d / "save_models" / "CRNN.pth"))
    import yaml
    cfg = yaml.safe_load(open(os.path.join(os.path.dirname(__file__), "..", "lightly_ocr_b200", "config.yml")))
    cfg["prediction"] = head
    cfg["num_classes"] = 37 if head == "CTC" else 38
    yaml.safe_dump(cfg, open(str(d / "config.yml"), "w"))
    os.environ["LOCR_OCR_DIR"] = str(d)
    os.environ["LOCR_EAGER"] = "1"
    import lightly_ocr_b200.net as net
    for e in getattr(net, "_ENGINES", {}).values():
        e.close()
    return importlib.reload(net)


def test_craft_api_against_reference_goldens(tmp_path):
    from lightly_ocr_b200.synth import receipts
    g = np.load(os.path.join(GOLD, "ref_ctc.npz"))
    net = _setup(tmp_path, "CTC")
    det = net.CRAFT(device=net.DEVICE)
    win = g["craft_win"]
    # preproc: bit-identical tensor and ratios (net.py:71-80)
    x, rw, rh = det.preproc(win)
    assert np.array_equal(x.numpy(), g["craft_x"]) and (rw, rh) == tuple(g["craft_ratio"])
    # y, feature = self.net(x) (net.py:103): score maps within the 1e-2 gate of the reference's
    y, feat = det.net(x)
    assert tuple(y.shape) == (1,) + g["craft_text"].shape + (2,)
    e_t = np.abs(y[0, :, :, 0].numpy() - g["craft_text"]).max()
    e_l = np.abs(y[0, :, :, 1].numpy() - g["craft_link"]).max()
    print("score maps vs the live reference: text %.2e link %.2e" % (e_t, e_l))
    assert e_t < 1e-2 and e_l < 1e-2
    # getCoords on the REFERENCE's maps: rects bit-exact (det_utils.py getDetBoxes + adjustResultCoordinates + net.py:82-98)
    rects = det.getCoords([g["craft_text"], g["craft_link"]], rw, rh)
    assert np.array_equal(np.array(rects, np.int32).reshape(-1, 4), g["craft_rects"])
    # process(): the crops of the reference (shapes, in the reference's reading order)
    roi = det.process(win)
    assert np.array_equal(np.array([r.shape[:2] for r in roi], np.int32).reshape(-1, 2), g["craft_roi_shapes"])
    # synthetic score maps: boxes, rects and the reading-order sort of the live reference, bit for bit
    for seed in (1, 2):
        t, l = receipts.score_maps(seed)
        out = det.engine.postproc(np.stack([t, l], -1)[None], 1.0, 1.0, want_labels=False)[0]
        assert np.array_equal(out["boxes"], g["maps%d_boxes" % seed])
        assert np.array_equal(out["rects"], g["maps%d_rects" % seed])
        assert np.array_equal(np.array(net.sort_rects(out["rects"].tolist()), np.int32), g["maps%d_sorted" % seed])


@pytest.mark.parametrize("head", ["CTC", "Attention"])
def test_crnn_api_against_reference_goldens(tmp_path, head):
    from lightly_ocr_b200.synth import receipts
    g = np.load(os.path.join(GOLD, "ref_%s.npz" % head.lower()))
    net = _setup(tmp_path, head)
    rec = net.CRNN(device=net.DEVICE)
    crops = [np.random.default_rng(0).integers(0, 256, (32, 100), dtype=np.uint8)] + receipts.crops(15, seed=3)
    same, errs = 0, []
    for i, gray in enumerate(crops):
        with contextlib.redirect_stdout(io.StringIO()):
            raw, preds = rec.getPreds(gray)
        assert tuple(preds.shape) == (1, 26, g["crnn_preds"].shape[2])
        same += raw[0] == str(g["crnn_raw"][i])
        errs.append(float(np.abs(preds[0].numpy() - g["crnn_preds"][i])[0].max()))   # first step: no greedy feedback yet
        # transformer attribute: the reference's ResizeNormalize, byte-exact resize
        from PIL import Image
        tt = rec.transformer(Image.fromarray(gray).convert("L"))
        assert np.array_equal((tt[0] * 0.5 + 0.5).mul(255).round().to(torch.uint8).numpy(), g["crnn_u8"][i])
    print("%s getPreds vs the live reference: %d / %d raw strings identical, step-0 logit max-abs %.3f" %
          (head, same, len(crops), max(errs)))
    assert same >= len(crops) - 1
