"""TEST INFRASTRUCTURE - generates tests/golden/*.npz by running the REAL reference (this container only).

/root/reference is read-only and absent on the GPU box, so the live reference is imported here once, fed the
synthetic checkpoints of oracle/weights.py, and its outputs are committed as small fixtures.  The reference sources
are staged in a scratch directory under /tmp by oracle/ref_env.py (shims and staging are documented there).

Usage:  python -m oracle.make_golden [--keep-calib] [--checkpoint fp32] | --poly

--checkpoint fp32 records the same fixtures for the plain-fp32-trained synthetic recogniser
(lightly_ocr_b200/synth/calib_crnn_*_fp32.npz) as tests/golden/ref_{ctc,attention}_fp32.npz.
--poly records only tests/golden/ref_poly.npz: the live reference's getDetBoxes(..., poly=True) (ocr/tools/det_utils.py
:97-256) on the curved synthetic score maps.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def poly_goldens(scratch):
    """getDetBoxes(textmap, linkmap, 0.7, 0.4, 0.4, poly=True) of the live reference on curved_score_maps(0..5)."""
    from oracle import receipts, ref_env, weights
    dst = ref_env.stage("CTC", {"x": torch.zeros(1)}, {"x": torch.zeros(1)}, scratch)
    out = {}
    with ref_env.imported(dst):
        import tools as ref_tools
        for seed in range(6):
            t, l = receipts.curved_score_maps(seed)
            boxes, polys = ref_tools.getDetBoxes(t, l, 0.7, 0.4, 0.4, True)
            out["s%d_boxes" % seed] = np.array(boxes, np.float32).reshape(-1, 4, 2)
            out["s%d_valid" % seed] = np.array([p is not None for p in polys], np.int32)
            out["s%d_polys" % seed] = np.array([p if p is not None else np.zeros((14, 2)) for p in polys],
                                               np.float64).reshape(-1, 14, 2)
    np.savez_compressed(os.path.join(GOLDEN, "ref_poly.npz"), **out)
    print("poly", {k: v.shape for k, v in out.items() if k.endswith("valid")},
          sum(int(v.sum()) for k, v in out.items() if k.endswith("valid")), "polygons")


def main():
    sys.path.insert(0, ROOT)
    from oracle import ocr_ref, receipts, ref_env, weights
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLDEN, exist_ok=True)
    scratch = "/tmp/locr_ref_scratch"
    os.makedirs(scratch, exist_ok=True)
    if "--poly" in sys.argv:
        poly_goldens(scratch)
        return

    ckpt = sys.argv[sys.argv.index("--checkpoint") + 1] if "--checkpoint" in sys.argv else "trained"
    trained = "fp32" if ckpt == "fp32" else True
    suffix = "_fp32" if ckpt == "fp32" else ""
    # ---- calibrations of the synthetic checkpoints (committed so every machine loads identical tensors)
    if "--keep-calib" not in sys.argv and ckpt != "fp32":
        weights.build_calibrations()
    img0 = receipts.receipt(0)
    craft_calibrated = weights.craft_calibrated

    for head in ("CTC", "Attention"):
        craft_sd = craft_calibrated(0, ink=True)
        crnn_sd = weights.crnn_calibrated(1, head=head, trained=trained)
        dst = ref_env.stage(head, craft_sd, crnn_sd, scratch)
        with ref_env.imported(dst):
            import net as ref_net
            import pipeline as ref_pipeline
            import tools as ref_tools
            out = {}
            detector, recognizer = ref_pipeline.prepModel(ref_pipeline.CONFIG, docker=True)

            if head == "CTC":
                # -- config 2 at reduced size: reference CRAFT on a 256x192 window of receipt(0) (ratio 1.5 resize path)
                win = np.ascontiguousarray(img0[40:296, 40:232])
                with torch.no_grad():
                    xt, rw, rh = detector.preproc(win)
                    y, feat = detector.net(xt)
                text = y[0, :, :, 0].numpy().copy()
                link = y[0, :, :, 1].numpy().copy()
                rects = detector.getCoords([text, link], rw, rh)
                out["craft_win"] = win
                out["craft_x"] = xt.numpy()
                out["craft_text"] = text
                out["craft_link"] = link
                out["craft_feature"] = feat.numpy().astype(np.float16)
                out["craft_rects"] = np.array(rects, np.int32).reshape(-1, 4)
                out["craft_ratio"] = np.array([rw, rh], np.float64)
                roi = detector.process(win)
                out["craft_roi_shapes"] = np.array([r.shape[:2] for r in roi], np.int32).reshape(-1, 2)
                # -- second window with native-resolution maps: identity-resize path (target == canvas)
                # -- post-processing on synthetic score maps (inputs regenerate from the seed; outputs stored)
                for seed in (1, 2):
                    t, l = receipts.score_maps(seed)
                    boxes, polys = ref_tools.getDetBoxes(t, l, 0.7, 0.4, 0.4, False)
                    out["maps%d_boxes" % seed] = np.array(boxes, np.float32).reshape(-1, 4, 2)
                    rr = detector.getCoords([t, l], 1.0, 1.0)
                    out["maps%d_rects" % seed] = np.array(rr, np.int32).reshape(-1, 4)
                    from functools import cmp_to_key
                    srt = sorted(rr, key=cmp_to_key(ref_tools.compare_rects))
                    out["maps%d_sorted" % seed] = np.array(srt, np.int32).reshape(-1, 4)
                # -- CTC decode known answers of the reference's own unit test (ocr/test/utils_test.py:37-43)
                import string
                conv = ref_tools.CTCLabelConverter(string.ascii_lowercase)
                out["kat_fifa"] = np.array(conv.decode(torch.IntTensor([6, 9, 6, 1]), torch.IntTensor([4])))
                out["kat_ea"] = np.array(conv.decode(torch.IntTensor([5, 5, 0, 1]), torch.IntTensor([4])))

            # -- recognizer: config 1 crop + ragged crops
            crop0 = np.random.default_rng(0).integers(0, 256, (32, 100), dtype=np.uint8)
            crop_list = [crop0] + receipts.crops(15, seed=3)
            preds_all, raw_all, conf_all, u8_all = [], [], [], []
            for g in crop_list:
                with contextlib.redirect_stdout(io.StringIO()):
                    raw, preds = recognizer.getPreds(g)
                    res = {}
                    try:
                        raw_p, res = recognizer.process(res, g)
                    except IndexError:      # reference quirk: [s] at position 0 -> cumprod of an empty tensor
                        res = {-2.0: "IndexError"}
                preds_all.append(preds.numpy()[0])
                raw_all.append(raw[0])
                conf_all.append([float(k) for k in res.keys()] or [-1.0])
                from PIL import Image
                tt = recognizer.transformer(Image.fromarray(g).convert("L"))
                u8_all.append((tt[0] * 0.5 + 0.5).mul(255).round().to(torch.uint8).numpy())
            out["crnn_preds"] = np.stack(preds_all)
            out["crnn_raw"] = np.array(raw_all)
            out["crnn_conf"] = np.array([c[0] for c in conf_all], np.float32)
            out["crnn_res_text"] = np.array([str(list(r)) for r in [raw_all]])
            out["crnn_u8"] = np.stack(u8_all)
            # intermediate taps of the reference network for layer-wise parity (first two crops)
            with torch.no_grad():
                xb = torch.cat([recognizer.transformer(Image.fromarray(g).convert("L")).unsqueeze(0)
                                for g in crop_list[:2]], 0)
                tr = recognizer.net.Transformation(xb)
                vf = recognizer.net.FeatureExtraction(tr)
            out["crnn_rectified"] = tr.numpy()
            out["crnn_visual"] = vf.numpy().astype(np.float32)

            # -- end to end: getText on full receipts through the reference pipeline (pipeline.py:65-87)
            import cv2
            texts, confs, counts = [], [], []
            for rid in ((1, 2, 3) if head == "CTC" else (2,)):
                path = os.path.join(scratch, "receipt%d.png" % rid)
                cv2.imwrite(path, receipts.receipt(rid))
                with contextlib.redirect_stdout(io.StringIO()):
                    res = ref_pipeline.getText(path, detector, recognizer, write=False)
                vals = [v[0] if isinstance(v, list) else v for v in res.values()]
                texts.extend(vals)
                confs.extend(float(k) for k in res.keys())
                counts.append(len(vals))
            out["e2e_receipts"] = np.array((1, 2, 3) if head == "CTC" else (2,), np.int32)
            out["e2e_counts"] = np.array(counts, np.int32)
            out["e2e_text"] = np.array(texts)
            out["e2e_conf"] = np.array(confs, np.float32)
            if suffix:      # the detector side does not depend on the recogniser's checkpoint: keep the file small
                out = {k: v for k, v in out.items() if k.startswith(("crnn_", "e2e_"))}
            np.savez_compressed(os.path.join(GOLDEN, "ref_%s%s.npz" % (head.lower(), suffix)), **out)
            print(head, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
