"""Parity of the CUDA CRAFT / CRNN forward passes against the CPU oracle (fp32 torch) on the same inputs and the same
synthetic checkpoints.

Tolerances (floating point; north_star: score maps and logits within 1e-2 max-abs, 16-bit operands / fp32 accumulate):
  * CRAFT score maps: 1e-2 max-abs against the fp32 oracle with fp16 storage (measured 4.6e-3; maps span [-0.2, 3]).
    bf16 storage measures 3.4e-2 and does NOT meet the gate on random-init weights (SURVEY.md 7.4 predicted 4e-2 to
    6e-2), which is why fp16 is the default activation type; the bf16 case is kept as a bounded regression check.
  * CRNN logits: the default CTC checkpoint has a trained read-out (tools/train_synth_crnn.py) and therefore logits
    of a trained recogniser's scale: std ~7.7, |logit| up to ~40.  Measured with fp16 storage: max-abs 0.092 (0.23% of
    the logit range, 1.2% of the std), arg-max agreement 99.9%.  An absolute 1e-2 would need logits of unit scale;
    the test bounds the error at 3% of the logit std (fp16), the arg-max agreement at 99.5%, and pins every
    intermediate (fiducials, rectified crop, visual and contextual features) layer-wise.  The string gate of the
    north star (>= 99.5% identical strings end to end) is tested in test_pipeline_gpu.py.
  * integer results (token ids, strings, confidences) are compared exactly GIVEN the CUDA logits (decode parity).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ACT = {"bf16": 1, "f16": 0}


@pytest.fixture(scope="module")
def oracle_mods():
    from oracle import ocr_ref, receipts, weights
    torch.set_num_threads(os.cpu_count())
    return ocr_ref, receipts, weights


@pytest.mark.parametrize("act", ["f16", "bf16"])
def test_craft_score_maps(oracle_mods, act):
    ocr_ref, receipts, weights = oracle_mods
    from lightly_ocr_b200 import bridge
    sd = weights.craft_calibrated(0, ink=True)
    eng = bridge.Engine(act_dtype=ACT[act])
    eng.load_state_dict(bridge.MODEL_CRAFT, sd)
    img = np.ascontiguousarray(receipts.receipt(0)[40:360, 40:296])      # 320 x 256 window, two images in a batch
    img2 = np.ascontiguousarray(receipts.receipt(2)[400:720, 300:556])
    batch = np.stack([img, img2])
    got = eng.craft_scores(batch)
    taps = {}
    with torch.no_grad():
        x = torch.cat([ocr_ref.craft_preproc(i, canvas_size=10 ** 6, mag_ratio=1.0)[0] for i in batch], 0)
        ref = ocr_ref.craft_forward(sd, x, taps).numpy()
    # layer-wise: relative error of every tap (normalised by the tap's own max) localises a broken layer
    for name in ("slice1.0", "relu2_2", "relu3_2", "relu4_3", "relu5_3", "fc7", "feature"):
        g = eng.debug_read(name)
        r = taps[name].permute(0, 2, 3, 1).numpy()
        if g.shape[2] == r.shape[2] + 3:       # row-padded tensor: one zero pixel left, two right
            assert not g[:, :, 0].any() and not g[:, :, -2:].any()
            g = g[:, :, 1:-2]
        rel = np.abs(g - r).max() / max(np.abs(r).max(), 1e-6)
        print("%s %-9s rel max err %.4g" % (act, name, rel))
        assert rel < (0.05 if act == "bf16" else 0.01), (name, rel)
    err = np.abs(got - ref).max()
    flips = int(((got[..., 0] > 0.4) != (ref[..., 0] > 0.4)).sum() + ((got[..., 1] > 0.4) != (ref[..., 1] > 0.4)).sum())
    print("%s score max-abs err %.4g, range [%.3f, %.3f], threshold flips %d / %d" %
          (act, err, ref.min(), ref.max(), flips, ref.size))
    assert err < (1e-2 if act == "f16" else 6e-2)
    eng.close()


def test_craft_score_maps_full_canvas(oracle_mods):
    """BASELINE-size parity (reference ocr/net.py:100-107): a batch of eight 1280x960 receipts in one CRAFT pass, so the
    M = 256 tiles, the haloed-patch path at full width and the multi-wave persistent grids all run; the score maps of
    two of them are compared with the fp32 oracle (gate: 1e-2 max-abs), and the threshold flips are reported."""
    ocr_ref, receipts, weights = oracle_mods
    from lightly_ocr_b200 import bridge
    sd = weights.craft_calibrated(0, ink=True)
    eng = bridge.Engine(act_dtype=ACT["f16"])
    eng.load_state_dict(bridge.MODEL_CRAFT, sd)
    batch = np.stack([receipts.receipt(i) for i in range(8)])
    got = eng.craft_scores(batch)
    assert got.shape == (8, 640, 480, 2) and np.isfinite(got).all()
    for i in (0, 5):
        with torch.no_grad():
            x, _, _ = ocr_ref.craft_preproc(batch[i])
            ref = ocr_ref.craft_forward(sd, x)[0].numpy()
        err = np.abs(got[i] - ref).max()
        flips = int(((got[i] > 0.4) != (ref > 0.4)).sum() + ((got[i][..., 0] > 0.7) != (ref[..., 0] > 0.7)).sum())
        print("receipt %d at 1280x960 (batch of 8): score max-abs err %.4g, range [%.3f, %.3f], threshold flips %d / %d"
              % (i, err, ref.min(), ref.max(), flips, ref.size))
        assert err < 1e-2
        assert flips <= ref.size // 2000          # < 0.05 % of the map pixels sit within rounding distance of a threshold
    # one canvas alone (B = 1: other tile shapes / fewer waves) gives the same maps as in the batch
    alone = eng.craft_scores(batch[5:6])
    assert np.array_equal(alone[0], got[5])
    eng.close()


@pytest.mark.parametrize("shape", [(1, 480, 352), (3, 448, 352), (5, 352, 288), (2, 736, 416)], ids=str)
def test_craft_score_maps_odd_shapes(oracle_mods, shape):
    """Canvas sizes and batch sizes that give odd tile counts, ragged 16 x 16 tiles and last CTA pairs without a partner
    in the pair / haloed-patch forms of the conv kernel (165, 462, 99 x 5 and 2 x 299 tiles at half resolution): score maps
    within 1e-2 of the fp32 oracle, and within 2e-3 of the same image run alone."""
    ocr_ref, receipts, weights = oracle_mods
    from lightly_ocr_b200 import bridge
    B, H, W = shape
    sd = weights.craft_calibrated(0, ink=True)
    eng = bridge.Engine(act_dtype=ACT["f16"])
    eng.load_state_dict(bridge.MODEL_CRAFT, sd)
    batch = np.stack([np.ascontiguousarray(receipts.receipt(10 + i)[60:60 + H, 40:40 + W]) for i in range(B)])
    got = eng.craft_scores(batch)
    assert got.shape == (B, H // 2, W // 2, 2) and np.isfinite(got).all()
    with torch.no_grad():
        x = torch.cat([ocr_ref.craft_preproc(i, canvas_size=10 ** 6, mag_ratio=1.0)[0] for i in batch], 0)
        ref = ocr_ref.craft_forward(sd, x).numpy()
    err = float(np.abs(got - ref).max())
    print("%s: score max-abs err %.4g" % (shape, err))
    assert err < 1e-2
    if B > 1:
        # (not bit-identical in general: below 148 tiles slice1.10 takes the generic pair form, whose fp32 summation
        # order over the taps differs from the haloed-patch form's)
        alone = eng.craft_scores(batch[B - 1:B])
        assert float(np.abs(alone[0] - got[B - 1]).max()) < 2e-3
    eng.close()


@pytest.mark.parametrize("head", ["CTC", "Attention"])
@pytest.mark.parametrize("act", ["f16", "bf16"])
def test_crnn_logits_and_decode(oracle_mods, act, head):
    ocr_ref, receipts, weights = oracle_mods
    from lightly_ocr_b200 import bridge
    sd = weights.crnn_calibrated(1, head)
    eng = bridge.Engine(act_dtype=ACT[act], head=head)
    eng.load_state_dict(bridge.MODEL_CRNN, sd)
    crops = [np.random.default_rng(0).integers(0, 256, (32, 100), dtype=np.uint8)] + receipts.crops(39, seed=3)
    u8 = np.stack([ocr_ref.crop_to_tensor(g)[0] for g in crops])
    out = eng.crnn_on_resized(u8)
    taps = {}
    with torch.no_grad():
        x = torch.cat([ocr_ref.crop_to_tensor(g)[1] for g in crops], 0)
        ref = ocr_ref.crnn_forward(sd, x, head, taps).numpy()
    for name, tol in (("fiducials", 1e-4), ("rectified", 5e-3), ("visual", 1e-2), ("contextual", 1e-2)):
        g = eng.debug_read(name)
        r = taps[name].numpy().reshape(g.shape)
        rel = np.abs(g - r).max() / max(np.abs(r).max(), 1e-6)
        print("%s/%s %-10s rel max err %.4g" % (act, head, name, rel))
        assert rel < tol * (8 if act == "bf16" else 1), (name, rel)
    # decode parity given identical logits: recompute ids / strings / confidence from the CUDA logits on the host
    lg = torch.from_numpy(out["logits"])
    ids = lg.max(2)[1].numpy()
    assert np.array_equal(ids, out["ids"])
    probs = torch.softmax(lg, 2).max(2)[0]
    for i in range(len(crops)):
        if head == "CTC":
            assert out["text"][i] == ocr_ref.ctc_decode(ids[i])
            assert out["has_eos"][i] == 1
            want = float(probs[i].cumprod(0)[-1])
        else:
            s = ocr_ref.attn_decode_tokens(ids[i])
            pos = s.find("[s]")
            if pos < 0:
                assert out["has_eos"][i] == 0
                continue
            if pos == 0:
                assert out["has_eos"][i] == -1
                continue
            assert out["has_eos"][i] == 1 and out["text"][i] == s[:pos]
            want = float(probs[i][:pos].cumprod(0)[-1])
        assert abs(out["conf"][i] - want) <= 1e-4 * want + 1e-30
    if head == "CTC":
        scale = max(1.0, float(ref.std()))
        err = np.abs(out["logits"] - ref).max()
        agree = (ids == ref.argmax(2)).mean()
        same = np.mean([out["text"][i] == ocr_ref.ctc_decode(ref[i].argmax(1)) for i in range(len(crops))])
        print("%s CTC logits max-abs err %.4g (std %.3f), argmax agreement %.4f, string agreement %.3f" %
              (act, err, ref.std(), agree, same))
        assert err < 0.03 * scale * (1 if act == "f16" else 8)
        assert agree > (0.995 if act == "f16" else 0.97)
    else:
        # greedy feedback: compare the first step (no feedback yet) tightly, and report token agreement
        scale = max(1.0, float(ref.std()))
        err0 = np.abs(out["logits"][:, 0] - ref[:, 0]).max()
        agree = (ids == ref.argmax(2)).mean()
        print("%s Attention step-0 logits max-abs err %.4g (std %.3f), token agreement %.4f" %
              (act, err0, ref.std(), agree))
        assert err0 < 0.03 * scale * (1 if act == "f16" else 8)
        assert agree > (0.99 if act == "f16" else 0.9)
    eng.close()


def test_fp16_range_saturation_and_audit(oracle_mods):
    """fp16 storage has 5 exponent bits: a checkpoint whose activations exceed 65504 cannot be represented.  The path
    (a) never produces infinities / NaNs - conversions saturate -, (b) says where the range ran out (locr_audit), and
    (c) runs the same checkpoint with bf16 storage (LOCR_ACT_BF16, 8 exponent bits) within the bf16 tolerance.
    Checkpoint: the synthetic CRAFT weights with the first layer's BatchNorm affine scaled by 2^15 and the (fp32) last
    layer scaled back by 2^-15, so every 16-bit activation in between is ~3e4 times larger than usual."""
    ocr_ref, receipts, weights = oracle_mods
    from lightly_ocr_b200 import bridge
    base = weights.craft_calibrated(0, ink=True)
    sd = {k: v.clone() for k, v in base.items()}
    k = 15
    sd["basenet.slice1.1.weight"] *= 2.0 ** k
    sd["basenet.slice1.1.bias"] *= 2.0 ** k
    sd["conv_cls.8.weight"] *= 2.0 ** -k
    img = np.ascontiguousarray(receipts.receipt(0)[40:360, 40:296])
    with torch.no_grad():
        x = ocr_ref.craft_preproc(img, canvas_size=10 ** 6, mag_ratio=1.0)[0]
        taps = {}
        ref = ocr_ref.craft_forward(sd, x, taps).numpy()
    assert float(taps["slice1.0"].abs().max()) > 6.5e4            # beyond fp16 in the fp32 reference
    assert np.isfinite(ref).all()
    # (a) + (b): fp16 storage saturates instead of overflowing, and the audit names the layers that ran out of range
    e16 = bridge.Engine(act_dtype=ACT["f16"])
    e16.load_state_dict(bridge.MODEL_CRAFT, sd)
    e16.audit(True)
    got16 = e16.craft_scores(img[None])
    rows = e16.audit_read()
    e16.audit(False)
    assert np.isfinite(got16).all()
    amax = dict(rows)
    assert len(rows) >= 24 and amax["basenet.slice1.0"] == 65504.0
    saturated = [n for n, v in rows if v >= 65504.0]
    print("fp16 storage: %d of %d audited layers saturated (first: %s); outputs finite, max-abs diff to fp32 %.3g" %
          (len(saturated), len(rows), saturated[0], np.abs(got16 - ref).max()))
    e16.close()
    # the unscaled checkpoint has plenty of head-room (what the audit is for)
    e0 = bridge.Engine(act_dtype=ACT["f16"])
    e0.load_state_dict(bridge.MODEL_CRAFT, base)
    e0.audit(True)
    e0.craft_scores(img[None])
    rows0 = e0.audit_read()
    print("unscaled checkpoint: largest stored activation %.1f (%s)" % max((v, n) for n, v in rows0))
    assert max(v for _, v in rows0) < 65504.0 / 16
    e0.close()
    # (c) the same checkpoint with bf16 storage: in the bf16 tolerance of the fp32 reference
    eb = bridge.Engine(act_dtype=ACT["bf16"])
    eb.load_state_dict(bridge.MODEL_CRAFT, sd)
    gotb = eb.craft_scores(img[None])
    errb = np.abs(gotb - ref).max()
    print("bf16 storage on the large-activation checkpoint: score max-abs err %.4g" % errb)
    assert np.isfinite(gotb).all() and errb < 6e-2
    eb.close()


def _word_crops(receipts, seed, n):
    """Gray crops around the first n rendered words of a synthetic receipt (ground-truth boxes + a small margin)."""
    import cv2
    img, words = receipts.receipt(seed, return_words=True)
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    out = []
    for (_, x, y, tw, th) in words[:n]:
        out.append(np.ascontiguousarray(gray[max(y - 4, 0):y + th + 4, max(x - 6, 0):x + tw + 6]))
    return out


@pytest.mark.parametrize("ckpt", ["trained", "fp32"])
@pytest.mark.parametrize("head", ["CTC", "Attention"])
@pytest.mark.parametrize("prec", ["exact", "fast"])
def test_crnn_probability_gate(oracle_mods, prec, head, ckpt):
    """The float gate on a scale-free quantity: what CRNN.process consumes is softmax(preds) (reference ocr/net.py:177-190),
    so the CUDA path's per-step class probabilities are compared with the fp32 oracle's, max-abs over all crops, steps
    and classes, on BOTH synthetic checkpoints (`trained`: conditioned on 16-bit storage during training; `fp32`:
    trained the plain way, never saw this repository's rounding), on 40 ragged crops + 80 word crops.
      exact (LOCR_PREC_EXACT, split-precision recogniser): probabilities within 1e-2 - the north-star tolerance;
      fast  (one 16-bit tensor-core pass per layer): 16-bit operand rounding accumulated over ~40 layers moves trained-
            scale logits (|logit| up to ~45) by ~0.1-0.2, i.e. a probability near a tie by several 1e-2: bounded here
            at 1e-1 and reported; arg-max agreement >= 99.5 %.
    Attention: greedy feedback makes later steps depend on earlier decisions, so the gate is applied to the steps up
    to the first disagreement of the arg-max sequence (all steps when the sequences agree)."""
    ocr_ref, receipts, weights = oracle_mods
    from lightly_ocr_b200 import bridge
    if ckpt == "fp32" and not weights.has_fp32_checkpoint(head):
        pytest.skip("fp32-trained checkpoint not generated")
    sd = weights.crnn_calibrated(1, head, trained="fp32" if ckpt == "fp32" else True)
    eng = bridge.Engine(act_dtype=ACT["f16"], head=head,
                        precision=bridge.PREC_EXACT if prec == "exact" else bridge.PREC_FAST)
    eng.load_state_dict(bridge.MODEL_CRNN, sd)
    crops = [np.random.default_rng(0).integers(0, 256, (32, 100), dtype=np.uint8)] + receipts.crops(39, seed=3) + \
        _word_crops(receipts, 30, 80)
    u8 = np.stack([ocr_ref.crop_to_tensor(g)[0] for g in crops])
    out = eng.crnn_on_resized(u8)
    taps = {}
    with torch.no_grad():
        x = torch.cat([ocr_ref.crop_to_tensor(g)[1] for g in crops], 0)
        ref = ocr_ref.crnn_forward(sd, x, head, taps)
    got = torch.from_numpy(out["logits"])
    vis = eng.debug_read("visual")
    vref = taps["visual"].numpy().reshape(vis.shape)
    ctx = eng.debug_read("contextual")
    cref = taps["contextual"].numpy().reshape(ctx.shape)
    valid = torch.ones(got.shape[:2], dtype=torch.bool)
    if head == "Attention":
        same = (got.argmax(2) == ref.argmax(2))
        first_bad = torch.where(same.all(1), torch.full((len(crops),), 26), (~same).float().argmax(1))
        valid = torch.arange(26)[None, :] <= first_bad[:, None]
    pg, pr = torch.softmax(got, 2), torch.softmax(ref, 2)
    perr = float(((pg - pr).abs().max(2)[0] * valid).max())
    lerr = float(((got - ref).abs().max(2)[0] * valid).max())
    agree = float((got.argmax(2) == ref.argmax(2))[valid].float().mean())
    print("%s/%s/%s: probabilities max-abs err %.4g, logits max-abs err %.4g (|logit| max %.1f, std %.2f), visual rel "
          "%.3g, contextual rel %.3g, arg-max agreement %.5f over %d steps" %
          (prec, head, ckpt, perr, lerr, float(ref.abs().max()), float(ref.std()),
           np.abs(vis - vref).max() / np.abs(vref).max(), np.abs(ctx - cref).max() / np.abs(cref).max(), agree,
           int(valid.sum())))
    if prec == "exact":
        assert perr < 1e-2
        assert agree >= 0.999
    else:
        assert perr < 1e-1
        assert agree >= 0.995
    eng.close()


def test_crnn_seed_only_checkpoint(oracle_mods):
    """The purely seed-generated CRNN checkpoint (`trained=False`: random weights, prediction head along the principal
    directions of the features, SURVEY 8d recipe) through the same CUDA path.  Its top-1 / top-2 margins have mass at
    zero, so the 16-bit storage noise that leaves the trained-like checkpoint's strings untouched flips ~1% of the
    arg-maxes here: the kernels are the same, the gates that depend on decision margins are not reachable on it
    (precision).  Bounds: logit error < 25% of the logit std, arg-max agreement > 98%; intermediates as tight
    as for the default checkpoint."""
    ocr_ref, receipts, weights = oracle_mods
    from lightly_ocr_b200 import bridge
    sd = weights.crnn_calibrated(1, "CTC", trained=False)
    eng = bridge.Engine(act_dtype=ACT["f16"], head="CTC")
    eng.load_state_dict(bridge.MODEL_CRNN, sd)
    crops = [np.random.default_rng(0).integers(0, 256, (32, 100), dtype=np.uint8)] + receipts.crops(39, seed=3)
    u8 = np.stack([ocr_ref.crop_to_tensor(g)[0] for g in crops])
    out = eng.crnn_on_resized(u8)
    taps = {}
    with torch.no_grad():
        x = torch.cat([ocr_ref.crop_to_tensor(g)[1] for g in crops], 0)
        ref = ocr_ref.crnn_forward(sd, x, "CTC", taps).numpy()
    for name, tol in (("fiducials", 1e-4), ("rectified", 5e-3), ("visual", 1e-2), ("contextual", 1e-2)):
        g = eng.debug_read(name)
        r = taps[name].numpy().reshape(g.shape)
        assert np.abs(g - r).max() / max(np.abs(r).max(), 1e-6) < tol, name
    err = np.abs(out["logits"] - ref).max()
    agree = (out["ids"] == ref.argmax(2)).mean()
    same = np.mean([out["text"][i] == ocr_ref.ctc_decode(ref[i].argmax(1)) for i in range(len(crops))])
    print("seed-only checkpoint: logits max-abs err %.4g (std %.3f), argmax agreement %.4f, string agreement %.3f" %
          (err, ref.std(), agree, same))
    assert err < 0.25 * max(1.0, float(ref.std())) and agree > 0.98
    eng.close()
