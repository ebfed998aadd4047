"""Edge cases of the product entry points through the C ABI: images without text, ragged batches, tiny images, empty and
out-of-range crops, capacity / argument errors, determinism, and the batched path == the one-image-at-a-time path the
reference runs (ocr/pipeline.py:65-87 is strictly one image, one crop at a time)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def runner():
    from lightly_ocr_b200 import bridge
    from lightly_ocr_b200.synth import weights
    r = bridge.OcrRunner(device_id=0, act_dtype=bridge.ACT_F16, head="CTC")
    r.load_state_dict(bridge.MODEL_CRAFT, weights.craft_calibrated(0, ink=True))
    r.load_state_dict(bridge.MODEL_CRNN, weights.crnn_calibrated(1, "CTC"))
    yield r
    r.close()


def test_blank_image_has_no_boxes(runner):
    white = np.full((1280, 960, 3), 255, np.uint8)     # canvas-sized: no padding strip (which the ink path would see)
    per_image, out = runner.ocr([white])
    assert per_image == [[]] and out["text"] == []
    tiny = np.full((96, 64, 3), 255, np.uint8)         # 144 x 96 after the 1.5x magnification, canvas 160 x 96
    per_image, out = runner.ocr([tiny, white])
    assert len(per_image) == 2 and per_image[1] == []


def test_ragged_batch_equals_one_image_at_a_time(runner):
    from lightly_ocr_b200.synth import receipts
    full = receipts.receipt(5)
    imgs = [full, np.ascontiguousarray(full[:640, :480]), np.ascontiguousarray(full[100:400, 50:550]),
            receipts.receipt(6), np.ascontiguousarray(full[:333, :777])]
    per_image, out = runner.ocr(imgs, want_logits=True)
    texts, k = [], 0
    for rects in per_image:
        texts.append(out["text"][k:k + len(rects)])
        k += len(rects)
    for img, rects, tx in zip(imgs, per_image, texts):
        one_rects, one = runner.ocr([img])
        assert one_rects[0] == rects
        assert one["text"] == tx
    assert sum(len(r) for r in per_image) == len(out["text"]) > 100


def test_deterministic(runner):
    from lightly_ocr_b200.synth import receipts
    imgs = [receipts.receipt(7), receipts.receipt(8)]
    a = runner.ocr(imgs, want_logits=True)
    b = runner.ocr(imgs, want_logits=True)
    assert a[0] == b[0] and a[1]["text"] == b[1]["text"]
    assert np.array_equal(a[1]["logits"], b[1]["logits"]) and np.array_equal(a[1]["conf"], b[1]["conf"])


def test_crops_of_extreme_shapes(runner):
    from oracle import ocr_ref
    rng = np.random.default_rng(0)
    crops = [rng.integers(0, 256, s, dtype=np.uint8) for s in ((1, 1), (1, 300), (200, 1), (5, 7), (32, 100), (64, 600))]
    out = runner.recognize(crops)
    assert len(out["text"]) == len(crops) and np.isfinite(out["logits"]).all()
    # the resize stage stays byte-exact against PIL for these shapes (checked through the debug tap of the last call)
    u8 = runner.debug_read("crop_u8").astype(np.uint8)
    for i, g in enumerate(crops):
        assert np.array_equal(u8[i], ocr_ref.crop_to_tensor(g)[0]), crops[i].shape


def test_empty_and_out_of_range_boxes(runner):
    from lightly_ocr_b200.synth import receipts
    img = receipts.receipt(9)
    rects, _, _ = runner.detect([img])
    good = rects[0][0].tolist()
    boxes = [good, [10, 10, 10, 40], [-5, -5, 30, 80], [1270, 900, 5000, 5000], good]
    out = runner.recognize_boxes([0] * len(boxes), boxes)
    assert out["has_eos"][1] == -2                      # empty slice: the reference's cv2.cvtColor raises on it
    assert out["text"][0] == out["text"][4] and out["has_eos"][0] == 1
    # numpy slice semantics (net.py:109-111): image[-5:30, -5:80] is EMPTY (negative starts count from the end),
    # image[1270:5000, 900:5000] is clipped to the 10 x 60 corner
    assert out["has_eos"][2] == -2 and out["has_eos"][3] == 1
    assert img[-5:30, -5:80].size == 0 and img[1270:5000, 900:5000].shape[:2] == (10, 60)


def test_argument_and_capacity_errors(runner):
    from lightly_ocr_b200 import bridge
    from lightly_ocr_b200.synth import receipts
    L = runner.L
    img = receipts.receipt(10)
    rects = np.empty((4, 4), np.int32)
    counts = np.zeros(1, np.int32)
    ptrs = (C.c_void_p * 1)(img.ctypes.data)
    hs, ws = (C.c_int * 1)(img.shape[0]), (C.c_int * 1)(img.shape[1])
    rc = L.locr_detect(runner.h, ptrs, hs, ws, None, 1, 4, rects.ctypes.data_as(C.c_void_p), None,
                       counts.ctypes.data_as(C.c_void_p), None)
    assert rc == -4 and b"max_boxes_total" in L.locr_last_error(runner.h)
    assert L.locr_detect(runner.h, ptrs, hs, ws, None, 0, 4, rects.ctypes.data_as(C.c_void_p), None,
                         counts.ctypes.data_as(C.c_void_p), None) == -1
    assert L.locr_detect(None, ptrs, hs, ws, None, 1, 4, rects.ctypes.data_as(C.c_void_p), None,
                         counts.ctypes.data_as(C.c_void_p), None) == -1
    with pytest.raises(bridge.LocrError):
        runner.recognize_boxes([3], [[0, 0, 10, 10]])    # image index 3 is not resident
    with pytest.raises(bridge.LocrError):
        runner.recognize([np.zeros((4, 4, 2), np.uint8)])  # 2 channels
    # the handle stays usable after errors
    per_image, out = runner.ocr([img])
    assert len(out["text"]) > 50


def test_unfinalized_model_is_an_error():
    from lightly_ocr_b200 import bridge
    e = bridge.OcrRunner(device_id=0, act_dtype=bridge.ACT_F16, head="CTC")
    with pytest.raises(bridge.LocrError):
        e.ocr([np.zeros((64, 64, 3), np.uint8)])
    with pytest.raises(bridge.LocrError):
        e.load_state_dict(bridge.MODEL_CRNN, {"Prediction.weight": np.zeros((37, 256), np.float32)})
    e.close()


def test_baseline_config4_full_size_lanes_equal_one_at_a_time(runner):
    """BASELINE config 4 at its full size - 256 distinct 1280x960 receipts - the way bench.py issues them (three host
    lanes with their own handle / stream on one GPU, 8 receipts per call, all lanes in flight at once) against the
    one-image-at-a-time order of the reference (ocr/pipeline.py:65-87): rects, their order and every string must be
    identical, confidences equal to float rounding; and a second run of the lanes reproduces the first bit for bit."""
    from concurrent.futures import ThreadPoolExecutor
    from lightly_ocr_b200 import bridge
    from lightly_ocr_b200.synth import receipts, weights
    images = [receipts.receipt(2000 + i) for i in range(256)]
    craft_sd, crnn_sd = weights.craft_calibrated(0, ink=True), weights.crnn_calibrated(1, "CTC")
    lanes = []
    for _ in range(3):
        r = bridge.OcrRunner(device_id=0, act_dtype=bridge.ACT_F16, head="CTC")
        r.load_state_dict(bridge.MODEL_CRAFT, craft_sd)
        r.load_state_dict(bridge.MODEL_CRNN, crnn_sd)
        lanes.append(r)
    batches = [images[i:i + 8] for i in range(0, len(images), 8)]

    def run_lanes():
        def work(k):
            return [lanes[k].ocr(b) for b in batches[k::3]]
        with ThreadPoolExecutor(max_workers=3) as ex:
            parts = list(ex.map(work, range(3)))
        outs = [None] * len(batches)
        for k in range(3):
            outs[k::3] = parts[k]
        return outs

    first, second = run_lanes(), run_lanes()
    n_crops = 0
    for bi, ((per_image, out), (per_image2, out2)) in enumerate(zip(first, second)):
        assert per_image == per_image2 and out["text"] == out2["text"] and np.array_equal(out["conf"], out2["conf"])
        k = 0
        for img, rects in zip(batches[bi], per_image):
            one_rects, one = runner.ocr([img])
            assert one_rects[0] == rects
            assert one["text"] == out["text"][k:k + len(rects)]
            np.testing.assert_allclose(one["conf"], out["conf"][k:k + len(rects)], rtol=1e-4, atol=1e-6)
            k += len(rects)
        assert k == len(out["text"])
        n_crops += k
    assert n_crops > 256 * 50
    for r in lanes:
        r.close()
