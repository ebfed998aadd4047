"""GPU PNG reader (locr_imdecode / locr_detect_encoded on PNG files) against the live cv2.imdecode: every pixel
identical over all colour types, bit depths, filters and both interlace methods; the fused encoded path gives the same
rects / strings as decoding with OpenCV first, also for batches that mix JPEG and PNG files."""
import struct
import zlib

import cv2
import numpy as np
import pytest

from test_png_oracle import assemble, chunk, hand_made_cases, library_cases, random_scanlines

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pipe():
    from lightly_ocr_b200 import bridge
    p = bridge.OcrRunner(device_id=0, act_dtype=bridge.ACT_F16, head="CTC")
    yield p
    p.close()


def test_imdecode_matches_cv2(pipe):
    n = 0
    for name, data in list(hand_made_cases()) + list(library_cases()):
        want = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        got = pipe.imdecode(data)
        assert got.shape == want.shape, name
        assert np.array_equal(got, want), "%s: %d bytes differ from cv2.imdecode" % (name, int((got != want).sum()))
        n += 1
    assert n > 150


def test_tall_images_and_every_filter_across_sweeps(pipe):
    """More than 1024 scanlines (several sweeps of the un-filter kernel, the first thread of a sweep reads the previous
    sweep's last line) with random filter types per line, and the BASELINE-size receipt as OpenCV writes it."""
    from lightly_ocr_b200.synth import receipts
    rng = np.random.default_rng(9)
    for (w, h, depth, color) in ((7, 2500, 8, 2), (40, 1100, 16, 6), (33, 2049, 1, 0), (1300, 3, 8, 6)):
        data = assemble(w, h, depth, color, random_scanlines(rng, w, h, depth, color, 0))
        want = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        assert np.array_equal(pipe.imdecode(data), want), (w, h, depth, color)
    img = receipts.receipt(0)
    for params in ([], [cv2.IMWRITE_PNG_COMPRESSION, 1], [cv2.IMWRITE_PNG_STRATEGY, cv2.IMWRITE_PNG_STRATEGY_FILTERED]):
        data = cv2.imencode(".png", img, params)[1].tobytes()
        assert np.array_equal(pipe.imdecode(data), img)


def test_bad_image_data_is_refused(pipe):
    from lightly_ocr_b200 import bridge
    raw = bytearray(random_scanlines(np.random.default_rng(1), 16, 8, 8, 2, 0))
    raw[3 * (16 * 3 + 1)] = 9                                    # filter type 9 on the fourth line
    with pytest.raises(bridge.LocrError, match="filter"):
        pipe.imdecode(assemble(16, 8, 8, 2, raw))
    idx = b"".join(b"\0" + bytes([1, 2, 3, 200]) for _ in range(3))
    with pytest.raises(bridge.LocrError, match="palette"):
        pipe.imdecode(assemble(4, 3, 8, 3, idx, chunk(b"PLTE", bytes(range(30)))))
    assert cv2.imdecode(np.frombuffer(assemble(16, 8, 8, 2, raw), np.uint8), cv2.IMREAD_COLOR) is None


def test_encoded_path_with_mixed_formats_equals_decoded_path(pipe):
    """getText over files: PNG and JPEG uploads in one batch through locr_detect_encoded == the same pixels decoded by
    OpenCV and handed to locr_detect."""
    from lightly_ocr_b200.synth import receipts, weights
    pipe.load_state_dict(0, weights.craft_calibrated(0, ink=True))
    pipe.load_state_dict(1, weights.crnn_calibrated(1, "CTC"))
    imgs = [receipts.receipt(i) for i in range(4)]
    blobs = [cv2.imencode(".png", imgs[0])[1].tobytes(),
             cv2.imencode(".jpg", imgs[1], [cv2.IMWRITE_JPEG_QUALITY, 92])[1].tobytes(),
             cv2.imencode(".png", cv2.cvtColor(imgs[2], cv2.COLOR_BGR2GRAY))[1].tobytes(),
             cv2.imencode(".png", np.ascontiguousarray(imgs[3][:700, :500]), [cv2.IMWRITE_PNG_COMPRESSION, 9])[1].tobytes()]
    decoded = [cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR) for b in blobs]
    per_a, out_a, sizes = pipe.ocr_encoded(blobs)
    per_b, out_b = pipe.ocr(decoded)
    assert sizes == [d.shape[:2] for d in decoded]
    assert per_a == per_b and out_a["text"] == out_b["text"] and len(out_a["text"]) > 200
    assert np.array_equal(out_a["conf"], out_b["conf"])
