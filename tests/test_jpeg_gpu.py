"""GPU JPEG reader (locr_imdecode / locr_detect_encoded) against the live cv2.imdecode and the oracle restatement:
every pixel identical; the fused encoded path gives the same rects / strings as decoding with OpenCV first."""
import cv2
import numpy as np
import pytest

from test_jpeg_oracle import SF, cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pipe():
    from lightly_ocr_b200 import bridge
    p = bridge.Pipeline(device_id=0, act_dtype=bridge.ACT_F16, head="CTC")
    yield p
    p.close()


def test_imdecode_matches_cv2_and_oracle(pipe):
    from oracle import jpeg_ref
    n = 0
    for name, buf in cases():
        want = cv2.imdecode(buf, cv2.IMREAD_COLOR)
        got = pipe.imdecode(buf.tobytes())
        assert got.shape == want.shape, name
        assert np.array_equal(got, want), "%s: %d bytes differ from cv2.imdecode" % (name, int((got != want).sum()))
        assert np.array_equal(got, jpeg_ref.imdecode(buf.tobytes())), name
        n += 1
    assert n >= 40


def test_imdecode_small_and_ragged_sizes(pipe):
    rng = np.random.default_rng(1)
    for h in (1, 2, 3, 7, 8, 9, 15, 16, 17, 33):
        for w in (1, 2, 3, 4, 5, 8, 9, 16, 17, 31, 32, 33, 49):
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            for name, sf in SF.items():
                ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 88, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf])
                assert np.array_equal(pipe.imdecode(buf.tobytes()), cv2.imdecode(buf, cv2.IMREAD_COLOR)), (h, w, name)


@pytest.mark.parametrize("params", [[cv2.IMWRITE_JPEG_QUALITY, 90], [cv2.IMWRITE_JPEG_QUALITY, 60, cv2.IMWRITE_JPEG_RST_INTERVAL, 60],
                                    [cv2.IMWRITE_JPEG_QUALITY, 95, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SF["444"]]], ids=["q90_420", "q60_rst", "q95_444"])
def test_imdecode_full_receipts(pipe, params):
    """BASELINE-size inputs (1280x960 receipts): identical to cv2.imdecode; also a photo-like noisy variant."""
    from lightly_ocr_b200.synth import receipts
    rng = np.random.default_rng(5)
    for seed in (0, 1):
        img = receipts.receipt(seed)
        if seed == 1:
            img = np.clip(img.astype(np.int16) + rng.integers(-20, 21, img.shape), 0, 255).astype(np.uint8)
        ok, buf = cv2.imencode(".jpg", img, params)
        got = pipe.imdecode(buf.tobytes())
        assert np.array_equal(got, cv2.imdecode(buf, cv2.IMREAD_COLOR))


def test_imread_and_errors(pipe, tmp_path):
    from lightly_ocr_b200 import bridge
    img = np.random.default_rng(0).integers(0, 256, (37, 53, 3), dtype=np.uint8)
    path = str(tmp_path / "a.jpg")
    cv2.imwrite(path, img)
    assert np.array_equal(pipe.imread(path), cv2.imread(path))
    ok, png = cv2.imencode(".png", img)
    assert np.array_equal(pipe.imdecode(png.tobytes()), img)          # PNG files take the PNG reader (test_png_gpu.py)
    with pytest.raises(bridge.LocrError, match="not a JPEG"):
        pipe.imdecode(b"GIF89a" + bytes(64))
    ok, good = cv2.imencode(".jpg", img)
    data = good.tobytes()
    with pytest.raises(bridge.LocrError):
        pipe.imdecode(data[:len(data) // 2])
    # the handle keeps working after a refused file
    assert np.array_equal(pipe.imdecode(data), cv2.imdecode(good, cv2.IMREAD_COLOR))


def test_ocr_encoded_equals_decode_then_ocr():
    """getText on JPEG files decoded on the GPU == getText on the same files decoded by OpenCV: same rects, same strings,
    same confidences (the decoded pixels are identical, so everything downstream is)."""
    from lightly_ocr_b200 import bridge
    from lightly_ocr_b200.synth import receipts, weights
    r = bridge.OcrRunner(device_id=0, act_dtype=bridge.ACT_F16, head="CTC")
    r.load_state_dict(bridge.MODEL_CRAFT, weights.craft_calibrated(0, ink=True))
    r.load_state_dict(bridge.MODEL_CRNN, weights.crnn_calibrated(1, "CTC"))
    blobs = []
    for seed in range(3):
        ok, buf = cv2.imencode(".jpg", receipts.receipt(seed), [cv2.IMWRITE_JPEG_QUALITY, 92])
        blobs.append(buf.tobytes())
    ok, buf = cv2.imencode(".jpg", receipts.receipt(3)[:640, :480], [cv2.IMWRITE_JPEG_QUALITY, 80])   # ragged batch
    blobs.append(buf.tobytes())
    per_a, out_a, sizes = r.ocr_encoded(blobs)
    imgs = [cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR) for b in blobs]
    per_b, out_b = r.ocr(imgs)
    assert sizes == [im.shape[:2] for im in imgs]
    assert [list(map(list, p)) for p in per_a] == [list(map(list, p)) for p in per_b]
    assert out_a["text"] == out_b["text"] and len(out_a["text"]) > 150
    assert np.array_equal(out_a["conf"], out_b["conf"])
    r.close()


def test_imdecode_pillow_encoded_files(pipe):
    from test_jpeg_oracle import pillow_cases
    for name, data in pillow_cases():
        if name == "cmyk":
            continue
        assert np.array_equal(pipe.imdecode(data), cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)), name


def test_imdecode_exif_orientations(pipe):
    from oracle import jpeg_ref
    rng = np.random.default_rng(2)
    for shape, sf in (((37, 53, 3), "420"), ((64, 48, 3), "444"), ((21, 80), "gray")):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        params = [cv2.IMWRITE_JPEG_QUALITY, 90] + ([] if sf == "gray" else [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SF[sf]])
        ok, buf = cv2.imencode(".jpg", img, params)
        for o in range(1, 9):
            d = jpeg_ref.with_exif_orientation(buf.tobytes(), o, little_endian=(o % 2 == 0))
            want = cv2.imdecode(np.frombuffer(d, np.uint8), cv2.IMREAD_COLOR)
            got = pipe.imdecode(d)
            assert got.shape == want.shape and np.array_equal(got, want), (shape, o)


def test_imdecode_progressive_files(pipe):
    from test_jpeg_oracle import progressive_cases
    for name, data in progressive_cases():
        assert np.array_equal(pipe.imdecode(data), cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)), name
    from lightly_ocr_b200.synth import receipts
    ok, buf = cv2.imencode(".jpg", receipts.receipt(2), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1, cv2.IMWRITE_JPEG_QUALITY, 90])
    assert np.array_equal(pipe.imdecode(buf.tobytes()), cv2.imdecode(buf, cv2.IMREAD_COLOR))
