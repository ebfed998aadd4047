"""CPU tests of the JPEG reader (`cv2.imread` of reference ocr/pipeline.py:68):
* the oracle restatement (oracle/jpeg_ref.py) is pinned against the live cv2.imdecode, bit for bit;
* the host half of the product path (marker parsing + Huffman entropy decoding in liblocr, no GPU involved) yields the
  same quantised coefficients as the oracle;
* files outside the covered subset fail loudly instead of decoding differently from OpenCV."""
import zlib

import cv2
import numpy as np
import pytest

from oracle import jpeg_ref

SF = {"420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, "444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444,
      "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, "440": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440,
      "411": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411}


def sample_images():
    from lightly_ocr_b200.synth import receipts
    rng = np.random.default_rng(7)
    rec = receipts.receipt(0)
    yield "receipt_window", np.ascontiguousarray(rec[100:260, 60:310])
    yield "noise", rng.integers(0, 256, (67, 93, 3), dtype=np.uint8)
    yield "smooth", cv2.GaussianBlur(rng.integers(0, 256, (120, 161, 3), dtype=np.uint8), (0, 0), 3)
    yield "saturated", (rng.integers(0, 2, (40, 56, 3)) * 255).astype(np.uint8)


def encodings(img):
    for q in (95, 75, 30, 5):
        yield "q%d" % q, [cv2.IMWRITE_JPEG_QUALITY, q]
    for name, sf in SF.items():
        yield "q90_" + name, [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf]
    yield "rst7", [cv2.IMWRITE_JPEG_QUALITY, 85, cv2.IMWRITE_JPEG_RST_INTERVAL, 7]
    yield "rst1", [cv2.IMWRITE_JPEG_QUALITY, 85, cv2.IMWRITE_JPEG_RST_INTERVAL, 1]
    yield "optimize", [cv2.IMWRITE_JPEG_QUALITY, 85, cv2.IMWRITE_JPEG_OPTIMIZE, 1]
    yield "q100_444", [cv2.IMWRITE_JPEG_QUALITY, 100, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SF["444"]]


def cases():
    for iname, img in sample_images():
        for ename, params in encodings(img):
            ok, buf = cv2.imencode(".jpg", img, params)
            assert ok
            yield iname + "/" + ename, buf
        ok, buf = cv2.imencode(".jpg", cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), [cv2.IMWRITE_JPEG_QUALITY, 85])
        yield iname + "/gray", buf


def test_oracle_matches_cv2_imdecode():
    n = 0
    for name, buf in cases():
        want = cv2.imdecode(buf, cv2.IMREAD_COLOR)
        got = jpeg_ref.imdecode(buf.tobytes())
        assert got.shape == want.shape and np.array_equal(got, want), name
        n += 1
    assert n >= 40


def test_oracle_matches_cv2_on_small_and_ragged_sizes():
    rng = np.random.default_rng(1)
    sizes = [(h, w) for h in (1, 2, 3, 7, 8, 9, 15, 16, 17, 33) for w in (1, 2, 3, 4, 5, 8, 9, 16, 17, 31, 32, 33, 49)]
    for h, w in sizes:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for name, sf in SF.items():
            ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 88, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf])
            assert np.array_equal(jpeg_ref.imdecode(buf.tobytes()), cv2.imdecode(buf, cv2.IMREAD_COLOR)), (h, w, name)


def test_library_entropy_decoder_matches_oracle():
    """liblocr's host half (no GPU): same coefficient blocks as the oracle for every sample file."""
    from lightly_ocr_b200 import bridge
    for name, buf in cases():
        data = buf.tobytes()
        info = jpeg_ref.parse(data)
        want, (hmax, vmax, mcux, mcuy) = jpeg_ref.decode_coefficients(info)
        got, geo = bridge.jpeg_coefficients(data)
        assert (geo["height"], geo["width"]) == (info["height"], info["width"]), name
        assert (geo["hmax"], geo["vmax"], geo["mcux"], geo["mcuy"]) == (hmax, vmax, mcux, mcuy), name
        assert len(got) == len(want), name
        for g, w in zip(got, want):
            assert g.shape == w.shape and np.array_equal(g, w), name
        assert bridge.jpeg_info(data)[:2] == (info["height"], info["width"])


def test_library_entropy_decoder_full_receipt():
    from lightly_ocr_b200 import bridge
    from lightly_ocr_b200.synth import receipts
    ok, buf = cv2.imencode(".jpg", receipts.receipt(3), [cv2.IMWRITE_JPEG_QUALITY, 90])
    data = buf.tobytes()
    want, _ = jpeg_ref.decode_coefficients(jpeg_ref.parse(data))
    got, geo = bridge.jpeg_coefficients(data)
    assert (geo["height"], geo["width"]) == (1280, 960)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    # checksum of checksums over the whole file, so that a silent change of the fixture generator shows up too
    assert zlib.crc32(np.concatenate([g.ravel() for g in got]).tobytes()) == zlib.crc32(
        np.concatenate([w.ravel() for w in want]).tobytes())


def test_unsupported_files_fail_loudly():
    from lightly_ocr_b200 import bridge
    img = np.random.default_rng(0).integers(0, 256, (32, 48, 3), dtype=np.uint8)
    ok, prog = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(jpeg_ref.JpegError):
        jpeg_ref.imdecode(prog.tobytes())
    with pytest.raises(bridge.LocrError, match="progressive"):
        bridge.jpeg_info(prog.tobytes())
    ok, png = cv2.imencode(".png", img)
    with pytest.raises(bridge.LocrError, match="not a JPEG"):
        bridge.jpeg_info(png.tobytes())
    ok, good = cv2.imencode(".jpg", img)
    data = good.tobytes()
    with pytest.raises(bridge.LocrError):
        bridge.jpeg_coefficients(data[:len(data) // 3] + b"\xff\xd9")   # truncated inside a marker segment or the scan
    # EXIF orientation 6 (cv2.imread would rotate): refused rather than returned unrotated
    exif = b"Exif\x00\x00MM\x00\x2a\x00\x00\x00\x08\x00\x01\x01\x12\x00\x03\x00\x00\x00\x01\x00\x06\x00\x00\x00\x00\x00\x00"
    seg = b"\xff\xe1" + (len(exif) + 2).to_bytes(2, "big") + exif
    rotated = data[:2] + seg + data[2:]
    assert cv2.imdecode(np.frombuffer(rotated, np.uint8), cv2.IMREAD_COLOR).shape == (48, 32, 3)
    with pytest.raises(bridge.LocrError, match="orientation"):
        bridge.jpeg_info(rotated)
    with pytest.raises(jpeg_ref.JpegError):
        jpeg_ref.imdecode(rotated)


def pillow_cases():
    """Files from a second encoder front end (Pillow): other marker layouts (JFIF density, comments, multi-segment ICC
    profiles), restart intervals in blocks / rows, custom quantisation presets, optimised tables."""
    import io
    from PIL import Image
    rng = np.random.default_rng(3)
    img = cv2.GaussianBlur(rng.integers(0, 256, (97, 131, 3), dtype=np.uint8), (0, 0), 1.5)
    pil = Image.fromarray(img[..., ::-1])
    for kw in (dict(quality=90), dict(quality=50, optimize=True), dict(quality=95, subsampling=0),
               dict(quality=80, subsampling=1), dict(quality=80, subsampling=2), dict(quality=70, restart_marker_blocks=5),
               dict(quality=70, restart_marker_rows=1), dict(quality=85, dpi=(300, 300), comment=b"hello"),
               dict(quality=3), dict(quality=100, subsampling=0), dict(quality=60, qtables="web_low"),
               dict(quality=75, icc_profile=b"x" * 70000)):
        for mode in ("RGB", "L"):
            bio = io.BytesIO()
            pil.convert(mode).save(bio, "JPEG", **kw)
            yield "%s/%s" % (sorted(kw.items()), mode), bio.getvalue()
    bio = io.BytesIO()
    pil.convert("CMYK").save(bio, "JPEG")
    yield "cmyk", bio.getvalue()


def test_pillow_encoded_files():
    from lightly_ocr_b200 import bridge
    n = 0
    for name, data in pillow_cases():
        if name == "cmyk":
            with pytest.raises(bridge.LocrError, match="component"):
                bridge.jpeg_info(data)
            continue
        want = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        assert np.array_equal(jpeg_ref.imdecode(data), want), name
        got, _ = bridge.jpeg_coefficients(data)
        ref, _ = jpeg_ref.decode_coefficients(jpeg_ref.parse(data))
        assert all(np.array_equal(a, b) for a, b in zip(got, ref)), name
        n += 1
    assert n == 24
