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
    ok, good = cv2.imencode(".jpg", img)
    arith = bytearray(good.tobytes())
    arith[arith.index(b"\xff\xc0") + 1] = 0xC9               # SOF9: arithmetic-coded sequential DCT
    with pytest.raises(jpeg_ref.JpegError):
        jpeg_ref.imdecode(bytes(arith))
    with pytest.raises(bridge.LocrError, match="arithmetic"):
        bridge.jpeg_info(bytes(arith))
    ok, png = cv2.imencode(".png", img)
    with pytest.raises(bridge.LocrError, match="not a JPEG"):
        bridge.jpeg_info(png.tobytes())
    ok, good = cv2.imencode(".jpg", img)
    data = good.tobytes()
    with pytest.raises(bridge.LocrError):
        bridge.jpeg_coefficients(data[:len(data) // 3] + b"\xff\xd9")   # truncated inside a marker segment or the scan


def _strip_app0(data):
    """The file without its JFIF APP0 segment."""
    i = data.index(b"\xff\xe0")
    n = (data[i + 2] << 8) | data[i + 3]
    return data[:i] + data[i + 2 + n:]


def test_colour_space_guess_follows_libjpeg():
    """libjpeg's default_decompress_parms: JFIF implies YCbCr; without JFIF an Adobe APP14 marker with transform 0, or
    component ids 'R','G','B', mean untransformed RGB - which this reader refuses (the caller falls back to OpenCV)
    instead of pushing RGB samples through the YCbCr conversion.  APP14 precedes the frame header."""
    from lightly_ocr_b200 import bridge
    img = np.random.default_rng(1).integers(0, 256, (24, 40, 3), dtype=np.uint8)
    ok, good = cv2.imencode(".jpg", img)
    data = good.tobytes()
    assert bridge.jpeg_info(data) == (24, 40, 3)

    def adobe(transform):
        return b"\xff\xee\x00\x0eAdobe\x00\x64\x00\x00\x00\x00" + bytes([transform])

    bare = _strip_app0(data)
    assert bridge.jpeg_info(bare) == (24, 40, 3)                                     # ids 1,2,3, no marker: YCbCr
    with pytest.raises(bridge.LocrError, match="RGB"):
        bridge.jpeg_info(bare[:2] + adobe(0) + bare[2:])
    assert bridge.jpeg_info(bare[:2] + adobe(1) + bare[2:]) == (24, 40, 3)           # Adobe YCbCr
    assert bridge.jpeg_info(data[:2] + adobe(0) + data[2:]) == (24, 40, 3)           # JFIF wins over Adobe
    # what OpenCV (libjpeg) makes of the same files: identical pixels whenever this reader accepts the file
    ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
    for blob in (bare, bare[:2] + adobe(1) + bare[2:], data[:2] + adobe(0) + data[2:]):
        assert np.array_equal(cv2.imdecode(np.frombuffer(blob, np.uint8), cv2.IMREAD_COLOR), ref)
        assert np.array_equal(jpeg_ref.imdecode(blob), ref)
    rgb0 = cv2.imdecode(np.frombuffer(bare[:2] + adobe(0) + bare[2:], np.uint8), cv2.IMREAD_COLOR)
    assert not np.array_equal(rgb0, ref)                                             # libjpeg really treats it as RGB
    sof = bare.index(b"\xff\xc0")
    ids = bytearray(bare)
    ids[sof + 10], ids[sof + 13], ids[sof + 16] = ord("R"), ord("G"), ord("B")
    sos = bytes(ids).index(b"\xff\xda")
    ids[sos + 5], ids[sos + 7], ids[sos + 9] = ord("R"), ord("G"), ord("B")
    with pytest.raises(bridge.LocrError, match="RGB"):
        bridge.jpeg_info(bytes(ids))


def test_header_cannot_demand_unbounded_memory():
    """A tiny file may declare 65535 x 65535 pixels: it is refused from the header alone (pixel limit of cv2.imread,
    and fewer entropy-coded bits than declared blocks), before any buffer is sized from it."""
    from lightly_ocr_b200 import bridge
    img = np.random.default_rng(2).integers(0, 256, (16, 16, 3), dtype=np.uint8)
    ok, good = cv2.imencode(".jpg", img)
    data = bytearray(good.tobytes())
    sof = bytes(data).index(b"\xff\xc0")
    data[sof + 5:sof + 9] = b"\xff\xff\xff\xff"
    with pytest.raises(bridge.LocrError, match="pixel limit"):
        bridge.jpeg_info(bytes(data))
    data[sof + 5:sof + 9] = b"\x40\x00\x40\x00"            # 16384 x 16384: under the limit, but the scan is 200 bytes
    with pytest.raises(bridge.LocrError, match="premature end"):
        bridge.jpeg_info(bytes(data))


def test_exif_orientation_like_cv2():
    """cv2.imread rotates by the EXIF orientation tag; the oracle does the same and the library reports the rotated size."""
    from lightly_ocr_b200 import bridge
    img = cv2.GaussianBlur(np.random.default_rng(0).integers(0, 256, (37, 53, 3), dtype=np.uint8), (0, 0), 1.2)
    ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 90])
    for little in (False, True):
        for o in range(0, 10):
            d = jpeg_ref.with_exif_orientation(buf.tobytes(), o, little)
            want = cv2.imdecode(np.frombuffer(d, np.uint8), cv2.IMREAD_COLOR)
            assert np.array_equal(jpeg_ref.imdecode(d), want), (little, o)
            assert bridge.jpeg_info(d)[:2] == want.shape[:2], (little, o)

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


def progressive_cases():
    rng = np.random.default_rng(5)
    imgs = [cv2.GaussianBlur(rng.integers(0, 256, (67, 93, 3), dtype=np.uint8), (0, 0), 1.0),
            rng.integers(0, 256, (33, 49, 3), dtype=np.uint8), rng.integers(0, 256, (8, 8, 3), dtype=np.uint8),
            rng.integers(0, 256, (17, 5, 3), dtype=np.uint8)]
    for img in imgs:
        for sn, sf in SF.items():
            for q in (30, 90):
                for extra in ([], [cv2.IMWRITE_JPEG_RST_INTERVAL, 3], [cv2.IMWRITE_JPEG_OPTIMIZE, 1]):
                    ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1, cv2.IMWRITE_JPEG_QUALITY, q,
                                                         cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf] + extra)
                    yield "%s/%s/q%d/%s" % (img.shape, sn, q, extra), buf.tobytes()
        ok, buf = cv2.imencode(".jpg", cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
        yield "%s/gray" % (img.shape,), buf.tobytes()
    import io
    from PIL import Image
    pil = Image.fromarray(imgs[0][..., ::-1])
    for kw in (dict(progressive=True, quality=85), dict(progressive=True, quality=60, subsampling=0),
               dict(progressive=True, optimize=True, quality=95, subsampling=1)):
        for mode in ("RGB", "L"):
            bio = io.BytesIO()
            pil.convert(mode).save(bio, "JPEG", **kw)
            yield "pillow/%s/%s" % (sorted(kw.items()), mode), bio.getvalue()


def test_progressive_files():
    """Progressive (SOF2) files - spectral selection, successive approximation, EOB runs, refinement scans: the oracle
    equals cv2.imdecode and the library's entropy decoder equals the oracle's coefficients."""
    from lightly_ocr_b200 import bridge
    n = 0
    for name, data in progressive_cases():
        want = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        assert np.array_equal(jpeg_ref.imdecode(data), want), name
        ref, _ = jpeg_ref.decode_scans(jpeg_ref.parse_scans(data))
        got, _ = bridge.jpeg_coefficients(data)
        assert len(got) == len(ref) and all(np.array_equal(a, b) for a, b in zip(got, ref)), name
        n += 1
    assert n == 130


def test_multiscan_decoder_equals_single_scan_decoder_on_baseline_files():
    """The general scan decoder of the oracle, run on baseline files, gives the coefficients of the baseline decoder."""
    for name, buf in cases():
        data = buf.tobytes()
        a, _ = jpeg_ref.decode_coefficients(jpeg_ref.parse(data))
        b, _ = jpeg_ref.decode_scans(jpeg_ref.parse_scans(data))
        assert all(np.array_equal(x, y) for x, y in zip(a, b)), name


def test_corrupt_huffman_table_is_refused():
    """A DHT segment whose code lengths over-subscribe the code space (found by the ASan fuzz harness, tools/fuzz_jpeg.cu)
    must be refused, not turned into look-up tables."""
    from lightly_ocr_b200 import bridge
    img = np.random.default_rng(0).integers(0, 256, (16, 16, 3), dtype=np.uint8)
    data = bytearray(cv2.imencode(".jpg", img)[1].tobytes())
    p = data.index(b"\xff\xc4")
    assert data[p + 5] == 0 and data[p + 7] >= 3     # standard luminance DC table: no 1-bit code, five 3-bit codes
    data[p + 5] = 3            # three codes of length 1 (the code space has two) ...
    data[p + 7] -= 3           # ... with the symbol count of the segment unchanged
    with pytest.raises(bridge.LocrError, match="Huffman table"):
        bridge.jpeg_info(bytes(data))
    # random corruption never crashes the host decoder: every outcome is a decoded image or a LocrError
    rng = np.random.default_rng(1)
    good = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])[1].tobytes()
    for _ in range(2000):
        d = bytearray(good)
        for _ in range(int(rng.integers(1, 5))):
            d[int(rng.integers(2, len(d)))] = int(rng.integers(0, 256))
        try:
            h, w, _ = bridge.jpeg_info(bytes(d))
            if h * w <= 1 << 20:
                bridge.jpeg_coefficients(bytes(d))
        except bridge.LocrError:
            pass
