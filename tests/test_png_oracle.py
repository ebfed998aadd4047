"""CPU tests of the PNG reader (`cv2.imread` of reference ocr/pipeline.py:68 for the `.png` uploads
ocr/server.py:11 accepts):
* the oracle restatement (oracle/png_ref.py) is pinned against the live cv2.imdecode, byte for byte, over every colour
  type / bit depth / filter / interlace combination, files written by OpenCV and by Pillow, and hand-assembled files;
* the host half of the product path (chunk walk + CRC + zlib inflate in liblocr, no GPU involved) yields the same
  filtered scanlines as Python's zlib;
* files outside the covered subset fail loudly instead of decoding differently from OpenCV."""
import io
import struct
import zlib

import cv2
import numpy as np
import pytest

from oracle import png_ref


def chunk(kind, body):
    return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xffffffff)


def assemble(w, h, depth, color, raw, extra=b"", interlace=0, level=6, split=0):
    """A PNG file around the given filtered scanline bytes."""
    z = zlib.compress(bytes(raw), level)
    idat = chunk(b"IDAT", z) if not split else b"".join(chunk(b"IDAT", z[i:i + split]) for i in range(0, len(z), split))
    return (png_ref.SIGNATURE + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color, 0, 0, interlace)) + extra +
            idat + chunk(b"IEND", b""))


def random_scanlines(rng, w, h, depth, color, interlace, filters=(0, 1, 2, 3, 4)):
    """Random filtered scanlines (every byte pattern is a valid image) with a random filter type per line."""
    bits = depth * png_ref.CHANNELS[color]
    passes = png_ref.ADAM7 if interlace else ((0, 0, 1, 1),)
    raw = bytearray()
    for xs, ys, dx, dy in passes:
        pw, ph = (w - xs + dx - 1) // dx, (h - ys + dy - 1) // dy
        if pw <= 0 or ph <= 0:
            continue
        rb = (pw * bits + 7) // 8
        for _ in range(ph):
            raw += bytes([int(rng.choice(filters))]) + rng.integers(0, 256, rb, dtype=np.uint8).tobytes()
    return raw


def hand_made_cases():
    rng = np.random.default_rng(11)
    pal = chunk(b"PLTE", rng.integers(0, 256, 768, dtype=np.uint8).tobytes())
    for color, depths in ((0, (1, 2, 4, 8, 16)), (2, (8, 16)), (3, (1, 2, 4, 8)), (4, (8, 16)), (6, (8, 16))):
        for depth in depths:
            for interlace in (0, 1):
                for (w, h) in ((1, 1), (3, 2), (21, 13), (64, 9)):
                    raw = random_scanlines(rng, w, h, depth, color, interlace)
                    extra = pal if color == 3 else b""
                    yield ("c%d d%d i%d %dx%d" % (color, depth, interlace, w, h),
                           assemble(w, h, depth, color, raw, extra, interlace))
    raw = random_scanlines(rng, 33, 17, 8, 2, 0)
    yield "gAMA", assemble(33, 17, 8, 2, raw, chunk(b"gAMA", struct.pack(">I", 45455)))
    yield "sRGB+bKGD", assemble(33, 17, 8, 2, raw, chunk(b"sRGB", b"\0") + chunk(b"bKGD", struct.pack(">HHH", 1, 2, 3)))
    yield "tRNS rgb", assemble(33, 17, 8, 2, raw, chunk(b"tRNS", struct.pack(">HHH", 5, 6, 7)))
    yield "split IDAT", assemble(33, 17, 8, 2, raw, split=7)
    yield "stored blocks", assemble(33, 17, 8, 2, raw, level=0)
    yield "short palette", assemble(9, 5, 8, 3, b"".join(b"\0" + bytes([i % 7] * 9) for i in range(5)),
                                    chunk(b"PLTE", bytes(range(21))))
    yield "palette tRNS", assemble(9, 5, 4, 3, b"".join(b"\0" + bytes([0x01, 0x23, 0x45, 0x60, 0x12]) for _ in range(5)),
                                   chunk(b"PLTE", bytes(range(24))) + chunk(b"tRNS", bytes([0, 128, 255])))


def library_cases():
    from PIL import Image
    rng = np.random.default_rng(5)
    from lightly_ocr_b200.synth import receipts
    imgs = {"noise": rng.integers(0, 256, (37, 53, 3), dtype=np.uint8),
            "smooth": cv2.GaussianBlur(rng.integers(0, 256, (61, 83, 3), dtype=np.uint8), (0, 0), 2),
            "receipt": np.ascontiguousarray(receipts.receipt(0)[100:260, 60:310])}
    for nm, arr in imgs.items():
        yield "cv2 bgr8 " + nm, cv2.imencode(".png", arr)[1].tobytes()
        yield "cv2 gray8 " + nm, cv2.imencode(".png", arr[..., 0])[1].tobytes()
        yield "cv2 bgra8 " + nm, cv2.imencode(".png", np.dstack([arr, arr[..., :1]]))[1].tobytes()
        yield "cv2 bgr16 " + nm, cv2.imencode(".png", arr.astype(np.uint16) * 257 + 13)[1].tobytes()
        yield "cv2 gray16 " + nm, cv2.imencode(".png", arr[..., 0].astype(np.uint16) * 201 + 7)[1].tobytes()
        for lvl in (0, 1, 9):
            yield "cv2 level %d %s" % (lvl, nm), cv2.imencode(".png", arr, [cv2.IMWRITE_PNG_COMPRESSION, lvl])[1].tobytes()
        for st in (cv2.IMWRITE_PNG_STRATEGY_FILTERED, cv2.IMWRITE_PNG_STRATEGY_RLE, cv2.IMWRITE_PNG_STRATEGY_FIXED):
            yield "cv2 strategy %d %s" % (st, nm), cv2.imencode(".png", arr, [cv2.IMWRITE_PNG_STRATEGY, st])[1].tobytes()
        im = Image.fromarray(arr[..., ::-1].copy())

        def pil(i, **kw):
            b = io.BytesIO()
            i.save(b, "PNG", **kw)
            return b.getvalue()

        yield "pil RGB " + nm, pil(im)
        yield "pil RGB optimize " + nm, pil(im, optimize=True)
        yield "pil L " + nm, pil(im.convert("L"))
        yield "pil LA " + nm, pil(im.convert("LA"))
        yield "pil RGBA " + nm, pil(im.convert("RGBA"))
        yield "pil P " + nm, pil(im.convert("P"))
        yield "pil 1 " + nm, pil(im.convert("1"))
        yield "pil I;16 " + nm, pil(Image.fromarray(arr[..., 0].astype(np.uint16) * 255))
        yield "pil P transparency " + nm, pil(im.quantize(16), transparency=3)
        yield "pil L transparency " + nm, pil(im.convert("L"), transparency=100)
        for bits in (1, 2, 4):
            yield "pil P %d-bit %s" % (bits, nm), pil(im.quantize(1 << bits), bits=bits)


def test_oracle_matches_cv2_imdecode():
    n = 0
    for name, data in list(hand_made_cases()) + list(library_cases()):
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        assert ref is not None, name
        got = png_ref.imdecode(data)
        assert got.shape == ref.shape and np.array_equal(got, ref), name
        n += 1
    assert n > 150


def test_library_host_half_matches_zlib():
    """liblocr's chunk walk + inflate (host only) against Python's zlib on the same files, and the header query."""
    from lightly_ocr_b200 import bridge
    for name, data in list(hand_made_cases())[::3] + list(library_cases())[::4]:
        hdr = png_ref.parse(data)
        h, w, c, fmt = bridge.image_info(data)
        assert (h, w, c, fmt) == (hdr["height"], hdr["width"], png_ref.CHANNELS[hdr["color"]], bridge.FORMAT_PNG), name
        raw = zlib.decompress(hdr["idat"])
        got = bridge.png_scanlines(data)
        assert got.tobytes() == raw[:len(got)], name
    ok, jpg = cv2.imencode(".jpg", np.zeros((8, 8, 3), np.uint8))
    assert bridge.image_info(jpg.tobytes())[3] == bridge.FORMAT_JPEG


def test_damaged_and_unsupported_files_fail_loudly():
    """Whatever cv2.imdecode refuses must be refused (the caller then falls back to cv2.imread like the reference);
    nothing may decode to different pixels."""
    from lightly_ocr_b200 import bridge
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (40, 56, 3), dtype=np.uint8)
    good = cv2.imencode(".png", img)[1].tobytes()
    assert np.array_equal(png_ref.imdecode(good), img)

    def refused(data, match=None):
        with pytest.raises(png_ref.PngError):
            png_ref.imdecode(data)
        with pytest.raises(bridge.LocrError, match=match):
            bridge.png_scanlines(data)

    refused(good[:len(good) * 6 // 10], "truncated")
    refused(good[:-12], "IEND")
    assert cv2.imdecode(np.frombuffer(good[:-12], np.uint8), cv2.IMREAD_COLOR) is None
    flipped = bytearray(good)
    flipped[len(flipped) // 2] ^= 0x55
    refused(bytes(flipped), "CRC")
    assert cv2.imdecode(np.frombuffer(bytes(flipped), np.uint8), cv2.IMREAD_COLOR) is None
    # valid CRCs around a corrupt deflate stream
    hdr = png_ref.parse(good)
    z = bytearray(hdr["idat"])
    z[len(z) // 2] ^= 0xff
    broken = (png_ref.SIGNATURE + chunk(b"IHDR", struct.pack(">IIBBBBB", 56, 40, 8, 2, 0, 0, 0)) + chunk(b"IDAT", bytes(z)) +
              chunk(b"IEND", b""))
    with pytest.raises(bridge.LocrError):
        bridge.png_scanlines(broken)
    # a tiny file that declares a huge image: refused from the header alone
    huge = png_ref.SIGNATURE + chunk(b"IHDR", struct.pack(">IIBBBBB", 30000, 30000, 8, 2, 0, 0, 0)) + \
        chunk(b"IDAT", zlib.compress(b"\0" * 100)) + chunk(b"IEND", b"")
    with pytest.raises(bridge.LocrError, match="not enough image data"):
        bridge.image_info(huge)
    toobig = png_ref.SIGNATURE + chunk(b"IHDR", struct.pack(">IIBBBBB", 40000, 40000, 8, 2, 0, 0, 0)) + \
        chunk(b"IDAT", zlib.compress(b"\0" * 100)) + chunk(b"IEND", b"")
    with pytest.raises(bridge.LocrError, match="pixel limit"):
        bridge.image_info(toobig)
    # animated PNG: OpenCV 4.13 has its own APNG reader; this one hands the file back
    apng = good[:33] + chunk(b"acTL", struct.pack(">II", 1, 0)) + good[33:]
    with pytest.raises(bridge.LocrError, match="animated"):
        bridge.image_info(apng)
    with pytest.raises(bridge.LocrError, match="not a PNG|no SOI|not a JPEG"):
        bridge.image_info(b"GIF89a" + b"\0" * 64)
    # random corruption never crashes the host half: every outcome is scanlines or a LocrError
    for k in range(300):
        d = bytearray(good)
        for _ in range(int(rng.integers(1, 4))):
            d[int(rng.integers(8, len(d)))] = int(rng.integers(0, 256))
        try:
            bridge.png_scanlines(bytes(d))
        except bridge.LocrError:
            pass
