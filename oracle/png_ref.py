"""TEST INFRASTRUCTURE - CPU restatement of `cv2.imdecode(buf, cv2.IMREAD_COLOR)` for PNG files.

The reference reads its input with `cv2.imread` (ocr/pipeline.py:68) and its server accepts `.png` uploads
(ocr/server.py:11).  OpenCV 4.13 hands PNG files to libpng 1.6 (grfmt_png.cpp); neither lives under /root/reference, so
the published format (PNG specification, 2nd edition / RFC 2083 + RFC 1950/1951 through Python's zlib) is restated here
with numpy, together with the transformations OpenCV requests for IMREAD_COLOR:

    16-bit samples      -> the high byte                      (png_set_strip_16)
    palette             -> RGB through PLTE                    (png_set_palette_to_rgb)
    gray 1 / 2 / 4 bit  -> scaled to 8 bit (x255, x85, x17)    (png_set_expand_gray_1_2_4_to_8)
    alpha / tRNS        -> dropped, no blending                (png_set_strip_alpha)
    gray                -> replicated to three channels        (png_set_gray_to_rgb)
    RGB                 -> BGR                                 (png_set_bgr)
    gAMA / sRGB / iCCP  -> ignored (OpenCV does not ask libpng for gamma correction)

Pinned against the live cv2.imdecode by tests/test_png_oracle.py.  Only tests/, __graft_entry__.smoke() and bench.py's
CPU legs may import this module; the product path is lightly_ocr_b200/csrc/png.cu.
"""
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89PNG\r\n\x1a\n"
CHANNELS = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}
# Adam7: (x start, y start, x step, y step) per pass
ADAM7 = ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2))


class PngError(ValueError):
    pass


def parse(data):
    """Chunk walk: returns dict(width, height, depth, color, interlace, plte, trns, idat)."""
    data = bytes(data)
    if data[:8] != SIGNATURE:
        raise PngError("not a PNG file")
    pos = 8
    hdr = None
    plte = trns = None
    idat = []
    seen_iend = False
    while pos + 8 <= len(data):
        n, kind = struct.unpack(">I4s", data[pos:pos + 8])
        if pos + 12 + n > len(data):
            raise PngError("truncated chunk")
        body = data[pos + 8:pos + 8 + n]
        crc = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0]
        critical = not (kind[0] & 0x20)
        if critical and zlib.crc32(kind + body) & 0xffffffff != crc:
            raise PngError("CRC mismatch in %r" % kind)
        pos += 12 + n
        if kind == b"IHDR":
            w, h, depth, color, comp, flt, inter = struct.unpack(">IIBBBBB", body)
            if comp != 0 or flt != 0 or inter > 1 or color not in CHANNELS or w == 0 or h == 0:
                raise PngError("bad IHDR")
            if depth not in {0: (1, 2, 4, 8, 16), 2: (8, 16), 3: (1, 2, 4, 8), 4: (8, 16), 6: (8, 16)}[color]:
                raise PngError("bad bit depth")
            hdr = dict(width=w, height=h, depth=depth, color=color, interlace=inter)
        elif hdr is None:
            raise PngError("IHDR is not the first chunk")
        elif kind == b"PLTE":
            if n % 3 or n == 0 or n > 768:
                raise PngError("bad PLTE")
            plte = np.frombuffer(body, np.uint8).reshape(-1, 3)
        elif kind == b"tRNS":
            trns = body
        elif kind == b"IDAT":
            idat.append(body)
        elif kind == b"IEND":
            seen_iend = True
            break
    if hdr is None or not idat:
        raise PngError("missing IHDR / IDAT")
    if not seen_iend:
        raise PngError("missing IEND (truncated file: cv2.imdecode returns None)")
    if hdr["color"] == 3 and plte is None:
        raise PngError("palette image without PLTE")
    hdr.update(plte=plte, trns=trns, idat=b"".join(idat), iend=seen_iend)
    return hdr


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    if pa <= pb and pa <= pc:
        return a
    return b if pb <= pc else c


def unfilter(raw, height, rowbytes, bpp):
    """Filtered scanlines [height][1 + rowbytes] -> uint8 [height][rowbytes] (PNG spec, section 9)."""
    out = np.zeros((height, rowbytes), np.uint8)
    prev = np.zeros(rowbytes, np.int32)
    for y in range(height):
        ft = raw[y * (rowbytes + 1)]
        line = np.frombuffer(raw, np.uint8, rowbytes, y * (rowbytes + 1) + 1).astype(np.int32)
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        elif ft == 1:
            cur = line.copy()
            for i in range(bpp, rowbytes):
                cur[i] = (cur[i] + cur[i - bpp]) & 255
        elif ft == 3:
            cur = line.copy()
            for i in range(rowbytes):
                left = cur[i - bpp] if i >= bpp else 0
                cur[i] = (cur[i] + ((left + prev[i]) >> 1)) & 255
        elif ft == 4:
            cur = line.copy()
            for i in range(rowbytes):
                a = int(cur[i - bpp]) if i >= bpp else 0
                c = int(prev[i - bpp]) if i >= bpp else 0
                cur[i] = (cur[i] + _paeth(a, int(prev[i]), c)) & 255
        else:
            raise PngError("bad filter type %d" % ft)
        out[y] = cur
        prev = cur
    return out


def _samples(rows, width, depth, channels):
    """Unfiltered scanline bytes -> uint16 samples [height][width][channels]."""
    h = rows.shape[0]
    if depth == 8:
        return rows[:, :width * channels].reshape(h, width, channels).astype(np.uint16)
    if depth == 16:
        b = rows[:, :width * channels * 2].reshape(h, width, channels, 2).astype(np.uint16)
        return (b[..., 0] << 8) | b[..., 1]
    per = 8 // depth
    bits = np.unpackbits(rows, axis=1).reshape(h, -1, depth)
    vals = np.zeros(bits.shape[:2], np.uint16)
    for k in range(depth):
        vals = (vals << 1) | bits[..., k]
    assert vals.shape[1] >= width and per
    return vals[:, :width].reshape(h, width, 1)


def decode_samples(hdr):
    """uint16 samples [H][W][channels] of the image (both interlace methods)."""
    w, h, depth, color = hdr["width"], hdr["height"], hdr["depth"], hdr["color"]
    ch = CHANNELS[color]
    bits_pp = depth * ch
    bpp = max(1, bits_pp // 8)
    try:
        d = zlib.decompressobj()
        raw = d.decompress(hdr["idat"])
    except zlib.error as e:
        raise PngError("inflate failed: %s" % e)
    if hdr["interlace"] == 0:
        rowbytes = (w * bits_pp + 7) // 8
        if len(raw) < h * (rowbytes + 1):
            raise PngError("not enough image data")
        return _samples(unfilter(raw, h, rowbytes, bpp), w, depth, ch)
    out = np.zeros((h, w, ch), np.uint16)
    pos = 0
    for xs, ys, dx, dy in ADAM7:
        pw, ph = (w - xs + dx - 1) // dx, (h - ys + dy - 1) // dy
        if pw <= 0 or ph <= 0:
            continue
        rowbytes = (pw * bits_pp + 7) // 8
        need = ph * (rowbytes + 1)
        if len(raw) < pos + need:
            raise PngError("not enough image data")
        out[ys::dy, xs::dx] = _samples(unfilter(raw[pos:pos + need], ph, rowbytes, bpp), pw, depth, ch)
        pos += need
    return out


def imdecode(data):
    """cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR): uint8 [H][W][3] BGR."""
    hdr = parse(data)
    s = decode_samples(hdr)
    depth, color = hdr["depth"], hdr["color"]
    if color == 3:
        idx = s[..., 0]
        if int(idx.max()) >= len(hdr["plte"]):
            raise PngError("palette index out of range")
        rgb = hdr["plte"][idx]
    else:
        if depth == 16:
            s = s >> 8
        elif depth < 8:
            s = s * (255 // ((1 << depth) - 1))
        s = s.astype(np.uint8)
        rgb = np.repeat(s[..., :1], 3, axis=2) if color in (0, 4) else s[..., :3]
    return np.ascontiguousarray(rgb[..., ::-1])
