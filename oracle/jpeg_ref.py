"""CPU restatement of `cv2.imread` / `cv2.imdecode(..., IMREAD_COLOR)` for baseline JPEG files — TEST INFRASTRUCTURE ONLY
(imported by tests/ only; the product path is lightly_ocr_b200/csrc/jpeg.cu).

The reference reads its input with `cv2.imread(path)` (ocr/pipeline.py:68).  OpenCV hands JPEG files to libjpeg(-turbo),
which is not part of /root/reference; the published algorithm restated here is the one libjpeg runs with its defaults
(the ones OpenCV leaves untouched): Huffman entropy decoding of a baseline sequential scan (ITU-T T.81 annex F),
`jidctint.c` "ISLOW" integer inverse DCT (13-bit constants, 2 extra bits after the column pass), `jdsample.c` "fancy"
triangle up-sampling of the chroma planes (h2v1 / h2v2), and `jdcolor.c` YCbCr -> RGB with 16-bit fixed-point tables.
Everything is integer arithmetic, so parity is bit-exact.

Pinned: tests/test_jpeg_oracle.py compares this file against the live cv2.imdecode on synthetic receipts and noise
images over sizes, qualities, chroma sub-samplings, restart intervals, optimised tables and grayscale files.
"""
import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7,
                   14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39,
                   46, 53, 60, 61, 54, 47, 55, 62, 63], np.int32)


class JpegError(ValueError):
    pass


def parse(data):
    """Marker segments of a baseline (SOF0 / SOF1, 8-bit, Huffman) file with one interleaved scan.
    Returns a dict: height, width, comps [(id, h, v, tq, td, ta)], qt {id: int32[64] natural order},
    dc / ac {id: (bits[16], vals)}, restart_interval, scan (entropy-coded bytes, markers still inside)."""
    data = bytes(data)
    if len(data) < 4 or data[0] != 0xFF or data[1] != 0xD8:
        raise JpegError("not a JPEG file (no SOI)")
    pos = 2
    out = {"qt": {}, "dc": {}, "ac": {}, "restart_interval": 0, "orientation": 1}
    frame = None
    while True:
        if pos + 4 > len(data):
            raise JpegError("truncated file")
        if data[pos] != 0xFF:
            raise JpegError("marker expected at byte %d" % pos)
        while data[pos + 1] == 0xFF:
            pos += 1
        m = data[pos + 1]
        pos += 2
        if m == 0xD9:
            raise JpegError("EOI before any scan")
        if m == 0x01 or 0xD0 <= m <= 0xD7:
            continue
        n = (data[pos] << 8) | data[pos + 1]
        seg = data[pos + 2:pos + n]
        if len(seg) != n - 2:
            raise JpegError("truncated segment")
        pos += n
        if m == 0xDB:
            i = 0
            while i < len(seg):
                pq, tq = seg[i] >> 4, seg[i] & 15
                i += 1
                if pq:
                    vals = [(seg[i + 2 * k] << 8) | seg[i + 2 * k + 1] for k in range(64)]
                    i += 128
                else:
                    vals = list(seg[i:i + 64])
                    i += 64
                q = np.zeros(64, np.int32)
                q[ZIGZAG] = vals
                out["qt"][tq] = q
        elif m in (0xC0, 0xC1):
            if seg[0] != 8:
                raise JpegError("only 8-bit samples are supported")
            out["height"] = (seg[1] << 8) | seg[2]
            out["width"] = (seg[3] << 8) | seg[4]
            nc = seg[5]
            frame = [(seg[6 + 3 * k], seg[7 + 3 * k] >> 4, seg[7 + 3 * k] & 15, seg[8 + 3 * k]) for k in range(nc)]
        elif m in (0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise JpegError("unsupported JPEG process (SOF%d): only baseline sequential Huffman" % (m - 0xC0))
        elif m == 0xC4:
            i = 0
            while i < len(seg):
                tc, th = seg[i] >> 4, seg[i] & 15
                bits = list(seg[i + 1:i + 17])
                cnt = sum(bits)
                vals = list(seg[i + 17:i + 17 + cnt])
                i += 17 + cnt
                (out["ac"] if tc else out["dc"])[th] = (bits, vals)
        elif m == 0xDD:
            out["restart_interval"] = (seg[0] << 8) | seg[1]
        elif m == 0xE1 and seg[:6] == b"Exif\x00\x00":
            out["orientation"] = _exif_orientation(seg[6:])
        elif m == 0xDA:
            if frame is None:
                raise JpegError("SOS before SOF")
            ns = seg[0]
            if ns != len(frame):
                raise JpegError("non-interleaved scans are not supported")
            comps = []
            for k in range(ns):
                cid, tabs = seg[1 + 2 * k], seg[2 + 2 * k]
                f = [c for c in frame if c[0] == cid]
                if not f:
                    raise JpegError("scan component not in frame")
                comps.append((cid, f[0][1], f[0][2], f[0][3], tabs >> 4, tabs & 15))
            out["comps"] = comps
            out["scan"] = data[pos:]
            return out


def _exif_orientation(tiff):
    try:
        le = tiff[:2] == b"II"
        rd = (lambda b: int.from_bytes(b, "little")) if le else (lambda b: int.from_bytes(b, "big"))
        off = rd(tiff[4:8])
        n = rd(tiff[off:off + 2])
        for k in range(n):
            e = tiff[off + 2 + 12 * k:off + 14 + 12 * k]
            if rd(e[0:2]) == 0x0112:
                v = rd(e[8:10])
                return v if 1 <= v <= 8 else 1
    except Exception:
        pass
    return 1


def _huff_lookup(bits, vals):
    """16-bit look-ahead tables (code length, symbol) of one DHT table (T.81 annex C)."""
    length = np.zeros(65536, np.uint8)
    symbol = np.zeros(65536, np.uint8)
    code, k = 0, 0
    for ln in range(1, 17):
        for _ in range(bits[ln - 1]):
            lo = code << (16 - ln)
            hi = (code + 1) << (16 - ln)
            length[lo:hi] = ln
            symbol[lo:hi] = vals[k]
            k += 1
            code += 1
        code <<= 1
    return length, symbol


def decode_coefficients(info):
    """Entropy decoding: returns per component an int16 array [blocks_v][blocks_h][64] of quantised coefficients in
    natural order (padded to whole MCUs), plus the geometry."""
    comps = info["comps"]
    if len(comps) == 1:   # T.81 A.2.2: a single-component scan is not interleaved, its sampling factors do not matter
        comps = info["comps"] = [(comps[0][0], 1, 1) + tuple(comps[0][3:])]
    hmax = max(c[1] for c in comps)
    vmax = max(c[2] for c in comps)
    H, W = info["height"], info["width"]
    mcux = (W + 8 * hmax - 1) // (8 * hmax)
    mcuy = (H + 8 * vmax - 1) // (8 * vmax)
    planes = [np.zeros((mcuy * c[2], mcux * c[1], 64), np.int16) for c in comps]
    dct = {k: _huff_lookup(*v) for k, v in info["dc"].items()}
    act = {k: _huff_lookup(*v) for k, v in info["ac"].items()}
    scan = info["scan"]
    # split into restart segments, remove byte stuffing
    segs, cur, i = [], bytearray(), 0
    while i < len(scan):
        b = scan[i]
        if b != 0xFF:
            cur.append(b)
            i += 1
            continue
        nb = scan[i + 1] if i + 1 < len(scan) else 0xD9
        if nb == 0x00:
            cur.append(0xFF)
            i += 2
        elif 0xD0 <= nb <= 0xD7:
            segs.append(bytes(cur))
            cur = bytearray()
            i += 2
        elif nb == 0xFF:
            i += 1
        else:
            break   # EOI or another marker ends the scan
    segs.append(bytes(cur))
    ri = info["restart_interval"]
    total = mcux * mcuy
    mcu = 0
    for seg in segs:
        if mcu >= total:
            break
        acc = int.from_bytes(seg + b"\x00" * 8, "big")     # zero fill after the data like libjpeg's warning path
        nbits = (len(seg) + 8) * 8
        posb = 0
        pred = [0] * len(comps)
        count = ri if ri else total
        for _ in range(count):
            if mcu >= total:
                break
            my, mx = divmod(mcu, mcux)
            for ci, (cid, ch, cv, tq, td, ta) in enumerate(comps):
                dl, ds = dct[td]
                al, as_ = act[ta]
                for by in range(cv):
                    for bx in range(ch):
                        blk = planes[ci][my * cv + by, mx * ch + bx]
                        peek = (acc >> (nbits - posb - 16)) & 0xFFFF
                        ln = int(dl[peek])
                        if ln == 0:
                            raise JpegError("bad Huffman code")
                        posb += ln
                        s = int(ds[peek])
                        diff = 0
                        if s:
                            v = (acc >> (nbits - posb - s)) & ((1 << s) - 1)
                            posb += s
                            diff = v if v >= (1 << (s - 1)) else v - (1 << s) + 1
                        pred[ci] += diff
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            peek = (acc >> (nbits - posb - 16)) & 0xFFFF
                            ln = int(al[peek])
                            if ln == 0:
                                raise JpegError("bad Huffman code")
                            posb += ln
                            rs = int(as_[peek])
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r == 15:
                                    k += 16
                                    continue
                                break
                            k += r
                            v = (acc >> (nbits - posb - s)) & ((1 << s) - 1)
                            posb += s
                            if k > 63:
                                raise JpegError("coefficient index out of range")
                            blk[ZIGZAG[k]] = v if v >= (1 << (s - 1)) else v - (1 << s) + 1
                            k += 1
                        if posb > len(seg) * 8:
                            # libjpeg would warn and carry on with a partly grey image; restated as an error
                            raise JpegError("premature end of the entropy-coded data")
            mcu += 1
    return planes, (hmax, vmax, mcux, mcuy)


def parse_scans(data):
    """Every marker segment of a Huffman-coded 8-bit file, progressive (SOF2) or sequential, any number of scans.
    Returns the frame description plus `scans`: a list of dicts (comps [(component index, td, ta)], ss, se, ah, al,
    dc / ac tables and restart interval in force, entropy-coded bytes)."""
    data = bytes(data)
    if len(data) < 4 or data[0] != 0xFF or data[1] != 0xD8:
        raise JpegError("not a JPEG file (no SOI)")
    pos = 2
    out = {"qt": {}, "orientation": 1, "scans": [], "progressive": False}
    dc, ac, ri, frame = {}, {}, 0, None
    while pos + 4 <= len(data):
        if data[pos] != 0xFF:
            raise JpegError("marker expected at byte %d" % pos)
        while data[pos + 1] == 0xFF:
            pos += 1
        m = data[pos + 1]
        pos += 2
        if m == 0xD9:
            break
        if m == 0x01 or 0xD0 <= m <= 0xD7:
            continue
        n = (data[pos] << 8) | data[pos + 1]
        seg = data[pos + 2:pos + n]
        if len(seg) != n - 2:
            raise JpegError("truncated segment")
        pos += n
        if m == 0xDB:
            i = 0
            while i < len(seg):
                pq, tq = seg[i] >> 4, seg[i] & 15
                i += 1
                vals = [(seg[i + 2 * k] << 8) | seg[i + 2 * k + 1] for k in range(64)] if pq else list(seg[i:i + 64])
                i += 128 if pq else 64
                q = np.zeros(64, np.int32)
                q[ZIGZAG] = vals
                out["qt"][tq] = q
        elif m in (0xC0, 0xC1, 0xC2):
            if seg[0] != 8:
                raise JpegError("only 8-bit samples are supported")
            out["progressive"] = m == 0xC2
            out["height"] = (seg[1] << 8) | seg[2]
            out["width"] = (seg[3] << 8) | seg[4]
            frame = [(seg[6 + 3 * k], seg[7 + 3 * k] >> 4, seg[7 + 3 * k] & 15, seg[8 + 3 * k]) for k in range(seg[5])]
            if len(frame) == 1:
                frame = [(frame[0][0], 1, 1, frame[0][3])]
            out["frame"] = frame
        elif m in (0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise JpegError("unsupported JPEG process (SOF%d)" % (m - 0xC0))
        elif m == 0xC4:
            i = 0
            while i < len(seg):
                tc, th = seg[i] >> 4, seg[i] & 15
                bits = list(seg[i + 1:i + 17])
                cnt = sum(bits)
                (ac if tc else dc)[th] = (bits, list(seg[i + 17:i + 17 + cnt]))
                i += 17 + cnt
        elif m == 0xDD:
            ri = (seg[0] << 8) | seg[1]
        elif m == 0xE1 and seg[:6] == b"Exif\x00\x00":
            out["orientation"] = _exif_orientation(seg[6:])
        elif m == 0xDA:
            if frame is None:
                raise JpegError("SOS before SOF")
            ns = seg[0]
            comps = []
            for k in range(ns):
                idx = [j for j, c in enumerate(frame) if c[0] == seg[1 + 2 * k]]
                if not idx:
                    raise JpegError("scan component not in frame")
                comps.append((idx[0], seg[2 + 2 * k] >> 4, seg[2 + 2 * k] & 15))
            ss, se, a = seg[1 + 2 * ns], seg[2 + 2 * ns], seg[3 + 2 * ns]
            end = pos
            while end + 1 < len(data):     # the entropy-coded segment runs up to the next real marker
                if data[end] == 0xFF and data[end + 1] != 0x00 and not (0xD0 <= data[end + 1] <= 0xD7) and data[end + 1] != 0xFF:
                    break
                end += 1
            else:
                end = len(data)
            out["scans"].append(dict(comps=comps, ss=ss, se=se, ah=a >> 4, al=a & 15, dc=dict(dc), ac=dict(ac), ri=ri,
                                     data=data[pos:end]))
            pos = end
    if frame is None or not out["scans"]:
        raise JpegError("no image data")
    return out


class _Bits:
    """Bit reader over one restart segment (byte stuffing already removed); zero bits after the end."""

    def __init__(self, seg):
        self.acc = int.from_bytes(seg + b"\x00" * 8, "big")
        self.n = (len(seg) + 8) * 8
        self.pos = 0
        self.limit = len(seg) * 8

    def bits(self, k):
        if k == 0:
            return 0
        v = (self.acc >> (self.n - self.pos - k)) & ((1 << k) - 1)
        self.pos += k
        return v

    def symbol(self, table):
        length, sym = table
        peek = (self.acc >> (self.n - self.pos - 16)) & 0xFFFF
        ln = int(length[peek])
        if ln == 0:
            raise JpegError("bad Huffman code")
        self.pos += ln
        return int(sym[peek])


def _extend(v, s):
    return v if s == 0 or v >= (1 << (s - 1)) else v - (1 << s) + 1


def _split_restarts(scan):
    segs, cur, i = [], bytearray(), 0
    while i < len(scan):
        b = scan[i]
        if b != 0xFF:
            cur.append(b)
            i += 1
            continue
        nb = scan[i + 1] if i + 1 < len(scan) else 0xD9
        if nb == 0x00:
            cur.append(0xFF)
            i += 2
        elif 0xD0 <= nb <= 0xD7:
            segs.append(bytes(cur))
            cur = bytearray()
            i += 2
        elif nb == 0xFF:
            i += 1
        else:
            break
    segs.append(bytes(cur))
    return segs


def decode_scans(info):
    """Entropy decoding of all scans (T.81 annex F sequential, annex G progressive with spectral selection and successive
    approximation; restated from the standard's flow charts / jdphuff.c).  Same return value as decode_coefficients."""
    frame = info["frame"]
    hmax = max(c[1] for c in frame)
    vmax = max(c[2] for c in frame)
    H, W = info["height"], info["width"]
    mcux = (W + 8 * hmax - 1) // (8 * hmax)
    mcuy = (H + 8 * vmax - 1) // (8 * vmax)
    planes = [np.zeros((mcuy * c[2], mcux * c[1], 64), np.int32) for c in frame]
    for sc in info["scans"]:
        dct = {k: _huff_lookup(*v) for k, v in sc["dc"].items()}
        act = {k: _huff_lookup(*v) for k, v in sc["ac"].items()}
        ss, se, ah, al = sc["ss"], sc["se"], sc["ah"], sc["al"]
        if not info["progressive"] and (ss != 0 or se != 63 or ah or al):
            raise JpegError("bad sequential scan header")
        if info["progressive"] and ss == 0 and se != 0:
            raise JpegError("bad progressive scan header")
        if ss > 0 and len(sc["comps"]) != 1:
            raise JpegError("AC scans must have one component")
        # the units of the scan: whole MCUs for interleaved scans, single blocks of the component's true extent otherwise
        if len(sc["comps"]) > 1:
            units = [[(ci, my * frame[ci][2] + by, mx * frame[ci][1] + bx, td, ta)
                      for (ci, td, ta) in sc["comps"] for by in range(frame[ci][2]) for bx in range(frame[ci][1])]
                     for my in range(mcuy) for mx in range(mcux)]
        else:
            ci, td, ta = sc["comps"][0]
            bw = ((W * frame[ci][1] + hmax - 1) // hmax + 7) // 8
            bh = ((H * frame[ci][2] + vmax - 1) // vmax + 7) // 8
            units = [[(ci, by, bx, td, ta)] for by in range(bh) for bx in range(bw)]
        segs = _split_restarts(sc["data"])
        ri = sc["ri"] if sc["ri"] else len(units)
        u = 0
        for seg in segs:
            if u >= len(units):
                break
            br = _Bits(seg)
            pred = [0] * len(frame)
            eobrun = 0
            for unit in units[u:u + ri]:
                for (ci, by, bx, td, ta) in unit:
                    blk = planes[ci][by, bx]
                    if ss == 0:
                        if ah == 0:                                   # DC first scan (or the DC of a sequential scan)
                            s = br.symbol(dct[td])
                            pred[ci] += _extend(br.bits(s), s)
                            blk[0] = pred[ci] * (1 << al)
                        elif br.bits(1):                              # DC refinement
                            blk[0] |= 1 << al
                        if se == 0:
                            continue
                    k0 = max(ss, 1)
                    if ah == 0:                                       # AC first scan / sequential AC
                        if eobrun > 0:
                            eobrun -= 1
                            continue
                        k = k0
                        while k <= se:
                            rs = br.symbol(act[ta])
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r == 15:
                                    k += 16
                                    continue
                                if info["progressive"]:
                                    eobrun = (1 << r) + br.bits(r) - 1
                                break
                            k += r
                            if k > 63:
                                raise JpegError("coefficient index out of range")
                            blk[ZIGZAG[k]] = _extend(br.bits(s), s) * (1 << al)
                            k += 1
                    else:                                             # AC refinement (jdphuff.c decode_mcu_AC_refine)
                        p1, m1 = 1 << al, -(1 << al)
                        k = k0
                        if eobrun == 0:
                            while k <= se:
                                rs = br.symbol(act[ta])
                                r, s = rs >> 4, rs & 15
                                val = 0
                                if s:
                                    val = p1 if br.bits(1) else m1
                                elif r != 15:
                                    eobrun = (1 << r) + br.bits(r)
                                    break
                                while k <= se:
                                    z = ZIGZAG[k]
                                    if blk[z] != 0:
                                        if br.bits(1) and (blk[z] & p1) == 0:
                                            blk[z] += p1 if blk[z] >= 0 else m1
                                    else:
                                        r -= 1
                                        if r < 0:
                                            break
                                    k += 1
                                if val and k <= 63:
                                    blk[ZIGZAG[k]] = val
                                k += 1
                        if eobrun > 0:
                            while k <= se:
                                z = ZIGZAG[k]
                                if blk[z] != 0 and br.bits(1) and (blk[z] & p1) == 0:
                                    blk[z] += p1 if blk[z] >= 0 else m1
                                k += 1
                            eobrun -= 1
                    if br.pos > br.limit:
                        raise JpegError("premature end of the entropy-coded data")
            u += ri
    return [p.astype(np.int16) for p in planes], (hmax, vmax, mcux, mcuy)


# jidctint.c: CONST_BITS = 13, PASS1_BITS = 2
_F = dict(f0298=2446, f0390=3196, f0541=4433, f0765=6270, f0899=7373, f1175=9633, f1501=12299, f1847=15137, f1961=16069,
          f2053=16819, f2562=20995, f3072=25172)


def _idct_1d(x, shift):
    """One pass of jpeg_idct_islow over the LAST axis of int64 x [..., 8]; descale by `shift` bits with rounding."""
    x0, x1, x2, x3, x4, x5, x6, x7 = [x[..., i] for i in range(8)]
    z1 = (x2 + x6) * _F["f0541"]
    tmp2 = z1 + x6 * (-_F["f1847"])
    tmp3 = z1 + x2 * _F["f0765"]
    tmp0 = (x0 + x4) << 13
    tmp1 = (x0 - x4) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    t0, t1, t2, t3 = x7, x5, x3, x1
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * _F["f1175"]
    t0 = t0 * _F["f0298"]
    t1 = t1 * _F["f2053"]
    t2 = t2 * _F["f3072"]
    t3 = t3 * _F["f1501"]
    z1 = z1 * (-_F["f0899"])
    z2 = z2 * (-_F["f2562"])
    z3 = z3 * (-_F["f1961"]) + z5
    z4 = z4 * (-_F["f0390"]) + z5
    t0 = t0 + z1 + z3
    t1 = t1 + z2 + z4
    t2 = t2 + z2 + z3
    t3 = t3 + z1 + z4
    rnd = 1 << (shift - 1)
    outs = [tmp10 + t3, tmp11 + t2, tmp12 + t1, tmp13 + t0, tmp13 - t0, tmp12 - t1, tmp11 - t2, tmp10 - t3]
    return np.stack([(o + rnd) >> shift for o in outs], -1)


def idct_islow(coef, q):
    """coef int16 [..., 64] natural order, q int32 [64] -> uint8 samples [..., 8, 8] (jidctint.c jpeg_idct_islow)."""
    x = (coef.astype(np.int64) * q.astype(np.int64)).reshape(coef.shape[:-1] + (8, 8))
    # pass 1: columns (transpose so that the column index is last)
    ws = _idct_1d(np.swapaxes(x, -1, -2), 13 - 2)        # [..., col, row-out]
    ws = np.swapaxes(ws, -1, -2)                            # [..., row, col]
    # the C code keeps the workspace in 32-bit ints; values stay far inside that range for 8-bit data
    out = _idct_1d(ws, 13 + 2 + 3)
    out = out & 1023
    out = np.where(out >= 512, out - 1024, out) + 128       # range_limit[x & RANGE_MASK]
    return np.clip(out, 0, 255).astype(np.uint8)


def _plane(samples):
    """[by][bx][8][8] -> [by*8][bx*8]"""
    by, bx = samples.shape[:2]
    return samples.transpose(0, 2, 1, 3).reshape(by * 8, bx * 8)


def upsample_h2v1(p):
    """jdsample.c h2v1_fancy_upsample on a plane [h][w] -> [h][2w]"""
    p = p.astype(np.int32)
    h, w = p.shape
    out = np.zeros((h, 2 * w), np.int32)
    left = np.concatenate([p[:, :1], p[:, :-1]], 1)
    right = np.concatenate([p[:, 1:], p[:, -1:]], 1)
    out[:, 0::2] = (3 * p + left + 1) >> 2
    out[:, 1::2] = (3 * p + right + 2) >> 2
    out[:, 0] = p[:, 0]
    out[:, -1] = p[:, -1]
    return out.astype(np.uint8)


def upsample_h2v2(p):
    """jdsample.c h2v2_fancy_upsample on a plane [h][w] -> [2h][2w] (rows above / below the plane replicate its edge)."""
    p = p.astype(np.int32)
    h, w = p.shape
    up = np.concatenate([p[:1], p[:-1]], 0)
    dn = np.concatenate([p[1:], p[-1:]], 0)
    out = np.zeros((2 * h, 2 * w), np.int32)
    for v, other in ((0, up), (1, dn)):
        cs = 3 * p + other                                   # thiscolsum
        last = np.concatenate([cs[:, :1], cs[:, :-1]], 1)
        nxt = np.concatenate([cs[:, 1:], cs[:, -1:]], 1)
        even = (3 * cs + last + 8) >> 4
        odd = (3 * cs + nxt + 7) >> 4
        even[:, 0] = (4 * cs[:, 0] + 8) >> 4
        odd[:, -1] = (4 * cs[:, -1] + 7) >> 4
        out[v::2, 0::2] = even
        out[v::2, 1::2] = odd
    return out.astype(np.uint8)


def upsample_h1v2(p):
    """jdsample.c h1v2_fancy_upsample (libjpeg-turbo): [h][w] -> [2h][w]"""
    p = p.astype(np.int32)
    up = np.concatenate([p[:1], p[:-1]], 0)
    dn = np.concatenate([p[1:], p[-1:]], 0)
    out = np.zeros((2 * p.shape[0], p.shape[1]), np.int32)
    out[0::2] = (3 * p + up + 1) >> 2
    out[1::2] = (3 * p + dn + 2) >> 2
    return out.astype(np.uint8)


def ycc_to_bgr(y, cb, cr):
    """jdcolor.c ycc_rgb_convert (SCALEBITS = 16) -> uint8 [h][w][3] in B, G, R order."""
    y = y.astype(np.int32)
    xb = cb.astype(np.int32) - 128
    xr = cr.astype(np.int32) - 128
    fix = lambda v: int(v * 65536 + 0.5)
    r = y + ((fix(1.40200) * xr + 32768) >> 16)
    b = y + ((fix(1.77200) * xb + 32768) >> 16)
    g = y + (((-fix(0.34414)) * xb + 32768 + (-fix(0.71414)) * xr) >> 16)
    return np.clip(np.stack([b, g, r], -1), 0, 255).astype(np.uint8)


def imdecode(data):
    """cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR) for a Huffman-coded 8-bit JPEG (baseline, extended
    sequential or progressive): uint8 [H][W][3] BGR, rotated by the EXIF orientation like OpenCV does."""
    multi = parse_scans(data)
    if not multi["progressive"] and len(multi["scans"]) == 1 and len(multi["scans"][0]["comps"]) == len(multi["frame"]):
        info = parse(data)                                   # the common baseline file: one interleaved scan
        planes, (hmax, vmax, mcux, mcuy) = decode_coefficients(info)
        comps = info["comps"]
    else:
        info = multi
        planes, (hmax, vmax, mcux, mcuy) = decode_scans(info)
        comps = [c + (0, 0) for c in info["frame"]]
    H, W = info["height"], info["width"]
    full = []
    for (cid, ch, cv, tq, td, ta), coef in zip(comps, planes):
        pl = _plane(idct_islow(coef, info["qt"][tq]))
        # the up-samplers see the component at its true down-sampled size (jdsample.c uses downsampled_width, and the
        # main controller replicates the last real row below the image)
        dw = (W * ch + hmax - 1) // hmax
        dh = (H * cv + vmax - 1) // vmax
        pl = pl[:dh, :dw]
        # jdsample.c jinit_upsampler: fancy (triangle) filters for 2:1 ratios when the component is more than two
        # samples wide, plain replication (int_upsample) for every other integral ratio
        if ch == hmax and cv == vmax:
            pass
        elif ch * 2 == hmax and cv == vmax and dw > 2:
            pl = upsample_h2v1(pl)
        elif ch * 2 == hmax and cv * 2 == vmax and dw > 2:
            pl = upsample_h2v2(pl)
        elif ch == hmax and cv * 2 == vmax:
            pl = upsample_h1v2(pl)
        elif hmax % ch == 0 and vmax % cv == 0:
            pl = np.repeat(np.repeat(pl, vmax // cv, 0), hmax // ch, 1)
        else:
            raise JpegError("unsupported sampling factors")
        full.append(pl[:H, :W])
    if len(full) == 1:
        img = np.repeat(full[0][..., None], 3, -1)
    elif len(full) == 3:
        img = ycc_to_bgr(*full)
    else:
        raise JpegError("unsupported number of components")
    return exif_transform(img, info["orientation"])


def exif_transform(img, orientation):
    """OpenCV's ExifTransform (modules/imgcodecs/src/loadsave.cpp): cv2.imread / imdecode rotate the decoded image by the
    EXIF orientation tag (IFD0 tag 0x0112) unless IMREAD_IGNORE_ORIENTATION is given; other values leave it alone."""
    if orientation == 2:
        img = img[:, ::-1]
    elif orientation == 3:
        img = img[::-1, ::-1]
    elif orientation == 4:
        img = img[::-1]
    elif orientation == 5:
        img = img.transpose(1, 0, 2)
    elif orientation == 6:
        img = img.transpose(1, 0, 2)[:, ::-1]
    elif orientation == 7:
        img = img.transpose(1, 0, 2)[::-1, ::-1]
    elif orientation == 8:
        img = img.transpose(1, 0, 2)[::-1]
    return np.ascontiguousarray(img)


def with_exif_orientation(data, orientation, little_endian=False):
    """Test helper: the same JPEG file with an APP1 Exif segment carrying the given orientation inserted after SOI."""
    if little_endian:
        tiff = b"II\x2a\x00\x08\x00\x00\x00" + b"\x01\x00" + b"\x12\x01\x03\x00\x01\x00\x00\x00" + \
            int(orientation).to_bytes(2, "little") + b"\x00\x00" + b"\x00\x00\x00\x00"
    else:
        tiff = b"MM\x00\x2a\x00\x00\x00\x08" + b"\x00\x01" + b"\x01\x12\x00\x03\x00\x00\x00\x01" + \
            int(orientation).to_bytes(2, "big") + b"\x00\x00" + b"\x00\x00\x00\x00"
    exif = b"Exif\x00\x00" + tiff
    seg = b"\xff\xe1" + (len(exif) + 2).to_bytes(2, "big") + exif
    data = bytes(data)
    return data[:2] + seg + data[2:]
