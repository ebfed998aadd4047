"""TEST INFRASTRUCTURE - CPU restatement of the reference's polygon path: `poly_core` (ocr/tools/det_utils.py:97-245),
reached through `getDetBoxes(..., poly=True)` (:248-256) when `CRAFT.enablePoly` is set.

The reference delegates four steps to OpenCV / LAPACK, none of which live under /root/reference:
  cv2.getPerspectiveTransform (8x8 linear solve), cv2.warpPerspective(INTER_NEAREST), np.linalg.inv (3x3) and
  cv2.line (thickness 1, 8-connected).
They are restated here without the libraries (blueprints of lightly_ocr_b200/csrc/polys.cu) and pinned against the live
libraries in tests/test_poly_oracle.py:
  * warp_nearest  == cv2.warpPerspective(int32 labels, INTER_NEAREST) pixel for pixel, given the same matrix;
  * line_pixels   == the pixels cv2.line sets;
  * perspective   == cv2.getPerspectiveTransform to ~1e-10 relative.  OpenCV solves the 8x8 system through LAPACK
    (this build: OpenBLAS dgesv), whose operation order is not reproducible bit for bit - nor is it the same from one
    OpenCV build to the next - so the matrix is the one quantity of this path that is compared with a tolerance; the
    decisions derived from it (which label each warped pixel takes) only change where a source coordinate lands within
    ~1e-9 of a rounding tie.
`poly_core` below follows the reference's control flow statement for statement (its Python float / numpy semantics
included) on top of those pieces.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import math

import numpy as np

NUM_CP = 5
MAX_LEN_RATIO = 0.7
EXPAND_RATIO = 1.45
MAX_R = 2.0
STEP_R = 0.2


def perspective(src, dst):
    """cv2.getPerspectiveTransform(src, dst) (float32 [4,2] each) -> float64 [3,3]: the 8x8 system of OpenCV's
    imgwarp.cpp solved by Gaussian elimination with partial pivoting in double precision (OpenCV's own LUImpl order)."""
    a = np.zeros((8, 8))
    b = np.zeros(8)
    for i in range(4):
        sx, sy, dx, dy = float(src[i][0]), float(src[i][1]), float(dst[i][0]), float(dst[i][1])
        a[i, 0] = a[i + 4, 3] = sx
        a[i, 1] = a[i + 4, 4] = sy
        a[i, 2] = a[i + 4, 5] = 1.0
        a[i, 6] = -sx * dx
        a[i, 7] = -sy * dx
        a[i + 4, 6] = -sx * dy
        a[i + 4, 7] = -sy * dy
        b[i] = dx
        b[i + 4] = dy
    n = 8
    for i in range(n):
        k = i
        for j in range(i + 1, n):
            if abs(a[j, i]) > abs(a[k, i]):
                k = j
        if abs(a[k, i]) < 2.220446049250313e-16 * 100:
            return None
        if k != i:
            a[[i, k]] = a[[k, i]]
            b[[i, k]] = b[[k, i]]
        d = -1.0 / a[i, i]
        for j in range(i + 1, n):
            alpha = a[j, i] * d
            for kk in range(i + 1, n):
                a[j, kk] = a[j, kk] + alpha * a[i, kk]
            b[j] = b[j] + alpha * b[i]
    x = np.zeros(n)
    for i in range(n - 1, -1, -1):
        s = b[i]
        for k in range(i + 1, n):
            s = s - a[i, k] * x[k]
        x[i] = s / a[i, i]
    return np.append(x, 1.0).reshape(3, 3)


def invert3(m):
    """Inverse of a 3x3 double matrix by cofactors (cv::invert's closed form for 3x3; np.linalg.inv to ~1e-16)."""
    det = (m[0, 0] * (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) - m[0, 1] * (m[1, 0] * m[2, 2] - m[1, 2] * m[2, 0]) +
           m[0, 2] * (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]))
    if det == 0.0:
        return None
    d = 1.0 / det
    t = np.empty((3, 3))
    t[0, 0] = (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) * d
    t[0, 1] = (m[0, 2] * m[2, 1] - m[0, 1] * m[2, 2]) * d
    t[0, 2] = (m[0, 1] * m[1, 2] - m[0, 2] * m[1, 1]) * d
    t[1, 0] = (m[1, 2] * m[2, 0] - m[1, 0] * m[2, 2]) * d
    t[1, 1] = (m[0, 0] * m[2, 2] - m[0, 2] * m[2, 0]) * d
    t[1, 2] = (m[0, 2] * m[1, 0] - m[0, 0] * m[1, 2]) * d
    t[2, 0] = (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]) * d
    t[2, 1] = (m[0, 1] * m[2, 0] - m[0, 0] * m[2, 1]) * d
    t[2, 2] = (m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]) * d
    return t


def warp_nearest(src, m, w, h):
    """cv2.warpPerspective(src, m, (w, h), flags=cv2.INTER_NEAREST) for a single-channel image: the matrix is inverted,
    every destination pixel (x, y) reads src at (cvRound(X / W), cvRound(Y / W)) (round half to even), 0 outside."""
    mi = invert3(m)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    x0 = mi[0, 0] * xs + mi[0, 1] * ys + mi[0, 2]
    y0 = mi[1, 0] * xs + mi[1, 1] * ys + mi[1, 2]
    ww = mi[2, 0] * xs + mi[2, 1] * ys + mi[2, 2]
    ww = np.where(ww != 0, 1.0 / np.where(ww != 0, ww, 1.0), 0.0)
    fx = np.clip(x0 * ww, -2.0 ** 31, 2.0 ** 31 - 1)
    fy = np.clip(y0 * ww, -2.0 ** 31, 2.0 ** 31 - 1)
    xi = np.rint(fx).astype(np.int64)
    yi = np.rint(fy).astype(np.int64)
    ok = (xi >= 0) & (xi < src.shape[1]) & (yi >= 0) & (yi < src.shape[0])
    out = np.zeros((h, w), src.dtype)
    out[ok] = src[yi[ok], xi[ok]]
    return out


def _clip_line(w, h, p1, p2):
    """cv::clipLine(Size(w, h), pt1, pt2) on 64-bit integers (drawing.cpp): returns None when fully outside."""
    x1, y1, x2, y2 = int(p1[0]), int(p1[1]), int(p2[0]), int(p2[1])
    right, bottom = w - 1, h - 1
    if w <= 0 or h <= 0:
        return None
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += _cdiv((a - y1) * (x2 - x1), (y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += _cdiv((a - y2) * (x2 - x1), (y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += _cdiv((a - x1) * (y2 - y1), (x2 - x1))
                x1 = a
                c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += _cdiv((a - x2) * (y2 - y1), (x2 - x1))
                x2 = a
                c2 = 0
    if (c1 | c2) != 0:
        return None
    return (x1, y1), (x2, y2)


def _cdiv(a, b):
    """C++ integer division (truncation toward zero)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def line_pixels(w, h, p1, p2):
    """The pixels cv2.line(img[h, w], p1, p2, color, thickness=1) sets (8-connected; cv::LineIterator on the clipped
    segment): list of (x, y)."""
    cl = _clip_line(w, h, p1, p2)
    if cl is None:
        return []
    (x1, y1), (x2, y2) = cl
    dx, dy = x2 - x1, y2 - y1
    if dx < 0:      # cv::Line asks the iterator to walk left to right: it starts from the other end point
        x1, y1, x2, y2 = x2, y2, x1, y1
        dx, dy = -dx, -dy
    sx = 1 if dx >= 0 else -1
    sy = 1 if dy >= 0 else -1
    dx, dy = abs(dx), abs(dy)
    pts = []
    if dx >= dy:
        err = dx - 2 * dy
        x, y = x1, y1
        for _ in range(dx + 1):
            pts.append((x, y))
            if err < 0:
                y += sy
                err += 2 * dx
            err -= 2 * dy
            x += sx
    else:
        err = dy - 2 * dx
        x, y = x1, y1
        for _ in range(dy + 1):
            pts.append((x, y))
            if err < 0:
                x += sx
                err += 2 * dy
            err -= 2 * dx
            y += sy
    return pts


def _f32norm(a, b):
    """np.linalg.norm(a - b) for float32 points: squares, sum and square root all in float32."""
    d0 = np.float32(a[0]) - np.float32(b[0])
    d1 = np.float32(a[1]) - np.float32(b[1])
    return np.sqrt(np.float32(d0 * d0) + np.float32(d1 * d1), dtype=np.float32)


def poly_core(boxes, labels, mapper):
    """poly_core (det_utils.py:97-245): one polygon ([14, 2] float64, score-map coordinates) or None per box."""
    polys = []
    for k, box in enumerate(boxes):
        box = np.asarray(box, np.float32)
        w = int(_f32norm(box[0], box[1]) + np.float32(1))
        h = int(_f32norm(box[1], box[2]) + np.float32(1))
        if w < 10 or h < 10:
            polys.append(None)
            continue
        tar = np.float32([[0, 0], [w, 0], [w, h], [0, h]])
        m = perspective(box, tar)
        minv = invert3(m) if m is not None else None
        if m is None or minv is None:
            polys.append(None)
            continue
        word = (warp_nearest(labels, m, w, h) == mapper[k]).astype(np.uint8)
        cp = []
        max_len = -1
        for i in range(w):
            region = np.nonzero(word[:, i])[0]
            if len(region) < 2:
                continue
            cp.append((i, int(region[0]), int(region[-1])))
            max_len = max(max_len, int(region[-1] - region[0] + 1))
        if h * MAX_LEN_RATIO < max_len:
            polys.append(None)
            continue
        tot_seg = NUM_CP * 2 + 1
        seg_w = w / tot_seg
        pp = [None] * NUM_CP
        cp_section = [[0, 0]] * tot_seg
        seg_height = [0] * NUM_CP
        seg_num = 0
        num_sec = 0
        prev_h = -1
        for (x, sy, ey) in cp:
            if (seg_num + 1) * seg_w <= x and seg_num <= tot_seg:
                if num_sec == 0:
                    break
                cp_section[seg_num] = [cp_section[seg_num][0] / num_sec, cp_section[seg_num][1] / num_sec]
                num_sec = 0
                seg_num += 1
                prev_h = -1
            cy = (sy + ey) * 0.5
            cur_h = ey - sy + 1
            cp_section[seg_num] = [cp_section[seg_num][0] + x, cp_section[seg_num][1] + cy]
            num_sec += 1
            if seg_num % 2 == 0:
                continue
            if prev_h < cur_h:
                pp[(seg_num - 1) // 2] = (x, cy)
                seg_height[(seg_num - 1) // 2] = cur_h
                prev_h = cur_h
        if num_sec != 0:
            cp_section[-1] = [cp_section[-1][0] / num_sec, cp_section[-1][1] / num_sec]
        if None in pp or seg_w < max(seg_height) * 0.25:
            polys.append(None)
            continue
        half_char_h = float(np.median(seg_height)) * EXPAND_RATIO / 2
        new_pp = []
        for i, (x, cy) in enumerate(pp):
            dx = cp_section[i * 2 + 2][0] - cp_section[i * 2][0]
            dy = cp_section[i * 2 + 2][1] - cp_section[i * 2][1]
            if dx == 0:
                new_pp.append([x, cy - half_char_h, x, cy + half_char_h])
                continue
            rad = -math.atan2(dy, dx)
            c, s = half_char_h * math.cos(rad), half_char_h * math.sin(rad)
            new_pp.append([x - s, cy - c, x + s, cy + c])
        found_s = found_e = False
        spp = epp = None
        grad_s = (pp[1][1] - pp[0][1]) / (pp[1][0] - pp[0][0]) + (pp[2][1] - pp[1][1]) / (pp[2][0] - pp[1][0])
        grad_e = (pp[-2][1] - pp[-1][1]) / (pp[-2][0] - pp[-1][0]) + (pp[-3][1] - pp[-2][1]) / (pp[-3][0] - pp[-2][0])
        for r in np.arange(0.5, MAX_R, STEP_R):
            dx = 2 * half_char_h * r
            if not found_s:
                dy = grad_s * dx
                p = np.array(new_pp[0]) - np.array([dx, dy, dx, dy])
                hit = any(word[y, x] for (x, y) in line_pixels(w, h, (int(p[0]), int(p[1])), (int(p[2]), int(p[3]))))
                if not hit or r + 2 * STEP_R >= MAX_R:
                    spp = p
                    found_s = True
            if not found_e:
                dy = grad_e * dx
                p = np.array(new_pp[-1]) + np.array([dx, dy, dx, dy])
                hit = any(word[y, x] for (x, y) in line_pixels(w, h, (int(p[0]), int(p[1])), (int(p[2]), int(p[3]))))
                if not hit or r + 2 * STEP_R >= MAX_R:
                    epp = p
                    found_e = True
            if found_s and found_e:
                break
        if not (found_s and found_e):
            polys.append(None)
            continue

        def unwarp(px, py):
            out = minv @ np.array([px, py, 1.0])
            return [out[0] / out[2], out[1] / out[2]]

        poly = [unwarp(spp[0], spp[1])]
        for p in new_pp:
            poly.append(unwarp(p[0], p[1]))
        poly.append(unwarp(epp[0], epp[1]))
        poly.append(unwarp(epp[2], epp[3]))
        for p in reversed(new_pp):
            poly.append(unwarp(p[2], p[3]))
        poly.append(unwarp(spp[2], spp[3]))
        polys.append(np.array(poly))
    return polys
