"""TEST INFRASTRUCTURE - library-free numpy restatements of the cv2 / PIL routines the reference path calls on
integer / float32 data, written from the published algorithms and pinned against the live libraries
(cv2 4.13, Pillow 12.2 - the versions of this image; the reference pins 4.2 / 7.1.2) by
tests/test_oracle_exact.py.  These are the blueprints of the CUDA kernels in csrc/postproc.cu and csrc/imgops.cu.

  bgr2gray            cv2.cvtColor(COLOR_BGR2GRAY)                      reference ocr/pipeline.py:75
  pil_bicubic_resize  PIL.Image.resize(BICUBIC) for mode "L"            reference ocr/tools/dataset.py:44
  cv_resize_linear    cv2.resize(INTER_LINEAR) for uint8                reference ocr/tools/imgproc.py:51
  label4              cv2.connectedComponentsWithStats(connectivity=4)  reference ocr/tools/det_utils.py:45
  convex_hull / min_area_rect / box_points   cv2.minAreaRect + cv2.boxPoints   det_utils.py:75-76
  det_boxes           det_boxes_core without cv2                        det_utils.py:35-94
"""
import math

import numpy as np

f32 = np.float32


# ------------------------------------------------------------------------------------------------ colour
def bgr2gray(bgr):
    """OpenCV's 8-bit BGR->gray: fixed point with 15 fractional bits, coefficients 3735 / 19235 / 9798."""
    b = bgr[..., 0].astype(np.uint32)
    g = bgr[..., 1].astype(np.uint32)
    r = bgr[..., 2].astype(np.uint32)
    return ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(np.uint8)


# ------------------------------------------------------------------------------------------------ PIL bicubic
def _bicubic(x, a=-0.5):
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    if x < 2.0:
        return (((x - 5.0) * x + 8.0) * x - 4.0) * a
    return 0.0


def pil_coeffs(in_size, out_size):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc (22 fractional bits): per output index (xmin, int weights)."""
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 2.0 * fscale
    inv = 1.0 / fscale
    out = []
    for i in range(out_size):
        center = (i + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((j + xmin - center + 0.5) * inv) for j in range(xmax)]
        total = sum(w)
        if total != 0.0:
            w = [v / total for v in w]
        k = [int(-0.5 + v * (1 << 22)) if v < 0 else int(0.5 + v * (1 << 22)) for v in w]
        out.append((xmin, np.array(k, np.int64)))
    return out


def _pil_pass(img, coeffs, axis):
    src = img.astype(np.int64)
    if axis == 1:
        out = np.empty((img.shape[0], len(coeffs)), np.uint8)
        for i, (xmin, k) in enumerate(coeffs):
            acc = (1 << 21) + (src[:, xmin:xmin + len(k)] * k[None, :]).sum(1)
            out[:, i] = np.clip(acc >> 22, 0, 255)
    else:
        out = np.empty((len(coeffs), img.shape[1]), np.uint8)
        for i, (ymin, k) in enumerate(coeffs):
            acc = (1 << 21) + (src[ymin:ymin + len(k), :] * k[:, None]).sum(0)
            out[i, :] = np.clip(acc >> 22, 0, 255)
    return out


def pil_bicubic_resize(gray, out_w=100, out_h=32):
    """Two passes, horizontal first, uint8 intermediate; a pass is skipped when that dimension is unchanged."""
    img = gray
    if img.shape[1] != out_w:
        img = _pil_pass(img, pil_coeffs(img.shape[1], out_w), 1)
    if img.shape[0] != out_h:
        img = _pil_pass(img, pil_coeffs(img.shape[0], out_h), 0)
    return img


# ------------------------------------------------------------------------------------------------ cv2 bilinear
def _cv_linear_tab(in_size, out_size, clamp=True):
    """cv2 resize INTER_LINEAR coefficient table for 8-bit data: source index and two 11-bit weights.
    Only the horizontal table clamps out-of-range taps (weight moved onto the border pixel); the vertical pass keeps
    the fractional weights and clips the ROW indices instead, which rounds differently."""
    scale = in_size / out_size
    idx = np.empty(out_size, np.int64)
    wts = np.empty((out_size, 2), np.int64)
    for d in range(out_size):
        fx = f32((d + 0.5) * scale - 0.5)
        sx = int(math.floor(fx))
        fx = f32(fx - f32(sx))
        if clamp and sx < 0:
            sx, fx = 0, f32(0)
        if clamp and sx >= in_size - 1:
            sx, fx = in_size - 1, f32(0)
        w0 = int(np.rint(f32(f32(1.0) - fx) * f32(2048)))
        w1 = int(np.rint(fx * f32(2048)))
        idx[d] = sx
        wts[d] = (w0, w1)
    return idx, wts


def cv_resize_linear(img, out_w, out_h):
    """cv2.resize(img, (out_w,out_h), INTER_LINEAR) for uint8 HxWxC: 11-bit fixed-point weights in both passes."""
    if img.shape[0] == out_h and img.shape[1] == out_w:
        return img.copy()
    h, w = img.shape[:2]
    xi, xw = _cv_linear_tab(w, out_w)
    yi, yw = _cv_linear_tab(h, out_h, clamp=False)
    src = img.astype(np.int64)
    x1 = np.minimum(xi + 1, w - 1)
    rows = src[:, xi] * xw[:, 0][None, :, None] + src[:, x1] * xw[:, 1][None, :, None]     # [h,out_w,C] scaled 2^11
    y0 = np.clip(yi, 0, h - 1)
    y1 = np.clip(yi + 1, 0, h - 1)
    s0, s1 = rows[y0], rows[y1]
    b0, b1 = yw[:, 0][:, None, None], yw[:, 1][:, None, None]
    out = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


# ------------------------------------------------------------------------------------------------ labelling
def label4(mask):
    """4-connected labelling; label ids follow the raster order of each component's first pixel (as cv2's SAUF).
    Returns (n_labels incl. background, labels int32, stats [n,5] = left, top, width, height, area)."""
    h, w = mask.shape
    parent = np.arange(h * w, dtype=np.int64)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    fg = mask != 0
    for y in range(h):
        for x in range(w):
            if not fg[y, x]:
                continue
            i = y * w + x
            if x > 0 and fg[y, x - 1]:
                ra, rb = find(i), find(i - 1)
                if ra != rb:
                    parent[max(ra, rb)] = min(ra, rb)
            if y > 0 and fg[y - 1, x]:
                ra, rb = find(i), find(i - w)
                if ra != rb:
                    parent[max(ra, rb)] = min(ra, rb)
    labels = np.zeros((h, w), np.int32)
    ids = {}
    stats = [[0, 0, w, h, 0]]
    for y in range(h):
        for x in range(w):
            if not fg[y, x]:
                stats[0][4] += 1
                continue
            r = find(y * w + x)
            if r not in ids:
                ids[r] = len(ids) + 1
                stats.append([x, y, x, y, 0])
            k = ids[r]
            labels[y, x] = k
            s = stats[k]
            s[0] = min(s[0], x); s[1] = min(s[1], y); s[2] = max(s[2], x); s[3] = max(s[3], y); s[4] += 1
    st = np.array(stats, np.int32)
    st[1:, 2] = st[1:, 2] - st[1:, 0] + 1
    st[1:, 3] = st[1:, 3] - st[1:, 1] + 1
    return len(stats), labels, st


# ------------------------------------------------------------------------------------------------ hull + calipers
def convex_hull(points):
    """Convex hull of integer points in cv2.convexHull(clockwise=False) vertex order: starts at the point with the
    largest x (largest y among those), walks the y-max side towards smaller x, and returns along the y-min side;
    collinear points are dropped."""
    pts = sorted(set((int(x), int(y)) for x, y in points))
    if len(pts) <= 1:
        return pts

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lower, upper = [], []
    for p in pts:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], p) <= 0:
            lower.pop()
        lower.append(p)
    for p in reversed(pts):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], p) <= 0:
            upper.pop()
        upper.append(p)
    # lower: min-x -> max-x along small y;  upper: max-x -> min-x along large y  (y grows downwards in images)
    ring = upper[:-1] + lower[:-1]      # starts at max-x/max-y ... -> min-x ... -> back along small y
    return _cyclic_shift(ring, points)


def _cyclic_shift(ring, points):
    """OpenCV's last step: rotate the hull so that the ORIGINAL point indices form an ascending or descending
    sequence when that is possible (always for triangles, practically never for dilated pixel sets)."""
    nout = len(ring)
    if nout < 3:
        return ring
    first = {}
    for i, (x, y) in enumerate(points):
        first.setdefault((int(x), int(y)), i)
    # cv2 sorts pointers with std::sort (unstable), duplicates are rare in practice: first occurrence is used here
    hb = [first[p] for p in ring]
    min_idx = max_idx = lt = 0
    for i in range(1, nout):
        idx = hb[i]
        lt += hb[i - 1] < idx
        if 1 < lt <= i - 2:
            break
        if idx < hb[min_idx]:
            min_idx = i
        if idx > hb[max_idx]:
            max_idx = i
    mmdist = abs(max_idx - min_idx)
    if (mmdist == 1 or mmdist == nout - 1) and (lt <= 1 or lt >= nout - 2):
        ascending = (max_idx + 1) % nout == min_idx
        i0 = min_idx if ascending else max_idx
        if i0 > 0:
            j = i0
            stack = []
            ok = True
            for i in range(nout):
                cur = hb[j]
                stack.append(ring[j])
                nj = j + 1 if j + 1 < nout else 0
                if i < nout - 1 and (ascending != (cur < hb[nj])):
                    ok = False
                    break
                j = nj
            if ok:
                return stack
    return ring


def rotating_calipers(hull):
    """OpenCV's rotatingCalipers(CALIPERS_MINAREARECT) in float32, statement for statement.  hull: list of (x, y).
    Returns out[6] = corner (x, y), first edge vector (x, y), second edge vector (x, y)."""
    n = len(hull)
    px = np.array([p[0] for p in hull], f32)
    py = np.array([p[1] for p in hull], f32)
    vx = np.empty(n, f32); vy = np.empty(n, f32); inv = np.empty(n, f32)
    left = bottom = right = top = 0
    left_x = right_x = px[0]
    top_y = bottom_y = py[0]
    for i in range(n):
        if px[i] < left_x:
            left_x, left = px[i], i
        if px[i] > right_x:
            right_x, right = px[i], i
        if py[i] > top_y:
            top_y, top = py[i], i
        if py[i] < bottom_y:
            bottom_y, bottom = py[i], i
        j = i + 1 if i + 1 < n else 0
        dx = float(px[j]) - float(px[i])
        dy = float(py[j]) - float(py[i])
        vx[i] = f32(dx); vy[i] = f32(dy)
        inv[i] = f32(1.0 / math.sqrt(dx * dx + dy * dy))
    orientation = f32(0)
    ax, ay = float(vx[n - 1]), float(vy[n - 1])
    for i in range(n):
        bx, by = float(vx[i]), float(vy[i])
        conv = ax * by - ay * bx
        if conv != 0:
            orientation = f32(1) if conv > 0 else f32(-1)
            break
        ax, ay = bx, by
    base_a, base_b = orientation, f32(0)
    seq = [bottom, right, top, left]
    minarea = f32(np.finfo(np.float32).max)
    best = None
    for _ in range(n):
        dp = [
            f32(f32(base_a * vx[seq[0]]) + f32(base_b * vy[seq[0]])),
            f32(f32(-base_b * vx[seq[1]]) + f32(base_a * vy[seq[1]])),
            f32(f32(-base_a * vx[seq[2]]) - f32(base_b * vy[seq[2]])),
            f32(f32(base_b * vx[seq[3]]) - f32(base_a * vy[seq[3]])),
        ]
        maxcos = f32(dp[0] * inv[seq[0]])
        main = 0
        for i in range(1, 4):
            c = f32(dp[i] * inv[seq[i]])
            if c > maxcos:
                main, maxcos = i, c
        p = seq[main]
        lead_x = f32(vx[p] * inv[p]); lead_y = f32(vy[p] * inv[p])
        if main == 0:
            base_a, base_b = lead_x, lead_y
        elif main == 1:
            base_a, base_b = lead_y, f32(-lead_x)
        elif main == 2:
            base_a, base_b = f32(-lead_x), f32(-lead_y)
        else:
            base_a, base_b = f32(-lead_y), lead_x
        seq[main] = 0 if seq[main] + 1 == n else seq[main] + 1
        dx = f32(px[seq[1]] - px[seq[3]]); dy = f32(py[seq[1]] - py[seq[3]])
        width = f32(f32(dx * base_a) + f32(dy * base_b))
        dx = f32(px[seq[2]] - px[seq[0]]); dy = f32(py[seq[2]] - py[seq[0]])
        height = f32(f32(-dx * base_b) + f32(dy * base_a))
        area = f32(width * height)
        if area <= minarea:
            minarea = area
            best = (seq[3], base_a, width, base_b, height, seq[0])
    li, A1, width, B1, height, bi = best
    A2, B2 = f32(-B1), A1
    C1 = f32(f32(A1 * px[li]) + f32(py[li] * B1))
    C2 = f32(f32(A2 * px[bi]) + f32(py[bi] * B2))
    idet = f32(f32(1) / f32(f32(A1 * B2) - f32(A2 * B1)))
    ox = f32(f32(f32(C1 * B2) - f32(C2 * B1)) * idet)
    oy = f32(f32(f32(A1 * C2) - f32(A2 * C1)) * idet)
    return [ox, oy, f32(A1 * width), f32(B1 * width), f32(A2 * height), f32(B2 * height)]


def min_area_rect(points):
    """cv2.minAreaRect(points) -> ((cx, cy), (w, h), angle) with float32 fields (OpenCV 4.13 angle convention:
    the result is rotated by -90 degrees with width/height swapped until the angle lies in [-90, 0))."""
    hull = convex_hull(points)
    n = len(hull)
    if n > 2:
        o = rotating_calipers(hull)
        cx = f32(o[0] + f32(f32(o[2] + o[4]) * f32(0.5)))
        cy = f32(o[1] + f32(f32(o[3] + o[5]) * f32(0.5)))
        w = f32(math.sqrt(float(o[2]) * float(o[2]) + float(o[3]) * float(o[3])))
        h = f32(math.sqrt(float(o[4]) * float(o[4]) + float(o[5]) * float(o[5])))
        ang = math.atan2(float(o[3]), float(o[2]))
    elif n == 2:
        cx = f32(f32(f32(hull[0][0]) + f32(hull[1][0])) * f32(0.5))
        cy = f32(f32(f32(hull[0][1]) + f32(hull[1][1])) * f32(0.5))
        dx = float(hull[1][0]) - float(hull[0][0]); dy = float(hull[1][1]) - float(hull[0][1])
        w = f32(math.sqrt(dx * dx + dy * dy)); h = f32(0)
        ang = math.atan2(dy, dx)
    else:
        cx, cy = (f32(hull[0][0]), f32(hull[0][1])) if n == 1 else (f32(0), f32(0))
        w = h = f32(0); ang = 0.0
    # OpenCV 4.13 keeps the angle in double until the very end and folds it into [-90, 0)
    ang = ang * 180.0 / math.pi
    while ang >= 0.0:
        ang -= 90.0
        w, h = h, w
    while ang < -90.0:
        ang += 90.0
        w, h = h, w
    return (cx, cy), (w, h), f32(ang)


def box_points(rect):
    """cv2.boxPoints (RotatedRect::points): float32 corner arithmetic, cos/sin evaluated in double."""
    (cx, cy), (w, h), ang = rect
    a_ = float(ang) * math.pi / 180.0
    b = f32(f32(math.cos(a_)) * f32(0.5))
    a = f32(f32(math.sin(a_)) * f32(0.5))
    p0x = f32(f32(cx - f32(a * h)) - f32(b * w))
    p0y = f32(f32(cy + f32(b * h)) - f32(a * w))
    p1x = f32(f32(cx + f32(a * h)) - f32(b * w))
    p1y = f32(f32(cy - f32(b * h)) - f32(a * w))
    p2x = f32(f32(f32(2) * cx) - p0x); p2y = f32(f32(f32(2) * cy) - p0y)
    p3x = f32(f32(f32(2) * cx) - p1x); p3y = f32(f32(f32(2) * cy) - p1y)
    return np.array([[p0x, p0y], [p1x, p1y], [p2x, p2y], [p3x, p3y]], f32)


# ------------------------------------------------------------------------------------------------ det_boxes_core
def label4_fast(mask):
    """Same result as label4 (raster-first-pixel label order) using scipy's labelling + an explicit re-ranking."""
    from scipy import ndimage
    lab, n = ndimage.label(mask != 0, structure=[[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    flat = lab.ravel()
    idx = np.nonzero(flat)[0]
    vals, first = np.unique(flat[idx], return_index=True)
    order = np.argsort(idx[first], kind="stable")
    remap = np.zeros(n + 1, np.int32)
    remap[vals[order]] = np.arange(1, n + 1, dtype=np.int32)
    return n + 1, remap[lab].astype(np.int32)


def dilated_row_extents(rows_min, rows_max, y0, niter, sx, ex, sy, ey):
    """Per-row [min x, max x] of (S dilated by the (1+niter)^2 rectangle with OpenCV's anchor k//2) clipped to the ROI
    [sx, ex) x [sy, ey).  rows_min/rows_max: extents of S for rows y0.. (min > max marks an empty row)."""
    k = 1 + niter
    a = k // 2
    lo, hi = k - 1 - a, a          # a pixel p spreads to [p - lo, p + hi] in both axes
    n = len(rows_min)
    out = []
    for Y in range(max(sy, y0 - lo), min(ey, y0 + n + hi)):
        mn, mx = 1 << 30, -1
        for py in range(Y - hi, Y + lo + 1):
            j = py - y0
            if 0 <= j < n and rows_min[j] <= rows_max[j]:
                mn = min(mn, rows_min[j] - lo)
                mx = max(mx, rows_max[j] + hi)
        if mx >= 0:
            mn, mx = max(mn, sx), min(mx, ex - 1)
            if mn <= mx:
                out.append((Y, mn, mx))
    return out


def det_boxes(textmap, linkmap, text_threshold=0.7, link_threshold=0.4, low_text=0.4):
    """det_boxes_core (reference ocr/tools/det_utils.py:35-94) with every cv2 call replaced by the restatements
    above and the dilation never materialised (hull of the dilated set from per-row extents).  Returns
    (boxes float32 [n,4,2], kept label ids, labels)."""
    img_h, img_w = textmap.shape
    text_score = textmap > f32(low_text)            # cv2.threshold THRESH_BINARY is a strict >
    link_score = linkmap > f32(link_threshold)
    n, labels = label4_fast(np.logical_or(text_score, link_score))
    boxes, kept = [], []
    link_only = np.logical_and(link_score, ~text_score)
    ys_all, xs_all = np.nonzero(labels)
    lab_all = labels[ys_all, xs_all]
    order = np.argsort(lab_all, kind="stable")
    ys_all, xs_all, lab_all = ys_all[order], xs_all[order], lab_all[order]
    starts = np.searchsorted(lab_all, np.arange(1, n + 1))
    for k in range(1, n):
        ys, xs = ys_all[starts[k - 1]:starts[k]], xs_all[starts[k - 1]:starts[k]]
        area = len(ys)
        if area < 10:
            continue
        if textmap[ys, xs].max() < f32(text_threshold):
            continue
        x, y = int(xs.min()), int(ys.min())
        w, h = int(xs.max()) - x + 1, int(ys.max()) - y + 1
        niter = int(math.sqrt(area * min(w, h) / (w * h)) * 2)
        sx, ex, sy, ey = max(x - niter, 0), min(x + w + niter + 1, img_w), max(y - niter, 0), min(y + h + niter + 1, img_h)
        keep = ~link_only[ys, xs]
        rows_min = np.full(h, 1 << 30, np.int64)
        rows_max = np.full(h, -1, np.int64)
        np.minimum.at(rows_min, ys[keep] - y, xs[keep])
        np.maximum.at(rows_max, ys[keep] - y, xs[keep])
        ext = dilated_row_extents(rows_min, rows_max, y, niter, sx, ex, sy, ey)
        # raster order (row by row, left then right) exactly like np.where on the dilated map restricted to extremes
        pts = []
        for (Y, mn, mx) in ext:
            pts.append((mn, Y))
            if mx != mn:
                pts.append((mx, Y))
        rect = min_area_rect(pts)
        box = box_points(rect)
        ew = f32(np.linalg.norm(box[0] - box[1]))
        eh = f32(np.linalg.norm(box[1] - box[2]))
        ratio = f32(max(ew, eh) / f32(min(ew, eh) + f32(1e-5)))
        if abs(f32(1) - ratio) <= f32(0.1):
            l = min(p[0] for p in pts); r = max(p[0] for p in pts)
            t = min(p[1] for p in pts); b = max(p[1] for p in pts)
            box = np.array([[l, t], [r, t], [r, b], [l, b]], f32)
        start = int(np.argmin(box.sum(axis=1)))
        boxes.append(np.roll(box, 4 - start, 0))
        kept.append(k)
    return np.array(boxes, f32).reshape(-1, 4, 2), kept, labels


def rects_from_boxes(boxes, ratio_w, ratio_h, ratio_net=2):
    """adjustResultCoordinates (det_utils.py:259-265) + CRAFT.getCoords (net.py:92-97): the in-place `*=` of a float32
    array by a tuple of Python floats multiplies in float64 and rounds back to float32; then truncation toward zero
    and min/max over the corners -> [min_y, min_x, max_y, max_x]."""
    out = []
    for box in boxes:
        bx = (box[:, 0].astype(np.float64) * (ratio_w * ratio_net)).astype(f32)
        by = (box[:, 1].astype(np.float64) * (ratio_h * ratio_net)).astype(f32)
        xi = np.trunc(bx).astype(np.int32)
        yi = np.trunc(by).astype(np.int32)
        out.append([int(yi.min()), int(xi.min()), int(yi.max()), int(xi.max())])
    return out
