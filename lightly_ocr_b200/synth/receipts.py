"""Synthetic inputs for the BASELINE.json configs.

receipt(seed)        white 1280x960 BGR "receipt" with ~70 rendered words (cv2.putText), config 2/4/5
score_maps(seed)     synthetic region/affinity score-map pairs for bit-exact post-processing tests
curved_score_maps(seed)  the same for words on wavy base lines (the polygon path, det_utils.py:97-245)
crops(n, seed)       gray uint8 crops of ragged sizes, config 3
"""
import cv2
import numpy as np

ALPHABET = "0123456789abcdefghijklmnopqrstuvwxyz"


def receipt(seed=0, height=1280, width=960, margin=60, scale=1.0, thickness=2, return_words=False):
    rng = np.random.default_rng(1000 + int(seed))
    img = np.full((height, width, 3), 255, np.uint8)
    words = []
    y = margin + 30
    while y < height - margin:
        x = margin + int(rng.integers(0, 40))
        while True:
            n = int(rng.integers(3, 9))
            word = "".join(ALPHABET[int(i)] for i in rng.integers(0, len(ALPHABET), n))
            (tw, th), base = cv2.getTextSize(word, cv2.FONT_HERSHEY_SIMPLEX, scale, thickness)
            if x + tw > width - margin:
                break
            shade = int(rng.integers(0, 60))
            cv2.putText(img, word, (x, y), cv2.FONT_HERSHEY_SIMPLEX, scale, (shade, shade, shade), thickness,
                        cv2.LINE_AA)
            words.append((word, x, y - th, tw, th + base))
            x += tw + int(rng.integers(40, 121))
        y += int(rng.integers(50, 81))
    if return_words:
        return img, words
    return img


def score_maps(seed=1, height=640, width=480, n_text=150, n_link=80):
    """Max of anisotropic Gaussians: text blobs (peak up to ~1) and link blobs (x0.8), incl. border-touching ones,
    near-square ("diamond" branch) ones and 1-pixel bridges."""
    rng = np.random.default_rng(int(seed))
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)

    def blobs(n, amp):
        m = np.zeros((height, width), np.float32)
        for i in range(n):
            cy = rng.uniform(-2, height + 2) if i % 10 == 0 else rng.uniform(8, height - 8)
            cx = rng.uniform(-2, width + 2) if i % 10 == 1 else rng.uniform(8, width - 8)
            sy = rng.uniform(2, 5)
            sx = rng.uniform(2, 5) if i % 7 == 0 else rng.uniform(4, 20)
            th = rng.uniform(-0.6, 0.6) if i % 3 == 0 else 0.0
            a = amp * rng.uniform(0.45, 1.0)
            dx, dy = xx - np.float32(cx), yy - np.float32(cy)
            u = np.float32(np.cos(th)) * dx + np.float32(np.sin(th)) * dy
            v = -np.float32(np.sin(th)) * dx + np.float32(np.cos(th)) * dy
            g = np.float32(a) * np.exp(-(u * u / np.float32(2 * sx * sx) + v * v / np.float32(2 * sy * sy)))
            m = np.maximum(m, g.astype(np.float32))
        return m

    text = blobs(n_text, 1.0)
    link = blobs(n_link, 0.8)
    # thin bridges between neighbouring blobs
    for _ in range(12):
        y = int(rng.integers(4, height - 4))
        x0 = int(rng.integers(4, width - 60))
        link[y, x0:x0 + int(rng.integers(10, 50))] = np.float32(0.55)
    return text, link


_RECEIPT_CACHE = {}


def crops(n=512, seed=3):
    """Ragged gray crops cut from the synthetic receipts: h in [16,48], w in [32,256], positioned on rendered words
    (with a random offset so that text is partially cut, as detector boxes do)."""
    rng = np.random.default_rng(int(seed))
    out = []
    while len(out) < n:
        rid = int(rng.integers(0, 8))
        if rid not in _RECEIPT_CACHE:
            img, words = receipt(rid, return_words=True)
            gray = ((img[..., 0].astype(np.uint32) * 3735 + img[..., 1].astype(np.uint32) * 19235
                     + img[..., 2].astype(np.uint32) * 9798 + 16384) >> 15).astype(np.uint8)
            _RECEIPT_CACHE[rid] = (gray, words)
        gray, words = _RECEIPT_CACHE[rid]
        word, x, y, tw, th = words[int(rng.integers(0, len(words)))]
        h = int(rng.integers(16, 49))
        w = int(rng.integers(32, 257))
        y0 = int(np.clip(y + th // 2 - h // 2 + rng.integers(-6, 7), 0, gray.shape[0] - h))
        x0 = int(np.clip(x + rng.integers(-10, 11), 0, gray.shape[1] - w))
        out.append(np.ascontiguousarray(gray[y0:y0 + h, x0:x0 + w]))
    return out


def curved_score_maps(seed=0, height=480, width=640, n_words=12):
    """Score-map pairs of words whose characters follow a sine-shaped base line: character blobs in the text map,
    blobs between neighbouring characters in the link map.  About half of the resulting boxes are curved enough for the
    reference's poly_core to produce a polygon, the others exercise its early exits."""
    rng = np.random.default_rng(int(seed))
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    text = np.zeros((height, width), np.float32)
    link = np.zeros((height, width), np.float32)
    for _ in range(n_words):
        cx0, cy0 = rng.uniform(60, width - 200), rng.uniform(60, height - 60)
        nchar = int(rng.integers(4, 10))
        pitch = rng.uniform(14, 22)
        amp = rng.uniform(0, 14) * (1 if rng.random() < 0.7 else 0)
        period = rng.uniform(80, 200)
        ch = rng.uniform(5, 9)
        pts = []
        for i in range(nchar):
            x = cx0 + i * pitch
            y = cy0 + amp * np.sin(2 * np.pi * i * pitch / period)
            pts.append((x, y))
            g = np.exp(-(((xx - np.float32(x)) ** 2) / np.float32(2 * (pitch * 0.28) ** 2) +
                         ((yy - np.float32(y)) ** 2) / np.float32(2 * ch ** 2)))
            text = np.maximum(text, (np.float32(rng.uniform(0.8, 1.0)) * g).astype(np.float32))
        for (x0, y0), (x1, y1) in zip(pts[:-1], pts[1:]):
            x, y = (x0 + x1) / 2, (y0 + y1) / 2
            g = np.exp(-(((xx - np.float32(x)) ** 2) / np.float32(2 * (pitch * 0.3) ** 2) +
                         ((yy - np.float32(y)) ** 2) / np.float32(2 * (ch * 0.6) ** 2)))
            link = np.maximum(link, (np.float32(0.8) * g).astype(np.float32))
    return text, link
