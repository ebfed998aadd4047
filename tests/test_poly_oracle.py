"""CPU tests of the polygon path (reference ocr/tools/det_utils.py:97-245 `poly_core`):
* oracle/poly_ref.py is pinned against the goldens recorded from the LIVE reference (oracle/make_golden.py --poly):
  the same boxes get polygons, the 14 points agree to 1e-9;
* its library-free pieces are pinned against the live OpenCV: cv2.line pixel for pixel (incl. clipping),
  cv2.warpPerspective(INTER_NEAREST) pixel for pixel given the same matrix, cv2.getPerspectiveTransform to 1e-9
  (OpenCV solves that system through LAPACK: not bit-reproducible, see the oracle's header)."""
import os

import cv2
import numpy as np
import pytest

from lightly_ocr_b200.synth import receipts
from oracle import ocr_ref, poly_ref

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_poly.npz"))


@pytest.mark.parametrize("seed", range(6))
def test_poly_core_matches_live_reference(seed):
    t, l = receipts.curved_score_maps(seed)
    boxes, labels, mapper = ocr_ref.det_boxes(t, l)
    assert np.array_equal(np.array(boxes, np.float32).reshape(-1, 4, 2), GOLD["s%d_boxes" % seed])
    polys = poly_ref.poly_core(boxes, labels, mapper)
    valid = np.array([p is not None for p in polys], np.int32)
    assert np.array_equal(valid, GOLD["s%d_valid" % seed])
    for k, p in enumerate(polys):
        if p is not None:
            assert p.shape == (14, 2) and np.abs(p - GOLD["s%d_polys" % seed][k]).max() < 1e-9


def test_goldens_hold_polygons_and_early_exits():
    n = sum(int(GOLD["s%d_valid" % s].sum()) for s in range(6))
    m = sum(len(GOLD["s%d_valid" % s]) for s in range(6))
    assert n >= 15 and m - n >= 15


def test_line_pixels_match_cv2_line():
    rng = np.random.default_rng(0)
    for _ in range(4000):
        w, h = int(rng.integers(5, 60)), int(rng.integers(5, 40))
        p1 = (int(rng.integers(-30, w + 30)), int(rng.integers(-30, h + 30)))
        p2 = (int(rng.integers(-30, w + 30)), int(rng.integers(-30, h + 30)))
        want = np.zeros((h, w), np.uint8)
        cv2.line(want, p1, p2, 1, thickness=1)
        got = np.zeros((h, w), np.uint8)
        for (x, y) in poly_ref.line_pixels(w, h, p1, p2):
            got[y, x] = 1
        assert np.array_equal(want, got), (w, h, p1, p2)


def test_warp_and_perspective_match_cv2():
    rng = np.random.default_rng(1)
    H, W = 240, 320
    labels = rng.integers(0, 5, (H // 8, W // 8)).repeat(8, 0).repeat(8, 1).astype(np.int32)
    worst = 0.0
    for it in range(80):
        c = rng.uniform([60, 60], [W - 60, H - 60])
        ang = rng.uniform(-0.6, 0.6)
        bw, bh = rng.uniform(30, 110), rng.uniform(12, 40)
        rot = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        box = (np.array([[-bw, -bh], [bw, -bh], [bw, bh], [-bw, bh]]) / 2 @ rot.T + c).astype(np.float32)
        w = int(np.linalg.norm(box[0] - box[1]) + 1)
        h = int(np.linalg.norm(box[1] - box[2]) + 1)
        assert w == int(poly_ref._f32norm(box[0], box[1]) + np.float32(1))
        tar = np.float32([[0, 0], [w, 0], [w, h], [0, h]])
        m = cv2.getPerspectiveTransform(box, tar)
        mine = poly_ref.perspective(box, tar)
        worst = max(worst, float(np.abs(m - mine).max() / np.abs(m).max()))
        assert np.array_equal(poly_ref.warp_nearest(labels, m, w, h),
                              cv2.warpPerspective(labels, m, (w, h), flags=cv2.INTER_NEAREST))
        assert np.abs(poly_ref.invert3(m) - np.linalg.inv(m)).max() <= 1e-12 * np.abs(np.linalg.inv(m)).max()
    assert worst < 1e-9
