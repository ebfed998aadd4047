"""Pins oracle/exact.py (library-free restatements of the cv2 / PIL routines) against the live libraries of this image
(cv2 4.13, Pillow 12.2).  CPU only."""
import numpy as np
import pytest

from oracle import exact, ocr_ref, receipts


def test_bgr2gray_matches_cv2():
    import cv2
    img = np.random.default_rng(0).integers(0, 256, (64, 80, 3), dtype=np.uint8)
    assert np.array_equal(exact.bgr2gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


def test_pil_bicubic_matches_pillow():
    from PIL import Image
    rng = np.random.default_rng(1)
    for _ in range(60):
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 400))
        g = rng.integers(0, 256, (h, w), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(g).convert("L").resize((100, 32), Image.BICUBIC))
        assert np.array_equal(exact.pil_bicubic_resize(g), ref), (h, w)


def test_cv_resize_linear_matches_cv2():
    import cv2
    rng = np.random.default_rng(2)
    for _ in range(40):
        h, w = int(rng.integers(8, 200)), int(rng.integers(8, 200))
        g = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        r = rng.uniform(0.4, 2.5)
        oh, ow = max(2, int(h * r)), max(2, int(w * r))
        assert np.array_equal(exact.cv_resize_linear(g, ow, oh), cv2.resize(g, (ow, oh), interpolation=cv2.INTER_LINEAR))


def test_label4_order_matches_cv2():
    import cv2
    rng = np.random.default_rng(3)
    for shape, p in (((40, 50), 0.5), ((64, 48), 0.3), ((30, 30), 0.7)):
        m = (rng.random(shape) < p).astype(np.uint8)
        n, lab, stats, _ = cv2.connectedComponentsWithStats(m, connectivity=4)
        n2, lab2, st2 = exact.label4(m)
        assert n == n2 and np.array_equal(lab, lab2) and np.array_equal(stats[1:], st2[1:])
        n3, lab3 = exact.label4_fast(m)
        assert n == n3 and np.array_equal(lab, lab3)


def test_min_area_rect_and_box_points_match_cv2():
    import cv2
    rng = np.random.default_rng(4)
    bad = 0
    for _ in range(1500):
        n, s = int(rng.integers(1, 60)), int(rng.integers(2, 60))
        pts = np.unique(rng.integers(0, s, (n, 2)).astype(np.int32), axis=0)
        rng.shuffle(pts)
        ref = cv2.minAreaRect(pts)
        got = exact.min_area_rect(pts)
        a = np.array([ref[0][0], ref[0][1], ref[1][0], ref[1][1], ref[2]], np.float32)
        b = np.array([got[0][0], got[0][1], got[1][0], got[1][1], got[2]], np.float32)
        bad += not np.array_equal(a, b)
        assert np.array_equal(cv2.boxPoints(ref), exact.box_points(ref))
    assert bad <= 2, "restated rotating calipers disagree with cv2 on %d / 1500 point sets" % bad


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_det_boxes_without_cv2_matches_cv2_path(seed):
    t, l = receipts.score_maps(seed)
    ref_boxes, ref_labels, kept = ocr_ref.det_boxes(t.copy(), l.copy())
    boxes, kept2, labels = exact.det_boxes(t, l)
    assert np.array_equal(labels, ref_labels) and kept == kept2
    ref = np.array(ref_boxes, np.float32).reshape(-1, 4, 2)
    ulp = np.abs(boxes.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
    assert ulp.max() <= 1 and (ulp > 0).any(axis=(1, 2)).sum() <= 1
    for rw in (1.0, 1 / 1.5):
        assert exact.rects_from_boxes(boxes, rw, rw) == ocr_ref.rects_from_boxes(list(ref), rw, rw)
