"""The polygon path on the GPU (polys.cu behind locr_get_det_boxes / bridge.Pipeline.get_det_boxes = the reference's
tools.getDetBoxes(..., poly=True), ocr/tools/det_utils.py:248-256) against the goldens recorded from the LIVE reference
and against the CPU oracle: identical boxes (bit-exact float32), the same boxes get polygons, points within 1e-6 px."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_poly.npz"))


@pytest.fixture(scope="module")
def pipe():
    from lightly_ocr_b200 import bridge
    p = bridge.Pipeline(device_id=0, act_dtype=bridge.ACT_F16, head="CTC")
    yield p
    p.close()


@pytest.mark.parametrize("seed", range(6))
def test_polys_match_live_reference_goldens(pipe, seed):
    from lightly_ocr_b200.synth import receipts
    t, l = receipts.curved_score_maps(seed)
    boxes, polys = pipe.get_det_boxes(t, l, 0.7, 0.4, 0.4, poly=True)
    assert np.array_equal(np.array(boxes, np.float32).reshape(-1, 4, 2), GOLD["s%d_boxes" % seed])
    valid = np.array([p is not None for p in polys], np.int32)
    assert np.array_equal(valid, GOLD["s%d_valid" % seed])
    worst = 0.0
    for k, p in enumerate(polys):
        if p is not None:
            worst = max(worst, float(np.abs(p - GOLD["s%d_polys" % seed][k]).max()))
    print("seed %d: %d boxes, %d polygons, worst point difference %.3g px" % (seed, len(boxes), int(valid.sum()), worst))
    assert worst < 1e-6
    b2, p2 = pipe.get_det_boxes(t, l, 0.7, 0.4, 0.4, poly=False)
    assert len(b2) == len(boxes) and all(p is None for p in p2)


def test_polys_match_oracle_on_more_maps(pipe):
    """Seeds without goldens, and the straight-text maps of the box tests (every box takes an early exit there)."""
    from lightly_ocr_b200.synth import receipts
    from oracle import ocr_ref, poly_ref
    n_poly = 0
    for maps in [receipts.curved_score_maps(s) for s in range(6, 14)] + [receipts.score_maps(1)]:
        t, l = maps
        ob, labels, mapper = ocr_ref.det_boxes(t, l)
        want = poly_ref.poly_core(ob, labels, mapper)
        boxes, polys = pipe.get_det_boxes(t, l, 0.7, 0.4, 0.4, poly=True)
        assert len(polys) == len(want)
        for a, b in zip(polys, want):
            assert (a is None) == (b is None)
            if a is not None:
                n_poly += 1
                assert np.abs(a - b).max() < 1e-6
    assert n_poly >= 20
