"""TEST INFRASTRUCTURE ONLY - the CPU oracle of the lightly-ocr detect-then-recognize path.

Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import it, and only as the checker / the timed CPU baseline.  The product path (lightly_ocr_b200/) never imports it.

Contents
  specs.py / receipts.py / weights.py   re-exports of the synthetic input generators that live in
                lightly_ocr_b200/synth/ (data only); weights.py adds the one-off calibration that needs the oracle
  ocr_ref.py    restatement of the reference path with the reference's own third-party arithmetic
                (torch CPU fp32, cv2, PIL) - the parity oracle and the CPU baseline ("port")
  exact.py      library-free numpy restatements of the cv2 / PIL integer and float32 routines the path calls
                (BGR2GRAY, PIL bicubic, 4-connected labelling order, dilate extents, minAreaRect, boxPoints)
  make_golden.py  imports the REAL reference from /root/reference (this container only) and writes tests/golden/

Pinning: oracle/ocr_ref.py is pinned against the live reference (ocr/net.py, ocr/model.py, ocr/tools/*) by the
fixtures under tests/golden/ that make_golden.py generated here; oracle/exact.py is pinned against live cv2 / PIL in
tests/test_oracle_exact.py.  The reference's own tests pin only CTC decode (ocr/test/utils_test.py:37-43), which
tests/test_oracle_pin.py reproduces.
"""
