"""TEST INFRASTRUCTURE - re-export of the synthetic input generators (lightly_ocr_b200/synth/receipts.py)."""
from lightly_ocr_b200.synth.receipts import *  # noqa: F401,F403
from lightly_ocr_b200.synth.receipts import crops, receipt, score_maps  # noqa: F401
