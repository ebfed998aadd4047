"""TEST INFRASTRUCTURE - re-export of the layer tables shared with the synthetic-checkpoint generator."""
from lightly_ocr_b200.synth.specs import *  # noqa: F401,F403
from lightly_ocr_b200.synth.specs import FE, LOC  # noqa: F401
