"""The oracle pin, re-checked on the GPU box: cv2 must be the pinned 4.13.0 there too, and the C restatement must
equal it live (the same checks as tests/test_oracle_golden.py, which the driver only runs in the CPU container).
Marked gpu only so that the GPU-box record (GPUTEST) shows them; they use no GPU."""
import pytest

import test_oracle_golden as tog

pytestmark = pytest.mark.gpu


def test_cv2_is_the_pinned_build():
    from oracle import cv2_ref
    assert cv2_ref.have_cv2(), "cv2 missing on the GPU box: the oracle pin cannot be re-checked here"
    assert cv2_ref.cv2_pinned().__version__ == cv2_ref.PINNED_VERSION


def test_oracle_equals_cv2_live_on_this_box(orc):
    tog.test_bm_oracle_vs_cv2_random_params(orc)
    tog.test_bm_oracle_vs_cv2_720p(orc)
    tog.test_sgbm_oracle_vs_cv2_random_params(orc)
    tog.test_oracle_vs_cv2_min_disparity(orc)
