"""The two scikit-image metrics the reference's utils/eval_utils.py:8-11 calls, implemented with numpy / scipy
(scikit-image is not installed in this image).  Same defaults as skimage 0.19: 7x7 uniform window, K1 = 0.01, K2 = 0.03,
sample covariance; `multichannel=True` (the keyword the reference uses) averages the per-channel SSIM."""
from . import metrics, measure  # noqa: F401
