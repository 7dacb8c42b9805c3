import numpy as np
from scipy.ndimage import uniform_filter


def peak_signal_noise_ratio(image_true, image_test, data_range=None):
    image_true = np.asarray(image_true, dtype=np.float64)
    image_test = np.asarray(image_test, dtype=np.float64)
    if data_range is None:
        data_range = image_true.max() - image_true.min()
    err = np.mean((image_true - image_test) ** 2)
    return 10 * np.log10((data_range ** 2) / err)


def _ssim_2d(x, y, data_range, win=7):
    K1, K2 = 0.01, 0.03
    cov_norm = win * win / (win * win - 1.0)
    ux, uy = uniform_filter(x, win), uniform_filter(y, win)
    uxx, uyy, uxy = uniform_filter(x * x, win), uniform_filter(y * y, win), uniform_filter(x * y, win)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2))
    pad = (win - 1) // 2
    return S[pad:-pad, pad:-pad].mean()


def structural_similarity(im1, im2, multichannel=False, channel_axis=None, data_range=None, win_size=7, **kwargs):
    im1 = np.asarray(im1, dtype=np.float64)
    im2 = np.asarray(im2, dtype=np.float64)
    if data_range is None:
        data_range = im1.max() - im1.min()
    if multichannel or channel_axis is not None:
        ax = -1 if channel_axis is None else channel_axis
        return float(np.mean([_ssim_2d(np.take(im1, c, axis=ax), np.take(im2, c, axis=ax), data_range, win_size)
                              for c in range(im1.shape[ax])]))
    return float(_ssim_2d(im1, im2, data_range, win_size))
