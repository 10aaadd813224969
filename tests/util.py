import numpy as np


def rel_rmse(img, ref):
    """sqrt(mean((img-ref)^2)) / mean(ref) over rgb - the relative RMSE used for radiance parity."""
    a, b = np.asarray(img, np.float64)[..., :3], np.asarray(ref, np.float64)[..., :3]
    return float(np.sqrt(np.mean((a - b) ** 2)) / max(np.mean(b), 1e-12))


def pixel_mismatch_fraction(img, ref, tol=1e-3):
    a, b = np.asarray(img, np.float64)[..., :3], np.asarray(ref, np.float64)[..., :3]
    bad = (np.abs(a - b) > tol * (1 + np.abs(b))).any(axis=-1)
    return float(bad.mean())
