"""Statistical image comparison helpers (SURVEY.md §8c parity protocol, item 2)."""
import numpy as np


def mean_var(sum_, sumsq, n):
    mu = sum_ / n
    var = np.maximum(sumsq / n - mu * mu, 0.0) * n / max(n - 1, 1)
    return mu, var


def three_sigma_check(mu_a, var_a, n_a, mu_b, var_b, n_b):
    """per channel: mean_px |mu_a - mu_b| <= 3 * mean_px sqrt(var_a/n_a + var_b/n_b).
    Returns (ok, per-channel delta, per-channel bound)."""
    se = np.sqrt(var_a / n_a + var_b / n_b)
    d = np.abs(mu_a - mu_b).reshape(-1, 3).mean(0)
    b = 3.0 * se.reshape(-1, 3).mean(0)
    return bool(np.all(d <= b)), d, b


def zscores(mu_a, var_a, n_a, mu_b, var_b, n_b):
    se = np.sqrt(var_a / n_a + var_b / n_b)
    m = se > 0
    return ((mu_a - mu_b)[m] / se[m])


def psnr(a, b, peak=1.0):
    mse = np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2)
    return 10.0 * np.log10(peak * peak / max(mse, 1e-30))


def gamma(x):
    return np.sqrt(np.clip(x, 0.0, 1.0))


def batch_variance(batches):
    """batches: [K][H][W][3] per-batch MEANS of equal sample counts -> (mean, variance
    of the overall mean) estimated from the spread of the batch means."""
    b = np.asarray(batches, np.float64)
    k = b.shape[0]
    mu = b.mean(0)
    var_of_mean = b.var(0, ddof=1) / k
    return mu, var_of_mean
