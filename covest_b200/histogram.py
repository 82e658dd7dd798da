"""Histogram pre-processing on the host: moment-based first guess, down-sampling, trimming
(reference: covest/histogram.py).  Runs once per invocation on a few hundred bins; it decides how
many bins the device kernels see, it is not part of the accelerated path.
"""
import math
import random
from collections import defaultdict

import numpy as np

from . import constants
from .utils import estimate_p, fix_coverage, kmer_to_read_coverage, verbose_print


def poisson_dist(lam, max_j):
    """[Poisson(lam).pmf(j) for j = 1..max_j] -- the host stand-in for covest_poisson.poisson_dist
    (c_src/covest_poissonmodule.c:64-108).  Evaluated in the log domain, so unlike the reference
    it stays correct for lam > 200 (there the reference subtracts 200 from the rate inside its bin
    loop, c:92-95)."""
    if lam == 0 or lam != lam:
        return [0.0] * max_j
    j = np.arange(1, max_j + 1, dtype=np.float64)
    from scipy.special import gammaln
    return np.exp(j * math.log(lam) - gammaln(j + 1.0) - lam).tolist()


def compute_coverage_apx(hist, k, r):
    """(coverage, error_rate) guessed from the moments of the histogram (histogram.py:12-44):
    the mean abundance of bins >= 2 fixes the k-mer coverage, the excess of singletons over the
    Poisson expectation fixes the share of erroneous k-mers."""
    ones = hist.get(1, 0)
    kmers = sum(j * h for j, h in hist.items())
    distinct = sum(hist.values())
    if distinct == 0:
        return 0.0, 1.0
    kmers -= ones
    multi = distinct - ones
    try:
        cov = fix_coverage(kmers / multi)
        multi /= (1.0 - math.exp(-cov) - cov * math.exp(-cov))
        expect_ones = multi * cov * math.exp(-cov)
        expect_zeros = multi * math.exp(-cov)
        alpha = max(0.0, ones - expect_ones) / (distinct + expect_zeros)
        p_ok = max(0.0, estimate_p(cov, alpha))
        err = 1 - p_ok ** (1.0 / k)
        if p_ok > 0:
            return float(kmer_to_read_coverage(cov / p_ok, k, r)), float(err)
        return 0.0, float(err)
    except ZeroDivisionError:
        return 0.0, 1.0


def sample_histogram(hist, factor=2, trim=None):
    """The histogram expected after keeping every read with probability 1/factor
    (histogram.py:47-74): a k-mer seen i times is seen Binomial(i, 1/factor) times (Poisson
    approximation from i = 100 on); fractional counts are rounded at random."""
    from scipy.stats import binom
    if trim is None:
        trim = get_trim(hist) if len(hist) > 300 else max(hist)
    else:
        trim = min(max(hist), trim * factor)
    kept = {j: h for j, h in hist.items() if j < trim}
    acc = defaultdict(int)
    prob = 1.0 / factor
    for i, h in kept.items():
        if i < 100:
            dist = binom(i, prob)
            pmf = [dist.pmf(j) for j in range(1, i + 1)]
        else:
            pmf = poisson_dist(i * prob, i)
        for j, p in enumerate(pmf):
            acc[j + 1] += h * p
    out = dict(acc)
    for j, v in out.items():
        frac = v - round(v)
        out[j] = math.ceil(v) if random.random() < frac else math.floor(v)
    return {j: v for j, v in out.items() if v > 0}


def auto_sample_hist(hist, k, r, trim=None):
    """Smallest sampling factor that brings the guessed coverage under
    AUTO_SAMPLE_TARGET_COVERAGE: doubling search, then bisection (histogram.py:77-102)."""
    best = dict(hist)
    factor, stride = 1, 1
    c, e = compute_coverage_apx(hist, k, r)
    while c > constants.AUTO_SAMPLE_TARGET_COVERAGE:
        factor += stride
        stride *= 2
        best = sample_histogram(hist, factor=factor, trim=trim)
        c, e = compute_coverage_apx(best, k, r)
    stride //= 4
    probe = factor - stride
    while stride >= 1:
        cand = sample_histogram(hist, factor=probe, trim=trim)
        c, e = compute_coverage_apx(cand, k, r)
        if c > constants.AUTO_SAMPLE_TARGET_COVERAGE:
            probe += stride
        else:
            best, factor = cand, probe
            probe -= stride
        stride //= 2
    return best, factor, c, e


def remove_noise(hist):
    total = sum(hist.values())
    return {j: h for j, h in hist.items() if h / total > constants.NOISE_THRESHOLD}


def get_trim(hist, ignore_last=False):
    """First bin at which the cumulative share of (de-noised) counts rounds to 1 at
    AUTO_TRIM_PRECISION digits (histogram.py:111-124)."""
    hist = remove_noise(hist)
    total = float(sum(hist.values()))
    if ignore_last:
        total -= hist[max(hist)]
    run = 0.0
    trim = max(hist)
    for j, h in sorted(hist.items()):
        run += h
        if round(run / total, constants.AUTO_TRIM_PRECISION) >= 1:
            trim = j
            break
    return trim


def trim_hist(hist, threshold):
    """(bins below `threshold` without empty ones, total count at or above it)
    (histogram.py:127-134)."""
    if threshold >= max(hist):
        return hist, 0
    tail = sum(h for j, h in hist.items() if j >= threshold)
    return {j: h for j, h in hist.items() if j < threshold and h > 0}, tail


def process_histogram(hist, k, r, trim=None, sample_factor=None, max_notrim=constants.MAX_NOTRIM):
    """-> (hist, tail, sample_factor, guessed coverage, guessed error rate)
    (histogram.py:137-162)."""
    hist = dict(hist)
    tail = 0
    if sample_factor is not None and sample_factor > 1:
        verbose_print('Sampling histogram {}x...'.format(sample_factor))
        hist = sample_histogram(hist, sample_factor, trim)
    if sample_factor is None and max(hist) > max_notrim:
        verbose_print('Sampling histogram...')
        hist, sample_factor, c, e = auto_sample_hist(hist, k, r, trim=trim)
        if sample_factor > 1:
            verbose_print('Histogram sampled with factor {}.'.format(sample_factor))
        else:
            verbose_print('No sampling necessary')
    else:
        c, e = compute_coverage_apx(hist, k, r)
        if sample_factor is None:
            sample_factor = 1
    if trim is None:
        if max(hist) > max_notrim:
            trim = get_trim(hist, ignore_last=True)
            verbose_print('Trimming at: {}'.format(trim))
            hist, tail = trim_hist(hist, trim)
    elif trim > 0:
        verbose_print('Trimming at: {}'.format(trim))
        hist, tail = trim_hist(hist, trim)
    return hist, tail, sample_factor, c, e
