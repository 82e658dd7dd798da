"""Small host helpers with the reference's names and semantics (covest/utils.py)."""
import math
import subprocess
import sys

from . import constants
from .inverse import inverse


def verbose_print(message):
    """Progress messages go to stderr while constants.VERBOSE is set (utils.py:15-18)."""
    if constants.VERBOSE:
        sys.stderr.write(message + '\n')


def print_wrap(x, label='', cond=True):
    if cond:
        print(label, x)
    return x


def safe_int(x):
    return None if x == float('inf') else int(x)


def fix_zero(x, val=1):
    """`val` instead of an exact zero (utils.py:25-29)."""
    return val if x == 0 else x


def safe_log(x):
    """log with log(x <= 0) = -inf (utils.py:32-35)."""
    if x is None or x <= 0:
        return -constants.INF
    return math.log(x)


def estimate_p(cc, alpha):
    return (cc * (alpha - 1)) / (alpha * cc - alpha - cc)


def kmer_to_read_coverage(coverage, k, r):
    return coverage * r / (r - k + 1)


def _truncated_mean(c):
    # mean of a Poisson(c) conditioned on being >= 2
    return (c - c * math.exp(-c)) / (1 - math.exp(-c) - c * math.exp(-c))


def fix_coverage(coverage):
    """The Poisson rate whose >=2-truncated mean is `coverage` (utils.py:47-48)."""
    return inverse(_truncated_mean)(coverage)


def nonefloat(x):
    try:
        return float(x)
    except ValueError:
        return None


def run(command, shell=False, output=None, verbose=False):
    if verbose:
        print(command, file=sys.stderr)
    f = open(output, 'w') if output else None
    try:
        if not shell:
            command = command.split()
        return subprocess.call(command, shell=shell, stdout=f)
    finally:
        if f:
            f.close()
