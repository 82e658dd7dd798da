"""Tunable constants, same names and values as the reference's covest/constants.py so that code
and command lines written against it behave identically."""
import os

INF = float('inf')
VERBOSE = True

# numerics of the reference's Poisson module (c_src/covest_poissonmodule.c:5)
MAX_EXP = 200

# grid search (covest/grid.py)
GRID_DEPTH = 3
STEP = 1.1
INITIAL_GRID_COUNT = 20
INITIAL_GRID_STEP = 3

# optimiser
OPTIMIZATION_METHOD = 'L-BFGS-B'

# command-line defaults
DEFAULT_ERR_SCALE = 1
DEFAULT_K = 21
DEFAULT_READ_LENGTH = 100
DEFAULT_REPEAT_MODEL = 0
DEFAULT_MIN_SINGLECOPY_RATIO = 0.3
MAX_ERRORS = 8

# histogram pre-processing
AUTO_SAMPLE_TARGET_COVERAGE = 12
AUTO_TRIM_PRECISION = 6
NOISE_THRESHOLD = 10 ** -6
MAX_NOTRIM = 25

PLOT_LOG_SCALE = True
USE_BIGFLOAT = False

# accepted by the APIs that took a process count in the reference; the device path ignores it
DEFAULT_THREAD_COUNT = os.cpu_count() or 2
