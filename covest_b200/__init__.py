"""covest_b200 -- CovEst's likelihood hot path on NVIDIA B200 (sm_100a).

Host-side mirror of the reference's interface (models, estimator, grid, CLI) over a C-ABI
library of hand-written CUDA kernels (covest_b200/csrc, include/covest_b200.h).  There is no CPU
implementation of the likelihood in this package: without the compiled library and a CUDA device
every evaluation raises.
"""
__version__ = '0.1.0'
# `covest --version`; 0.5.6 is the CovEst release whose interface and numerics are mirrored
version_string = 'CovEst 0.5.6 (covest_b200 %s)' % __version__
