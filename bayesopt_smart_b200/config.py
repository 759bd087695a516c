"""Constants of the hot path.  Same names and values as the reference's bayesopt/config.py
(:16-83): they define the numerics (jitters, variance clamp) and the defaults of the public API.
"""
import os

import numpy as np

# bayesopt/config.py:16 -- kept for API compatibility; there is no Numba here, so the flag only
# silences nothing and selects nothing.
DEBUG_MODE = os.environ.get("BAYESIAN_DEBUG", "False").lower() in ("true", "1", "yes")

RANDOM_SEED = 42  # config.py:22
np.random.seed(RANDOM_SEED)  # config.py:25 (seeds the NumPy LHS initialisation)

DEFAULT_PRIOR_MEAN = 0.0  # 0.0 means "compute from the initial samples" (config.py:32-35)
DEFAULT_PRIOR_VARIANCE = 1.0
DEFAULT_LENGTH_SCALE = 1.0
DEFAULT_BETA = 1.0
DEFAULT_BATCH_SIZE = 3
DEFAULT_INITIAL_SAMPLES = 3

NUMBA_FLOAT_TYPE = np.float64  # config.py:54; the CUDA path is FP64 only

KERNEL_JITTER = 1e-6  # absolute, added to K before factorisation (config.py:64)
CHOLESKY_JITTER = 1e-8  # added to the normalised Gram matrix inside the MLL (config.py:65)
MIN_VARIANCE = 1e-10  # absolute clamp of the posterior variance (config.py:66)

HYPERPARAM_METHOD = "Powell"  # config.py:73-83
HYPERPARAM_XTOL = 1e-3
HYPERPARAM_FTOL = 1e-4
HYPERPARAM_MAXITER = 1000
HYPERPARAM_MIN_BOUND = 1e-5

DEFAULT_PLOT_ENABLED = True
