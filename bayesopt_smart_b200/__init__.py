"""bayesopt_smart_b200 -- B200-native (sm_100a) implementation of BayesOpt_smart's acquisition hot
path behind the reference's own Python API.  Export list mirrors the reference's
``bayesopt/__init__.py:32-62``.  Importing the package does not touch the GPU; any compute call
requires the in-tree CUDA library (``python -m bayesopt_smart_b200.build``) and a CUDA device.
"""

__version__ = "0.1.0"

from .bayesian_optimization import BayesianOptimization
from .callbacks import OptimizationLogger, PerformanceMonitor, PlotterCallback, ProgressLogger
from .pareto import compute_pareto_front, is_pareto_efficient, print_pareto_analysis
from .acquisition import select_next_batch
from .config import (
    DEBUG_MODE,
    DEFAULT_BATCH_SIZE,
    DEFAULT_BETA,
    DEFAULT_INITIAL_SAMPLES,
    DEFAULT_LENGTH_SCALE,
    DEFAULT_PRIOR_MEAN,
    DEFAULT_PRIOR_VARIANCE,
    RANDOM_SEED,
)
from .engine import DeviceGP

__all__ = [
    "BayesianOptimization",
    "PlotterCallback",
    "ProgressLogger",
    "OptimizationLogger",
    "PerformanceMonitor",
    "select_next_batch",
    "is_pareto_efficient",
    "compute_pareto_front",
    "print_pareto_analysis",
    "DeviceGP",
    "__version__",
]
