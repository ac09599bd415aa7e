"""tzddpc_b200 -- B200-native (sm_100a) implementation of the TZDDPC hot path.

Public names mirror the reference package (`tzddpc/__init__.py:1-23`); the zonotope value
types the reference imports from `pyzonotope` are shipped here as well.
Importing the package does not need a GPU; using it does (no CPU fallback).
"""
from .objects import (Data, DataDrivenDataset, OptimizationProblem, OptimizationProblemVariables,  # noqa: F401
                      SystemZonotopes, Theta)
from .ops import SolverOptions  # noqa: F401
from .program import BoxConstraint, StageCost  # noqa: F401
from .tzddpc import TZDDPC, TubeHandle  # noqa: F401
from .ensemble import TZDDPCEnsemble  # noqa: F401
from .zonotope import Interval, MatrixZonotope, Zonotope, concatenate_zonotope  # noqa: F401
from .utils import compute_theta, compute_A_B, compute_control_gain, is_gain_robust, spectral_radius  # noqa: F401

__version__ = "0.1.0"
__reference__ = "https://github.com/rssalessio/TZDDPC (0.0.3)"
