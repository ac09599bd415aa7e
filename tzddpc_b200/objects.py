"""Data containers of the controller -- same names and fields as the reference's
`tzddpc/objects.py:7-72`, importable without cvxpy (the reference imports cvxpy at module
load, `tzddpc/objects.py:3-4`)."""
from __future__ import annotations

from typing import Any, List, NamedTuple

import numpy as np

from .zonotope import Zonotope


class OptimizationProblemVariables(NamedTuple):
    """tzddpc/objects.py:7-18 (never instantiated by the reference; kept for API parity)."""
    y0: Any
    u: Any
    y: Any
    s_l: Any
    s_u: Any
    beta_u: Any


class OptimizationProblem(NamedTuple):
    """tzddpc/objects.py:20-30."""
    variables: OptimizationProblemVariables
    constraints: List[Any]
    objective_function: Any
    problem: Any


class Data(NamedTuple):
    """Input/state data, each T x dim (tzddpc/objects.py:33-40)."""
    u: np.ndarray
    x: np.ndarray


class DataDrivenDataset(NamedTuple):
    """X+, X-, U- split (tzddpc/objects.py:43-52)."""
    Xp: np.ndarray
    Xm: np.ndarray
    Um: np.ndarray
    original_data: Data


class SystemZonotopes(NamedTuple):
    """tzddpc/objects.py:55-67."""
    X0: Zonotope
    U: Zonotope
    X: Zonotope
    W: Zonotope


class Theta(NamedTuple):
    """tzddpc/objects.py:69-72."""
    K: np.ndarray
    deltaA: np.ndarray
    deltaB: np.ndarray
