"""Optimization plugins of the hot path; same public names as learning/optimize/__init__.py:28-40."""
from .problem import Problem
from .initialstep import IncrPrevStep, FOChangeInitialStep, QuadraticInitialStep
from .linesearch import SetStepSize, BacktrackingLineSearch, WolfeLineSearch
from .optimizer import make_optimizer, SteepestDescent, SteepestDescentMomentum, BFGS, LBFGS
