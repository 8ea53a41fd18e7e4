"""Synthetic belief states for the covariance-propagation kernel (config C4): the generator lives in the package
(bench.py's `c4_sweep` uses it too)."""
from drl_graph_exploration_b200.synth import synth_states  # noqa: F401
