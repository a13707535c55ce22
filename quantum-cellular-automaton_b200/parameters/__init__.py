"""Rule and run parameters with the attribute names of the reference's
``parameters.Rules`` (rules.py:1-6) and ``parameters.Parser`` (parser.py:182-212), so that
the algorithms accept either this ``Args`` or the reference's own ``Parser`` instance."""
from __future__ import annotations

import argparse
from dataclasses import dataclass, field


class Rules(object):
    """parameters/rules.py:1-6: ``activation_interval`` is a ``range`` (upper bound excluded)."""

    def __init__(self, ncells: int, activation_interval: range, distance: int, periodic: bool = False) -> None:
        if periodic:
            raise NotImplementedError("periodic boundaries are disabled in the reference as well (parser.py:53-58)")
        self.ncells = ncells
        self.activation_interval = activation_interval
        self.distance = distance
        self.periodic = periodic


@dataclass
class Args:
    """The fields of ``Parser`` the time-evolution path reads (parser.py:182-212)."""
    rules: Rules
    num_steps: int = 10000
    step_size: float = 0.005
    algorithm: str = "exact"
    max_bond_dim: int = 32
    svd_epsilon: float = 0.00005
    plot_frequency: float = 1.0
    approximative_evolution_method: str = "taylor"
    taylor_steps: int = 5
    initial_states: list = field(default_factory=list)
    initial_state_files: list = field(default_factory=list)

    @property
    def plot_step_interval(self) -> int:  # parser.py:209
        return int(1 / (self.plot_frequency * self.step_size))

    @property
    def plot_steps(self) -> int:  # parser.py:210-212
        steps = self.num_steps // self.plot_step_interval
        return steps + 1 if self.num_steps % self.plot_step_interval > 0 else steps

    @classmethod
    def from_argv(cls, argv=None) -> "Args":
        """Same flags, defaults and destinations as main.py's parser for the evolution path
        (parser.py:21-124); the plot flags of parser.py:125-178 belong to plot.py and are
        accepted and ignored."""
        p = argparse.ArgumentParser(description="quantum game of life, B200 time-evolution path")
        p.add_argument("--num-cells", dest="NUM_CELLS", type=int, default=9)
        p.add_argument("--distance", dest="DISTANCE", type=int, default=1)
        p.add_argument("--activation-interval", dest="INTERVAL", type=int, nargs=2, default=(1, 2),
                       metavar=("LOWER", "UPPER"))
        p.add_argument("--num-steps", dest="NUM_STEPS", type=int, default=10000)
        p.add_argument("--initial-states", dest="INITIAL_STATES", nargs="*", default=[])
        p.add_argument("--initial-state-files", dest="INITIAL_STATE_FILES", nargs="*", default=[])
        p.add_argument("--algorithm", dest="ALGORITHM", default="exact",
                       choices=["exact", "1tdvp", "2tdvp", "a1tdvp"])
        p.add_argument("--approximative-evolution-method", dest="APPROX", default="taylor",
                       choices=["taylor", "expm_multiply", "exact_exponential"])
        p.add_argument("--taylor-steps", dest="TAYLOR_STEPS", type=int, default=5)
        p.add_argument("--step-size", dest="STEP_SIZE", type=float, default=.005)
        p.add_argument("--max-bond-dim", dest="MAX_BOND_DIM", type=int, default=32)
        p.add_argument("--svd-epsilon", dest="SVD_EPSILON", type=float, default=0.00005)
        p.add_argument("--plotting-frequency", dest="PLOTTING_FREQUENCY", type=float, default=1.0)
        a, _ = p.parse_known_args(argv)
        if a.ALGORITHM == "a1tdvp":
            # the flag value exists in the reference (parser.py:77-83); its adaptive bond-dimension heuristics
            # (tdvp.py:190-268, marked experimental there) are outside the B200 path
            p.error("--algorithm a1tdvp is not implemented on the B200 path (use exact, 1tdvp or 2tdvp)")
        return cls(
            rules=Rules(ncells=a.NUM_CELLS, activation_interval=range(a.INTERVAL[0], a.INTERVAL[1]),
                        distance=a.DISTANCE, periodic=False),
            num_steps=a.NUM_STEPS, step_size=a.STEP_SIZE, algorithm=a.ALGORITHM,
            max_bond_dim=a.MAX_BOND_DIM, svd_epsilon=a.SVD_EPSILON, plot_frequency=a.PLOTTING_FREQUENCY,
            approximative_evolution_method=a.APPROX, taylor_steps=a.TAYLOR_STEPS,
            initial_states=list(a.INITIAL_STATES or []), initial_state_files=list(a.INITIAL_STATE_FILES or []))
