"""`tfmpc` -- the reference's import name, served by tfmpc_b200.

A user of thiagopbueno/tf-mpc keeps `from tfmpc.solvers import ilqr`, `from tfmpc import envs`,
`tfmpc.envs.navigation.Navigation`, ... : every reference module path resolves to the B200-native
implementation in tfmpc_b200.
"""
import importlib
import sys

import tfmpc_b200

__version__ = tfmpc_b200.__version__

_ALIASES = ["solvers", "solvers.lqr", "solvers.ilqr", "envs", "envs.diffenv", "envs.gymenv", "envs.navigation", "envs.reservoir",
            "envs.hvac", "envs.lqr", "envs.lqr.navigation", "envs.synthetic", "utils", "utils.trajectory", "agents", "agents.mpc",
            "runners", "launchers"]
for _name in _ALIASES:
    _mod = importlib.import_module(f"tfmpc_b200.{_name}")
    sys.modules[f"tfmpc.{_name}"] = _mod
    if "." not in _name:
        globals()[_name] = _mod
