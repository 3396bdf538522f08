from .mpc import MPC  # noqa: F401
