"""tfmpc_b200 -- B200-native batched LQR / iLQR solver with the API surface of thiagopbueno/tf-mpc.

Importing the package is cheap and GPU-free; the CUDA library (tfmpc_b200/lib/libtfmpc_b200.so,
built by `python -m tfmpc_b200.build`) is loaded on first use and there is no CPU fallback.
"""
__version__ = "0.1.0"
