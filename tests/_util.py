"""Shared helpers for the test-suite."""
import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name, prec):
    return np.load(os.path.join(GOLDEN, f"{name}_{prec}.npz"))


def golden_names(prefix, prec="f32"):
    return sorted(os.path.basename(f)[: -len(f"_{prec}.npz")] for f in glob.glob(os.path.join(GOLDEN, f"{prefix}*_{prec}.npz")))


def cfg_of(d):
    return json.loads(str(d["env_json"]))


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))


def tol(prec, f32, f64):
    return f32 if prec == "f32" else f64
