"""ctypes binding of libtfmpc_b200 (include/tfmpc_b200.h).

The CUDA library is the ONLY compute path of this package: there is no CPU or eager-PyTorch
fallback.  If the shared library is missing, cannot be loaded, or a tensor is not on a CUDA
device, the call raises -- loudly -- instead of computing something elsewhere.

Tensors cross the boundary zero-copy: each torch tensor is exported as a DLPack capsule
(torch.utils.dlpack.to_dlpack), the capsule's DLManagedTensor* is validated by the library
(tfmpc_dl_unpack: device, dtype, rank, contiguity) which returns the raw device pointer, and the
compute entry points receive plain pointers and sizes.
"""
import ctypes as C
import os
import threading

import torch
from torch.utils import dlpack as _dlpack

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBDIR = os.path.join(_HERE, "lib")
_LIBS = {}
_LOCK = threading.Lock()

ENV_KIND = {"NavigationLQR": 0, "Navigation": 1, "Reservoir": 2, "HVAC": 3}
STATUS = {0: "converged", 1: "max_iterations", 2: "non_pd", 3: "regularisation_loop", 4: "nan", 5: "aborted", 6: "tick_budget"}


class TfmpcError(RuntimeError):
    pass


class IlqrOpts(C.Structure):
    _fields_ = [("atol", C.c_double), ("max_iterations", C.c_int32), ("mu_min", C.c_double), ("delta_0", C.c_double),
                ("c1", C.c_double), ("alpha_min", C.c_double)]


_PyCapsule_GetPointer = C.pythonapi.PyCapsule_GetPointer
_PyCapsule_GetPointer.restype = C.c_void_p
_PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]


def lib_path(precision="f32"):
    # TFMPC_B200_LIBDIR: load another build of the same sources (A/B experiments); never a different implementation
    return os.path.join(os.environ.get("TFMPC_B200_LIBDIR", _LIBDIR), "libtfmpc_b200.so" if precision == "f32" else "libtfmpc_b200_f64.so")


def load(precision="f32"):
    """Load (once) and return the ctypes handle of the requested build."""
    with _LOCK:
        if precision in _LIBS:
            return _LIBS[precision]
        path = lib_path(precision)
        if not os.path.exists(path):
            raise TfmpcError(
                f"{path} is missing: the CUDA extension has not been built (run `python -m tfmpc_b200.build` or "
                "__graft_entry__.build()).  tfmpc_b200 has no CPU fallback.")
        lib = C.CDLL(path)
        lib.tfmpc_last_error.restype = C.c_char_p
        lib.tfmpc_kernel_launch_count.restype = C.c_int64
        lib.tfmpc_ilqr_workspace_bytes.restype = C.c_int64
        lib.tfmpc_ilqr_workspace_bytes.argtypes = [C.c_void_p, C.c_int64, C.c_int]
        lib.tfmpc_set_option.argtypes = [C.c_char_p, C.c_int]
        if lib.tfmpc_abi_version() != 1:
            raise TfmpcError("libtfmpc_b200 ABI version mismatch")
        want = 4 if precision == "f32" else 8
        if lib.tfmpc_real_bytes() != want:
            raise TfmpcError("libtfmpc_b200 precision mismatch")
        _LIBS[precision] = lib
        return lib


def check(lib, rc):
    if rc != 0:
        raise TfmpcError(f"libtfmpc_b200 error {rc}: {lib.tfmpc_last_error().decode()}")


def dtype_of(precision):
    return torch.float32 if precision == "f32" else torch.float64


def require_cuda():
    if not torch.cuda.is_available():
        raise TfmpcError("tfmpc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Ptr:
    """Keeps the DLPack capsule (and thus the tensor) alive while the pointer is in use."""

    def __init__(self, lib, tensor, want_device=2, want_code=0):
        if tensor is None:
            self.capsule, self.value = None, None
            return
        self.capsule = _dlpack.to_dlpack(tensor)
        managed = _PyCapsule_GetPointer(self.capsule, b"dltensor")
        data = C.c_void_p()
        check(lib, lib.tfmpc_dl_unpack(C.c_void_p(managed), want_device, want_code, C.byref(data), None, None, None))
        self.value = data.value

    @property
    def p(self):
        return C.c_void_p(self.value)


def dev_ptr(lib, tensor, int32=False):
    return _Ptr(lib, tensor, 2, 1 if int32 else 0)


def host_ptr(lib, tensor, int32=False):
    return _Ptr(lib, tensor, 1, 1 if int32 else 0)


class Env:
    """Owner of a tfmpc_env_t handle."""

    def __init__(self, precision, kind, n, m, nz, params):
        require_cuda()
        self.lib = load(precision)
        self.precision = precision
        self.kind, self.n, self.m, self.nz = kind, n, m, nz
        arr = (C.c_double * len(params))(*[float(v) for v in params])
        h = C.c_void_p()
        check(self.lib, self.lib.tfmpc_env_create(kind, n, m, nz, arr, C.c_int64(len(params)), C.byref(h)))
        self.handle = h

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                self.lib.tfmpc_env_destroy(h)
            except Exception:
                pass
            self.handle = None


_WS_CACHE = {}


def workspace(device, nbytes):
    """Grow-only per-(device, stream) scratch buffer handed to tfmpc_ilqr_solve."""
    key = (device, torch.cuda.current_stream().cuda_stream)
    buf = _WS_CACHE.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _WS_CACHE[key] = buf
    return buf


def kernel_launch_count(precision="f32"):
    return int(load(precision).tfmpc_kernel_launch_count())
