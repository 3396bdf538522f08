"""Compile the CUDA sources in tfmpc_b200/csrc into the two in-tree shared libraries

    tfmpc_b200/lib/libtfmpc_b200.so       tfmpc_real = float   (product build)
    tfmpc_b200/lib/libtfmpc_b200_f64.so   tfmpc_real = double  (verification build, no FMA contraction)
    tfmpc_b200/lib_ieee/libtfmpc_b200.so  tfmpc_real = float, IEEE division / sqrt / exp, no flush-to-zero (A/B build: what
                                          --use_fast_math changes in the product build; loaded only via TFMPC_B200_LIBDIR)

for sm_100a only.  nvcc cross-compiles without a GPU, so this runs in the build container;
the .so files are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")
LIB_IEEE = os.path.join(HERE, "lib_ieee")
OBJ = os.path.join(HERE, "build")
SOURCES = ["api.cu", "ilqr_small.cu", "ilqr_queue.cu", "ilqr_warp.cu", "env_ops.cu", "lqr.cu", "peak.cu", "backward_dense.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# fp32 product build: fast intrinsics (MUFU-based division / exp / sqrt, FTZ).  Measured on B200 ("numerics"):
# 1.5x faster end to end and statistically indistinguishable parity -- same-iteration-count agreement with the fp32 oracle
# 96.7% with and 96.6% without, against an fp32-vs-fp64 noise band of 96.1%.  The fp64 verification build stays IEEE.
F32_FLAGS = "--use_fast_math"
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return nvcc


def _host_compiler_flags():
    # the image exports CC/CXX pointing at a wrapper without the system specs; use the system g++
    return ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []


def _deps_mtime():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "tfmpc_b200.h")]
    return max(os.path.getmtime(f) for f in files)


def _compile(args):
    src, obj, extra, verbose = args
    cmd = [_nvcc()] + ARCH + COMMON + _host_compiler_flags() + extra + ["-c", src, "-o", obj]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False, precisions=("f32", "f64", "f32ieee")):
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(LIB_IEEE, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    newest = _deps_mtime()
    outputs = {}
    jobs = []
    for prec in precisions:
        name = "libtfmpc_b200_f64.so" if prec == "f64" else "libtfmpc_b200.so"
        out = os.path.join(LIB_IEEE if prec == "f32ieee" else LIB, name)
        outputs[prec] = out
        if not force and os.path.exists(out) and os.path.getmtime(out) >= newest:
            continue
        # fp64 verification build: no FMA contraction, so that two code shapes of the same expressions (thread-per-problem,
        # warp-cooperative, tick kernels) and the gcc-built oracle round identically
        extra = (["-DTFMPC_F64", "-fmad=false"] if prec == "f64" else [] if prec == "f32ieee" else
                 [f for f in os.environ.get("TFMPC_F32_FLAGS", F32_FLAGS).split() if f])
        for s in SOURCES:
            jobs.append((os.path.join(CSRC, s), os.path.join(OBJ, f"{os.path.splitext(s)[0]}_{prec}.o"), extra, verbose))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, rc, log in ex.map(_compile, jobs):
                if verbose or rc:
                    sys.stderr.write(f"--- {os.path.basename(src)}\n{log}\n")
                if rc:
                    raise RuntimeError(f"nvcc failed on {src}")
        for prec in precisions:
            objs = [j[1] for j in jobs if j[1].endswith(f"_{prec}.o")]
            if not objs:
                continue
            cmd = [_nvcc()] + ARCH + _host_compiler_flags() + ["-shared", "-o", outputs[prec]] + objs
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("link failed")
    return outputs


if __name__ == "__main__":
    out = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    for k, v in out.items():
        print(k, v)
