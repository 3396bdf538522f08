#!/usr/bin/env python
"""A/B builds of the fp32 library: recompile the named sources with extra nvcc flags and link them with the current
objects of the other sources into ab/<name>/libtfmpc_b200.so (selected at run time with TFMPC_B200_LIBDIR=ab/<name>).

    python scripts/build_variant.py wps20 -DTFMPC_QUEUE_MAXWPS=20 [--src ilqr_queue.cu]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tfmpc_b200 import build as B  # noqa: E402


def main():
    name, rest = sys.argv[1], sys.argv[2:]
    srcs = ["ilqr_queue.cu"]
    if "--src" in rest:
        i = rest.index("--src")
        srcs = rest[i + 1].split(",")
        rest = rest[:i] + rest[i + 2:]
    B.build(precisions=("f32",))
    out_dir = os.path.join(ROOT, "ab", name)
    os.makedirs(out_dir, exist_ok=True)
    objs = []
    for s in B.SOURCES:
        base = os.path.splitext(s)[0]
        if s in srcs:
            obj = os.path.join(out_dir, base + ".o")
            cmd = [B._nvcc()] + B.ARCH + B.COMMON + B._host_compiler_flags() + B.F32_FLAGS.split() + rest + ["-Xptxas", "-v", "-c", os.path.join(B.CSRC, s), "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
                raise SystemExit(1)
            for line in (r.stdout + r.stderr).splitlines():
                if "k_queue_solveILi101ELi2ELi2ELi2" in line and "Compiling" in line:
                    want = True
                elif "Used" in line and locals().get("want"):
                    print(name, line.strip())
                    want = False
                elif "spill" in line and locals().get("want"):
                    print(name, line.strip())
        else:
            obj = os.path.join(B.OBJ, f"{base}_f32.o")
        objs.append(obj)
    lib = os.path.join(out_dir, "libtfmpc_b200.so")
    subprocess.check_call([B._nvcc()] + B.ARCH + B._host_compiler_flags() + ["-shared", "-o", lib] + objs)
    print("built", lib)


if __name__ == "__main__":
    main()
