"""Builds libmtgp_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m multitreegp_b200.build [--verbose] [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libmtgp_b200.so")
SOURCES = ["mtgp_api.cu", "mtgp_encode.cu", "mtgp_rollout.cu", "mtgp_rollout_rk.cu", "mtgp_pointwise.cu",
           "mtgp_select.cu"]
HEADERS = ["mtgp_internal.cuh", "mtgp_program.cuh", "mtgp_threaded.cuh", "mtgp_threaded_r8.cuh", "mtgp_env.cuh", "mtgp_lower.cuh", "../../include/mtgp.h", "../../include/mtgp_math.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # the numerical contract (include/mtgp_math.h): no implicit a*b+c contraction, IEEE div/sqrt, no FTZ
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


# The rollout kernels' device code stays uncompressed in the fat binary: sass_patch.py (jump-table merge of the
# threaded-code interpreter) patches the embedded cubin in place.
PER_SOURCE_FLAGS = {"mtgp_rollout.cu": ["--no-compress"], "mtgp_rollout_rk.cu": ["--no-compress"],
                    "mtgp_pointwise.cu": ["--no-compress"]}
PATCHED_CUBINS = ("mtgp_rollout", "mtgp_pointwise")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(os.path.join(_HERE, "build"), exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(_HERE, "build", src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + PER_SOURCE_FLAGS.get(src, []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with ThreadPoolExecutor(max_workers=4) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr + "\n")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {cmd[-3]}")
    if force or jobs or _stale(LIB, objs):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
        from .sass_patch import patch_library
        st = patch_library(LIB, PATCHED_CUBINS, verbose=verbose)
        if not st.get("patched"):
            raise RuntimeError("sass_patch: no jump table was merged -- the interpreter's dispatch sites were not found")
    build_ffi_shim(verbose)
    return LIB


def build_ffi_shim(verbose: bool = False) -> str | None:
    """Builds libmtgp_b200_ffi.so (csrc/mtgp_xla_ffi.cc) when jax's XLA-FFI headers are available;
    returns None otherwise (this image has no jax)."""
    try:
        import jax
        inc = jax.ffi.include_dir()
    except Exception:  # noqa: BLE001 -- jax missing or too old: the shim is optional
        return None
    out = os.path.join(_HERE, "libmtgp_b200_ffi.so")
    src = os.path.join(CSRC, "mtgp_xla_ffi.cc")
    if _stale(out, [src, LIB]):
        cmd = [_nvcc(), "-x", "cu", "-std=c++17", "-O2", "-Xcompiler", "-fPIC", "-shared", "-I", inc, src, "-o", out,
               "-L", _HERE, "-lmtgp_b200", "-Xlinker", "-rpath", "-Xlinker", _HERE]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("building the XLA-FFI shim failed")
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
