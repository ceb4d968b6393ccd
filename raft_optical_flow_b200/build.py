"""Compiles raft_optical_flow_b200/csrc/*.cu into libraftcorr_b200.so (in-tree, next to this file).

    python -m raft_optical_flow_b200.build [--force] [--debug]

--debug additionally compiles libraftcorr_b200_debug.so with -DRCB_DEBUG: the same kernels plus the timing /
profiling hooks that tools/time_*.py drive through environment variables (RCB_TC_DEBUG_SKIP, RCB_TC_PROF_PTR,
RCB_LOOKUP_DEBUG, RCB_LCONV_* ...).  The release library ignores the environment entirely; the Python binding loads
the debug library only when RCB_USE_DEBUG_LIB=1 is set.

sm_100a only (B200): nvcc cross-compiles without a GPU.  The library exports the C ABI declared in
include/raft_corr_b200.h and nothing else; it links the CUDA runtime statically and resolves the one
driver symbol it needs (cuTensorMapEncodeTiled) through cudaGetDriverEntryPoint at run time.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libraftcorr_b200.so")
DEBUG_LIB = os.path.join(HERE, "libraftcorr_b200_debug.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-shared",
    "-ccbin", "/usr/bin/g++",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build(lib=LIB):
    if not os.path.exists(lib):
        return True
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps)


def build(force=False, verbose=False, debug=False, variant=None, defines=()):
    """variant="name" + defines=["-DRCB_X=1", ...] compiles libraftcorr_b200_<name>.so (a debug-hook build with
    compile-time switches, for A/B timing inside one GPU session; select it with RCB_LIB_VARIANT=name)."""
    lib = DEBUG_LIB if debug else LIB
    if variant:
        lib = os.path.join(HERE, f"libraftcorr_b200_{variant}.so")
        debug = True
    if not force and not needs_build(lib):
        return lib
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-DRCB_DEBUG"] if debug else []) + list(defines) + \
        (["-Xptxas", "-v"] if verbose else []) + ["-o", lib] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--debug" in sys.argv:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug=True))
    if "--variant" in sys.argv:
        name = sys.argv[sys.argv.index("--variant") + 1]
        print(build(force=True, verbose="-v" in sys.argv, variant=name, defines=[a for a in sys.argv if a.startswith("-D")]))
