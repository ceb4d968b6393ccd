"""Compile the REFERENCE's own alt_cuda_corr extension into oracle/_ref/ (test infrastructure).

The two source files are compiled where they lie under /root/reference/alt_cuda_corr (nothing is
copied into this repository) with explicit g++/nvcc command lines -- the reference's setup.py is not
run.  The result, oracle/_ref/alt_cuda_corr.so, is git-ignored but travels to the GPU box, where
tests/test_gpu_parity.py loads it as the on-device oracle for alt_cuda_corr.forward/backward and
bench.py --ref-gpu times it ("the same-box kernel to beat").  It can only *run* on a GPU.

Needs: torch headers/libs (present in the image), nvcc, g++.  Exits 0 and prints a note when the
reference checkout is absent (e.g. on the GPU box).
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RAFT_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "alt_cuda_corr")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "alt_cuda_corr.so")


def main():
    if not os.path.isdir(SRC):
        print(f"build_ref: {SRC} not present; keeping whatever is in {OUT_DIR}")
        return 0
    srcs = [os.path.join(SRC, "correlation.cpp"), os.path.join(SRC, "correlation_kernel.cu")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs):
        print(f"build_ref: {OUT} up to date")
        return 0
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = os.path.join(OUT_DIR, "obj")
    os.makedirs(tmp, exist_ok=True)
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    abi = f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"
    defs = ["-DTORCH_EXTENSION_NAME=alt_cuda_corr", "-DTORCH_API_INCLUDE_EXTENSION_H", abi]
    cxx = "/usr/bin/g++"
    o_cpp, o_cu = os.path.join(tmp, "correlation.o"), os.path.join(tmp, "correlation_kernel.o")
    cmds = [
        [cxx, "-c", srcs[0], "-o", o_cpp, "-O2", "-fPIC", "-std=c++17", "-w"] + defs + inc,
        ["nvcc", "-c", srcs[1], "-o", o_cu, "-O3", "-std=c++17", "-w", "-ccbin", cxx,
         "-gencode", "arch=compute_100,code=sm_100", "--expt-relaxed-constexpr",
         "-Xcompiler", "-fPIC"] + defs + inc,
        [cxx, "-shared", o_cpp, o_cu, "-o", OUT] + [f"-L{p}" for p in ce.library_paths("cuda")] +
        ["-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart"],
    ]
    for c in cmds:
        print("build_ref:", " ".join(c[:4]), "...")
        subprocess.run(c, check=True)
    print(f"build_ref: wrote {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
