"""Builds csrc/*.cu into libb200unet3d.so (in-tree, next to this file) with nvcc for sm_100a only."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libb200unet3d.so"
LIB_PATH = os.path.join(_HERE, LIB_NAME)
SOURCES = ["api.cu", "igemm.cu", "dmarch.cu", "wgrad_halo.cu", "bandwidth.cu", "probe.cu"]
HEADERS = ["igemm.cuh", "ptx.cuh", "bandwidth.cuh", "../../include/b200_unet3d.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    csrc = os.path.join(_HERE, "csrc")
    return any(os.path.getmtime(os.path.join(csrc, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the C-ABI library if it is missing or older than its sources; returns its path."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, cwd=os.path.join(_HERE, "csrc"), capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
