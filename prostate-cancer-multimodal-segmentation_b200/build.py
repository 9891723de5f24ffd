"""Builds csrc/*.cu into libb200unet3d.so (in-tree, next to this file) with nvcc for sm_100a only.

The library is rebuilt when the SHA-256 of its sources, headers and flags differs from the one recorded next to it
(`<lib>.hash`) — file times do not survive a checkout or a snapshot copy.  A build writes to a temporary file and
renames it into place under an exclusive file lock, so concurrent ranks (torchrun on a fresh checkout) never load a
half-written library and only one of them compiles.

`build(dev=True)` compiles the development variant `libb200unet3d_dev.so` with -DB200_DEV: ablation switches inside the
GEMM kernels and the tcgen05 micro-probes (csrc/probe.cu, tools/probe_*.py).  None of that is in the product library.
"""
import fcntl
import hashlib
import os
import subprocess
import tempfile

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libb200unet3d.so"
LIB_PATH = os.path.join(_HERE, LIB_NAME)
DEV_LIB_PATH = os.path.join(_HERE, "libb200unet3d_dev.so")
SOURCES = ["api.cu", "igemm.cu", "dmarch.cu", "dmarch2.cu", "wgrad_halo.cu", "conv1_march.cu", "bandwidth.cu"]
DEV_SOURCES = ["probe.cu"]
HEADERS = ["igemm.cuh", "ptx.cuh", "bandwidth.cuh", "launch.cuh", "../../include/b200_unet3d.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
] + os.environ.get("B200_EXTRA_NVCC_FLAGS", "").split()   # development aid (e.g. -DSOME_DEBUG_SWITCH); part of the hash


def _paths(dev):
    return (DEV_LIB_PATH if dev else LIB_PATH), SOURCES + (DEV_SOURCES if dev else [])


def source_hash(dev: bool = False) -> str:
    _, sources = _paths(dev)
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS + (["-DB200_DEV"] if dev else [])).encode())
    csrc = os.path.join(_HERE, "csrc")
    for f in sources + HEADERS:
        h.update(f.encode())
        with open(os.path.join(csrc, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _recorded_hash(lib):
    try:
        with open(lib + ".hash") as fh:
            return fh.read().strip()
    except OSError:
        return None


def is_stale(dev: bool = False) -> bool:
    lib, _ = _paths(dev)
    return not os.path.exists(lib) or _recorded_hash(lib) != source_hash(dev)


def build(force: bool = False, verbose: bool = False, dev: bool = False) -> str:
    """Compile the C-ABI library if it is missing or its sources changed; returns its path."""
    lib, sources = _paths(dev)
    if not force and not is_stale(dev):
        return lib
    with open(lib + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale(dev):   # another process built it while this one waited for the lock
                return lib
            want = source_hash(dev)
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            fd, tmp = tempfile.mkstemp(prefix=".build_", suffix=".so", dir=_HERE)
            os.close(fd)
            try:
                cmd = ([nvcc] + NVCC_FLAGS + (["-DB200_DEV"] if dev else []) + (["-Xptxas", "-v"] if verbose else [])
                       + ["-o", tmp] + sources)
                res = subprocess.run(cmd, cwd=os.path.join(_HERE, "csrc"), capture_output=True, text=True)
                if res.returncode != 0:
                    raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
                if verbose:
                    print(res.stderr)
                os.chmod(tmp, 0o755)
                os.replace(tmp, lib)
                with open(lib + ".hash.tmp", "w") as fh:
                    fh.write(want + "\n")
                os.replace(lib + ".hash.tmp", lib + ".hash")
            finally:
                if os.path.exists(tmp):
                    os.unlink(tmp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return lib


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose=True, dev="--dev" in sys.argv))
