"""Build liblocov_b200.so (sm_100a only) in-tree with nvcc.

``python -m locov_b200.build`` or ``locov_b200.build.build()``.  The shared object lands in
``locov_b200/lib/`` (git-ignored, shipped to the GPU box with the working tree).  nvcc cross-compiles
without a GPU, so this also runs in the CPU-only build container.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# developer experiments: LOCOV_B200_LIBDIR builds / loads a variant library in another directory (relative to the package),
# LOCOV_B200_NVCC_EXTRA adds compiler flags (e.g. -DLOCOV_EXP=3) to it
LIBDIR = os.path.join(HERE, os.environ.get("LOCOV_B200_LIBDIR", "lib"))
LIBNAME = "liblocov_b200.so"
SOURCES = ["api.cu", "roi_align.cu", "misc_kernels.cu", "tc_ops.cu", "box_infer.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr",
] + os.environ.get("LOCOV_B200_NVCC_EXTRA", "").split()


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: liblocov_b200 cannot be built (there is no CPU fallback)")
    return exe


def _digest():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/locov_b200.h"]
    for f in files:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


def is_current():
    stamp = os.path.join(LIBDIR, "build.stamp")
    return os.path.exists(lib_path()) and os.path.exists(stamp) and open(stamp).read().strip() == _digest()


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the C-ABI shared library.  Returns its path.

    Safe to call from several processes at once (one rank per GPU under torchrun): the build runs under an exclusive file lock with
    per-process object / temporary names, and a process that waited for the lock re-checks the stamp instead of building again."""
    if not force and is_current():
        return lib_path()
    nvcc = _nvcc()
    os.makedirs(LIBDIR, exist_ok=True)
    import fcntl
    with open(os.path.join(LIBDIR, "build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():             # another process built it while this one waited
                return lib_path()
            return _build_locked(nvcc, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(nvcc, verbose):
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = lib_path() + f".tmp{os.getpid()}"
    r = subprocess.run([nvcc, "-shared", "-o", tmp] + objs + ["-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, lib_path())
    with open(os.path.join(LIBDIR, "build.stamp"), "w") as fh:
        fh.write(_digest())
    return lib_path()


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
