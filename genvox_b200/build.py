"""Build the C-ABI CUDA library in-tree: genvox_b200/lib/libgenvox_b200.so (sm_100a only).

    python -m genvox_b200.build [--force] [--verbose]

The built .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libgenvox_b200.so")
STAMP = os.path.join(LIB_DIR, "build.stamp")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "--cudart", "shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint():
    """Hash of the sources and flags.  File NAMES only, not absolute paths: the repo is snapshotted to a different
    directory on the GPU box, and a path-dependent stamp made every process there rebuild the library - under torchrun
    several ranks at once, racing on the same .so (a rank then loaded a half-written file)."""
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(INCLUDE, "genvox_b200.h")]
    for p in files:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _marker(fingerprint):
    return ("GVXFP:" + fingerprint).encode()


def is_fresh():
    """True when the in-tree library was compiled from exactly the sources in the tree: the source fingerprint is
    compiled INTO the library (-DGVX_BUILD_FINGERPRINT, gvx_api.cu) and looked up in the file, so a library left over
    from other sources (e.g. after `git checkout` of the sources and of build.stamp) is never mistaken for a fresh one."""
    if not (os.path.isfile(LIB_PATH) and os.path.isfile(STAMP)):
        return False
    fp = _fingerprint()
    with open(STAMP) as fh:
        if fh.read().strip() != fp:
            return False
    with open(LIB_PATH, "rb") as fh:
        return _marker(fp) in fh.read()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library.  Returns the library path.
    Safe against concurrent callers (one process per GPU under torchrun): an exclusive file lock serialises builders, the
    freshness check is repeated under the lock, objects and the library are written to process-private names and the
    library is moved into place atomically."""
    if not force and is_fresh():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    import fcntl
    with open(os.path.join(LIB_DIR, "build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and is_fresh():        # another process built it while this one waited
            return LIB_PATH
        return _build_locked(verbose)


def _build_locked(verbose):
    tag = f".{os.getpid()}"
    objs = []
    fp_define = f'-DGVX_BUILD_FINGERPRINT="{_fingerprint()}"'
    for src in sources():
        obj = os.path.join(LIB_DIR, os.path.basename(src)[:-3] + tag + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + [fp_define] + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE, "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        objs.append(obj)
    tmp_lib = LIB_PATH + tag
    cmd = [_nvcc(), "-shared", "--cudart", "shared", "-o", tmp_lib] + objs + ["-L/usr/local/cuda/lib64", "-lcublas"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    for obj in objs:
        try:
            os.remove(obj)
        except OSError:
            pass
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp_lib, LIB_PATH)           # atomic: a concurrent loader sees the old or the new library, never a partial one
    tmp_stamp = STAMP + tag
    with open(tmp_stamp, "w") as fh:
        fh.write(_fingerprint())
    os.replace(tmp_stamp, STAMP)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
