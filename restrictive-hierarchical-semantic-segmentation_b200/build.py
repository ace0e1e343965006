"""In-tree build of librhseg_b200.so (nvcc, sm_100a only).  No torch headers involved: the
library is a plain C-ABI shared object (include/rhseg_b200.h) loaded with ctypes."""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
LIB_PATH = os.path.join(PKG_DIR, "librhseg_b200.so")
OBJ_DIR = os.path.join(PKG_DIR, "build")
STAMP = os.path.join(PKG_DIR, "librhseg_b200.sha256")  # next to the library: it travels to the GPU box with it

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-I", INCLUDE, "-I", CSRC,
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build librhseg_b200.so")
    return exe


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, INCLUDE):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current():
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == _digest()


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link librhseg_b200.so.  Returns the path."""
    if not force and is_current():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = _sources()

    headers = hashlib.sha256()
    for root in (CSRC, INCLUDE):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    headers.update(f.encode() + fh.read())
    headers.update(" ".join(NVCC_FLAGS).encode())

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        # per-object stamp (source + every header + flags): unchanged translation units are not recompiled
        h = headers.copy()
        with open(os.path.join(CSRC, src), "rb") as fh:
            h.update(fh.read())
        stamp = obj + ".sha256"
        if not force and not verbose and os.path.exists(obj) and os.path.exists(stamp):
            with open(stamp) as fh:
                if fh.read().strip() == h.hexdigest():
                    return obj
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        with open(stamp, "w") as fh:
            fh.write(h.hexdigest())
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, srcs))
    r = subprocess.run([nvcc, "-shared", "-o", LIB_PATH] + objs + ["-cudart", "static"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(STAMP, "w") as f:
        f.write(_digest())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
