"""Build libhctr_b200.so (hand-written sm_100a CUDA kernels + C ABI) in-tree with nvcc.

The library is built next to this file so that it travels with the repository snapshot to the GPU
box; nothing is JIT-compiled at run time.  `python -m hctr_b200.build` or `build_native()`.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(CSRC, "_build")
LIB_PATH = os.path.join(PKG_DIR, "libhctr_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(PKG_DIR), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(src, digest, verbose):
    obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
    stamp = obj + ".stamp"
    with open(os.path.join(CSRC, src), "rb") as fh:
        key = hashlib.sha256(fh.read() + digest.encode()).hexdigest()
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == key:
        return obj, False, ""
    cmd = [NVCC, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(stamp, "w") as fh:
        fh.write(key)
    with open(obj + ".ptxas.log", "w") as fh:
        fh.write(r.stderr)
    return obj, True, r.stderr if verbose else ""


def build_native(verbose=False, force=False):
    """Compile every .cu under csrc/ for sm_100a and link libhctr_b200.so. Returns the library path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    digest = _deps_digest()
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile_one(s, digest, verbose), srcs))
    objs = [r[0] for r in results]
    rebuilt = any(r[1] for r in results)
    if verbose:
        for r in results:
            if r[2]:
                sys.stderr.write(r[2])
    if rebuilt or not os.path.exists(LIB_PATH):
        cmd = [NVCC, "-shared", "-o", LIB_PATH, *objs, "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB_PATH


if __name__ == "__main__":
    path = build_native(verbose="-v" in sys.argv, force="--force" in sys.argv)
    print(path)
