"""In-tree build of the CUDA library (liblocr.so) and the oracle's C helpers.

nvcc cross-compiles for sm_100a without a GPU; the resulting .so files stay in-tree (git-ignored) so that they travel
to the GPU box with the repository snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "liblocr.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
] + os.environ.get("LOCR_NVCC_EXTRA", "").split()
# Translation units whose float arithmetic must match cv2 / PIL / numpy bit for bit: no FMA contraction.
EXACT_UNITS = {"postproc.cu", "imgops.cu", "polys.cu"}


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the CUDA path cannot be built (there is no CPU fallback)")
    return nvcc


def _digest(paths, extra):
    h = hashlib.sha256()
    h.update(" ".join(extra).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def build(verbose=False, force=False):
    os.makedirs(BUILD, exist_ok=True)
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    nvcc = _nvcc()
    stamp_path = os.path.join(BUILD, "stamp")
    stamp = _digest([os.path.join(CSRC, s) for s in sources] + headers, NVCC_FLAGS)
    if not force and os.path.exists(LIB) and os.path.exists(stamp_path) and open(stamp_path).read() == stamp:
        return LIB

    def compile_one(src):
        obj = os.path.join(BUILD, src[:-3] + ".o")
        flags = list(NVCC_FLAGS)
        if src in EXACT_UNITS:
            flags += ["--fmad=false"]
        cmd = [nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose and r.stderr.strip():
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                                                 "-lz"]      # zlib: inflate + CRC-32 of the PNG reader (png.cu)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp_path, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
