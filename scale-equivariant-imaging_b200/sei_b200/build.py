"""Build libsei_b200.so (hand-written sm_100a CUDA behind a C ABI) in-tree with nvcc.

    python -m sei_b200.build        (with scale-equivariant-imaging_b200/ on sys.path)

The library has no torch dependency; it is loaded with ctypes (sei_b200._lib).
"""
import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libsei_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]
OBJ_DIR = os.path.join(LIB_DIR, "obj")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(PKG_DIR, "..", "include", "sei_b200.h")]
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile_one(args):
    nvcc, src, obj, verbose = args
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src, res.returncode, res.stdout + res.stderr


def build(force=False, verbose=False):
    """one object per .cu, compiled in parallel (only the stale ones), then one link step"""
    if not force and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(PKG_DIR, "..", "include", "sei_b200.h")]
    newest_header = max(os.path.getmtime(h) for h in headers)
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header)
        if stale:
            jobs.append((nvcc, src, obj, verbose))
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 4))) as ex:
        for src, rc, out in ex.map(_compile_one, jobs):
            if rc != 0:
                sys.stderr.write(out)
                raise RuntimeError("nvcc failed compiling " + src)
            if verbose:
                sys.stderr.write(out)
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs,
                         capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libsei_b200.so")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
