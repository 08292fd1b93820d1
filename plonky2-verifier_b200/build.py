"""In-tree build of libp2v.so (CUDA kernels for sm_100a + C++ host side + the C ABI).

    python plonky2-verifier_b200/build.py [--force] [-v]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.
"""
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "libp2v.so")
OBJ = os.path.join(HERE, "build")
APPS = os.path.join(HERE, "apps")
# host programs above the C ABI (apps/<name>.cpp -> p2v_<name>): testmain = the reference's driver (src/testmain.hs),
# shard_demo = BASELINE config 5 from a host without Python or torch (one process per GPU, NCCL bootstrapped through the C ABI)
APP_NAMES = ["testmain", "shard_demo"]
APP_BINS = [os.path.join(HERE, "p2v_" + a) for a in APP_NAMES]
TESTMAIN = APP_BINS[0]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"),
]
NVCC_FLAGS += os.environ.get("P2V_EXTRA_NVCC", "").split()  # tuning experiments only (e.g. -DPOSEIDON_SBOX_GROUP=12)
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-I", os.path.join(ROOT, "include")]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libp2v cannot be built")


def _sources():
    cu, cpp = [], []
    for dirpath, _, files in os.walk(CSRC):
        for f in sorted(files):
            p = os.path.join(dirpath, f)
            if f.endswith(".cu"):
                cu.append(p)
            elif f.endswith(".cpp"):
                cpp.append(p)
    return cu, cpp


def _deps_hash():
    h = hashlib.sha256()
    for dirpath, _, files in sorted(os.walk(CSRC)):
        for f in sorted(files):
            if f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".h")):
                with open(os.path.join(dirpath, f), "rb") as fh:
                    h.update(f.encode() + fh.read())
    for f in sorted(os.listdir(APPS)):
        with open(os.path.join(APPS, f), "rb") as fh:
            h.update(f.encode() + fh.read())
    with open(os.path.join(ROOT, "include", "p2v.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS + CXX_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every translation unit under csrc/ and link libp2v.so.  Idempotent (content hash)."""
    stamp = os.path.join(OBJ, "stamp")
    want = _deps_hash()
    if not force and os.path.exists(OUT) and all(os.path.exists(b) for b in APP_BINS) and os.path.exists(stamp) and open(stamp).read() == want:
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    cu, cpp = _sources()
    jobs = []
    for src in cu:
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        jobs.append((src, obj, [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]))
    for src in cpp:
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        jobs.append((src, obj, ["g++"] + CXX_FLAGS + ["-I", "/usr/local/cuda/include", "-c", src, "-o", obj]))

    def run(job):
        src, obj, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("compile failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        objs = list(ex.map(run, jobs))
    link = [nvcc, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    # host programs above the C ABI: plain g++, include/p2v.h only, libp2v.so found next to the binary
    for name, out in zip(APP_NAMES, APP_BINS):
        app = ["g++"] + CXX_FLAGS + [os.path.join(APPS, name + ".cpp"), "-o", out, "-L", HERE, "-lp2v", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(app, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("p2v_%s failed to build:\n" % name + r.stdout + r.stderr)
    with open(stamp, "w") as fh:
        fh.write(want)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
