#!/usr/bin/env python3
"""Build compile-time variants of libp2v.so HERE (CPU container, nvcc cross-compiles) so that a GPU box only has to
run them:  python tools/build_variants.py tag1="-DFOO=1 -DBAR=2" tag2="..."   ->  plonky2-verifier_b200/variants/libp2v_<tag>.so
(the directory is git-ignored through *.so but travels with gpurun).  Load one with P2V_LIB_PATH=<path>."""
import concurrent.futures as cf
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "plonky2-verifier_b200")
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "variants")
BASE = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include")]


def build(tag, flags):
    d = os.path.join(OUT, "obj_" + tag)
    os.makedirs(d, exist_ok=True)
    objs = []
    procs = []
    for f in sorted(os.listdir(CSRC)):
        if f.endswith(".cu"):
            o = os.path.join(d, f + ".o")
            objs.append(o)
            procs.append(subprocess.Popen(["nvcc"] + BASE + flags.split() + ["-c", os.path.join(CSRC, f), "-o", o], stderr=subprocess.PIPE, text=True))
    for f in sorted(os.listdir(os.path.join(CSRC, "host"))):
        if f.endswith(".cpp"):
            o = os.path.join(d, f + ".o")
            objs.append(o)
            procs.append(subprocess.Popen(["g++", "-O2", "-std=c++17", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", "-c",
                                           os.path.join(CSRC, "host", f), "-o", o], stderr=subprocess.PIPE, text=True))
    for p in procs:
        err = p.communicate()[1]
        if p.returncode:
            return tag, "FAILED: " + err[-2000:]
    so = os.path.join(OUT, "libp2v_%s.so" % tag)
    r = subprocess.run(["nvcc", "-shared", "-o", so] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"], capture_output=True, text=True)
    if r.returncode:
        return tag, "LINK FAILED: " + r.stderr[-2000:]
    for o in objs:
        os.remove(o)
    os.rmdir(d)
    return tag, so


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    jobs = [a.split("=", 1) for a in sys.argv[1:]]
    with cf.ThreadPoolExecutor(max_workers=3) as ex:
        for tag, res in ex.map(lambda j: build(j[0], j[1]), jobs):
            print(tag, "->", res)
