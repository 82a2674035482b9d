"""Builds libmjhmc_b200.so in-tree with nvcc for sm_100a (no JIT, no torch extension).

    python -m mjhmc_b200.build [--force]

The objects are compiled in parallel; the fused-kernel instantiation unit is compiled
once per (dtype, dimension group) so the template instantiations do not serialise.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.environ.get("MJHMC_B200_BUILD_DIR") or os.path.join(HERE, "_build")
LIB = os.environ.get("MJHMC_B200_BUILD_OUT") or os.path.join(HERE, "libmjhmc_b200.so")
EXTRA = os.environ.get("MJHMC_B200_EXTRA_FLAGS", "").split()

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC"]

DIM_GROUPS = [(1, 2), (3, 4), (6, 8), (10, 16)]
STREAM_DTS = [2, 4, 6, 8, 10, 13, 16]     # dims per thread of the streaming kernels (one unit each)


def _units():
    units = []
    for name in ("api", "unfused", "analysis", "autocorr_fft", "dense", "dense_tc", "stream"):
        units.append((name, os.path.join(CSRC, name + ".cu"), []))
    for tname, ctype in (("f64", "double"), ("f32", "float")):
        for g, (da, db) in enumerate(DIM_GROUPS):
            tag = "%s_g%d" % (tname, g)
            units.append(("fused_inst_" + tag, os.path.join(CSRC, "fused_inst.cu"),
                          ["-DMJ_T=" + ctype, "-DMJ_TAG=" + tag, "-DMJ_DA=%d" % da, "-DMJ_DB=%d" % db]))
    for tname, ctype in (("f64", "double"), ("f32", "float")):
        for g, da in enumerate(STREAM_DTS):
            tag = "%s_g%d" % (tname, g)
            units.append(("stream_inst_" + tag, os.path.join(CSRC, "stream_inst.cu"),
                          ["-DMJ_T=" + ctype, "-DMJ_TAG=" + tag, "-DMJ_DA=%d" % da]))
    return units


def _source_digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for fn in sorted(os.listdir(root)):
            if fn.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, fn), "rb") as f:
                    h.update(fn.encode())
                    h.update(f.read())
    h.update(" ".join(FLAGS + EXTRA).encode())
    return h.hexdigest()


def _compile(unit):
    name, src, defs = unit
    obj = os.path.join(BUILD, name + ".o")
    cmd = [NVCC] + FLAGS + EXTRA + defs + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (name, " ".join(cmd), r.stderr))
    return obj


def build(force=False, verbose=True):
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.splitext(LIB)[0] + ".digest"
    digest = _source_digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError("nvcc not found at %s and no up-to-date %s present" % (NVCC, LIB))
    units = _units()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(_compile, units))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stderr)
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
