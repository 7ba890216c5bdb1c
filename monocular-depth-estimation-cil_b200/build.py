"""Builds libdepth_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libdepth_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + os.environ.get("DP_EXTRA_FLAGS", "").split()


def _newer(src, dst, extra=()):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(s) > t for s in (src,) + tuple(extra))


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = tuple(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    jobs = []
    objs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _newer(s, o, hdrs):
            jobs.append([NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(" ".join(cmd[-3:]) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {cmd[-3]}")
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    build_probe(verbose)
    with open(os.path.join(HERE, "build_id.txt"), "w") as f:
        f.write(source_id() + "\n")
    import json
    with open(os.path.join(HERE, "build_files.json"), "w") as f:
        json.dump(source_ids(), f, indent=1, sort_keys=True)
    return LIB


def source_ids():
    """per-file sha1 of the kernel sources (a capture stays valid for a kernel as long as ITS files are unchanged)"""
    import hashlib
    out = {}
    for f in sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                    glob.glob(os.path.join(HERE, "..", "include", "*.h"))):
        with open(f, "rb") as fh:
            out[os.path.basename(f)] = hashlib.sha1(fh.read()).hexdigest()[:12]
    return out


def source_id():
    """sha1 over the kernel sources: profiles/*.json captured with ncu carry it, and bench.py pairs a timing with a
    `traffic` figure only when the capture was taken from the same sources"""
    import hashlib
    h = hashlib.sha1()
    files = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                   glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:12]


def build_probe(verbose=False):
    """tests/probe/libdepth_b200_probe.so: the single-UMMA descriptor probe used by tests/test_umma_probe_gpu.py.  Test
    scaffolding - it links the product objects it needs (error string, tensor-map encoder) but is not part of the
    product library."""
    pdir = os.path.join(HERE, "..", "tests", "probe")
    src = os.path.join(pdir, "umma_probe.cu")
    out = os.path.join(pdir, "libdepth_b200_probe.so")
    if not os.path.exists(src):
        return None
    deps = [os.path.join(OBJ, "dp_core.o"), os.path.join(OBJ, "tmap.o")]
    hdrs = tuple(glob.glob(os.path.join(CSRC, "*.cuh"))) + (os.path.join(pdir, "probe.h"),)
    if _newer(src, out, hdrs + tuple(deps)):
        cmd = [NVCC] + FLAGS + ["-shared", src] + deps + ["-o", out, "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("probe build failed")
    return out


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
