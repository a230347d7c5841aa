"""Builds libcoma_b200.so in-tree with nvcc for sm_100a (B200). No JIT, no torch dependency.

    python -m coma_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("COMA_B200_LIB") or os.path.join(HERE, "libcoma_b200.so")   # override: A/B builds for tuning runs
SOURCES = ["capi.cu", "pair.cu", "orient.cu", "occupancy.cu", "nearest.cu", "nearest_dist.cu", "normals.cu", "readout.cu", "gemm.cu", "unet_ops.cu", "pipeline_ops.cu", "attention.cu", "conv_small.cu", "conv_halo.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off", "--fmad=true", "-Xptxas", "-v",
]
OBJ_DIR = os.path.join(HERE, "..", "build", "obj")


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "coma_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """One nvcc -c per translation unit (in parallel, only the stale ones), then one link into the shared library."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    env = dict(os.environ)
    env.pop("CC", None)  # the image's CC points at a wrapper nvcc should not use
    extra = os.environ.get("COMA_NVCC_EXTRA", "").split()
    tag = "" if LIB.endswith("libcoma_b200.so") and not extra else "_" + str(abs(hash((LIB, tuple(extra)))) % 10 ** 8)
    obj_dir = os.path.abspath(OBJ_DIR + tag)
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [os.path.join(HERE, "..", "include", "coma_b200.h")]
    hdr_t = max(os.path.getmtime(h) for h in headers)

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        srcp = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(srcp), hdr_t):
            return obj, "", 0
        cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-c", srcp, "-o", obj]
        r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return obj, " ".join(cmd) + "\n" + r.stdout, r.returncode

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    log = "".join(r[1] for r in results)
    if any(r[2] for r in results):
        raise RuntimeError("nvcc failed:\n" + "".join(r[1] for r in results if r[2]))
    link = [_nvcc(), "--shared", "-gencode", "arch=compute_100a,code=sm_100a"] + [r[0] for r in results] + ["-o", LIB]
    r = subprocess.run(link, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    if log:
        with open(os.path.join(HERE, "build.log"), "a" if not force else "w") as f:
            f.write(log)
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
