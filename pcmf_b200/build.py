"""Builds libpcmf_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

One translation unit per padded factor width (kernels_kp.cu -DPCMF_KP=n) + solver.cu + datagen.cu, compiled in
parallel, linked into pcmf_b200/libpcmf_b200.so.  The .so is git-ignored but travels to the GPU box with gpurun.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
SO = os.path.join(HERE, "libpcmf_b200.so")
KPS = [4, 8, 12, 16, 20, 32, 64]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# PCMF_EXTRA_NVCC="-DPCMF_GENE_MINB=3 ..." adds flags for kernel tuning experiments
FLAGS = os.environ.get("PCMF_EXTRA_NVCC", "").split() + ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-ccbin", "/usr/bin/g++"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build(force=False, verbose=False, ptxas_info=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = sorted(os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith(".cuh")) + \
              [os.path.join(os.path.dirname(HERE), "include", "pcmf_b200.h")]
    jobs = []
    objs = []
    extra = ["-Xptxas", "-v"] if ptxas_info else []
    for kp in KPS:
        o = os.path.join(OBJ, "kernels_kp%d.o" % kp)
        objs.append(o)
        src = os.path.join(CSRC, "kernels_kp.cu")
        if force or _newer(o, [src] + headers):
            jobs.append([NVCC] + FLAGS + extra + ["-DPCMF_KP=%d" % kp, "-c", src, "-o", o])
    for name in ("solver", "datagen", "multi"):
        o = os.path.join(OBJ, name + ".o")
        objs.append(o)
        src = os.path.join(CSRC, name + ".cu")
        if force or _newer(o, [src] + headers):
            jobs.append([NVCC] + FLAGS + extra + ["-c", src, "-o", o])
    logs = []
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(_run, jobs):
                logs.append(out)
                if verbose and out.strip():
                    print(out)
    if jobs or not os.path.exists(SO):
        _run([NVCC, "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-ldl"])
    return SO, logs


if __name__ == "__main__":
    so, logs = build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv)
    print("built", so)
