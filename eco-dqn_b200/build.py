"""Build libecodqn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Two outputs under eco-dqn_b200/lib/: the product library (every csrc/*.cu except the test kernels) and
libecodqn_b200_probe.so (csrc/tc_probe.cu: the tcgen05 building-block probe of tests/test_gpu_tc_probe.py, linked against
the product library, not part of it)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libecodqn_b200.so")
PROBE_LIB = os.path.join(LIBDIR, "libecodqn_b200_probe.so")
TEST_ONLY = ("tc_probe.cu",)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def stale():
    if not os.path.exists(LIB) or not os.path.exists(PROBE_LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(ROOT, "include", "ecodqn_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s\n" % src)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    arch = ["-gencode", "arch=compute_100a,code=sm_100a"]
    product = [o for o, src in zip(objs, sources()) if os.path.basename(src) not in TEST_ONLY]
    probe = [o for o, src in zip(objs, sources()) if os.path.basename(src) in TEST_ONLY]
    subprocess.check_call([NVCC, "-shared", "-o", LIB] + product + arch)
    subprocess.check_call([NVCC, "-shared", "-o", PROBE_LIB] + probe + arch +
                          ["-L", LIBDIR, "-lecodqn_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
