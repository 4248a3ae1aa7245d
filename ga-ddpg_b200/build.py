"""Build libgaddpg_b200.so in-tree with nvcc for sm_100a (explicit -gencode; no torch arch list involved).

The .so lands in ga-ddpg_b200/lib/ (git-ignored, but it travels to the GPU box with the gpurun snapshot).
"""
import glob
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libgaddpg_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-Xptxas", "-warn-spills", "-DGADDPG_NT_MINB=2",
]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stamp():
    h = hashlib.sha1()
    for f in _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(CSRC, "*.h"))) + [
        os.path.join(HERE, "..", "include", "gaddpg_b200.h")
    ]:
        h.update(open(f, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build_variant(name, defines):
    """Experimental build with extra -D flags into lib/libgaddpg_b200_<name>.so (selected with GADDPG_LIB)."""
    os.makedirs(LIBDIR, exist_ok=True)
    out = os.path.join(LIBDIR, "libgaddpg_b200_%s.so" % name)
    cmd = [NVCC] + FLAGS + ["-D%s" % d for d in defines] + ["-shared", "-o", out] + _sources() + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    subprocess.run(cmd, check=True)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    stamp_file = os.path.join(LIBDIR, "build.stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if out.strip():
            print("[nvcc %s]\n%s" % (os.path.basename(src), out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    # the link step gets the same -gencode: without it nvcc adds an (empty) device-link stub for its default architecture
    subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs +
                   ["-Xcompiler", "-fvisibility=hidden", "-lcudart_static", "-lpthread", "-ldl", "-lrt"], check=True)
    open(stamp_file, "w").write(stamp)
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
