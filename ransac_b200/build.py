"""Build libusac_gpu.so for sm_100a with nvcc (in-tree; the .so travels to the GPU box with the snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libusac_gpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
         "-Xptxas", "-v", "-lcudart", "-ldl"]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(HERE, "..", "include", "usac_gpu.h")]


def build(force=False, verbose=False):
    srcs = sources()
    if not force and os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs):
        return LIB
    cmd = [NVCC] + FLAGS + ["-o", LIB, os.path.join(CSRC, "usac_gpu.cu")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = res.stdout
    with open(os.path.join(HERE, "build.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if verbose or res.returncode:
        print(log)
    if res.returncode:
        raise RuntimeError("nvcc failed, see ransac_b200/build.log")
    return LIB


if __name__ == "__main__":
    build(force="-f" in sys.argv, verbose=True)
