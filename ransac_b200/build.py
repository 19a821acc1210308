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


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: tuning variants, e.g. build(defines=["USAC_PPI=1"], out="libusac_gpu_ppi1.so") (tools/ only)."""
    srcs = sources()
    lib = os.path.join(HERE, out) if out else LIB
    if not force and os.path.exists(lib) and all(os.path.getmtime(s) <= os.path.getmtime(lib) for s in srcs):
        return lib
    cmd = [NVCC] + FLAGS + ["-D" + d for d in defines] + ["-o", lib, os.path.join(CSRC, "usac_gpu.cu")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = res.stdout
    with open(os.path.join(HERE, "build.log" if not out else out + ".log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if verbose or res.returncode:
        print(log)
    if res.returncode:
        raise RuntimeError("nvcc failed, see ransac_b200/build.log")
    return lib


if __name__ == "__main__":
    build(force="-f" in sys.argv, verbose=True)
