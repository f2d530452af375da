"""In-tree build of libgraphaligner_b200.so: nvcc for sm_100a (cross-compiles without a GPU) + g++.
The .so lands next to this file so that it travels to the GPU box with the repo snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgraphaligner_b200.so")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")
HOST_SOURCES = ["ga_host.cpp", "alignment_graph.cpp", "bigraph_to_digraph.cpp", "vg_codec.cpp", "ga_capi.cpp", "aligner_wrapper.cpp"]
CUDA_SOURCES = ["ga_kernels.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
CXX_FLAGS = ["-std=c++17", "-O2", "-fPIC", "-Wall", "-I" + os.path.join(CUDA_HOME, "include")]
# profiling builds: GA_EXTRA_FLAGS="-DGA_PHASE_TIMING" python graphaligner_b200/build.py --force (both compilers see the flags: shared structs)
_EXTRA = os.environ.get("GA_EXTRA_FLAGS", "").split()
NVCC_FLAGS += _EXTRA
CXX_FLAGS += _EXTRA


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(verbose=False, force=False):
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "graphaligner_b200.h"))
    objs = []
    for src in CUDA_SOURCES:
        obj = os.path.join(objdir, src + ".o")
        path = os.path.join(CSRC, src)
        if force or _newer(obj, [path] + headers):
            cmd = [NVCC] + NVCC_FLAGS + ["-c", path, "-o", obj]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed for " + src)
            with open(os.path.join(objdir, src + ".ptxas.log"), "w") as f:
                f.write(res.stdout + res.stderr)
        objs.append(obj)
    for src in HOST_SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(objdir, src + ".o")
        if force or _newer(obj, [path] + headers):
            res = subprocess.run(["g++"] + CXX_FLAGS + ["-c", path, "-o", obj], capture_output=True, text=True)
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode != 0:
                raise RuntimeError("g++ failed for " + src)
        objs.append(obj)
    if force or _newer(LIB, objs):
        cmd = ["g++", "-shared", "-o", LIB] + objs + ["-L" + os.path.join(CUDA_HOME, "lib64"), "-lcudart", "-lz", "-Wl,-rpath," + os.path.join(CUDA_HOME, "lib64")]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    # the reference-compatible command line, linked against the in-tree library
    bindir = os.path.join(HERE, "bin")
    os.makedirs(bindir, exist_ok=True)
    exe = os.path.join(bindir, "Aligner")
    main_src = os.path.join(CSRC, "aligner_main.cpp")
    if force or _newer(exe, [main_src, LIB] + headers):
        cmd = ["g++"] + CXX_FLAGS + ["-o", exe, main_src, "-L" + HERE, "-lgraphaligner_b200", "-Wl,-rpath," + HERE, "-Wl,-rpath," + os.path.join(CUDA_HOME, "lib64"), "-pthread"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("Aligner link failed")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
