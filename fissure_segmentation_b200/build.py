"""In-tree nvcc build of libfissure_b200.so (sm_100a only).

The library is a plain C-ABI shared object (include/fissure_b200.h); it links the static CUDA
runtime, so it loads on a machine without a GPU (symbol checks) and travels with the source tree.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfissure_b200.so")
STAMP = os.path.join(HERE, "csrc", ".build_stamp")
SOURCES = ["knn.cu", "knn_tc.cu", "edgeconv.cu", "edgeconv_smem.cu", "edge3.cu", "edge2.cu", "dense.cu", "heads.cu", "pool_gemm.cu", "chamfer.cu", "pointops.cu", "infer.cu", "microbench.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def _source_files():
    files = [os.path.join(CSRC, f) for f in SOURCES if os.path.exists(os.path.join(CSRC, f))]
    headers = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "fissure_b200.h"))
    return files, headers


def _digest(paths):
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into objects and link libfissure_b200.so. Returns the path."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    files, headers = _source_files()
    digest = _digest(files + headers)
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == digest:
                return LIB
    if not os.path.exists(nvcc):
        if os.path.exists(LIB):
            return LIB  # GPU box without a toolkit: use the prebuilt library that travelled with the tree
        raise RuntimeError("nvcc not found and no prebuilt libfissure_b200.so present")
    objs = []
    procs = []
    logs = []
    for src in files:
        obj = src[:-3] + ".o"
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        logs.append("==== %s\n%s" % (os.path.basename(src), out))
        if p.returncode != 0:
            failed = True
    log_text = "\n".join(logs)
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write(log_text)
    if failed:
        sys.stderr.write(log_text)
        raise RuntimeError("nvcc failed; see csrc/build.log")
    if verbose:
        print(log_text)
    link = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    subprocess.check_call(link)
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
