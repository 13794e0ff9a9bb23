"""ctypes binding of libfissure_b200.so — the C ABI declared in include/fissure_b200.h.

The prototypes are parsed from the header itself, so the Python side cannot drift from the C
declarations. There is no CPU fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, "..", "include", "fissure_b200.h")
LIB_PATH = os.path.join(_HERE, "libfissure_b200.so")

FS_F32 = 0
FS_BF16 = 1
FS_MAX_K = 127

_CTYPES = {
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "size_t": ctypes.c_size_t,
    "fs_stream_t": ctypes.c_void_p,
}


def parse_header(path=HEADER):
    """Return {name: (restype, [(ctype, argname), ...])} for every prototype in the header."""
    with open(path) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char\*|int|size_t)\s+(fs_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        arglist = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    arglist.append((ctypes.c_void_p, a.split("*")[-1].strip()))
                else:
                    ty, nm = a.rsplit(" ", 1)
                    arglist.append((_CTYPES[ty.replace("const ", "").strip()], nm))
        restype = {"int": ctypes.c_int, "size_t": ctypes.c_size_t}.get(ret, ctypes.c_char_p)
        protos[name] = (restype, arglist)
    return protos


_lib = None
_protos = None


def load():
    """Load the shared library (building is the job of build.py / __graft_entry__.build)."""
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libfissure_b200.so is missing (%s). Build it with `python -m fissure_segmentation_b200.build`; "
            "there is no CPU or PyTorch fallback for the CUDA hot path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    protos = parse_header()
    for name, (restype, args) in protos.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = [a[0] for a in args]
    _lib, _protos = lib, protos
    return lib


def error_string(code):
    return load().fs_error_string(int(code)).decode()


def dtype_code(t):
    if t.dtype == torch.float32:
        return FS_F32
    if t.dtype == torch.bfloat16:
        return FS_BF16
    raise TypeError("fissure_b200 kernels take float32 or bfloat16 tables, got %s" % t.dtype)


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x.data_ptr()
    return x


# kernels launched per entry point (for bench.py's gpu_launches claim)
_LAUNCHES = {"fs_knn_feat": 2, "fs_knn_feat_tc": 5, "fs_knn3d_tc": 5, "fs_edge2_fwd": 2, "fs_edge2_bwd": 3, "fs_edge3_bn_coef": 2, "fs_edge3_bwd": 2, "fs_bn_act_bwd": 2,
             "fs_pool_reduce": 2, "fs_colsum": 2, "fs_pool_lin_bwd_dx_sparse": 3, "fs_reverse_graph": 3}
launch_count = 0          # kernels of this library launched so far in this process
timed = {}                # name -> list of (start_event, end_event); filled only for names in `time_calls`
time_calls = set()


def call(name, ref, *args):
    """Invoke fs_<name>(device, stream, *args) on the device / current stream of tensor `ref`."""
    global launch_count
    lib = load()
    if not ref.is_cuda:
        raise RuntimeError("fissure_b200.%s needs CUDA tensors; there is no CPU path" % name)
    dev = ref.device.index
    stream = torch.cuda.current_stream(dev)
    timing = name in time_calls
    if timing:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
    rc = getattr(lib, name)(dev, stream.cuda_stream, *[_ptr(a) for a in args])
    if rc != 0:
        raise RuntimeError("%s failed: %s (code %d)" % (name, error_string(rc), rc))
    launch_count += _LAUNCHES.get(name, 1)
    if timing:
        e1.record(stream)
        timed.setdefault(name, []).append((e0, e1))
