"""ctypes binding of libgaot_b200.so (include/gaot_b200.h).  There is no fallback: if the library is
missing or a call fails, the caller gets an exception."""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgaot_b200.so")

_P, _I, _L, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
_KIND = {"p": _P, "i": _I, "l": _L, "f": _F}

HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "gaot_b200.h")


def parse_header(path: str = HEADER_PATH):
    """{function name: argument kinds} read from include/gaot_b200.h (p pointer, i int, l int64,
    f float).  The header is the single source of truth for the ABI; tests check that the library
    exports every name found here."""
    import re
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int|long long)\s+(gaot_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        kinds = ""
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    kinds += "p"
                elif a.startswith("int64_t") or a.startswith("long long"):
                    kinds += "l"
                elif a.startswith("float"):
                    kinds += "f"
                elif a.startswith("int"):
                    kinds += "i"
                else:
                    raise ValueError(f"unparsed argument {a!r} of {name}")
        out[name] = (ret, kinds)
    return out


_lib = None


class GaotLibraryError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GaotLibraryError(
                f"{LIB_PATH} is missing - build it with `python -m gaot_b200.build` (there is no CPU or eager fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (ret, sig) in parse_header().items():
            fn = getattr(L, name)
            fn.argtypes = [_KIND[c] for c in sig]
            fn.restype = ctypes.c_longlong if ret == "long long" else ctypes.c_int
        if L.gaot_abi_version() != 1:
            raise GaotLibraryError("libgaot_b200.so ABI version mismatch - rebuild")
        _lib = L
    return _lib


def launch_count() -> int:
    return int(lib().gaot_launch_count())


def note_launches(n: int):
    """Account for kernels that ran through a CUDA-graph replay (they bypass the per-call counter)."""
    lib().gaot_launch_counter_add(int(n))


_dummies = {}


def ptr(t):
    """Raw address for the ABI.  None -> NULL (optional argument).  A zero-element tensor (e.g. the edge arrays of a graph
    without any edge) has data_ptr() == 0 in torch; the library treats NULL as "argument missing", so it gets the address
    of a small per-device dummy buffer instead - it is never dereferenced because the matching size is 0."""
    if t is None:
        return 0
    if t.numel() == 0:
        d = _dummies.get(t.device)
        if d is None:
            d = _dummies[t.device] = torch.zeros(16, dtype=torch.int64, device=t.device)
        return d.data_ptr()
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


_profile = None   # when a list: (name, args, start_event, end_event) appended per call (bench.py kernel breakdown)


def set_profile(records):
    """records: a list to collect per-call CUDA-event timings into, or None to switch profiling off."""
    global _profile
    _profile = records


def call(name: str, *args):
    """Invoke an entry point on the current torch CUDA stream; raise on a non-zero status."""
    if _profile is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib(), name)(*args, stream())
        e1.record()
        _profile.append((name, args, e0, e1))
    else:
        rc = getattr(lib(), name)(*args, stream())
    if rc != 0:
        raise GaotLibraryError(f"{name} failed with status {rc}"
                               + (" (bad argument)" if rc == -1 else " (alignment)" if rc == -2 else " (CUDA error)"))


def check_cuda_f32(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise GaotLibraryError("gaot_b200 ops run on CUDA tensors only (no CPU fallback)")
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise GaotLibraryError(f"expected contiguous float32, got {t.dtype} contiguous={t.is_contiguous()}")
