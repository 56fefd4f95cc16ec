"""ctypes binding of libgigagan_sm100.so (the C-ABI in include/gigagan_sm100.h).

The prototypes are parsed from the public header so the Python side can never drift from it.  There is no
fallback: if the library is missing or a kernel reports an error this raises."""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GG_LIB") or os.path.join(_HERE, "libgigagan_sm100.so")     # GG_LIB: A/B a second build
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "gigagan_sm100.h")

_CT = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "float": ctypes.c_float, "gg_stream_t": ctypes.c_void_p}


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [argtypes])} for every function declared in the public header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|const char\*)\s+(gg_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    argtypes.append(_CT[a.split()[-2] if len(a.split()) > 1 else a])
        protos[name] = (ctypes.c_char_p if "char" in ret else ctypes.c_int, argtypes)
    return protos


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  gigagan_pytorch_b200 has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (ret, argtypes) in parse_header().items():
            fn = getattr(L, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype, fn.argtypes = ret, argtypes
        if os.environ.get("GG_FLAGS"):            # A/B switches for benchmarking (see gg_set_flags in the header)
            L.gg_set_flags(int(os.environ["GG_FLAGS"]))
        _lib = L
    return _lib


launch_count = 0


def call(name, *args):
    """Invoke a kernel entry point; raise on a non-zero status."""
    global launch_count
    L = lib()
    rc = getattr(L, name)(*args)
    launch_count += 1
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {L.gg_last_error().decode()}")
