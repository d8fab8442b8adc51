"""ctypes binding of libaccbpg_b200.so, generated from include/accbpg_b200.h at import time.

The prototypes are parsed out of the header, so the header is the single source of truth
for the C ABI: a symbol declared there but missing from the shared object (or the other
way round) fails here, loudly.  There is no CPU fallback: if the library is absent the
import raises and every operator in this package is unusable.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "accbpg_b200.h")
LIB_PATH = os.path.join(HERE, "libaccbpg_b200.so")

_CTYPES = {
    "void*": ctypes.c_void_p,
    "const void*": ctypes.c_void_p,
    "void**": ctypes.POINTER(ctypes.c_void_p),
    "void* const*": ctypes.POINTER(ctypes.c_void_p),
    "const double*": ctypes.c_void_p,     # device or host address passed as an integer
    "double*": ctypes.c_void_p,
    "uint32_t*": ctypes.POINTER(ctypes.c_uint32),
    "int": ctypes.c_int,
    "int*": ctypes.POINTER(ctypes.c_int),
    "int64_t": ctypes.c_int64,
    "int64_t*": ctypes.POINTER(ctypes.c_int64),
    "const int64_t*": ctypes.c_void_p,    # device address
    "const int*": ctypes.c_void_p,        # device address
    "uint64_t": ctypes.c_uint64,
    "size_t": ctypes.c_size_t,
    "double": ctypes.c_double,
    "const char*": ctypes.c_char_p,
}


def parse_header(path=HEADER):
    """Return {name: (restype_str, [argtype_str, ...])} and {macro: int} from the C header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    macros = {}
    for mname, mval in re.findall(r"^\s*#define\s+(ACCBPG_\w+)\s+(0x[0-9a-fA-F]+u?|\d+u?)\s*$", text, flags=re.M):
        macros[mname] = int(mval.rstrip("u"), 0)
    text = re.sub(r"^\s*#.*$", " ", text, flags=re.M)
    protos = {}
    for ret, name, args in re.findall(r"([\w\s\*]+?)\s*\b(accbpg_\w+)\s*\(([^)]*)\)\s*;", text):
        ret = " ".join(ret.split()).replace(" *", "*")
        ret = ret.replace('extern "C" {', "").strip()
        argtypes = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                m = re.match(r"^(.*?)(\w+)$", a)          # strip the parameter name
                typ = " ".join(m.group(1).split()).replace(" *", "*").strip()
                argtypes.append(typ)
        protos[name] = (ret.split()[-1] if ret.split()[-1] in ("int", "size_t", "uint64_t") else ret, argtypes)
    return protos, macros


PROTOS, MACROS = parse_header()


class NativeError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python accbpg_and_fw_b200/_build.py` "
            "(nvcc, sm_100a).  This package has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (ret, args) in PROTOS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ImportError(f"{LIB_PATH} does not export {name} declared in include/accbpg_b200.h") from e
        if ret == "double*":
            fn.restype = ctypes.c_void_p
        else:
            fn.restype = _CTYPES[ret]
        fn.argtypes = [_CTYPES[a] for a in args]
    return lib


lib = _load()

OK = MACROS["ACCBPG_OK"]
E_ARG = MACROS["ACCBPG_E_ARG"]
E_CUDA = MACROS["ACCBPG_E_CUDA"]
ST = {k[len("ACCBPG_ST_"):]: v for k, v in MACROS.items() if k.startswith("ACCBPG_ST_")}


def check(rc):
    """Map a C return code to the exception type the reference would raise for the same misuse."""
    if rc == OK:
        return
    msg = lib.accbpg_last_error().decode("utf-8", "replace")
    if rc == E_ARG:
        raise AssertionError(msg)
    raise NativeError(msg)


def launch_count():
    return int(lib.accbpg_launch_count())
