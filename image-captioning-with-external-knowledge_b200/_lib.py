"""
ctypes binding of csrc/libickb200.so.  The signatures are parsed from include/ickb200.h (the single source of truth
for the C ABI).  There is NO fallback: if the library is missing or a symbol is absent, importing fails loudly.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "ickb200.h")
# ICKB200_LIB overrides the library path (A/B runs of two builds of the same ABI); there is still no non-CUDA fallback
LIB_PATH = os.environ.get("ICKB200_LIB") or os.path.join(HERE, "csrc", "libickb200.so")

_CTYPES = {
    "int": ctypes.c_int,
    "unsigned": ctypes.c_uint,
    "long long": ctypes.c_longlong,
    "float": ctypes.c_float,
    "cudaStream_t": ctypes.c_void_p,
}


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[Tuple[str, str]]]]:
    """-> {name: (restype, [(ctype_name, arg_name), ...])} for every `ick_*` prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char\*|int)\s+(ick_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        params: List[Tuple[str, str]] = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                pm = re.match(r"(.*?)(\w+)$", a)
                typ, pname = pm.group(1).strip(), pm.group(2)
                params.append((typ, pname))
        protos[name] = (ret, params)
    return protos


def _ctype(typ: str):
    if "*" in typ:
        return ctypes.c_void_p
    return _CTYPES[typ]


class Library:
    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise ImportError(
                f"ickb200: kernel library {path} not found. Build it with `python -c 'import __graft_entry__ as g; "
                f"g.build()'` (nvcc, sm_100a). There is no CPU fallback."
            )
        self.path = path
        self.cdll = ctypes.CDLL(path)
        self.protos = parse_header()
        self.fn = {}
        for name, (ret, params) in self.protos.items():
            try:
                f = getattr(self.cdll, name)
            except AttributeError as e:
                raise ImportError(f"ickb200: {path} does not export {name} declared in {HEADER}") from e
            f.restype = ctypes.c_char_p if ret.startswith("const char") else ctypes.c_int
            f.argtypes = [_ctype(t) for t, _ in params]
            self.fn[name] = f
        ver = self.fn["ick_abi_version"]()
        m = re.search(r"#define\s+ICK_ABI_VERSION\s+(\d+)", open(HEADER).read())
        if m and int(m.group(1)) != ver:
            raise ImportError(f"ickb200: ABI version mismatch: header {m.group(1)} vs library {ver}; rebuild")

    @property
    def launches(self) -> int:
        """Kernels the library has launched in this process (counted at every cudaLaunchKernelEx of csrc/, bench.py's gpu_launches)."""
        n = ctypes.c_longlong(0)
        self.fn["ick_launch_count"](ctypes.addressof(n))
        return int(n.value)

    def call(self, name: str, *args) -> None:
        rc = self.fn[name](*args)
        if rc != 0:
            raise RuntimeError(f"{name} failed (rc={rc}): {self.fn['ick_last_error']().decode()}")


_LIB = None


def get() -> Library:
    global _LIB
    if _LIB is None:
        _LIB = Library()
    return _LIB
