"""ctypes binding of libposeb200.so.

The argument structs are generated from include/poseb200.h at import time, so the Python
mirror can never drift from the C ABI.  There is no CPU fallback: if the shared library is
missing or fails to load, every op raises (``LibraryMissing``).
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(PKG_DIR), "include", "poseb200.h")
LIB_PATH = os.path.join(PKG_DIR, "libposeb200.so")


class LibraryMissing(RuntimeError):
    pass


class PoseB200Error(RuntimeError):
    def __init__(self, fn: str, code: int, msg: str):
        super().__init__(f"{fn} failed ({code}): {msg}")
        self.code = code


_SCALARS = {
    "int8_t": ctypes.c_int8, "uint8_t": ctypes.c_uint8, "int32_t": ctypes.c_int32, "uint32_t": ctypes.c_uint32,
    "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64, "float": ctypes.c_float, "double": ctypes.c_double,
    "int": ctypes.c_int,
}


def _strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def parse_header(path: str = HEADER):
    """returns (defines, structs{name: [(field, ctype)]}, functions[name])"""
    with open(path) as fh:
        text = _strip_comments(fh.read())
    defines = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(PB_\w+)\s+(-?\d+)", text)}
    enums: Dict[str, int] = {}
    for m in re.finditer(r"typedef\s+enum\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        nxt = 0
        for item in m.group(1).split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                k, v = (s.strip() for s in item.split("="))
                nxt = int(v, 0)
            else:
                k = item
            enums[k] = nxt
            nxt += 1
    structs: Dict[str, type] = {}
    order: List[str] = []
    for m in re.finditer(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        body, name = m.group(1), m.group(2)
        fields: List[Tuple[str, object]] = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            if "*" in decl:
                fname = decl.split("*")[-1].strip()
                fields.append((fname, ctypes.c_void_p))
                continue
            toks = decl.replace("const ", "").split(" ", 1)
            tname, rest = toks[0], toks[1]
            base = _SCALARS.get(tname) or structs.get(tname)
            if base is None:
                raise ValueError(f"unknown type {tname!r} in {name}")
            for d in rest.split(","):
                d = d.strip()
                am = re.match(r"(\w+)\[(\w+)\]$", d)
                if am:
                    n = defines.get(am.group(2))
                    n = int(am.group(2)) if n is None else n
                    fields.append((am.group(1), base * n))
                else:
                    fields.append((d, base))
        structs[name] = type(name, (ctypes.Structure,), {"_fields_": fields})
        order.append(name)
    funcs = re.findall(r"\bint\s+(pb_\w+)\s*\(", text)
    funcs += re.findall(r"\bconst\s+char\s*\*\s*(pb_\w+)\s*\(", text)
    return defines, enums, structs, funcs


DEFINES, ENUMS, STRUCTS, FUNCTIONS = parse_header()
PB_F32, PB_BF16, PB_F16 = ENUMS["PB_F32"], ENUMS["PB_BF16"], ENUMS["PB_F16"]
PB_ACT_NONE, PB_ACT_LRELU, PB_ACT_MASKMUL, PB_ACT_GELU = (ENUMS[k] for k in
                                                         ("PB_ACT_NONE", "PB_ACT_LRELU", "PB_ACT_MASKMUL", "PB_ACT_GELU"))
PB_MAX_TAPS = DEFINES["PB_MAX_TAPS"]

_lib = None


def load() -> ctypes.CDLL:
    """dlopen libposeb200.so (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  The hot path has no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.pb_last_error_string.restype = ctypes.c_char_p
    lib.pb_last_error_string.argtypes = []
    lib.pb_abi_version.restype = ctypes.c_int
    if lib.pb_abi_version() != DEFINES["PB_ABI_VERSION"]:
        raise LibraryMissing("libposeb200.so ABI version does not match include/poseb200.h -- rebuild")
    _lib = lib
    return lib


def last_error() -> str:
    return load().pb_last_error_string().decode("utf-8", "replace")


def call(fn: str, args: ctypes.Structure, stream: int) -> None:
    """Invoke ``int pb_<fn>(const args*, stream)`` and raise on a non-zero status."""
    lib = load()
    f = getattr(lib, fn)
    rc = f(ctypes.byref(args), ctypes.c_void_p(stream))
    if rc != 0:
        raise PoseB200Error(fn, rc, last_error())


def exported_symbols() -> List[str]:
    return list(FUNCTIONS)
