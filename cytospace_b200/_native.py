"""cffi (ABI mode) binding of libcytospace_b200.so -- the C-ABI boundary.

The declarations are parsed from ``include/cytospace_b200.h`` itself so the Python
layer cannot drift from the header.  There is NO fallback: if the shared library
has not been built (``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C cytospace_b200/csrc``) importing a compute entry raises.
"""
from __future__ import annotations

import os
import re
import subprocess
import threading

import cffi

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
HEADER = os.path.join(_ROOT, "include", "cytospace_b200.h")
LIB_PATH = os.path.join(_PKG, "libcytospace_b200.so")
CSRC = os.path.join(_PKG, "csrc")
SOURCES = ("common.cu", "cost_build.cu", "metrics.cu", "lap_sap.cu", "lap_check.cu", "dist.cu", "stage.cu")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-ldl"]

_lock = threading.Lock()
_ffi = None
_lib = None


class NativeLibraryMissing(RuntimeError):
    """libcytospace_b200.so is not built / not loadable.  Never caught by the product path."""


def _cdef_from_header(text: str) -> str:
    out = []
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for line in text.splitlines():
        s = line.strip()
        if s.startswith("#"):
            m = re.match(r"#define\s+(CYB_\w+)\s+(-?\d+)\s*$", s)
            if m:
                out.append(f"#define {m.group(1)} {m.group(2)}")
            continue
        if s.startswith('extern "C"') or s == "}":
            continue
        out.append(line)
    return "\n".join(out)


def declared_symbols() -> list[str]:
    """Every function the header declares (used by the CPU-side export test)."""
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(cyb_\w+)\s*\(", text)))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libcytospace_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, "common.h"), HEADER]
    if not force and os.path.exists(LIB_PATH) and all(
            os.path.getmtime(d) <= os.path.getmtime(LIB_PATH) for d in deps):
        return LIB_PATH
    cmd = ["nvcc", *NVCC_FLAGS, "-I", os.path.join(_ROOT, "include"), "-I", CSRC, "-o", LIB_PATH, *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True)
    return LIB_PATH


def ffi() -> cffi.FFI:
    load()
    return _ffi


def load():
    """Returns the dlopen()ed library; raises NativeLibraryMissing when it is absent."""
    global _ffi, _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise NativeLibraryMissing(
                    f"{LIB_PATH} not found: build it with __graft_entry__.build() -- "
                    "cytospace_b200 has no CPU fallback")
            f = cffi.FFI()
            f.cdef(_cdef_from_header(open(HEADER).read()))
            try:
                lib = f.dlopen(LIB_PATH)
            except OSError as e:  # pragma: no cover
                raise NativeLibraryMissing(f"cannot load {LIB_PATH}: {e}") from e
            if lib.cyb_abi_version() != lib.CYB_ABI_VERSION:
                raise NativeLibraryMissing(
                    f"{LIB_PATH} has ABI {lib.cyb_abi_version()}, header wants {lib.CYB_ABI_VERSION}: rebuild")
            _ffi, _lib = f, lib
    return _lib


class CybError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cytospace_b200 native error {code}: {msg}")
        self.code = code


def check(rc: int):
    if rc != 0:
        raise CybError(rc, _ffi.string(_lib.cyb_last_error()).decode(errors="replace"))


def ptr(ctype: str, tensor):
    """Device (or host) pointer of a torch tensor as a cffi pointer; None -> NULL."""
    if tensor is None:
        return _ffi.NULL
    return _ffi.cast(ctype, tensor.data_ptr())


def stream_ptr(stream):
    return _ffi.cast("void *", stream.cuda_stream)
