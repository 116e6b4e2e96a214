"""cffi binding of libfrcnn_b200.so.  The cdef source is include/frcnn_b200.h itself (preprocessor lines and the
extern "C" braces stripped), so the Python host mirror binds exactly the ABI the LuaJIT glue binds."""
import os
import re

import cffi

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "frcnn_b200.h")
LIB_PATH = os.path.join(HERE, "libfrcnn_b200.so")

ffi = cffi.FFI()
_lib = None


def header_cdef():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    lines = [l for l in src.splitlines() if not l.lstrip().startswith("#") and l.strip() not in ('extern "C" {', "}")]
    return "\n".join(lines)


def declared_functions():
    """Names of every function the header declares."""
    return re.findall(r"\b(frcnn_[a-z0-9_]+)\s*\(", header_cdef())


ffi.cdef(header_cdef())


class FrcnnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("frcnn_b200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Loads the CUDA library.  There is no fallback: a missing library is a hard error."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FrcnnError(-1, "libfrcnn_b200.so is not built; run `python __graft_entry__.py build` "
                                 "(this package has no CPU or PyTorch fallback)")
        _lib = ffi.dlopen(LIB_PATH)
        # resolve every entry point now: cffi builds accessors lazily under a non-reentrant lock, and a finaliser
        # (Model.__del__ -> frcnn_destroy) that runs inside another cffi call would otherwise deadlock on it
        for name in declared_functions():
            getattr(_lib, name)
    return _lib


def check(ctx, code):
    if code != 0:
        msg = lib().frcnn_last_error(ctx if ctx is not None else ffi.NULL)
        raise FrcnnError(code, ffi.string(msg).decode() if msg != ffi.NULL else "?")
