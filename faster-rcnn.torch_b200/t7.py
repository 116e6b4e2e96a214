"""Torch7 object-stream (.t7) reader / writer and the snapshot functions of utilities.lua:113-134.

The reference saves and restores training snapshots with `torch.DiskFile(name, 'w'):writeObject{version = 0, weights,
options, stats}` (utilities.lua:113-134, main.lua:94-98,145-148) -- an ASCII torch.File object stream, `weights` being
the flat CudaTensor of `nn.Module.flatten`.  This module restates that serialisation grammar (torch7 `File.lua`,
`generic/Tensor.c`, `generic/Storage.c`, `lib/TH/THDiskFile.c`: un-vendored, unpinned, ~Nov 2015) so that the Python host
can exchange weights with a real Torch7 run of the reference:

  object   := TYPE_NIL
            | TYPE_NUMBER double | TYPE_BOOLEAN int | TYPE_STRING int(len) chars
            | TYPE_TABLE  index [int(size) {object(key) object(value)}*size]      (body only on first occurrence of index)
            | TYPE_TORCH  index [string("V 1") string(class) class-body]
  Tensor   := int(nDim) long[nDim](size) long[nDim](stride) long(storageOffset, 1-based) object(Storage | nil)
  Storage  := long(size) elem[size]

ASCII mode (`torch.DiskFile` default, what save_obj uses): every scalar on its own line, arrays blank-separated and
newline-terminated, doubles "%.17g", floats "%.9g", chars raw; binary mode (torch.save's default) little-endian
int32 / int64 / native element types.  Both are read (auto-detected) and written."""
import io
import struct

import numpy as np

TYPE_NIL, TYPE_NUMBER, TYPE_STRING, TYPE_TABLE, TYPE_TORCH, TYPE_BOOLEAN = 0, 1, 2, 3, 4, 5
TYPE_FUNCTION, LEGACY_TYPE_RECUR_FUNCTION, TYPE_RECUR_FUNCTION = 6, 7, 8

_ELEM = {  # Torch class stem -> (numpy dtype, ascii format)
    "Float": (np.float32, "%.9g"), "Double": (np.float64, "%.17g"), "Long": (np.int64, "%d"), "Int": (np.int32, "%d"),
    "Short": (np.int16, "%d"), "Byte": (np.uint8, None), "Char": (np.int8, None), "Cuda": (np.float32, "%.9g"),
}


class TorchObject:
    """A torch class instance the reader has no native mapping for (e.g. an nn module): its class name and the
    field table Torch serialised (File.lua: objects without a `write` method are stored as their field table)."""

    def __init__(self, typename, fields):
        self.typename, self.fields = typename, fields

    def __repr__(self):
        return "TorchObject(%s)" % self.typename


class Tensor(np.ndarray):
    """numpy array that remembers its Torch class ('torch.CudaTensor', 'torch.FloatTensor', ...)."""

    def __new__(cls, array, typename=None):
        obj = np.asarray(array).view(cls)
        obj.typename = typename
        return obj

    def __array_finalize__(self, obj):
        self.typename = getattr(obj, "typename", None)


def _stem(typename):
    name = typename.split(".")[-1]
    for suffix in ("Tensor", "Storage"):
        if name.endswith(suffix):
            return name[:-len(suffix)]
    raise ValueError("not a tensor / storage class: " + typename)


# ---------------------------------------------------------------------------------------------------------- reader
class _Reader:
    def __init__(self, data):
        self.d, self.pos, self.memo = data, 0, {}
        # ASCII streams start with the type tag as a decimal digit + newline; binary ones with a little-endian int32
        self.ascii = len(data) >= 2 and chr(data[0]).isdigit() and data[1:2] in (b"\n", b"\r", b" ")

    # -- primitive reads (THDiskFile.c READ_WRITE_METHODS)
    def _token(self):
        d, n = self.d, len(self.d)
        i = self.pos
        while i < n and d[i] in b" \t\r\n":
            i += 1
        j = i
        while j < n and d[j] not in b" \t\r\n":
            j += 1
        if i == j:
            raise EOFError("unexpected end of .t7 stream")
        self.pos = j
        return d[i:j]

    def _eat_newline(self):  # auto-spacing: one '\n' after every ASCII read
        if self.pos < len(self.d) and self.d[self.pos:self.pos + 1] == b"\n":
            self.pos += 1

    def int(self):
        if self.ascii:
            v = int(self._token())
            self._eat_newline()
            return v
        v = struct.unpack_from("<i", self.d, self.pos)[0]
        self.pos += 4
        return v

    def long(self):
        if self.ascii:
            v = int(self._token())
            self._eat_newline()
            return v
        v = struct.unpack_from("<q", self.d, self.pos)[0]
        self.pos += 8
        return v

    def double(self):
        if self.ascii:
            v = float(self._token())
            self._eat_newline()
            return v
        v = struct.unpack_from("<d", self.d, self.pos)[0]
        self.pos += 8
        return v

    def chars(self, n):
        v = bytes(self.d[self.pos:self.pos + n])
        if len(v) != n:
            raise EOFError("unexpected end of .t7 stream")
        self.pos += n
        if self.ascii and n > 0:
            self._eat_newline()
        return v

    def array(self, dtype, n):
        dtype = np.dtype(dtype)
        if n == 0:
            return np.zeros(0, dtype)
        if not self.ascii or dtype.itemsize == 1:
            v = np.frombuffer(self.d, dtype=dtype.newbyteorder("<"), count=n, offset=self.pos).astype(dtype)
            self.pos += n * dtype.itemsize
            if self.ascii:
                self._eat_newline()
            return v
        # n blank-separated numbers; find the end of the n-th token without tokenising in Python
        end = self.pos
        view = self.d
        # numbers never contain '\n'; one array is one line (THDiskFile writes " " between elements, "\n" after the last)
        nl = view.find(b"\n", end)
        if nl < 0:
            nl = len(view)
        line = bytes(view[end:nl])
        # C-speed text parse (strtod semantics incl. nan / inf); a 27 M-element weight vector is one ~320 MB line
        v = np.fromstring(line, dtype=np.float64 if dtype.kind == "f" else np.int64, sep=" ").astype(dtype)
        if v.size != n:
            raise ValueError("expected %d elements, found %d" % (n, v.size))
        self.pos = min(nl + 1, len(view))
        return v

    def string(self):
        return self.chars(self.int()).decode("latin-1")

    # -- File:readObject (File.lua)
    def object(self):
        t = self.int()
        if t == TYPE_NIL:
            return None
        if t == TYPE_NUMBER:
            v = self.double()
            return int(v) if (v == v and abs(v) < 2 ** 53 and v == int(v)) else v
        if t == TYPE_BOOLEAN:
            return self.int() == 1
        if t == TYPE_STRING:
            return self.string()
        if t in (TYPE_FUNCTION, LEGACY_TYPE_RECUR_FUNCTION, TYPE_RECUR_FUNCTION):
            raise ValueError("serialised Lua functions are not supported")
        if t not in (TYPE_TABLE, TYPE_TORCH):
            raise ValueError("unknown .t7 type tag %d at byte %d" % (t, self.pos))
        index = self.int()
        if index in self.memo:
            return self.memo[index]
        if t == TYPE_TABLE:
            size = self.int()
            tab = {}
            self.memo[index] = tab
            for _ in range(size):
                k = self.object()
                tab[k] = self.object()
            return tab
        version = self.string()
        if version.startswith("V "):
            typename = self.string()
        else:  # pre-versioning streams store the class name directly
            typename = version
        return self._torch(index, typename)

    def _torch(self, index, typename):
        if typename.endswith("Tensor"):
            ndim = self.int()
            size = self.array(np.int64, ndim)
            stride = self.array(np.int64, ndim)
            offset = self.long() - 1
            storage = self.object()
            if storage is None or ndim == 0:
                out = Tensor(np.zeros(0, _ELEM[_stem(typename)][0]), typename)
            else:
                out = Tensor(np.lib.stride_tricks.as_strided(storage[offset:], shape=tuple(int(x) for x in size),
                                                             strides=tuple(int(x) * storage.itemsize for x in stride)).copy(), typename)
            self.memo[index] = out
            return out
        if typename.endswith("Storage"):
            n = self.long()
            out = self.array(_ELEM[_stem(typename)][0], n)
            self.memo[index] = out
            return out
        obj = TorchObject(typename, None)
        self.memo[index] = obj
        obj.fields = self.object()
        return obj


def load(path_or_bytes):
    """load_obj (utilities.lua:119-124): the object stored in a .t7 file (ASCII or binary).  Lua tables come back as
    dicts (integral number keys as ints), tensors as `Tensor` (numpy) arrays, numbers as int when integral else float."""
    if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
        data = bytes(path_or_bytes)
    else:
        with open(path_or_bytes, "rb") as f:
            data = f.read()
    return _Reader(data).object()


# ---------------------------------------------------------------------------------------------------------- writer
class _Writer:
    def __init__(self, f, ascii=True, cuda=True):
        self.f, self.ascii, self.cuda, self.memo, self.n, self.keep = f, ascii, cuda, {}, 0, []

    def int(self, v):
        self.f.write(b"%d\n" % v if self.ascii else struct.pack("<i", v))

    def long(self, v):
        self.f.write(b"%d\n" % v if self.ascii else struct.pack("<q", v))

    def double(self, v):
        self.f.write(("%.17g\n" % v).encode() if self.ascii else struct.pack("<d", v))

    def chars(self, b):
        self.f.write(b)
        if self.ascii and len(b) > 0:
            self.f.write(b"\n")

    def string(self, s):
        b = s.encode("latin-1")
        self.int(len(b))
        self.chars(b)

    def array(self, a, fmt):
        a = np.ascontiguousarray(a)
        if a.size == 0:
            return
        if not self.ascii or fmt is None:
            self.f.write(a.astype(a.dtype.newbyteorder("<")).tobytes())
            if self.ascii:
                self.f.write(b"\n")
            return
        flat = a.reshape(-1)
        step = 1 << 16
        for i in range(0, flat.size, step):  # chunked: a 27 M-element weight vector is ~320 MB of text
            chunk = flat[i:i + step].tolist()
            self.f.write((" ".join([fmt % x for x in chunk])).encode())
            self.f.write(b" " if i + step < flat.size else b"\n")

    def _index(self, obj):
        key = id(obj)
        if key in self.memo:
            self.int(self.memo[key])
            return True
        self.n += 1
        self.memo[key] = self.n
        self.keep.append(obj)  # ids stay unique while the stream is open
        self.int(self.n)
        return False

    def object(self, obj):
        if obj is None:
            return self.int(TYPE_NIL)
        if isinstance(obj, (bool, np.bool_)):
            self.int(TYPE_BOOLEAN)
            return self.int(1 if obj else 0)
        if isinstance(obj, (int, float, np.integer, np.floating)):
            self.int(TYPE_NUMBER)
            return self.double(float(obj))
        if isinstance(obj, str):
            self.int(TYPE_STRING)
            return self.string(obj)
        if isinstance(obj, np.ndarray):
            return self._tensor(obj)
        if isinstance(obj, TorchObject):
            self.int(TYPE_TORCH)
            if self._index(obj):
                return
            self.string("V 1")
            self.string(obj.typename)
            return self.object(obj.fields)
        if isinstance(obj, (list, tuple)):
            obj_t = {i + 1: v for i, v in enumerate(obj)}   # Lua arrays are 1-based tables
            self.int(TYPE_TABLE)
            if self._index(obj):
                return
            return self._table_body(obj_t)
        if isinstance(obj, dict):
            self.int(TYPE_TABLE)
            if self._index(obj):
                return
            return self._table_body(obj)
        if hasattr(obj, "detach") and hasattr(obj, "cpu"):  # torch.Tensor
            t = obj.detach().cpu().numpy()
            return self._tensor(Tensor(t, "torch.CudaTensor" if (obj.is_cuda and t.dtype == np.float32) else None))
        raise TypeError("cannot serialise %r into a .t7 stream" % type(obj))

    def _table_body(self, tab):
        self.int(len(tab))
        for k, v in tab.items():
            self.object(k)
            self.object(v)

    def _tensor(self, a):
        typename = getattr(a, "typename", None)
        if typename is None:
            stem = {np.dtype(np.float32): "Cuda" if self.cuda else "Float", np.dtype(np.float64): "Double", np.dtype(np.int64): "Long",
                    np.dtype(np.int32): "Int", np.dtype(np.int16): "Short", np.dtype(np.uint8): "Byte", np.dtype(np.int8): "Char"}[a.dtype]
            typename = "torch.%sTensor" % stem
        stem = _stem(typename)
        dtype, fmt = _ELEM[stem]
        a = np.ascontiguousarray(np.asarray(a), dtype=dtype)
        self.int(TYPE_TORCH)
        if self._index(a):
            return
        self.string("V 1")
        self.string(typename)
        self.int(a.ndim)
        self.array(np.array(a.shape, dtype=np.int64), "%d")
        self.array(np.array([s // a.itemsize for s in a.strides], dtype=np.int64), "%d")
        self.long(1)  # storageOffset, 1-based
        # the storage object (generic/Storage.c write): long(size) + elements
        storage = a.reshape(-1)
        self.int(TYPE_TORCH)
        self.n += 1
        self.int(self.n)
        self.string("V 1")
        self.string("torch.%sStorage" % stem)
        self.long(storage.size)
        self.array(storage, fmt)


def save(path, obj, ascii=True, cuda=True):
    """save_obj (utilities.lua:113-117): `torch.DiskFile(file_name, 'w'):writeObject(obj)` -- ASCII by default, as the
    reference writes it.  cuda: float32 arrays without an explicit class are stored as torch.CudaTensor (the reference
    flattens the parameters after :cuda(), main.lua:86-92)."""
    if isinstance(path, io.IOBase):
        _Writer(path, ascii, cuda).object(obj)
        return
    with open(path, "wb") as f:
        _Writer(f, ascii, cuda).object(obj)


def dumps(obj, ascii=True, cuda=True):
    f = io.BytesIO()
    _Writer(f, ascii, cuda).object(obj)
    return f.getvalue()


# ---------------------------------------------------------------------------------------- snapshots (utilities.lua:126-134)
def save_model(file_name, weights, options, stats, bn_running_stats=None, ascii=True):
    """save_model(file_name, weights, options, stats) (utilities.lua:126-134): {version = 0, weights, options, stats}.
    `weights` is the flat parameter vector in `nn.Module.flatten` order (learnable parameters only).  The reference never
    stores the BatchNormalization running statistics (they are not parameters), so a restored cnet evaluates with
    mean 0 / var 1; `bn_running_stats` (dict name -> array) persists them under an extra key the reference ignores."""
    obj = {"version": 0, "weights": weights, "options": options, "stats": stats}
    if bn_running_stats is not None:
        obj["bn_running_stats"] = {k: np.asarray(v, dtype=np.float32) for k, v in bn_running_stats.items()}
    save(file_name, obj, ascii=ascii)


def load_model(file_name):
    """load_obj on a snapshot (main.lua:94-98): returns (weights float32 vector, options, stats, bn_running_stats | None)."""
    stored = load(file_name)
    w = np.asarray(stored["weights"], dtype=np.float32).reshape(-1)
    return w, stored.get("options"), stored.get("stats"), stored.get("bn_running_stats")
