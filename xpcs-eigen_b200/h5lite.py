"""ctypes front end of h5lite (xpcs-eigen_b200/host/h5lite.{h,cpp}): the from-scratch HDF5
subset reader/writer the host `corr` uses, exposed to Python so that tests and tools can write
configuration files and read result files without libhdf5 / h5py (neither exists in the image)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libh5lite.so")
_lib = None

_TYPES = ["int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64", "float32", "float64", "str"]
_CODE = {np.dtype(t): i for i, t in enumerate(_TYPES[:-1])}


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not found: run `make -C xpcs-eigen_b200 host`" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.h5l_error.restype = C.c_char_p
        L.h5l_new.restype = C.c_void_p
        L.h5l_open.restype = C.c_void_p
        L.h5l_open.argtypes = [C.c_char_p]
        L.h5l_save.argtypes = [C.c_void_p, C.c_char_p]
        L.h5l_close.argtypes = [C.c_void_p]
        L.h5l_list.restype = C.c_long
        L.h5l_list.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_long]
        L.h5l_info.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                               C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
        L.h5l_read.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_ulonglong]
        L.h5l_put.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_ulonglong), C.c_void_p,
                              C.c_ulonglong]
        L.h5l_set_storage.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_ulonglong), C.c_int, C.c_int]
        L.h5l_extra_count.argtypes = [C.c_void_p, C.c_char_p]
        L.h5l_extra_get.restype = C.c_long
        L.h5l_extra_get.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_long]
        L.h5l_report.restype = C.c_long
        L.h5l_report.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_long]
        _lib = L
    return _lib


class H5Error(RuntimeError):
    pass


class File:
    """In-memory HDF5 tree: File(path) loads, File() starts empty; save(path) rewrites the file."""

    def __init__(self, path=None):
        L = _load()
        self._h = L.h5l_open(path.encode()) if path else L.h5l_new()
        if not self._h:
            raise H5Error(L.h5l_error().decode())

    def close(self):
        if self._h:
            _load().h5l_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def save(self, path):
        if _load().h5l_save(self._h, path.encode()) != 0:
            raise H5Error(_load().h5l_error().decode())

    def list(self, group="/"):
        L = _load()
        n = L.h5l_list(self._h, group.encode(), None, 0)
        buf = C.create_string_buffer(n + 1)
        L.h5l_list(self._h, group.encode(), buf, n + 1)
        return [x for x in buf.value.decode().split("\n") if x]

    def kind(self, path):
        t, r, es, nb = C.c_int(), C.c_int(), C.c_ulonglong(), C.c_ulonglong()
        dims = (C.c_ulonglong * 8)()
        return ["absent", "group", "dataset"][_load().h5l_info(self._h, path.encode(), C.byref(t), C.byref(r), dims,
                                                              C.byref(es), C.byref(nb))]

    def get(self, path):
        """numpy array (or str for string datasets)."""
        L = _load()
        t, r, es, nb = C.c_int(), C.c_int(), C.c_ulonglong(), C.c_ulonglong()
        dims = (C.c_ulonglong * 8)()
        k = L.h5l_info(self._h, path.encode(), C.byref(t), C.byref(r), dims, C.byref(es), C.byref(nb))
        if k != 2:
            raise KeyError(path)
        raw = np.zeros(max(nb.value, 1), np.uint8)
        if L.h5l_read(self._h, path.encode(), raw.ctypes.data, raw.size) != 0:
            raise H5Error("read failed: " + path)
        raw = raw[: nb.value]
        if _TYPES[t.value] == "str":
            return raw.tobytes().split(b"\0")[0].decode().rstrip(" ")
        shape = tuple(int(dims[i]) for i in range(r.value))
        return raw.view(_TYPES[t.value]).reshape(shape).copy()

    def put(self, path, value):
        L = _load()
        if isinstance(value, str):
            b = value.encode() + b"\0"
            dims = (C.c_ulonglong * 1)(1)
            rc = L.h5l_put(self._h, path.encode(), 10, 1, dims, b, len(b))
        else:
            a = np.ascontiguousarray(value)
            code = _CODE[a.dtype]
            dims = (C.c_ulonglong * max(a.ndim, 1))(*a.shape)
            rc = L.h5l_put(self._h, path.encode(), code, a.ndim, dims, a.ctypes.data, 0)
        if rc != 0:
            raise H5Error(L.h5l_error().decode())

    def set_storage(self, path, chunk=None, deflate=0, shuffle=False):
        """How save() lays the dataset out: chunk = extents per dimension (None = contiguous), deflate level, shuffle."""
        ch = list(chunk or [])
        arr = (C.c_ulonglong * max(len(ch), 1))(*ch)
        if _load().h5l_set_storage(self._h, path.encode(), len(ch), arr, int(deflate), int(bool(shuffle))) != 0:
            raise KeyError(path)

    def extra(self, path):
        """[(message type, raw bytes)] of the attribute (0x0C) / comment (0x0D) messages carried by `path`."""
        L = _load()
        out = []
        for i in range(max(L.h5l_extra_count(self._h, path.encode()), 0)):
            t = C.c_int()
            n = L.h5l_extra_get(self._h, path.encode(), i, C.byref(t), None, 0)
            buf = C.create_string_buffer(max(n, 1))
            L.h5l_extra_get(self._h, path.encode(), i, C.byref(t), buf, n)
            out.append((t.value, buf.raw[:n]))
        return out

    def report(self, which="lossy"):
        """lines of File::lossy (content a save cannot reproduce) or File::notes (re-encoded content)."""
        L = _load()
        w = 0 if which == "lossy" else 1
        n = L.h5l_report(self._h, w, None, 0)
        buf = C.create_string_buffer(n + 1)
        L.h5l_report(self._h, w, buf, n + 1)
        return [x for x in buf.value.decode().split("\n") if x]

    def walk(self, group="/"):
        """{relative path: value} of every dataset below `group`."""
        out = {}
        base = group.rstrip("/")
        for n in self.list(group):
            p = base + "/" + n
            if self.kind(p) == "group":
                for k, v in self.walk(p).items():
                    out[n + "/" + k] = v
            else:
                out[n] = self.get(p)
        return out
