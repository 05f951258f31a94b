"""Device-resident tensor trains for the host mirror.

`Handle` owns one `ttipm_tt` of the C ABI (cores in HBM).  `TTList` is what the TT functions of `ttipm_b200.tt` return:
a `list` subclass that behaves exactly like the reference's `list[np.ndarray]` for the unchanged driver code
(`src/tt_ipm.py`, `psd_system/*`), but holds only the device handle until somebody looks at the cores.  A chain such as

    tt_rank_reduce(tt_add(X, tt_scale(a, tt_fast_mat_mat_mul(A, B, eps))), eps)

therefore runs upload-free and download-free: every intermediate stays in HBM, and the NumPy cores are materialised
with ONE device-to-host copy the first time the list is indexed, iterated, concatenated, compared, printed or copied.

Once materialised the NumPy arrays are the caller's (the reference mutates cores in place in a few spots); the handle is
reused afterwards only while a checksum of the cores (small trains) says they are unchanged, otherwise the train is
uploaded again.  `len()` and the TT ranks never materialise.
"""
import ctypes as C
import zlib

import numpy as np

from .runtime import TTIPMError, get_runtime

_CHECKSUM_LIMIT = 1 << 20          # bytes: above this a materialised train is simply uploaded again


class Handle:
    """Owner of one ttipm_tt."""

    __slots__ = ("ptr", "rt", "_shapes")

    def __init__(self, ptr, rt):
        if not ptr:
            raise TTIPMError("native TT operation failed: " + rt.lib.ttipm_last_error().decode())
        self.ptr = C.c_void_p(ptr)
        self.rt = rt
        self._shapes = None

    def __del__(self):
        try:
            self.rt.lib.ttipm_tt_destroy(self.ptr)
        except Exception:
            pass

    @staticmethod
    def from_numpy(cores, rt):
        d = len(cores)
        dims = np.empty(4 * d, dtype=np.int32)
        flat = []
        for k, c in enumerate(cores):
            a = np.ascontiguousarray(c, dtype=np.float64)
            if a.ndim == 3:
                dims[4 * k:4 * k + 4] = (a.shape[0], a.shape[1], 0, a.shape[2])
            elif a.ndim == 4:
                dims[4 * k:4 * k + 4] = a.shape
            else:
                raise ValueError(f"TT core with {a.ndim} axes")
            flat.append(a.reshape(-1))
        buf = np.concatenate(flat) if flat else np.zeros(0)
        h = Handle(rt.lib.ttipm_tt_create(d, rt.stream()), rt)
        rt.check(rt.lib.ttipm_tt_set_cores(h.ptr, buf.ctypes.data, dims.ctypes.data), "ttipm_tt_set_cores")
        return h

    def shapes(self):
        if self._shapes is None:
            d = int(self.rt.lib.ttipm_tt_length(self.ptr))
            dims = np.zeros(4 * d, dtype=np.int32)
            self.rt.lib.ttipm_tt_shapes(self.ptr, dims.ctypes.data)
            out = []
            for k in range(d):
                r, n1, n2, R = (int(v) for v in dims[4 * k:4 * k + 4])
                out.append((r, n1, n2, R) if n2 > 0 else (r, n1, R))
            self._shapes = out
        return self._shapes

    def to_numpy(self):
        shapes = self.shapes()
        sizes = [int(np.prod(s)) for s in shapes]
        buf = np.empty(sum(sizes))
        self.rt.check(self.rt.lib.ttipm_tt_get_cores(self.ptr, buf.ctypes.data), "ttipm_tt_get_cores")
        out, o = [], 0
        for s, n in zip(shapes, sizes):
            out.append(buf[o:o + n].reshape(s).copy())
            o += n
        return out

    def clone(self):
        return Handle(self.rt.lib.ttipm_tt_clone(self.ptr), self.rt)


def _checksum(cores):
    total = sum(c.nbytes for c in cores)
    if total > _CHECKSUM_LIMIT:
        return None
    acc = 0
    for c in cores:
        if not isinstance(c, np.ndarray) or c.dtype != np.float64:
            return None
        acc = zlib.crc32(np.ascontiguousarray(c).data, acc)
        acc = zlib.crc32(repr(c.shape).encode(), acc)
    return acc


def _mat(method):
    def wrapper(self, *a, **k):
        self._materialise()
        for other in a:                      # C-level list code reads a list-typed operand's storage directly
            if isinstance(other, TTList):
                other._materialise()
        return method(self, *a, **k)
    wrapper.__name__ = method.__name__
    return wrapper


def _mat_mut(method):
    def wrapper(self, *a, **k):
        self._materialise()
        for other in a:
            if isinstance(other, TTList):
                other._materialise()
        self._sum = None                     # the list itself changes: the handle no longer describes it
        self._h = None
        return method(self, *a, **k)
    wrapper.__name__ = method.__name__
    return wrapper


class TTList(list):
    """list[np.ndarray] whose cores live on the device until they are looked at."""

    def __init__(self, handle):
        super().__init__()
        self._h = handle
        self._live = False               # True once the NumPy cores are in the list storage
        self._sum = None

    # ---- device side ------------------------------------------------------------------------------------------
    def _materialise(self):
        if not self._live:
            self._live = True
            cores = self._h.to_numpy()
            list.extend(self, cores)
            self._sum = _checksum(cores)

    def handle(self):
        """A handle describing the current contents, or None if the train has to be uploaded again."""
        if self._h is None:
            return None
        if not self._live:
            return self._h
        if self._sum is not None and list.__len__(self) == len(self._h.shapes()) and \
                _checksum(list.__iter__(self)) == self._sum:
            return self._h
        self._h = None
        return None

    def rebind(self, handle):
        """In-place replacement of the whole train (functions that mutate their input list in the reference)."""
        list.clear(self)
        self._h = handle
        self._live = False
        self._sum = None

    def shapes(self):
        if self._live or self._h is None:
            return [tuple(c.shape) for c in list.__iter__(self)]
        return self._h.shapes()

    # ---- list protocol ----------------------------------------------------------------------------------------
    def __len__(self):
        return list.__len__(self) if self._live else len(self._h.shapes())

    def __bool__(self):
        return len(self) > 0

    __getitem__ = _mat(list.__getitem__)
    __iter__ = _mat(list.__iter__)
    __reversed__ = _mat(list.__reversed__)
    __contains__ = _mat(list.__contains__)
    __add__ = _mat(list.__add__)
    __mul__ = _mat(list.__mul__)
    __rmul__ = _mat(list.__rmul__)
    __eq__ = _mat(list.__eq__)
    __ne__ = _mat(list.__ne__)
    __lt__ = _mat(list.__lt__)
    __le__ = _mat(list.__le__)
    __gt__ = _mat(list.__gt__)
    __ge__ = _mat(list.__ge__)
    __repr__ = _mat(list.__repr__)
    index = _mat(list.index)
    count = _mat(list.count)
    __hash__ = None

    def __radd__(self, other):
        self._materialise()
        return list(other) + list(list.__iter__(self))

    def copy(self):
        self._materialise()
        return list(list.__iter__(self))

    def __reduce_ex__(self, protocol):
        self._materialise()
        return (list, (list(list.__iter__(self)),))

    __setitem__ = _mat_mut(list.__setitem__)
    __delitem__ = _mat_mut(list.__delitem__)
    __iadd__ = _mat_mut(list.__iadd__)
    __imul__ = _mat_mut(list.__imul__)
    append = _mat_mut(list.append)
    extend = _mat_mut(list.extend)
    insert = _mat_mut(list.insert)
    pop = _mat_mut(list.pop)
    remove = _mat_mut(list.remove)
    clear = _mat_mut(list.clear)
    reverse = _mat_mut(list.reverse)
    sort = _mat_mut(list.sort)


def as_handle(tt, rt=None):
    """Device handle of a train given as a TTList (reused when still valid) or any sequence of NumPy cores (uploaded)."""
    rt = rt or get_runtime()
    if isinstance(tt, TTList):
        h = tt.handle()
        if h is not None and h.rt is rt:
            return h
        return Handle.from_numpy(list(list.__iter__(tt)) if tt._live else tt._h.to_numpy(), rt)
    return Handle.from_numpy(list(tt), rt)


def shapes_of(tt):
    if isinstance(tt, TTList):
        return tt.shapes()
    return [tuple(c.shape) for c in tt]
