"""Resident mode of the per-function drop-in: the reference's unmodified Python driver with its arrays living on the GPU.

The reference's driver owns every array as a numpy array, passes them to ~21 ``fimera`` calls per step and mutates
them in Python in between (``J[:] = 0``, ``Rho += BckGrndRho``, ``gradRho_fb_prv[:] = gradRho_fb_nxt``, ``EB[:] = 0``,
``vec_fb[:] = J_fb``, ``resize``; reference moduls/chimera_main.py:110-190, species.py:234-256, solvers.py:318).  Copying
each argument in and out of every call costs 56 GB of PCIe traffic per LWFA step.  Resident mode removes the copies
without asking anything of the driver:

1. **numpy's data allocator** (NEP 49, ``PyDataMem_SetHandler``) is pointed at CUDA *managed* memory for blocks of
   1 MB and more (``csrc/npalloc.c`` -> ``chimera_managed_alloc`` of libchimera_b200.so).  Arrays the driver creates with
   ``np.zeros`` / arithmetic / ``resize`` are then plain numpy arrays that own their data (the driver's ``resize`` calls
   keep working) and whose pointer is valid on the device: every C-ABI entry point takes it as it is
   (``csrc/api_host.cu`` ``Call::up``), prefetches it -- a no-op once the pages are on the device -- and computes in place.
   Whatever the driver does to an array on the host stays correct, with no bookkeeping here: the CUDA driver migrates
   the touched pages on access.
2. **The driver's whole-array statements run on the device**: the shim returns :class:`ResidentArray` (an ``ndarray``
   subclass; views taken from it, e.g. ``EG_fb[:,:,:,3:]``, are ResidentArrays too) whose ``a[:] = scalar``,
   ``a[:] = other`` and ``a += other`` go to ``chimera_fill`` / ``chimera_copy`` / ``chimera_add_inplace`` when both
   sides are device-accessible and contiguous, so these statements do not pull the pages back to the host.  Anything
   else falls through to numpy on the same memory.

3. **Arrays the driver only ever passes as inputs** (``gradRho_fb_prv``, ``BckGrndRho``: never returned by a call, so the
   driver keeps its plain ``np.zeros`` object and ``prv[:] = nxt`` would be a CPU copy that drags both arrays to the
   host) are upgraded where they are held: the first time such an array is seen as an ``intent(in)`` argument, the
   entries of the dictionaries that reference it (``solver.Data``, an instance ``__dict__``) are rebound to a
   ResidentArray *view* of the same memory -- what the driver itself does with every in/out argument
   (``self.Data[k] = chimera.f(self.Data[k], ...)``).  Particle arrays are left alone (the driver resizes them, a view
   cannot be resized).  ``upgrade_inputs = False`` switches this off.

4. **Other numpy expressions on resident arrays** (the driver's ``(momenta[0] * weights).sum() / weights.sum()`` under
   'StaticKick', chimera_main.py:121-122; the reductions of moduls/diagnostics.py) would pull their operands to the host
   page by page.  ``ResidentArray.__array_ufunc__`` runs the common element-wise ufuncs and the add / max / min
   reductions on the device instead -- through torch, on the very same memory (``__cuda_array_interface__``; strided
   views such as ``momenta[0]`` included) -- when an operand is large; everything else, and everything small, is numpy on
   the host as before.  Results are new numpy arrays (managed memory) or scalars.  ``offload_ufuncs = False`` switches
   it off.

``enable()`` switches it on for the process (also: environment ``CHIMERA_B200_RESIDENT=1`` before importing
``chimera_b200.fimera``); without a CUDA device it raises.  In resident mode use the object a call RETURNS, as the
reference's driver does (``self.Data[k] = chimera.f(self.Data[k], ...)``): a particle array that is not yet resident
is replaced by a resident owner on its first in/out call.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
_state = {"on": False, "np": None}
THRESHOLD = 1 << 20
upgrade_inputs = True
offload_ufuncs = True
_seen_inputs = set()
_torch = {"mod": None, "map": None}


def _npalloc():
    if _state["np"] is None:
        path = os.path.join(_HERE, "_npalloc.so")
        if not os.path.exists(path):
            raise ImportError("%s not found -- run `python -m chimera_b200.build`" % path)
        h = ctypes.PyDLL(path)
        h.chb_npalloc_install.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_size_t]
        h.chb_npalloc_count.restype = ctypes.c_longlong
        h.chb_npalloc_bytes.restype = ctypes.c_longlong
        _state["np"] = h
    return _state["np"]


def _fnptr(lib, name):
    return ctypes.cast(getattr(lib, name), ctypes.c_void_p)


def enabled():
    return _state["on"]


def enable(threshold=THRESHOLD):
    """numpy allocates blocks >= ``threshold`` bytes in CUDA managed memory from now on; the shim returns ResidentArrays"""
    if _state["on"]:
        return
    lib = _lib.load()
    if _lib.device_count() == 0:
        raise RuntimeError("chimera_b200 resident mode needs a CUDA device (there is no CPU fallback)")
    lib.chimera_managed_touched.restype = None
    lib.chimera_managed_alloc.restype = ctypes.c_void_p
    lib.chimera_managed_realloc.restype = ctypes.c_void_p
    probe = lib.chimera_managed_alloc(ctypes.c_size_t(4096), 1)  # creates the context; fails early without managed memory
    if not probe:
        raise RuntimeError("cudaMallocManaged failed: resident mode is not available on this device")
    lib.chimera_managed_free(ctypes.c_void_p(probe))
    rc = _npalloc().chb_npalloc_install(_fnptr(lib, "chimera_managed_alloc"), _fnptr(lib, "chimera_managed_realloc"),
                                        _fnptr(lib, "chimera_managed_free"), _fnptr(lib, "chimera_managed_owns"),
                                        ctypes.c_size_t(int(threshold)))
    if rc != 0:
        raise RuntimeError("could not install the numpy data allocator (rc=%d)" % rc)
    _state["on"] = True


def disable():
    """back to numpy's own allocator; arrays allocated meanwhile stay valid (each remembers its allocator)"""
    if _state["on"]:
        _npalloc().chb_npalloc_uninstall()
        _state["on"] = False
        _lib.load().chimera_managed_trim()  # freed blocks are cached for reuse while the mode is on


def stats():
    h = _npalloc()
    return {"managed_blocks": int(h.chb_npalloc_count()), "managed_bytes": int(h.chb_npalloc_bytes())}


def accessible(a) -> bool:
    """is the array's memory visible to the device (managed or device memory)?"""
    return bool(a.size) and _lib.load().chimera_is_device_accessible(ctypes.c_void_p(a.ctypes.data)) != 0


def _contig(a):
    return a.flags.f_contiguous or a.flags.c_contiguous


def _full(key, ndim):
    if key is Ellipsis or (isinstance(key, slice) and key == slice(None)):
        return True
    if isinstance(key, tuple) and len(key) <= ndim:
        return all(k is Ellipsis or (isinstance(k, slice) and k == slice(None)) for k in key)
    return False


def _check(rc):
    if rc != 0:
        raise RuntimeError("libchimera_b200: status %d: %s" % (rc, _lib.load().chimera_last_error().decode()))


class _CudaView:
    """``__cuda_array_interface__`` of a numpy array in managed / device memory (shape, strides, dtype as they are)"""

    def __init__(self, a):
        self.__cuda_array_interface__ = {"shape": a.shape, "typestr": a.dtype.str, "data": (a.ctypes.data, False), "version": 2,
                                         "strides": a.strides if a.size else None}


def _torch_ops():
    if _torch["map"] is None:
        import torch

        _torch["mod"] = torch
        _torch["map"] = {np.add: torch.add, np.subtract: torch.sub, np.multiply: torch.mul, np.true_divide: torch.div,
                         np.negative: torch.neg, np.absolute: torch.abs, np.conjugate: torch.conj_physical,
                         np.square: torch.square, np.sqrt: torch.sqrt, np.exp: torch.exp, np.maximum: torch.maximum,
                         np.minimum: torch.minimum}
    return _torch["mod"], _torch["map"]


def _offload(ufunc, method, inputs, out, kwargs):
    """the ufunc on the device through torch, or NotImplemented"""
    if method not in ("__call__", "reduce") or not any(isinstance(x, np.ndarray) and x.nbytes >= THRESHOLD for x in inputs):
        return NotImplemented
    for x in inputs:
        if isinstance(x, np.ndarray):
            if x.dtype.kind not in "fc" or x.dtype.itemsize not in (8, 16) or (x.nbytes >= 4096 and not accessible(x)):
                return NotImplemented
        elif not isinstance(x, (int, float, complex, np.number)):
            return NotImplemented
    try:
        torch, ops = _torch_ops()
    except Exception:
        return NotImplemented
    as_t = lambda x: torch.as_tensor(_CudaView(x), device="cuda") if isinstance(x, np.ndarray) and x.nbytes >= 4096 else (  # noqa: E731
        torch.as_tensor(np.asarray(x), device="cuda") if isinstance(x, np.ndarray) else x)
    if method == "__call__":
        op = ops.get(ufunc)
        if op is None or set(kwargs) - {"casting", "order", "subok"}:
            return NotImplemented
        if ufunc is np.absolute and any(isinstance(x, np.ndarray) and x.dtype.kind == "c" for x in inputs):
            rdtype = np.dtype("float64")
        else:
            rdtype = np.result_type(*[x.dtype if isinstance(x, np.ndarray) else x for x in inputs])
        shape = np.broadcast_shapes(*[x.shape for x in inputs if isinstance(x, np.ndarray)])
        if out is not None:
            res = out[0]
            if not (isinstance(res, np.ndarray) and res.shape == shape and res.dtype == rdtype and accessible(res)):
                return NotImplemented
        else:
            forder = all(x.flags.f_contiguous for x in inputs if isinstance(x, np.ndarray) and x.ndim > 1)
            res = np.empty(shape, dtype=rdtype, order="F" if forder else "C").view(ResidentArray)
            if res.nbytes >= 4096 and not accessible(res):
                return NotImplemented
        op(*[as_t(x) for x in inputs], out=as_t(res))
        torch.cuda.current_stream().synchronize()
        return res
    # reductions: sum / max / min over all axes or the given ones
    red = {np.add: "sum", np.maximum: "amax", np.minimum: "amin"}.get(ufunc)
    if red is None or out is not None or set(kwargs) - {"axis", "keepdims", "dtype"} or kwargs.get("dtype") is not None:
        return NotImplemented
    (x,) = inputs
    t, axis = as_t(x), kwargs.get("axis", 0)
    dims = tuple(range(x.ndim)) if axis is None else (axis if isinstance(axis, tuple) else (axis,))
    if red != "sum" and x.dtype.kind == "c":
        return NotImplemented
    r = getattr(torch, red)(t, dim=dims, keepdim=bool(kwargs.get("keepdims", False)))
    if r.ndim == 0:
        return x.dtype.type(r.item())
    res = np.empty(tuple(r.shape), dtype=x.dtype, order="F").view(ResidentArray)
    if res.nbytes >= 4096 and accessible(res):
        as_t(res).copy_(r)
        torch.cuda.current_stream().synchronize()
        return res
    res[...] = r.cpu().numpy()
    return res


def _host_touch(*arrays):
    """the host is about to read or write these arrays: the library prefetches a managed block only once per
    (re)allocation (csrc/api_host.cu managed_needs_prefetch), so tell it which blocks went back to the host"""
    if not _state["on"]:
        return
    lib = _lib.load()
    for a in arrays:
        if isinstance(a, np.ndarray) and a.nbytes >= THRESHOLD:
            lib.chimera_managed_touched(ctypes.c_void_p(a.ctypes.data))


class ResidentArray(np.ndarray):
    """numpy array in CUDA managed memory whose whole-array statements run on the device (module docstring)"""

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        if _state["on"] and offload_ufuncs:
            r = _offload(ufunc, method, inputs, out, kwargs)
            if r is not NotImplemented:
                return r
        # numpy on the host, on the same memory (the CUDA driver migrates what is touched)
        _host_touch(*inputs)
        if out is not None:
            _host_touch(*out)
        args = [np.asarray(x) if isinstance(x, ResidentArray) else x for x in inputs]
        if out is not None:
            kwargs["out"] = tuple(np.asarray(o) if isinstance(o, ResidentArray) else o for o in out)
        r = getattr(ufunc, method)(*args, **kwargs)
        if out is not None:
            return out[0] if len(out) == 1 else out
        return r.view(ResidentArray) if isinstance(r, np.ndarray) and r.nbytes >= THRESHOLD else r

    def __setitem__(self, key, value):
        if _state["on"] and self.size and self.dtype.kind in "fc" and self.dtype.itemsize in (8, 16) and _contig(self) \
                and _full(key, self.ndim) and accessible(self):
            lib = _lib.load()
            if isinstance(value, (int, float, complex, np.number)):
                v = complex(value)
                if self.dtype.kind == "c" or v.imag == 0.0:
                    _check(lib.chimera_fill(ctypes.c_void_p(self.ctypes.data), ctypes.c_longlong(self.size), ctypes.c_double(v.real),
                                            ctypes.c_double(v.imag), int(self.dtype.kind == "c")))
                    return
            elif isinstance(value, np.ndarray) and value.shape == self.shape and value.dtype == self.dtype and \
                    value.flags.f_contiguous == self.flags.f_contiguous and _contig(value) and accessible(value):
                _check(lib.chimera_copy(ctypes.c_void_p(self.ctypes.data), ctypes.c_void_p(value.ctypes.data),
                                        ctypes.c_longlong(self.nbytes)))
                return
        _host_touch(self, value)
        super().__setitem__(key, value)

    def __iadd__(self, other):
        if _state["on"] and isinstance(other, np.ndarray) and other.shape == self.shape and other.dtype == self.dtype and \
                self.dtype.kind in "fc" and self.dtype.itemsize in (8, 16) and _contig(self) and _contig(other) and \
                other.flags.f_contiguous == self.flags.f_contiguous and self.size and accessible(self) and accessible(other):
            n = self.size * (2 if self.dtype.kind == "c" else 1)
            _check(_lib.load().chimera_add_inplace(ctypes.c_void_p(self.ctypes.data), ctypes.c_void_p(other.ctypes.data),
                                                   ctypes.c_longlong(n)))
            return self
        _host_touch(self, other)
        return super().__iadd__(other)


def _particle_like(a):
    """arrays the driver resizes (species.py:234-254, 394-398): (3|4|6, Np) and (Np,) real arrays"""
    return a.dtype == np.float64 and (a.ndim == 1 or (a.ndim == 2 and a.shape[0] in (3, 4, 6)))


def adopt(a):
    """what the shim hands back for an in/out or out array in resident mode: the array itself when it already is a
    ResidentArray; a ResidentArray view of a grid / spectral array (aliasing preserved); a resident OWNER (one device
    copy, once) of a particle array, which the driver is going to ``resize``"""
    if not _state["on"] or not isinstance(a, np.ndarray) or isinstance(a, ResidentArray) or a.nbytes < THRESHOLD:
        return a
    if _particle_like(a) and a.base is None and a.flags.owndata:
        if accessible(a):
            # already managed (allocated after enable()): an owner cannot change class, so move the data once
            new = ResidentArray(a.shape, dtype=a.dtype, order="F" if a.flags.f_contiguous else "C")
            _check(_lib.load().chimera_copy(ctypes.c_void_p(new.ctypes.data), ctypes.c_void_p(a.ctypes.data), ctypes.c_longlong(a.nbytes)))
            return new
        new = ResidentArray(a.shape, dtype=a.dtype, order="F" if a.flags.f_contiguous else "C")
        np.ndarray.__setitem__(new, Ellipsis, a)
        return new
    return a.view(ResidentArray)


def adopt_input(a):
    """intent(in) array in resident mode: rebind the dictionary entries that hold a plain, device-accessible grid /
    spectral array to a ResidentArray view of it (module docstring, 3.); once per buffer"""
    if not _state["on"] or not upgrade_inputs or type(a) is not np.ndarray or a.nbytes < THRESHOLD or _particle_like(a):
        return a
    key = (a.ctypes.data, a.nbytes)
    if key in _seen_inputs:
        return a
    _seen_inputs.add(key)
    if not (_contig(a) and accessible(a)):
        return a
    import gc

    view = a.view(ResidentArray)
    for ref in gc.get_referrers(a):
        if isinstance(ref, dict):
            for k, v in list(ref.items()):
                if v is a:
                    ref[k] = view
    return view
