"""ctypes binding of the C ABI in include/b2p.h (libb2p.so, hand-written sm_100a kernels).

This is the only way Python reaches the product: there is no Python/NumPy/torch implementation of
any operator behind it and no CPU fallback. Loading fails loudly when the library has not been
built (``python -c 'import __graft_entry__ as g; g.build()'``) and every entry point fails when no
sm_100 GPU is present. torch is used by the callers only to own device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2p.so")

CURLCURL, ND_MASS, CURLCURL_MASS, H1_DIFFUSION = 0, 1, 2, 3

_lib = None


class B2PError(RuntimeError):
    pass


class OpDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("p", C.c_int), ("ne", C.c_int), ("lsize", C.c_int64),
        ("idx", C.c_void_p), ("orient", C.c_void_p), ("dof_map", C.c_void_p),
        ("Bo", C.c_void_p), ("Bc", C.c_void_p), ("Gc", C.c_void_p),
        ("coeff_ctx", C.c_void_p), ("coeff_ctx_bytes", C.c_size_t), ("assemble_qdata", C.c_int),
    ]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B2PError(f"{LIB_PATH} not built: run __graft_entry__.build() (nvcc, sm_100a). No CPU fallback exists.")
        _lib = C.CDLL(LIB_PATH)
        _lib.b2p_last_error.restype = C.c_char_p
        _lib.b2p_last_error.argtypes = [C.c_void_p]
        _lib.b2p_op_lsize.restype = C.c_int64
        _lib.b2p_op_algorithmic_bytes.restype = C.c_int64
        for name in ("b2p_op_lsize", "b2p_op_algorithmic_bytes", "b2p_op_destroy", "b2p_geom_destroy", "b2p_ctx_destroy"):
            getattr(_lib, name).argtypes = [C.c_void_p]
        _lib.b2p_op_destroy.restype = None
        _lib.b2p_geom_destroy.restype = None
        _lib.b2p_ctx_destroy.restype = None
    return _lib


def _chk(rc, ctx=None):
    if rc != 0:
        msg = lib().b2p_last_error(ctx)
        raise B2PError(f"b2p error {rc}: {msg.decode() if msg else '?'}")


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _vp(x):
    """Device pointer of a torch tensor / int / None."""
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(int(x))


def _stream(stream):
    if stream is None:
        import torch

        return C.c_void_p(torch.cuda.current_stream().cuda_stream)
    return C.c_void_p(int(stream))


class Ctx:
    def __init__(self, device=0, nccl_uid=None, rank=0, nranks=1):
        h = C.c_void_p()
        if nranks > 1:
            buf = (C.c_char * 128).from_buffer_copy(bytes(nccl_uid))
            _chk(lib().b2p_ctx_create_dist(int(device), buf, int(rank), int(nranks), C.byref(h)))
        else:
            _chk(lib().b2p_ctx_create(int(device), C.byref(h)))
        self.h = h
        self.rank, self.nranks = rank, nranks

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = (C.c_char * 128)()
        _chk(lib().b2p_nccl_unique_id(buf))
        return bytes(buf)

    def close(self):
        if self.h:
            lib().b2p_ctx_destroy(self.h)
            self.h = None


class Geom:
    def __init__(self, ctx: Ctx, h):
        self.ctx, self.h = ctx, h

    @classmethod
    def hex(cls, ctx, xe, attr, mesh_order, q1d, nodeB, nodeG, qw1d):
        xe = _np(xe, np.float64)
        attr = _np(attr, np.int32)
        nodeB, nodeG, qw1d = _np(nodeB, np.float64), _np(nodeG, np.float64), _np(qw1d, np.float64)
        h = C.c_void_p()
        _chk(lib().b2p_geom_create_hex(ctx.h, int(xe.shape[0]), int(mesh_order), int(q1d), _ptr(xe), _ptr(nodeB), _ptr(nodeG),
                                       _ptr(qw1d), _ptr(attr), C.byref(h)), ctx.h)
        g = cls(ctx, h)
        g.ne, g.q1d = xe.shape[0], q1d
        return g

    @classmethod
    def from_qdata(cls, ctx, qdata, q1d):
        qdata = _np(qdata, np.float64)
        h = C.c_void_p()
        _chk(lib().b2p_geom_create_qdata(ctx.h, int(qdata.shape[0]), int(q1d), _ptr(qdata), C.byref(h)), ctx.h)
        g = cls(ctx, h)
        g.ne, g.q1d = qdata.shape[0], q1d
        return g

    def qdata(self):
        out = np.empty((self.ne, 11, self.q1d ** 3))
        _chk(lib().b2p_geom_get_qdata(self.h, _ptr(out)), self.ctx.h)
        return out

    def close(self):
        if self.h:
            lib().b2p_geom_destroy(self.h)
            self.h = None


def _desc(kind, p, lsize, idx, orient, dof_map, Bo, Bc, Gc, coeff, assemble):
    keep = []
    d = OpDesc()
    d.kind, d.p = int(kind), int(p)
    idx = _np(idx, np.int32)
    d.ne, d.lsize = idx.shape[0], int(lsize)
    keep.append(idx)
    d.idx = _ptr(idx)
    for name, arr, dt in (("orient", orient, np.int8), ("dof_map", dof_map, np.int32), ("Bo", Bo, np.float64),
                          ("Bc", Bc, np.float64), ("Gc", Gc, np.float64)):
        if arr is not None:
            a = _np(arr, dt)
            keep.append(a)
            setattr(d, name, _ptr(a))
    if coeff is not None:
        c = _np(coeff, np.float64)
        keep.append(c)
        d.coeff_ctx = _ptr(c)
        d.coeff_ctx_bytes = c.nbytes
    d.assemble_qdata = int(bool(assemble))
    return d, keep


class Op:
    """One local partially assembled operator (a ceed::Operator sub-operator)."""

    def __init__(self, ctx, h, lsize):
        self.ctx, self.h, self.lsize = ctx, h, lsize

    @classmethod
    def create(cls, ctx, geom, kind, p, lsize, idx, orient, dof_map, Bo, Bc, Gc, coeff, assemble=False):
        d, keep = _desc(kind, p, lsize, idx, orient, dof_map, Bo, Bc, Gc, coeff, assemble)
        h = C.c_void_p()
        _chk(lib().b2p_op_create(ctx.h, geom.h, C.byref(d), C.byref(h)), ctx.h)
        return cls(ctx, h, lsize)

    def coarsen(self, p, lsize, idx, orient, dof_map, Bo, Bc, Gc):
        d, keep = _desc(0, p, lsize, idx, orient, dof_map, Bo, Bc, Gc, None, False)
        h = C.c_void_p()
        _chk(lib().b2p_op_coarsen(self.h, C.byref(d), C.byref(h)), self.ctx.h)
        return Op(self.ctx, h, lsize)

    def apply(self, x, y, stream=None):
        _chk(lib().b2p_op_apply(self.h, _vp(x), _vp(y), _stream(stream)), self.ctx.h)

    def apply_add(self, x, y, stream=None):
        _chk(lib().b2p_op_apply_add(self.h, _vp(x), _vp(y), _stream(stream)), self.ctx.h)

    def apply_add_ex(self, alpha, x, y, masked=False, simple_kernel=False, stream=None):
        flags = (1 if masked else 0) | (2 if simple_kernel else 0)
        _chk(lib().b2p_op_apply_add_ex(self.h, C.c_double(alpha), _vp(x), _vp(y), flags, _stream(stream)), self.ctx.h)

    def set_essential(self, ess_ldofs):
        e = _np(ess_ldofs, np.int32)
        _chk(lib().b2p_op_set_essential(self.h, _ptr(e), C.c_int64(e.size)), self.ctx.h)

    def diag_add(self, d, stream=None):
        _chk(lib().b2p_op_diag_add(self.h, _vp(d), _stream(stream)), self.ctx.h)

    def set_coeff(self, coeff):
        c = _np(coeff, np.float64)
        _chk(lib().b2p_op_set_coeff(self.h, _ptr(c), C.c_size_t(c.nbytes)), self.ctx.h)

    def algorithmic_bytes(self):
        return int(lib().b2p_op_algorithmic_bytes(self.h))

    def close(self):
        if self.h:
            lib().b2p_op_destroy(self.h)
            self.h = None
