"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package (daisyriot_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("daisy_oracle.c", "daisy_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


class _Mesh(C.Structure):
    _fields_ = [("vertices", C.POINTER(C.c_float)), ("nv", C.c_int), ("normals", C.POINTER(C.c_float)),
                ("nn", C.c_int), ("tri", C.POINTER(C.c_int)), ("ntri", C.c_int)]


HIT_DTYPE = np.dtype([("t", np.float32), ("triangleId", np.int32), ("u", np.float32), ("v", np.float32)])


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
        L.orc_surface3.restype = C.c_float
        L.orc_surface3.argtypes = [fp, fp, fp]
        L.orc_surface_tri.restype = C.c_float
        L.orc_p2p_ff.restype = C.c_float
        L.orc_p2p_ff.argtypes = [C.POINTER(_Mesh), C.c_int, C.c_int, C.c_int]
        L.orc_point_ff.restype = C.c_float
        L.orc_point_ff.argtypes = [fp, fp, fp, fp, C.c_float, C.c_int]
        L.orc_unoccluded_rows.argtypes = [C.POINTER(_Mesh), C.c_int, C.c_int, C.c_int, fp]
        L.orc_pair_ray.argtypes = [C.POINTER(_Mesh), C.c_int, C.c_int, C.c_float, C.c_float, fp]
        L.orc_uv2xyz.argtypes = [C.POINTER(_Mesh), C.c_int, C.c_float, C.c_float, fp]
        L.orc_ray_tri.restype = C.c_int
        L.orc_ray_tri.argtypes = [fp, fp, fp, fp, fp, fp, fp]
        L.orc_bvh_build.restype = C.c_void_p
        L.orc_bvh_build.argtypes = [C.POINTER(_Mesh)]
        L.orc_bvh_free.argtypes = [C.c_void_p]
        L.orc_query_closest.argtypes = [C.POINTER(_Mesh), C.c_void_p, C.c_int, fp, C.c_void_p]
        L.orc_query_closest_brute.argtypes = [C.POINTER(_Mesh), C.c_int, fp, C.c_void_p]
        L.orc_radmat_rows.restype = C.c_int64
        L.orc_radmat_rows.argtypes = [C.POINTER(_Mesh), C.c_void_p, fp, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, fp, C.POINTER(C.c_uint64), C.c_int]
        L.orc_radmat_rowlist.restype = C.c_int64
        L.orc_radmat_rowlist.argtypes = [C.POINTER(_Mesh), C.c_void_p, fp, C.c_int, ip, C.c_int, C.c_int,
                                         C.c_int, C.c_int, fp, C.POINTER(C.c_uint64), C.c_int]
        L.orc_count_ties.restype = C.c_int64
        L.orc_count_ties.argtypes = [C.POINTER(_Mesh), C.c_void_p, fp, C.c_int, ip, C.c_int, C.c_int, C.c_int]
        L.orc_radmat_upper.restype = C.c_int64
        L.orc_radmat_upper.argtypes = [C.POINTER(_Mesh), C.c_void_p, fp, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp,
                                       C.POINTER(C.c_uint64), C.c_int]
        L.orc_gather_pass.argtypes = [fp, C.c_int64, C.c_int, C.c_int, fp, fp, fp, ip, C.c_int,
                                      C.POINTER(C.c_double), C.c_int]
        L.orc_num_procs.restype = C.c_int
        _LIB = L
    return _LIB


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Oracle:
    """CPU oracle bound to one mesh (arrays in the reference MeshS layout)."""

    FF_DEVICE, FF_HOST = 0, 1

    def __init__(self, vertices, normals, tri):
        self.L = lib()
        self.vertices = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
        self.normals = np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
        self.tri = np.ascontiguousarray(tri, np.int32).reshape(-1, 6)
        self.N = self.tri.shape[0]
        self.mesh = _Mesh(_fp(self.vertices), self.vertices.shape[0], _fp(self.normals), self.normals.shape[0],
                          self.tri.ctypes.data_as(C.POINTER(C.c_int)), self.N)
        self._bvh = None

    @classmethod
    def from_scene(cls, sc):
        return cls(sc.vertices, sc.normals, sc.tri)

    @property
    def bvh(self):
        if self._bvh is None:
            self._bvh = self.L.orc_bvh_build(C.byref(self.mesh))
        return self._bvh

    def __del__(self):
        try:
            if self._bvh is not None:
                self.L.orc_bvh_free(self._bvh)
        except Exception:
            pass

    def surface(self, tri):
        return float(self.L.orc_surface_tri(C.byref(self.mesh), int(tri)))

    def p2p_ff(self, o, d, variant=0):
        return np.float32(self.L.orc_p2p_ff(C.byref(self.mesh), int(o), int(d), int(variant)))

    def unoccluded_rows(self, row0, row1, variant=0):
        out = np.empty((row1 - row0, self.N), np.float32)
        self.L.orc_unoccluded_rows(C.byref(self.mesh), row0, row1, variant, _fp(out))
        return out

    def pair_ray(self, o, d, u, v):
        r = np.empty(6, np.float32)
        self.L.orc_pair_ray(C.byref(self.mesh), int(o), int(d), C.c_float(u), C.c_float(v), _fp(r))
        return r

    def pair_rays(self, o, d, uv):
        return np.stack([self.pair_ray(o, d, float(u), float(v)) for u, v in np.asarray(uv, np.float32)])

    def query_closest(self, rays6, brute=False):
        rays6 = np.ascontiguousarray(rays6, np.float32).reshape(-1, 6)
        hits = np.empty(rays6.shape[0], HIT_DTYPE)
        if brute:
            self.L.orc_query_closest_brute(C.byref(self.mesh), rays6.shape[0], _fp(rays6), hits.ctypes.data)
        else:
            self.L.orc_query_closest(C.byref(self.mesh), self.bvh, rays6.shape[0], _fp(rays6), hits.ctypes.data)
        return hits

    def radmat_rows(self, uv, row0, row1, variant=0, reciprocity=False, brute=False, want_masks=True, nthreads=0):
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        F = np.empty((row1 - row0, self.N), np.float32)
        masks = np.empty((row1 - row0, self.N), np.uint64) if want_masks else None
        rays = self.L.orc_radmat_rows(C.byref(self.mesh), self.bvh, _fp(uv), uv.shape[0], row0, row1, variant,
                                      int(reciprocity), int(brute), _fp(F),
                                      masks.ctypes.data_as(C.POINTER(C.c_uint64)) if want_masks else None, nthreads)
        return F, masks, int(rays)

    def radmat_rowlist(self, uv, rows, variant=0, reciprocity=False, brute=False, want_masks=True, nthreads=0):
        """radmat_rows for an arbitrary list of rows (one host thread per row)."""
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        rows = np.ascontiguousarray(rows, np.int32)
        F = np.empty((rows.size, self.N), np.float32)
        masks = np.empty((rows.size, self.N), np.uint64) if want_masks else None
        rays = self.L.orc_radmat_rowlist(C.byref(self.mesh), self.bvh, _fp(uv), uv.shape[0], rows.ctypes.data_as(C.POINTER(C.c_int)),
                                         rows.size, variant, int(reciprocity), int(brute), _fp(F),
                                         masks.ctypes.data_as(C.POINTER(C.c_uint64)) if want_masks else None, nthreads)
        return F, masks, int(rays)

    def count_ties(self, uv, rows, variant=0, nthreads=0):
        """Visibility rays of the listed rows whose closest-hit distance is shared exactly by two different triangles."""
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        rows = np.ascontiguousarray(rows, np.int32)
        return int(self.L.orc_count_ties(C.byref(self.mesh), self.bvh, _fp(uv), uv.shape[0], rows.ctypes.data_as(C.POINTER(C.c_int)),
                                         rows.size, variant, nthreads))

    def radmat_upper(self, uv, row0, row1, variant=0, nthreads=0):
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        F_rc = np.empty((row1 - row0, self.N), np.float32)
        F_cr = np.empty((row1 - row0, self.N), np.float32)
        masks = np.empty((row1 - row0, self.N), np.uint64)
        rays = self.L.orc_radmat_upper(C.byref(self.mesh), self.bvh, _fp(uv), uv.shape[0], row0, row1, variant, _fp(F_rc), _fp(F_cr),
                                       masks.ctypes.data_as(C.POINTER(C.c_uint64)), nthreads)
        return F_rc, F_cr, masks, int(rays)


def row_checksums(F_rows, mask_rows):
    """Order-free per-row digests used by the whole-matrix golden files:
    (sum over columns of mask*(2c+1) mod 2^64, float64 row sum of F, xor of the F bit patterns)."""
    N = F_rows.shape[1]
    w = (2 * np.arange(N, dtype=np.uint64) + np.uint64(1))
    with np.errstate(over="ignore"):
        mh = (mask_rows * w[None, :]).sum(axis=1, dtype=np.uint64)
    return mh, F_rows.astype(np.float64).sum(axis=1), np.bitwise_xor.reduce(F_rows.view(np.uint32), axis=1)


def ray_tri(ray6, a, b, c):
    L = lib()
    ray6, a, b, c = (np.ascontiguousarray(x, np.float32) for x in (ray6, a, b, c))
    t, u, v = C.c_float(), C.c_float(), C.c_float()
    ok = L.orc_ray_tri(_fp(ray6), _fp(a), _fp(b), _fp(c), C.byref(t), C.byref(u), C.byref(v))
    return bool(ok), t.value, u.value, v.value


def gather_pass(F, res, B, M, mat_idx, accum=1, nthreads=0):
    """In-place pass on band-major (K,N) float32 arrays; returns per-band residual sums (float64[K])."""
    L = lib()
    assert F.dtype == np.float32 and F.flags.c_contiguous
    K, N = res.shape
    assert res.dtype == np.float32 and B.dtype == np.float32 and res.flags.c_contiguous and B.flags.c_contiguous
    M = np.ascontiguousarray(M, np.float32)
    mat_idx = np.ascontiguousarray(mat_idx, np.int32)
    sums = np.zeros(K, np.float64)
    L.orc_gather_pass(_fp(F), F.shape[1], N, K, _fp(res), _fp(B), _fp(M), mat_idx.ctypes.data_as(C.POINTER(C.c_int)),
                      accum, sums.ctypes.data_as(C.POINTER(C.c_double)), nthreads)
    return sums


def num_procs():
    return int(lib().orc_num_procs())
