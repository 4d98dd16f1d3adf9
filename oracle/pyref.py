"""ctypes binding of oracle/_ref/libdaisy_ref.so -- the reference's OWN sources compiled by
oracle/ref_build/Makefile.  TEST INFRASTRUCTURE ONLY (pins the oracle; "reference" CPU baseline)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libdaisy_ref.so")
_LIB = None


def available() -> bool:
    return os.path.exists(SO)


def build() -> bool:
    """(Re)build from /root/reference when it is present; otherwise use the prebuilt file if any."""
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "ref_build")], stdout=subprocess.DEVNULL)
    return available()


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise RuntimeError("oracle/_ref/libdaisy_ref.so missing (build it where /root/reference exists)")
        L = C.CDLL(SO)
        fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
        L.ref_scene_load.restype = vp
        L.ref_scene_load.argtypes = [C.c_char_p, C.c_char_p, fp, C.c_int]
        L.ref_scene_from_arrays.restype = vp
        L.ref_scene_from_arrays.argtypes = [fp, C.c_int, fp, C.c_int, ip, C.c_int]
        L.ref_scene_free.argtypes = [vp]
        L.ref_scene_counts.argtypes = [vp, ip, ip, ip, ip]
        L.ref_scene_arrays.argtypes = [vp, fp, fp, ip, ip]
        L.ref_material.argtypes = [vp, C.c_int, fp, fp, fp, fp, fp]
        L.ref_material_set_M.argtypes = [vp, C.c_int, fp]
        L.ref_calculateSurface3.restype = C.c_float
        L.ref_calculateSurface3.argtypes = [fp, fp, fp]
        L.ref_calculateSurface.restype = C.c_float
        L.ref_calculateSurface.argtypes = [vp, C.c_int]
        for n in ("ref_calculateCentre", "ref_avgNormal", "ref_divideInFourTriangles"):
            getattr(L, n).argtypes = [vp, C.c_int, fp]
        L.ref_uv2xyz.argtypes = [vp, C.c_int, C.c_float, C.c_float, fp]
        L.ref_calcPointFormfactor.restype = C.c_float
        L.ref_calcPointFormfactor.argtypes = [fp, fp, fp, fp, C.c_float]
        L.ref_p2pFormfactor_unoccluded.restype = C.c_float
        L.ref_p2pFormfactor_unoccluded.argtypes = [vp, C.c_int, C.c_int]
        L.ref_pair_ray.argtypes = [vp, C.c_int, C.c_int, C.c_float, C.c_float, fp]
        L.ref_set_triplets.argtypes = [vp, ip, ip, C.POINTER(C.c_double), C.c_int64]
        L.ref_lightning_create.restype = C.c_int
        L.ref_lightning_create.argtypes = [vp, C.c_int, C.c_float]
        L.ref_lightning_reset.argtypes = [vp]
        L.ref_lightning_increment.restype = C.c_int
        L.ref_lightning_increment.argtypes = [vp]
        L.ref_lightning_converge.restype = C.c_int
        L.ref_lightning_converge.argtypes = [vp]
        L.ref_lightning_bands.restype = C.c_int
        L.ref_lightning_bands.argtypes = [vp]
        L.ref_lightning_read.argtypes = [vp, fp, fp]
        L.ref_lightning_color.argtypes = [vp, C.c_int, fp]
        L.ref_lightning_pass_only.argtypes = [vp]
        L.ref_radmat_nnz.restype = C.c_int64
        L.ref_radmat_nnz.argtypes = [vp]
        _LIB = L
    return _LIB


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class _quiet_stdout:
    """The reference's C code printf()s while loading tables; keep that off the caller's stdout (fd 1)."""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        lib().ref_flush_stdout()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.null)


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class RefScene:
    def __init__(self, handle, wavelengths=None):
        self.L = lib()
        self.h = handle
        self.wavelengths = wavelengths
        nv, nn, nt, nm = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self.L.ref_scene_counts(self.h, C.byref(nv), C.byref(nn), C.byref(nt), C.byref(nm))
        self.nv, self.nn, self.N, self.nmat = nv.value, nn.value, nt.value, nm.value

    @classmethod
    def load(cls, obj, mtl_dir, wavelengths, coeff_cwd):
        """MeshS::loadFromFile; ``coeff_cwd`` is a directory holding color_tables/srgb.coeff."""
        wl = np.ascontiguousarray(wavelengths, np.float32)
        old = os.getcwd()
        os.chdir(coeff_cwd)
        try:
          with _quiet_stdout():
            h = lib().ref_scene_load(os.path.abspath(obj).encode() if os.path.isabs(obj) else obj.encode(),
                                     mtl_dir.encode(), _fp(wl), wl.size)
        finally:
            os.chdir(old)
        return cls(h, wl)

    @classmethod
    def from_arrays(cls, vertices, normals, tri):
        v = np.ascontiguousarray(vertices, np.float32)
        n = np.ascontiguousarray(normals, np.float32)
        t = np.ascontiguousarray(tri, np.int32)
        return cls(lib().ref_scene_from_arrays(_fp(v), v.shape[0], _fp(n), n.shape[0], _ip(t), t.shape[0]))

    def close(self):
        if self.h:
            self.L.ref_scene_free(self.h)
            self.h = None

    def arrays(self):
        v = np.empty((self.nv, 3), np.float32)
        n = np.empty((self.nn, 3), np.float32)
        t = np.empty((self.N, 6), np.int32)
        m = np.empty(self.N, np.int32)
        self.L.ref_scene_arrays(self.h, _fp(v), _fp(n), _ip(t), _ip(m))
        return v, n, t, m

    def material(self, i):
        K = self.wavelengths.size
        rgb, em = np.empty(3, np.float32), np.empty(3, np.float32)
        sv, se, M = np.empty(K, np.float32), np.empty(K, np.float32), np.empty((K, K), np.float32)
        self.L.ref_material(self.h, i, _fp(rgb), _fp(em), _fp(sv), _fp(se), _fp(M))
        return dict(rgbcolor=rgb, emission=em, spectral_values=sv, spectral_emission=se, M=M.T.copy())

    def set_material_M(self, i, M):
        Mcm = np.ascontiguousarray(np.asarray(M, np.float32).T)
        self.L.ref_material_set_M(self.h, i, _fp(Mcm))

    def surface(self, tri):
        return np.float32(self.L.ref_calculateSurface(self.h, int(tri)))

    def centre(self, tri):
        o = np.empty(3, np.float32); self.L.ref_calculateCentre(self.h, int(tri), _fp(o)); return o

    def avg_normal(self, tri):
        o = np.empty(3, np.float32); self.L.ref_avgNormal(self.h, int(tri), _fp(o)); return o

    def divide4(self, tri):
        o = np.empty((4, 3, 3), np.float32); self.L.ref_divideInFourTriangles(self.h, int(tri), _fp(o)); return o

    def uv2xyz(self, tri, u, v):
        o = np.empty(3, np.float32); self.L.ref_uv2xyz(self.h, int(tri), C.c_float(u), C.c_float(v), _fp(o)); return o

    def p2p_unoccluded(self, o, d):
        return np.float32(self.L.ref_p2pFormfactor_unoccluded(self.h, int(o), int(d)))

    def pair_ray(self, row, col, u, v):
        o = np.empty(6, np.float32); self.L.ref_pair_ray(self.h, int(row), int(col), C.c_float(u), C.c_float(v), _fp(o)); return o

    # ---- Lightning.h ----
    def set_triplets_from_dense(self, F):
        """RadMat := the non-zeros of a dense float matrix (what setFromTriplets receives, as doubles)."""
        r, c = np.nonzero(F)
        vals = F[r, c].astype(np.float64)
        r32, c32 = np.ascontiguousarray(r, np.int32), np.ascontiguousarray(c, np.int32)
        self.L.ref_set_triplets(self.h, _ip(r32), _ip(c32), vals.ctypes.data_as(C.POINTER(C.c_double)), vals.size)
        return vals.size

    def lightning_create(self, method, emission_value):
        return int(self.L.ref_lightning_create(self.h, int(method), C.c_float(emission_value)))

    def lightning_reset(self):
        self.L.ref_lightning_reset(self.h)

    def lightning_increment(self):
        return int(self.L.ref_lightning_increment(self.h))

    def lightning_converge(self):
        return int(self.L.ref_lightning_converge(self.h))

    def lightning_pass_only(self):
        self.L.ref_lightning_pass_only(self.h)

    def lightning_read(self):
        K = int(self.L.ref_lightning_bands(self.h))
        B, R = np.empty((K, self.N), np.float32), np.empty((K, self.N), np.float32)
        self.L.ref_lightning_read(self.h, _fp(B), _fp(R))
        return B, R

    def lightning_color(self, patch):
        o = np.empty(3, np.float32); self.L.ref_lightning_color(self.h, int(patch), _fp(o)); return o


def calc_point_ff(op, on, dp, dn, surface):
    op, on, dp, dn = (np.ascontiguousarray(x, np.float32) for x in (op, on, dp, dn))
    return np.float32(lib().ref_calcPointFormfactor(_fp(op), _fp(on), _fp(dp), _fp(dn), C.c_float(surface)))


def calculate_surface3(a, b, c):
    a, b, c = (np.ascontiguousarray(x, np.float32) for x in (a, b, c))
    return np.float32(lib().ref_calculateSurface3(_fp(a), _fp(b), _fp(c)))


def camera_rays(width, height, supersampling, antialiasing):
    """Camera::gen_rays_for_screen of the reference's Camera.h -> (W*H*samples, 6) float32."""
    L = lib()
    samples = supersampling if antialiasing else 1
    out = np.empty((width * height * samples, 6), np.float32)
    L.ref_camera_rays.restype = C.c_int
    L.ref_camera_rays.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
    n = L.ref_camera_rays(width, height, supersampling, int(antialiasing), _fp(out))
    assert n == out.shape[0]
    return out


# ---- the reference's own CUDA kernel (parallellism.cu), compiled unmodified: GPU box only -------------------------
def cuda_kernel_available(nofma: bool = False) -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libdaisy_ref_cuda_nofma.so" if nofma else "libdaisy_ref_cuda.so"))


def cuda_run_calculate_radiosity_matrix(vertices, normals, tri, nofma: bool = False):
    """parallellism::runCalculateRadiosityMatrix on the current GPU -> (N x N float64 of Tripl.m_value, wall seconds)."""
    L = C.CDLL(os.path.join(_HERE, "_ref", "libdaisy_ref_cuda_nofma.so" if nofma else "libdaisy_ref_cuda.so"))
    f = L.ref_cuda_runCalculateRadiosityMatrix
    f.restype = C.c_double
    f.argtypes = [C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_double)]
    v = np.ascontiguousarray(vertices, np.float32)
    n = np.ascontiguousarray(normals, np.float32)
    t = np.ascontiguousarray(tri, np.int32)
    out = np.zeros((t.shape[0], t.shape[0]), np.float64)
    import sys
    sys.stdout.flush()
    saved, null = os.dup(1), os.open(os.devnull, os.O_WRONLY)
    os.dup2(null, 1)  # the reference draws a progress bar on stdout
    try:
        secs = f(_fp(v), v.shape[0], _fp(n), n.shape[0], _ip(t), t.shape[0], out.ctypes.data_as(C.POINTER(C.c_double)))
    finally:
        os.dup2(saved, 1); os.close(saved); os.close(null)
    return out, float(secs)
