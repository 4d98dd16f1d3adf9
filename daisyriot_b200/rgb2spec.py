"""RGB -> spectrum upsampling used by the material setup (reference ``visual studio/rgb2spec.cpp``).

Restates, in float32 with the reference's operation order,

* ``rgb2spec_load``          (``rgb2spec.cpp:11-47``; file = "SPEC", u32 res, float scale[res], float data[3*res^3*3])
* ``rgb2spec_find_interval`` (``rgb2spec.cpp:57-76``)
* ``rgb2spec_fetch``         (``rgb2spec.cpp:78-119``, trilinear lookup)
* ``rgb2spec_eval_precise``  (``rgb2spec.cpp:130-134``; the non-FMA branch of ``rgb2spec_fma``, which is what the
  reference's MSVC x64 build without /arch:AVX2 takes)

The coefficient tables themselves (``color_tables/*.coeff``) are missing from the reference checkout
(``.MISSING_LARGE_BLOBS``), so :func:`write_surrogate_table` writes a small table IN THE SAME BINARY FORMAT from a
closed-form recipe.  It is a stand-in for exercising the code path, not Jakob & Hanika's optimised table; spectra it
yields are smooth and bounded in (0,1) but are not colorimetrically exact.
"""
from __future__ import annotations

import struct

import numpy as np

f32 = np.float32
N_COEFFS = 3


class RGB2Spec:
    def __init__(self, res: int, scale: np.ndarray, data: np.ndarray):
        self.res = int(res)
        self.scale = np.ascontiguousarray(scale, np.float32)
        self.data = np.ascontiguousarray(data, np.float32).reshape(-1)

    @classmethod
    def load(cls, filename: str) -> "RGB2Spec":
        with open(filename, "rb") as f:
            if f.read(4) != b"SPEC":
                raise ValueError("not a rgb2spec coefficient file")
            (res,) = struct.unpack("<I", f.read(4))
            scale = np.frombuffer(f.read(4 * res), np.float32)
            n = res * res * res * 3 * N_COEFFS
            data = np.frombuffer(f.read(4 * n), np.float32)
            if scale.size != res or data.size != n:
                raise ValueError("truncated rgb2spec coefficient file")
        return cls(res, scale, data)

    def _find_interval(self, x: np.float32) -> int:
        values, size_ = self.scale, self.res
        left, last_interval = 0, size_ - 2
        size = last_interval
        while size > 0:
            half = size >> 1
            middle = left + half + 1
            if values[middle] < x:
                left = middle
                size -= half + 1
            else:
                size = half
        return min(left, last_interval)

    def fetch(self, rgb) -> np.ndarray:
        rgb = [f32(c) for c in rgb]
        res = self.res
        i = 0
        for j in range(1, 3):
            if rgb[j] >= rgb[i]:
                i = j
        with np.errstate(divide="ignore", invalid="ignore"):
            z = rgb[i]
            scale = f32(res - 1) / z
            x = rgb[(i + 1) % 3] * scale
            y = rgb[(i + 2) % 3] * scale
        # (uint32_t)x of a NaN (black input: 0 * inf) is undefined in C; x86-64 compilers emit a 64-bit cvttss2si
        # whose low half is 0.  The fetched coefficients are NaN either way, callers filter with `> 0`.
        def u32(v):
            if not np.isfinite(v):
                return 0
            return int(v) & 0xFFFFFFFF

        xi = min(u32(x), res - 2)
        yi = min(u32(y), res - 2)
        zi = self._find_interval(z)
        offset = (((i * res + zi) * res + yi) * res + xi) * N_COEFFS
        dx, dy, dz = N_COEFFS, N_COEFFS * res, N_COEFFS * res * res
        with np.errstate(invalid="ignore"):
            x1 = f32(x - f32(xi)); x0 = f32(f32(1.0) - x1)
            y1 = f32(y - f32(yi)); y0 = f32(f32(1.0) - y1)
            z1 = f32(f32(z - self.scale[zi]) / f32(self.scale[zi + 1] - self.scale[zi])); z0 = f32(f32(1.0) - z1)
            d = self.data
            out = np.zeros(N_COEFFS, np.float32)
            for j in range(N_COEFFS):
                o = offset + j
                a = f32(f32(f32(d[o] * x0) + f32(d[o + dx] * x1)) * y0)
                b = f32(f32(f32(d[o + dy] * x0) + f32(d[o + dy + dx] * x1)) * y1)
                c = f32(f32(f32(d[o + dz] * x0) + f32(d[o + dz + dx] * x1)) * y0)
                e = f32(f32(f32(d[o + dz + dy] * x0) + f32(d[o + dz + dy + dx] * x1)) * y1)
                out[j] = f32(f32(f32(a + b) * z0) + f32(f32(c + e) * z1))
        return out


def eval_precise(coeff, lam) -> np.float32:
    c0, c1, c2 = (f32(c) for c in coeff)
    lam = f32(lam)
    x = f32(f32(f32(f32(c0 * lam) + c1) * lam) + c2)
    y = f32(f32(1.0) / np.sqrt(f32(f32(x * x) + f32(1.0)), dtype=np.float32))
    return f32(f32(f32(f32(0.5) * x) * y) + f32(0.5))


def write_surrogate_table(path: str, res: int = 16) -> None:
    """Write a SURROGATE ``.coeff`` file (same binary layout as the real tables, closed-form content)."""
    # same non-linear z spacing idea as the original generator: smoothstep(smoothstep(k/(res-1)))
    t = np.arange(res, dtype=np.float64) / (res - 1)
    sm = lambda v: v * v * (3.0 - 2.0 * v)
    scale = sm(sm(t)).astype(np.float32)
    data = np.zeros((3, res, res, res, N_COEFFS), np.float32)
    centre = {0: 610.0, 1: 545.0, 2: 455.0}  # nm, rough primaries
    for i in range(3):
        for zi in range(res):
            z = float(scale[zi])
            for yi in range(res):
                for xi in range(res):
                    x = xi / (res - 1) * z
                    y = yi / (res - 1) * z
                    rgb = np.zeros(3)
                    rgb[i], rgb[(i + 1) % 3], rgb[(i + 2) % 3] = z, x, y
                    mean = float(np.clip(rgb.mean(), 1e-3, 1 - 1e-3))
                    lo = float(np.clip(rgb.min(), 1e-3, 1 - 1e-3))
                    hi = float(np.clip(rgb.max(), 1e-3, 1 - 1e-3))
                    # peak wavelength = colour-weighted mean of the primaries, curvature from saturation
                    w = rgb + 1e-6
                    lam0 = float((w * np.array([centre[0], centre[1], centre[2]])).sum() / w.sum())
                    sat = (hi - lo)
                    logit = lambda p: (2 * p - 1) / np.sqrt(max(1e-9, 1 - (2 * p - 1) ** 2))
                    top, base = logit(hi), logit(lo)
                    c0 = -(top - base) * sat / (120.0 ** 2)
                    c1 = -2.0 * c0 * lam0
                    c2 = top * sat + logit(mean) * (1 - sat) + c0 * lam0 * lam0
                    data[i, zi, yi, xi] = (c0, c1, c2)
    with open(path, "wb") as f:
        f.write(b"SPEC")
        f.write(struct.pack("<I", res))
        f.write(scale.tobytes())
        f.write(data.tobytes())
