"""Per-material spectra and K x K band-mixing matrices (reference ``visual studio/Material.cpp``).

Host-side setup that feeds the gather kernel: for every material the reflectance spectrum, the emission spectrum
and the column-major ``M`` (``Eigen::MatrixXf``) that ``SpectralLightning`` multiplies the bounced light with
(``visual studio/Lightning.h:205-218``).  Classification follows ``visual studio/MeshS.cpp:36-66``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import rgb2spec as r2s

f32 = np.float32


@dataclass
class Material:
    name: str
    kind: str  # "diffuse" | "uvlight" | "fluorescent"
    rgbcolor: np.ndarray  # float32[3]
    emission: np.ndarray  # float32[3]
    spectral_values: np.ndarray  # float32[K]
    spectral_emission: np.ndarray  # float32[K]
    M: np.ndarray  # float32[K,K], M[i,j] like Eigen M(i,j)


def _rgb_to_spectrum(model: r2s.RGB2Spec, rgb, wavelengths) -> np.ndarray:
    # Material::rgb_to_spectrum, Material.cpp:35-45
    coeff = model.fetch(rgb)
    return np.array([r2s.eval_precise(coeff, w) for w in wavelengths], np.float32)


def _bell(x: float) -> np.float32:
    # UVLightMaterial::sample_bell_curve, Material.cpp:68-78: float a=1,b=350,c=10; double math; returns float
    a, b, c = f32(1.0), f32(350.0), f32(10.0)
    d = float(f32(f32(x) - b))
    return f32(float(a) * math.exp(-1.0 * (d ** 2.0) / (2 * (float(c) ** 2.0))))


def make_material(name, Kd, Ke, Ks, wavelengths, model: r2s.RGB2Spec) -> Material:
    """One entry of ``MeshS::materials`` (MeshS.cpp:36-66)."""
    Kd, Ke, Ks = (np.asarray(v, np.float32) for v in (Kd, Ke, Ks))
    wl = np.asarray(wavelengths, np.float32)
    K = wl.size
    # float sums in source order: emission[0]+emission[1]+emission[2] > 0 && diffuse sum == 0.0
    blacklightsource = (f32(f32(Ke[0] + Ke[1]) + Ke[2]) > 0) and (f32(f32(Kd[0] + Kd[1]) + Kd[2]) == 0.0)
    fluorescent = f32(f32(Ks[0] + Ks[1]) + Ks[2]) > 0.0
    if blacklightsource:
        # UVLightMaterial({0,0,0},{0,0,0}, wavelengths), Material.cpp:47-66.  The base constructor runs on black
        # (NaN spectra), then emission is resampled from the bell curve and reflectance zeroed.
        # DEVIATION: the reference sets M(i,i) = emission[i] where `emission` is the glm::vec3 *parameter*
        # (0,0,0) indexed with i up to K-1 (Material.cpp:52-54) -- undefined behaviour for i >= 3.  Restated as
        # M = 0: a lamp reflects nothing, consistent with spectral_values[i] = 0 (Material.cpp:65).
        spec_e = np.array([_bell(float(w)) for w in wl], np.float32)
        return Material(name, "uvlight", np.zeros(3, np.float32), np.zeros(3, np.float32),
                        np.zeros(K, np.float32), spec_e, np.zeros((K, K), np.float32))
    spec = _rgb_to_spectrum(model, Kd, wl)
    spec_e = _rgb_to_spectrum(model, Ke, wl)
    if fluorescent:
        # FluorescentMaterial::set_fluorescent_matrix, Material.cpp:90-100: identity, UV columns := spectrum(Ks)
        spec_bl = _rgb_to_spectrum(model, Ks, wl)
        M = np.eye(K, dtype=np.float32)
        for i in range(K):
            if 300.0 < float(wl[i]) < 400.0:
                M[:, i] = spec_bl
        return Material(name, "fluorescent", Kd, Ke, spec, spec_e, M)
    M = np.eye(K, dtype=np.float32)  # Material.cpp:18-21
    for i in range(K):
        M[i, i] = spec[i]
    return Material(name, "diffuse", Kd, Ke, spec, spec_e, M)


def make_materials(scene_materials, wavelengths, model: r2s.RGB2Spec):
    return [make_material(m["name"], m["Kd"], m["Ke"], m["Ks"], wavelengths, model) for m in scene_materials]


# ---- inputs of the three Lightning flavours (Lightning.h:263-292, 359-382, 434-442) ---------------------------
def spectral_inputs(materials, mat_idx, emission_value):
    """E (K,N) band-major and M (nmat,K,K) column-major-per-matrix for SpectralLightning."""
    K = materials[0].spectral_values.size
    ev = f32(emission_value)
    Emat = np.zeros((len(materials), K), np.float32)
    for i, m in enumerate(materials):
        se = m.spectral_emission
        with np.errstate(invalid="ignore"):
            Emat[i] = np.where(se > 0.0, (se * ev).astype(np.float32), f32(0))
    E = np.ascontiguousarray(Emat[mat_idx].T)
    Mcm = np.stack([np.ascontiguousarray(m.M.T) for m in materials]).astype(np.float32)  # [mat][col][row]
    return E, Mcm


def rgb_inputs(materials, mat_idx, emission_value):
    """RGBLightning: residual_c = (F residual_c) * rho_c  ==  M = diag(rho) per material."""
    ev = f32(emission_value)
    nm = len(materials)
    Emat = np.zeros((nm, 3), np.float32)
    Mcm = np.zeros((nm, 3, 3), np.float32)
    for i, m in enumerate(materials):
        for c in range(3):
            if m.emission[c] > 0.0:
                Emat[i, c] = f32(m.emission[c] * ev)
            if m.rgbcolor[c] > 0.0:
                Mcm[i, c, c] = m.rgbcolor[c]
    return np.ascontiguousarray(Emat[mat_idx].T), Mcm


def bw_inputs(materials, mat_idx, emission_value):
    """BWLightning: residual = F residual (no reflectance at all, Lightning.h:410-417)."""
    ev = f32(emission_value)
    Emat = np.array([f32(m.emission[0] * ev) if m.emission[0] > 0.0 else f32(0) for m in materials], np.float32)
    Mcm = np.ones((len(materials), 1, 1), np.float32)
    return np.ascontiguousarray(Emat[mat_idx][None, :]), Mcm
