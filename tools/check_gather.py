"""Print the gather kernels' error against the FP64-accumulating oracle for several band counts (GPU box)."""
import ctypes as C
import sys
import numpy as np
sys.path.insert(0, ".")
import daisyriot_b200 as dz
from daisyriot_b200 import _lib, scenes
from oracle import pyoracle

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
sc = scenes.cornell_box(N)
uv = scenes.msvc_sample_pattern(1)
mesh = dz.MeshS.from_scene(sc)
p = dz.OptixPrimeFunctionality(mesh, rands=uv)
L = _lib.lib()
for K in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "9,12,16,32".split(","))]:
    rng = np.random.RandomState(K)
    F = (rng.uniform(0, 1, (N, N)) * (rng.uniform(0, 1, (N, N)) < 0.3) / N).astype(np.float32)
    np.fill_diagonal(F, 0)
    p.loadRadiosityMatrix(F)
    nmat = len(sc.materials)
    M = rng.uniform(0, 0.4, (nmat, K, K)).astype(np.float32)
    E = np.ascontiguousarray(rng.uniform(0, 3, (K, N)).astype(np.float32) * (rng.uniform(0, 1, (K, N)) < 0.2), np.float32)
    s = C.c_void_p()
    _lib.check(L.daisy_solver_create(p._ctx, K, _lib.fptr(E), _lib.fptr(M), nmat, _lib.iptr(sc.mat_idx), C.byref(s)))
    res, B = E.copy(), E.copy()
    for it in range(2):
        sums = np.zeros(K)
        _lib.check(L.daisy_solver_step(s, sums.ctypes.data_as(C.POINTER(C.c_double))))
        pyoracle.gather_pass(F, res, B, M, sc.mat_idx, accum=1)
        Bg, Rg = np.empty_like(E), np.empty_like(E)
        _lib.check(L.daisy_solver_read(s, _lib.fptr(Bg), _lib.fptr(Rg)))
        err = np.abs(Rg - res)
        rel = err.max() / np.abs(res).max()
        bad = np.argwhere(err > 1e-5 * np.abs(res).max())
        print(f"K={K} pass {it}: max abs err {err.max():.3e} rel-to-max {rel:.3e} bad {len(bad)} first {bad[:5].tolist()}", flush=True)
        if len(bad):
            k, i = bad[0]
            print("   got", Rg[k, i], "want", res[k, i], "per-band bad counts", np.bincount(bad[:, 0], minlength=K).tolist())
    L.daisy_solver_destroy(s)
p.close()
