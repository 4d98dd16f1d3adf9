"""Quick device-side timing probe (not the benchmark): LBVH, form-factor build and gather pass at a few sizes."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import daisyriot_b200 as dz
from daisyriot_b200 import _lib, scenes

sizes = [int(a) for a in sys.argv[1:]] or [8192, 32768]
uv = scenes.msvc_sample_pattern(1)
L = _lib.lib()
for N in sizes:
    sc = scenes.cornell_box(N)
    t = time.time()
    p = dz.OptixPrimeFunctionality(dz.MeshS.from_scene(sc), rands=uv)
    t1 = time.time()
    p.cudaCalculateRadiosityMatrix()
    t2 = time.time()
    st = p.stats()
    print(f"N={N} ctx {t1-t:.3f}s lbvh {st['lbvh_ms']:.2f}ms ff {st['ff_ms']:.1f}ms wall {t2-t1:.3f}s pairs {st['pairs_traced']} "
          f"({st['pairs_traced']/(N*(N-1)/2):.3f}) fallback {st['pairs_fallback']/max(1,st['pairs_traced']):.3f} rays/s {st['rays']/(st['ff_ms']*1e-3):.3e}", flush=True)
    for K in (9, 3, 1, 32):
        rng = np.random.RandomState(0)
        E = rng.uniform(0, 1, (K, N)).astype(np.float32)
        M = rng.uniform(0, 0.1, (len(sc.materials), K, K)).astype(np.float32)
        s = C.c_void_p()
        _lib.check(L.daisy_solver_create(p._ctx, K, _lib.fptr(E), _lib.fptr(M), M.shape[0], _lib.iptr(sc.mat_idx), C.byref(s)))
        sums = np.zeros(K)
        for _ in range(3):
            _lib.check(L.daisy_solver_step(s, sums.ctypes.data_as(C.POINTER(C.c_double))))
        ms = []
        for _ in range(10):
            _lib.check(L.daisy_solver_step(s, sums.ctypes.data_as(C.POINTER(C.c_double))))
            m = C.c_double()
            L.daisy_solver_last_step_ms(s, C.byref(m))
            ms.append(m.value)
        t = time.time()
        for _ in range(20):
            _lib.check(L.daisy_solver_step(s, None))
        L.daisy_solver_band_sums(s, sums.ctypes.data_as(C.POINTER(C.c_double)))
        wall = (time.time() - t) / 20
        byt = 4.0 * N * N + 16.0 * N * K
        print(f"   K={K}: step {np.median(ms):.3f} ms (min {min(ms):.3f})  {byt/np.median(ms)/1e6:.0f} GB/s  async wall/step {wall*1e3:.3f} ms", flush=True)
        L.daisy_solver_destroy(s)
    p.close()
