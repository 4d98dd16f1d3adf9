"""GPU box: the gather kernel exactly as ONE rank of an N-way row partition runs it (rank 0's row block of the workload, all
columns, zero matrix -- the traffic does not depend on the values), for the ncu DRAM-traffic captures behind
profiles/gather_traffic.json's "<workload>@<N>" entries.

    python tools/gather_rank_block.py cornell_128k 8
"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
import daisyriot_b200 as dz  # noqa: E402
from daisyriot_b200 import _lib  # noqa: E402

name, nranks = sys.argv[1], int(sys.argv[2])
sc, wl, E, M, tmp = bench.make_workload(name)
N, K, _ = bench.WORKLOADS[name]
p = dz.OptixPrimeFunctionality(dz.MeshS.from_scene(sc), rands=dz.msvc_sample_pattern(1), rank=0, nranks=nranks)
r0, r1 = p.row_range
L = dz.lib()
_lib.check(L.daisy_formfactors_write_rows(p._ctx, r0, 0, None))  # allocates (zero-filled) and marks the matrix present
s = C.c_void_p()
_lib.check(L.daisy_solver_create(p._ctx, K, _lib.fptr(E), _lib.fptr(M), M.shape[0], _lib.iptr(sc.mat_idx), C.byref(s)))
for _ in range(4):
    _lib.check(L.daisy_solver_step_local(s))
    _lib.check(L.daisy_solver_step_finish(s, None))
ms = C.c_double()
sums = np.zeros(K)
_lib.check(L.daisy_solver_band_sums(s, sums.ctypes.data_as(C.POINTER(C.c_double))))
_lib.check(L.daisy_solver_last_step_ms(s, C.byref(ms)))
print(name, "rank 0 of", nranks, "rows", r1 - r0, "last pass ms", ms.value, "algorithmic bytes", 4.0 * (r1 - r0) * N + 16.0 * N * K)
L.daisy_solver_destroy(s)
p.close()
