"""GPU box: whole-matrix digests of the form-factor build with coplanar skipping on and off at a size whose matrix does not
fit the host twice (rows are read back in chunks and reduced to an xor of the float bit patterns + a float64 sum per chunk)."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import daisyriot_b200 as dz
from daisyriot_b200 import scenes

N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
sc = scenes.cornell_box(N)
uv = scenes.msvc_sample_pattern(1)
dig = {}
for ring in ("1", "0"):
    os.environ["DAISY_FF_RING"] = ring
    p = dz.OptixPrimeFunctionality(dz.MeshS.from_scene(sc), rands=uv)
    t0 = time.time()
    rm = p.cudaCalculateRadiosityMatrix()
    print(f"ring={ring}: N={N} ff_ms={p.stats()['ff_ms']:.0f}", flush=True)
    xs, ss = [], []
    step = max(1, (512 << 20) // (4 * N))
    for r0 in range(0, N, step):
        rows = rm.rows(r0, min(step, N - r0))
        xs.append(np.bitwise_xor.reduce(rows.view(np.uint32), axis=1))
        ss.append(rows.sum(axis=1, dtype=np.float64))
    dig[ring] = (np.concatenate(xs), np.concatenate(ss))
    print(f"   digests in {time.time()-t0:.0f}s, nnz-ish xor of xors {np.bitwise_xor.reduce(dig[ring][0]):08x}, total {dig[ring][1].sum():.9e}", flush=True)
    p.close()
same = np.array_equal(dig["1"][0], dig["0"][0]) and np.array_equal(dig["1"][1], dig["0"][1])
print("row digests identical:", same)
sys.exit(0 if same else 1)
