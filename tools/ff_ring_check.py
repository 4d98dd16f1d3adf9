"""GPU box: the form-factor matrix with coplanar skipping on must equal the matrix with it off, bit for bit."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import daisyriot_b200 as dz
from daisyriot_b200 import scenes

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
sc = scenes.cornell_box(N)
uv = scenes.msvc_sample_pattern(1)
out = {}
for ring in ("1", "0"):
    os.environ["DAISY_FF_RING"] = ring
    p = dz.OptixPrimeFunctionality(dz.MeshS.from_scene(sc), rands=uv)
    t0 = time.time()
    F = p.cudaCalculateRadiosityMatrix().rows()
    st = p.stats()
    print(f"ring={ring}: N={N} ff_ms={st['ff_ms']:.1f} pairs={st['pairs_traced']} heavy={st.get('pairs_heavy')} wall={time.time()-t0:.2f}s", flush=True)
    out[ring] = F.copy()
    if ring == "1":
        m1 = p.visibilityMasks(0, min(N, 512))
    else:
        m0 = p.visibilityMasks(0, min(N, 512))
    p.close()
same = np.array_equal(out["1"], out["0"])
print("F identical:", same, "masks identical:", np.array_equal(m1, m0))
if not same:
    d = np.argwhere(out["1"] != out["0"])
    print("differing entries:", len(d), d[:10].tolist())
    sys.exit(1)
