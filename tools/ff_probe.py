"""GPU box: time the fused form-factor build at the given patch counts and print an exact digest of the whole matrix
(xor / sum of the per-row device digests), so that A/B builds (DAISY_B200_LIB, DAISY_FF_ORDER, DAISY_FF_RING) can be
compared for speed AND for bit-identical results in one go.  Not the benchmark."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import daisyriot_b200 as dz
from daisyriot_b200 import scenes

sizes = [int(a) for a in sys.argv[1:]] or [32768]
uv = scenes.msvc_sample_pattern(1)
for N in sizes:
    sc = scenes.cornell_box(N)
    p = dz.OptixPrimeFunctionality(dz.MeshS.from_scene(sc), rands=uv)
    t = time.time()
    rm = p.cudaCalculateRadiosityMatrix()
    wall = time.time() - t
    st = p.stats()
    x, w = rm.row_digest()
    with np.errstate(over="ignore"):
        print(f"N={N} lbvh {st['lbvh_ms']:.2f} ms  ff {st['ff_ms']:.1f} ms (wall {wall:.2f} s)  pairs {st['pairs_traced']}  fallback {st['pairs_fallback']}  "
              f"rays/s {st['rays'] / (st['ff_ms'] * 1e-3):.3e}  digest {int(np.bitwise_xor.reduce(x)):08x}/{int(w.sum(dtype=np.uint64)):016x}", flush=True)
    p.close()
