"""Regenerates the committed golden vectors with the CPU oracle (run in the build container).

    python tests/golden/make_golden.py small      # seconds: cornell-512 whole matrices + sampled fixture rows
    python tests/golden/make_golden.py full NAME  # ~10+ min: whole-matrix per-row digests of a fixture scene

Before writing anything the oracle is re-checked against the reference's own sources compiled in oracle/_ref
(triangle_math.cpp form factors, ray generation) on random pairs of the scene being frozen."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from daisyriot_b200 import scenes  # noqa: E402
from oracle import pyoracle, pyref  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
FIXTURE_ROWS = {"cornellbox_blacklight": [0, 1, 2, 777, 2560, 5119, 5120, 7679, 7711],
                "colorballs": [0, 1, 2, 777, 2560, 5119, 5120, 6367, 6399]}


def check_against_ref(sc, orc, uv, n=400):
    if not pyref.available():
        print("  (oracle/_ref not built: skipping the reference cross-check)")
        return
    rs = pyref.RefScene.from_arrays(sc.vertices, sc.normals, sc.tri)
    rng = np.random.RandomState(5)
    for _ in range(n):
        i, j = rng.randint(0, sc.numtriangles, 2)
        a, b = orc.p2p_ff(i, j, 1), rs.p2p_unoccluded(i, j)
        assert a == b or (np.isnan(a) and np.isnan(b)), (i, j, a, b)
        k = rng.randint(uv.shape[0])
        assert np.array_equal(orc.pair_ray(i, j, uv[k, 0], uv[k, 1]), rs.pair_ray(i, j, uv[k, 0], uv[k, 1]), equal_nan=True)
    rs.close()


def small():
    uv = scenes.msvc_sample_pattern(1)
    sc = scenes.cornell_box(512)
    orc = pyoracle.Oracle.from_scene(sc)
    check_against_ref(sc, orc, uv)
    F0, m0, rays0 = orc.radmat_rows(uv, 0, 512, variant=0, brute=True)
    F1, m1, _ = orc.radmat_rows(uv, 0, 512, variant=1, reciprocity=True, brute=True)
    U0 = orc.unoccluded_rows(0, 512, 0)
    # a fluorescent gather: K = 9, three passes, FP64 accumulation
    rng = np.random.RandomState(11)
    K = 9
    M = rng.uniform(0, 0.3, (len(sc.materials), K, K)).astype(np.float32)
    E = (rng.uniform(0, 7, (K, 512)) * (rng.uniform(0, 1, (K, 512)) < 0.1)).astype(np.float32)
    res, B = E.copy(), E.copy()
    sums = [pyoracle.gather_pass(F0, res, B, M, sc.mat_idx, accum=1) for _ in range(3)]
    out = dict(uv=uv, F_device=F0, masks_device=m0, rays_device=rays0, F_host_reciprocity=F1, unoccluded_device=U0,
               gather_M=M, gather_E=E, gather_B3=B, gather_res3=res, gather_sums=np.array(sums))
    hits_rays = np.concatenate([orc.pair_rays(i, j, uv) for i, j in [(0, 511), (3, 100), (17, 300), (40, 41)]])
    out["hit_rays"] = hits_rays
    out["hits"] = orc.query_closest(hits_rays, brute=True)
    np.savez_compressed(os.path.join(G, "cornell512_golden.npz"), **out)
    print("cornell512:", rays0, "rays,", int((F0 != 0).sum()), "nnz")
    for name, rows in FIXTURE_ROWS.items():
        sc = scenes.load_scene_npz(os.path.join(G, name + ".npz"))
        orc = pyoracle.Oracle.from_scene(sc)
        check_against_ref(sc, orc, uv)
        Fr, Mr = [], []
        for r in rows:
            F, m, _ = orc.radmat_rows(uv, r, r + 1)
            Fr.append(F[0]); Mr.append(m[0])
        np.savez_compressed(os.path.join(G, name + "_rows_golden.npz"), rows=np.array(rows), F=np.array(Fr), masks=np.array(Mr), uv=uv)
        print(name, "rows frozen")


def full(name):
    uv = scenes.msvc_sample_pattern(1)
    sc = scenes.load_scene_npz(os.path.join(G, name + ".npz"))
    N = sc.numtriangles
    orc = pyoracle.Oracle.from_scene(sc)
    check_against_ref(sc, orc, uv)
    w = (2 * np.arange(N, dtype=np.uint64) + np.uint64(1))
    mh = np.zeros(N, np.uint64); fs = np.zeros(N, np.float64); fx = np.zeros(N, np.uint32)
    pairs = 0
    t0 = time.time()
    CH = 64
    for r0 in range(0, N, CH):
        r1 = min(N, r0 + CH)
        F_rc, F_cr, m, rays = orc.radmat_upper(uv, r0, r1)
        pairs += rays // uv.shape[0]
        with np.errstate(over="ignore"):
            mh[r0:r1] += (m * w[None, :]).sum(axis=1, dtype=np.uint64)          # row r, columns c > r
            mh += (m * w[r0:r1, None]).sum(axis=0, dtype=np.uint64)             # row c, column r
        fs[r0:r1] += F_rc.astype(np.float64).sum(axis=1)
        fs += F_cr.astype(np.float64).sum(axis=0)
        fx[r0:r1] ^= np.bitwise_xor.reduce(F_rc.view(np.uint32), axis=1)
        fx ^= np.bitwise_xor.reduce(F_cr.view(np.uint32), axis=0)
        if (r0 // CH) % 10 == 0:
            print(f"  {name}: row {r0}/{N} {time.time()-t0:.0f}s", flush=True)
    np.savez_compressed(os.path.join(G, name + "_full_golden.npz"), mask_hash=mh, F_sum=fs, F_xor=fx, pairs=np.int64(pairs), uv=uv)
    print(name, "pairs", pairs, "time", time.time() - t0)


if __name__ == "__main__":
    if sys.argv[1] == "small":
        small()
    else:
        full(sys.argv[2])
