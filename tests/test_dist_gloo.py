"""World-size-2 test of the multi-GPU exchange protocol on CPU (gloo): row partition, exchange-buffer layout,
in-place all-gather, rank-ordered band-sum totals and identical stop decisions on every rank.  The local gather
kernel is played by the CPU oracle on this rank's rows; everything else is the product's host logic (dist.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    from daisyriot_b200 import dist, scenes
    from oracle import pyoracle
    g = np.load(os.path.join(GOLDEN, "cornell512_golden.npz"))
    F, M, E = g["F_device"], g["gather_M"], g["gather_E"]
    N, K = F.shape[0], E.shape[0]
    N_use = 510  # not a multiple of the block size: the last rank's block is short
    F, E = np.ascontiguousarray(F[:N_use, :N_use]), np.ascontiguousarray(E[:, :N_use])
    mat = scenes.cornell_box(512).mat_idx[:N_use]
    Kp = dist.padded_K(K)
    r0, r1, n = dist.partition(N_use, rank, world)
    bstride, off = dist.block_layout(Kp, n)
    bufs = [torch.zeros(world * bstride, dtype=torch.float32) for _ in range(2)]

    def to_exchange(buf, X):
        b = buf.numpy()
        b[:] = 0
        for k in range(K):
            for gk in range(world):
                a0, a1, _ = dist.partition(N_use, gk, world)
                b[gk * bstride + k * n: gk * bstride + k * n + (a1 - a0)] = X[k, a0:a1]

    def from_exchange(buf):
        b = buf.numpy()
        X = np.zeros((K, N_use), np.float32)
        for k in range(K):
            for gk in range(world):
                a0, a1, _ = dist.partition(N_use, gk, world)
                X[k, a0:a1] = b[gk * bstride + k * n: gk * bstride + k * n + (a1 - a0)]
        return X

    to_exchange(bufs[0], E)
    cur = 0
    B_loc = E[:, r0:r1].copy()
    sums = E.astype(np.float64).sum(1)
    passes = 0
    thr = 0.05 * sums.sum()
    while sums.sum() > thr and passes < 50:
        nxt = bufs[cur ^ 1]

        def step_local():
            res_full = from_exchange(bufs[cur])
            # this rank's rows only: bounced = F[r0:r1] @ res, then the per-material mix (FP64 oracle arithmetic)
            bounced = (F[r0:r1].astype(np.float64) @ res_full.T.astype(np.float64)).astype(np.float32)  # (nloc, K)
            new = np.zeros((K, r1 - r0), np.float32)
            for p in range(r1 - r0):
                Mp = M[mat[r0 + p]].astype(np.float64)  # [col][row]
                new[:, p] = (Mp.T @ bounced[p].astype(np.float64)).astype(np.float32)
            nb = nxt.numpy()
            blk = nb[rank * bstride:(rank + 1) * bstride]
            blk[:] = 0
            for k in range(K):
                blk[k * n: k * n + (r1 - r0)] = new[k]
            blk[off: off + 2 * Kp].view(np.float64)[:K] = new.astype(np.float64).sum(1)
            B_loc[:] = B_loc + new

        dist.exchange_pass(step_local, nxt, rank, world)
        cur ^= 1
        sums = dist.total_band_sums(bufs[cur], K, Kp, n, world)
        passes += 1
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), passes=passes, res=from_exchange(bufs[cur]), B=B_loc, r0=r0, r1=r1, sums=sums)
    tdist.destroy_process_group()


def test_two_rank_exchange_matches_single_process(tmp_path):
    world, port = 2, 29512 + os.getpid() % 200
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from daisyriot_b200 import scenes
    from oracle import pyoracle
    g = np.load(os.path.join(GOLDEN, "cornell512_golden.npz"))
    N_use = 510
    F = np.ascontiguousarray(g["F_device"][:N_use, :N_use])
    E = np.ascontiguousarray(g["gather_E"][:, :N_use])
    M = g["gather_M"]
    mat = np.ascontiguousarray(scenes.cornell_box(512).mat_idx[:N_use])
    res, B = E.copy(), E.copy()
    sums = E.astype(np.float64).sum(1)
    thr = 0.05 * sums.sum()
    passes = 0
    while sums.sum() > thr and passes < 50:
        sums = pyoracle.gather_pass(F, res, B, M, mat, accum=1)
        passes += 1
    outs = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert passes > 1 and all(int(o["passes"]) == passes for o in outs)
    assert np.array_equal(outs[0]["res"], outs[1]["res"])                      # every rank holds the same residual
    assert np.allclose(outs[0]["res"], res, rtol=2e-6, atol=1e-7 * np.abs(res).max())
    assert np.allclose(outs[0]["sums"], outs[1]["sums"], rtol=0, atol=0)       # identical stop inputs on all ranks
    assert np.allclose(outs[0]["sums"], sums, rtol=1e-6)
    Bcat = np.concatenate([o["B"] for o in outs], axis=1)
    assert [int(o["r0"]) for o in outs] == [0, 256] and int(outs[1]["r1"]) == N_use  # 255 -> 256 rows per block
    assert np.allclose(Bcat, B, rtol=2e-6, atol=1e-7 * np.abs(B).max())
