#!/usr/bin/env python
"""bench.py -- radiosity gather iterations/s (+ form-factor visibility rays/s) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path on the host cores

A step is ONE gather pass (residual <- M (F residual); B += residual, reference Lightning.h:196-226) over the
resident dense FP32 form-factor matrix of the workload.  Before the timed passes the matrix is built by the fused
form-factor/visibility kernel; that build is timed too and reported under "formfactor" (rays/s, BASELINE.json's other
metric).  Inputs are synthetic (subdivided Cornell box, BASELINE.json configs 3-5).  F is >= 4 GB, far larger than the
126 MB L2, so successive passes cannot be served from cache.

One JSON line is printed by rank 0.  Keys: see the contract in the task statement; extra: "formfactor",
"roofline" (gather kernel, HBM), "cpu_baseline".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (patches, bands, n_fluorescent)   -- BASELINE.json configs
    "cornell_32k": (32768, 9, 2),        # config 3: 32K patches, dense F ~4 GB, 1 B200
    "cornell_128k": (131072, 9, 2),      # config 4 / north-star target: 128K patches, F 68.7 GB, row-sharded
    "fluor_64k_k32": (65536, 32, 10),    # config 5: 64K patches, 32 bands, >= 8 fluorescent materials
    "cornell_8k": (8192, 9, 2),          # small smoke size
    # the reference's own example scenes (BASELINE.json configs 1-2), read from the committed fixtures in tests/golden/
    "cornellbox_blacklight": (7712, 9, 0),
    "colorballs": (6400, 9, 0),
}
FIXTURE_SCENES = ("cornellbox_blacklight", "colorballs")


def wavelengths_for(K):
    if K == 9:
        return np.arange(200, 601, 50).astype(np.float32)  # reference main.cpp:94
    return (200.0 + (400.0 / (K - 1)) * np.arange(K)).astype(np.float32) if K > 1 else np.array([350.0], np.float32)


def make_workload(name):
    from daisyriot_b200 import materials, rgb2spec, scenes
    N, K, nfl = WORKLOADS[name]
    if name in FIXTURE_SCENES:
        sc = scenes.load_scene_npz(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        assert sc.numtriangles == N
    else:
        sc = scenes.cornell_box(N, n_fluorescent=nfl)
    wl = wavelengths_for(K)
    tmp = tempfile.mkdtemp(prefix="daisy_bench_")
    os.makedirs(os.path.join(tmp, "color_tables"))
    coeff = os.path.join(tmp, "color_tables", "srgb.coeff")
    rgb2spec.write_surrogate_table(coeff, 16)
    mats = materials.make_materials(sc.materials, wl, rgb2spec.RGB2Spec.load(coeff))
    if K == 32:
        # config 5: full 32x32 re-emission matrices -- the Material.cpp:90-100 rule plus a dense perturbation
        rng = np.random.RandomState(0x5EED)
        for m in mats:
            if m.kind == "fluorescent":
                m.M = (m.M * 0.6 + rng.uniform(0, 0.4 / K, (K, K))).astype(np.float32)
    E, M = materials.spectral_inputs(mats, sc.mat_idx, 7.0)  # emission_value 7.0, config_example.ini:18
    return sc, wl, E, M, tmp


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_run(sc, wl, K, budget_s, steps, coeff_dir, log=lambda *a: None):
    """The reference's CPU path on the host cores, on a bounded row sample of the workload.

    form factors + visibility: the oracle's restatement of calculateAllVisibility (OptiX Prime itself is closed source),
    all host threads; gather: the reference's UNMODIFIED Lightning.h + vendored Eigen 3.2.10 compiled into
    oracle/_ref (single-threaded, as in the reference), fed the sampled rows; full-pass time = SpMV time scaled by
    N/rows + the measured per-patch loop."""
    from daisyriot_b200 import scenes
    from oracle import pyoracle, pyref
    N = sc.numtriangles
    uv = scenes.msvc_sample_pattern(1)
    orc = pyoracle.Oracle.from_scene(sc)
    cores = pyoracle.num_procs()
    orc.bvh
    # --- FF + visibility rays/s on sampled row batches spread over the scene, sized to the time budget;
    # the oracle's row loop runs one row per host thread
    rows_done, rays_done, t_ff, Fs = [], 0, 0.0, []
    batch = max(1, cores)
    starts = list(range(0, max(1, N - batch + 1), max(batch, N // 64)))
    np.random.RandomState(7).shuffle(starts)
    for s0 in starts:
        if t_ff >= budget_s or len(rows_done) >= 1024:
            break
        s1 = min(N, s0 + batch)
        t0 = time.time()
        F_b, _, rays = orc.radmat_rows(uv, s0, s1, want_masks=False)
        t_ff += time.time() - t0
        rays_done += rays
        rows_done += list(range(s0, s1))
        Fs += [F_b[i] for i in range(s1 - s0)]
    rays_per_s = rays_done / t_ff if t_ff > 0 else 0.0
    log(f"cpu ff: {len(rows_done)} rows, {rays_done} rays in {t_ff:.1f}s on {cores} threads")
    out = {"ff_rays_per_s": rays_per_s, "ff_rows": len(rows_done), "ff_rays": rays_done, "ff_seconds": t_ff, "cores_ff": cores,
           "gather_kind": None, "it_per_s": None}
    # --- gather: reference Lightning.h + Eigen on the sampled rows
    R = len(rows_done)
    if pyref.available():
        obj, _ = scenes.write_obj(sc, coeff_dir, "bench_scene")
        rs = pyref.RefScene.load(obj, coeff_dir + "/", wl, coeff_dir)
        for i in range(rs.nmat):  # the lamp's M is undefined behaviour in the reference; use the documented restatement
            m = rs.material(i)
            if not np.isfinite(m["M"]).all() or (m["spectral_values"] == 0).all():
                rs.set_material_M(i, np.zeros((K, K), np.float32))
        def run(rows_idx, rows_val):
            r_i, c_i, v = [], [], []
            for r, row in zip(rows_idx, rows_val):
                nz = np.nonzero(row)[0]
                r_i.append(np.full(nz.size, r, np.int32)); c_i.append(nz.astype(np.int32)); v.append(row[nz].astype(np.float64))
            r_i = np.concatenate(r_i) if r_i else np.zeros(0, np.int32)
            c_i = np.concatenate(c_i) if c_i else np.zeros(0, np.int32)
            v = np.concatenate(v) if v else np.zeros(0, np.float64)
            rs.L.ref_set_triplets(rs.h, r_i.ctypes.data_as(C.POINTER(C.c_int)), c_i.ctypes.data_as(C.POINTER(C.c_int)),
                                  v.ctypes.data_as(C.POINTER(C.c_double)), v.size)
            rs.lightning_create(2 if K > 3 else (1 if K == 3 else 0), 7.0)
            rs.lightning_reset()
            ts = []
            for _ in range(max(1, steps)):
                t0 = time.time(); rs.lightning_pass_only(); ts.append(time.time() - t0)
            return float(np.median(ts)), v.size
        t_empty, _ = run([], [])
        t_samp, nnz = run(rows_done, Fs)
        t_full = max(t_samp - t_empty, 1e-9) * (N / max(R, 1)) + t_empty
        out.update({"gather_kind": "reference", "it_per_s": 1.0 / t_full, "gather_sample_nnz": int(nnz), "t_pass_sample_s": t_samp,
                    "t_pass_empty_s": t_empty})
        rs.close()
    else:
        out.update({"gather_kind": "port", "it_per_s": None})  # oracle/_ref was not built: no Eigen pass to time
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    sc, wl, E, M, tmp = make_workload(name)
    N, K, _ = WORKLOADS[name]
    t0 = time.time()
    r = cpu_reference_run(sc, wl, K, budget_s=args.cpu_budget * max(1, args.steps) / 5.0, steps=args.steps + args.warmup, coeff_dir=tmp,
                          log=lambda *a: print(*a, file=sys.stderr))
    val = r["it_per_s"]
    line = {"impl": "reference", "metric": "radiosity_gather_iterations_per_s", "value": val, "unit": "iterations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": (1e3 / val) if val else None, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "patches": N, "bands": K, "rays_per_pair": 50},
            "cpu_baseline": {"value": val, "unit": "iterations/s", "cores": 1, "kind": r["gather_kind"],
                             "sample": f"Lightning.h+Eigen pass on {r['ff_rows']} of {N} matrix rows ({r.get('gather_sample_nnz')} nnz), SpMV time scaled by N/rows"},
            "e2e": {"value": val, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "formfactor": {"metric": "formfactor_visibility_rays_per_s", "value": r["ff_rays_per_s"], "unit": "rays/s", "cores": r["cores_ff"],
                           "kind": "port", "sample": f"{r['ff_rows']} rows, {r['ff_rays']} rays in {r['ff_seconds']:.1f} s"},
            "wall_s": time.time() - t0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import daisyriot_b200 as dz
    from daisyriot_b200 import _lib, dist as ddist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = dz.lib()  # raises without the CUDA library: there is no CPU fallback

    name = args.workload
    sc, wl, E, M, tmp = make_workload(name)
    N, K, _ = WORKLOADS[name]
    uv = dz.msvc_sample_pattern(1)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
        return float(t.item())

    # ---- form-factor stage (timed once; it is seconds long, so self-warming): host mesh -> LBVH -> fused FF/visibility
    barrier()
    t0 = time.time()
    mesh = dz.MeshS.from_scene(sc)
    optixP = dz.OptixPrimeFunctionality(mesh, device=local, rands=uv, rank=rank, nranks=world)
    if world > 1:
        ddist.build_formfactors_sharded(optixP, peer_tiles=not args.no_peer_tiles)  # mirrored tiles go to the peer's F over NVLink
    else:
        optixP.cudaCalculateRadiosityMatrix()
    torch.cuda.synchronize()
    ff_wall = allmax(time.time() - t0)
    st = optixP.stats()
    ff_ms = allmax(st["ff_ms"])
    lbvh_ms = allmax(st["lbvh_ms"])
    r0, r1 = optixP.row_range
    # unique facing pairs of the whole matrix: every rank counts pairs whose lower index it owns
    pairs_unique = allsum(float(st["pairs_owned"]))
    pairs_traced = allsum(float(st["pairs_traced"]))
    rays = pairs_unique * uv.shape[0]

    solver = ddist.PartitionedSolver(optixP, K, E, M, sc.mat_idx, fused=not args.no_fused)
    nloc = r1 - r0

    def one_step():
        solver.step(False)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # nvidia-smi needs a moment to start: sample from the warm-up through the kernel-timing loop
    for _ in range(max(3, args.warmup)):
        one_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    barrier()
    ms_total = allmax(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = 1e3 / ms_step

    # ---- kernel-only duration of the gather pass for the roofline (library's CUDA events around the pass's kernels)
    kms = []
    for _ in range(min(args.steps, 50)):
        solver.step(True)
        m = C.c_double()
        L.daisy_solver_last_step_ms(solver._s, C.byref(m))
        kms.append(m.value)
    k_ms = allmax(float(np.mean(kms)))
    # keep the same load up for ~0.6 s so that nvidia-smi (50 ms period) sees it; the pass count is derived from the
    # all-reduced step time, i.e. identical on every rank (the fused exchange needs all ranks to step in lockstep)
    for _ in range(int(min(20000, max(0, 600.0 / max(ms_step, 1e-3) - args.steps)))):
        solver.step(False)
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    alg_bytes = 4.0 * nloc * N + 16.0 * N * K  # F rows streamed once + residual in/out + B read/write
    alg_bytes = allmax(alg_bytes)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    # measured DRAM bytes per launch of the gather kernel (one ncu --set full capture per workload, profiles/gather_traffic.json)
    ff_ncu = None
    try:
        ff_ncu = json.load(open(os.path.join(ROOT, "profiles", "ff_ncu.json")))
    except Exception:
        pass
    traffic = None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "gather_traffic.json"))).get(name)
        if t and world == 1:
            traffic = t["dram_bytes_per_launch"]
    except Exception:
        pass

    # ---- end to end through the C-ABI with HOST buffers (single GPU): H2D residual+B, pass, D2H residual+B
    e2e = None
    if world == 1:
        hB = torch.empty((K, N), dtype=torch.float32, pin_memory=True).numpy()
        hR = torch.empty((K, N), dtype=torch.float32, pin_memory=True).numpy()
        hB[:] = E; hR[:] = E
        for _ in range(2):
            _lib.check(L.daisy_solver_write(solver._s, _lib.fptr(hB), _lib.fptr(hR)))
            _lib.check(L.daisy_solver_step(solver._s, None))
            _lib.check(L.daisy_solver_read(solver._s, _lib.fptr(hB), _lib.fptr(hR)))
        hB[:] = E; hR[:] = E
        torch.cuda.synchronize()
        nst = min(args.steps, 50)
        t0 = time.time()
        for _ in range(nst):
            _lib.check(L.daisy_solver_write(solver._s, _lib.fptr(hB), _lib.fptr(hR)))
            _lib.check(L.daisy_solver_step(solver._s, None))
            _lib.check(L.daisy_solver_read(solver._s, _lib.fptr(hB), _lib.fptr(hR)))
        torch.cuda.synchronize()
        e2e_s = (time.time() - t0) / nst
        e2e = {"value": 1.0 / e2e_s, "unit": "iterations/s", "h2d_bytes_per_step": int(2 * K * N * 4), "d2h_bytes_per_step": int(2 * K * N * 4)}
    else:
        # multi-GPU: every rank uploads the whole residual vector and its slice of B from pinned host memory, runs one pass
        # through the public API (exchange included) and reads back its slices plus the band sums
        hBl = torch.empty((K, nloc), dtype=torch.float32, pin_memory=True).numpy()
        hRl = torch.empty((K, nloc), dtype=torch.float32, pin_memory=True).numpy()
        hR = torch.empty((K, N), dtype=torch.float32, pin_memory=True).numpy()
        hBl[:] = E[:, r0:r1]; hR[:] = E
        def e2e_step():
            _lib.check(L.daisy_solver_write_partitioned(solver._s, _lib.fptr(hBl), _lib.fptr(hR)))
            solver.step(True)
            _lib.check(L.daisy_solver_read(solver._s, _lib.fptr(hBl), _lib.fptr(hRl)))
        for _ in range(2):
            e2e_step()
        hBl[:] = E[:, r0:r1]; hR[:] = E
        barrier()
        t0 = time.time()
        nst = min(args.steps, 50)
        for _ in range(nst):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = allmax((time.time() - t0) / nst)
        e2e = {"value": 1.0 / e2e_s, "unit": "iterations/s", "h2d_bytes_per_step": int(world * (K * N + K * nloc) * 4),
               "d2h_bytes_per_step": int(world * (2 * K * nloc * 4 + 8 * K * world))}

    # ---- whole solve with the reference's stop rule (SpectralLightning::converge_lightning, Lightning.h:145-151:
    # iterate while the residual summed over bands and patches exceeds 200), band sums read back every pass
    barrier()
    solver.reset()
    t0 = time.time()
    conv_passes = solver.converge(200.0, per_band=False, max_passes=2000)
    torch.cuda.synchronize()
    conv_s = allmax(time.time() - t0)

    cpu_baseline = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        try:
            r = cpu_reference_run(sc, wl, K, budget_s=args.cpu_budget, steps=3, coeff_dir=tmp)
            cpu_baseline = {"value": r["it_per_s"], "unit": "iterations/s", "cores": 1, "kind": r["gather_kind"],
                            "sample": f"Lightning.h+Eigen pass on {r['ff_rows']} of {N} matrix rows ({r.get('gather_sample_nnz')} nnz), SpMV time scaled by N/rows",
                            "formfactor": {"value": r["ff_rays_per_s"], "unit": "rays/s", "cores": r["cores_ff"], "kind": "port",
                                           "sample": f"{r['ff_rows']} rows, {r['ff_rays']} rays in {r['ff_seconds']:.1f} s"}}
        except Exception as ex:  # the baseline is a reported extra; never let it kill the measurement
            cpu_baseline = {"value": None, "unit": "iterations/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}

    if rank == 0:
        line = {
            "metric": "radiosity_gather_iterations_per_s", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "reference example scene (tests/golden fixture)" if name in FIXTURE_SCENES else "synthetic",
            "config": {"workload": name, "patches": N, "bands": K, "rays_per_pair": int(uv.shape[0]), "parallelism": f"rowshard{world}",
                       "cache": "F rows per GPU %.2f GB > 126 MB L2 (inputs larger than L2, no flush needed)" % (4.0 * nloc * N / 1e9)},
            "e2e": e2e, "gpu_launches": int((2 + (1 if K > 9 else 0) + (1 if world > 1 and not args.no_fused else 0)) * args.steps),
            "clocks": clocks,
            "roofline": {"kernel": ("k_gather_mma" if K > 9 else "k_gather_tma") + "+k_gather_epilogue", "bound": "hbm", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_note": "hbm_gbs is a copy (read+write) bandwidth; this kernel only reads, so frac can exceed 1", "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback",
                         "kernel_ms": k_ms, "algorithmic_bytes": alg_bytes},
            "cpu_baseline": cpu_baseline,
            "formfactor": {"metric": "formfactor_visibility_rays_per_s", "value": rays / (ff_ms * 1e-3), "unit": "rays/s",
                           "pairs_facing": int(pairs_unique), "pairs_traced_all_ranks": int(pairs_traced), "rays": int(rays), "kernel_ms": ff_ms,
                           "lbvh_build_ms": lbvh_ms, "e2e_wall_s": ff_wall,
                           "e2e_rays_per_s": rays / ff_wall,
                           "ncu": ff_ncu},  # pipe / cache utilisation of the traversal kernel from the committed ncu capture
            "converge": {"rule": "sum of residual over bands and patches <= 200 (Lightning.h:145-151)", "passes": int(conv_passes),
                         "seconds": conv_s, "total_with_formfactors_s": conv_s + ff_wall},
        }
        print(json.dumps(line))
    solver.close()
    optixP.close()
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("DAISY_WORKLOAD", "cornell_128k"), choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU form-factor sampling for the baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fused", action="store_true", help="multi-GPU: NCCL all-gather per pass instead of the epilogue kernel's peer stores")
    ap.add_argument("--no-peer-tiles", action="store_true", help="multi-GPU: trace every tile touching this rank's rows instead of exchanging mirrored tiles")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
