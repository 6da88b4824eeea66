#!/usr/bin/env python
"""Headline benchmark: order-50 spharm decompositions/sec of 128^3 volumes (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = the hot path (`z3d_decompose`) over one batch of 128 synthetic 128^3 volumes per GPU: two tiles of 64
volumes, each k_fold_tile_gen -> k_decompose_gen (tcgen05 3xTF32 contraction, operand T = R_nl * Q_lm generated inside the
kernel - no basis table exists) -> k_reduce_stream.  The table-fed kernels of round 1 (opt-in) are timed beside it.  N > 1 is launched by torchrun, one rank per GPU; the path shards by
independent volumes (weak scaling, no data-path collective); the voxel-sharded single-volume latency with its
NCCL all-reduce is reported alongside as `sharded_single_volume`.  Rank 0 prints ONE JSON line.

`--impl reference` times the CPU arm instead: the reference is pure Python that cannot run on the GPU box
(/root/reference is absent there and its basis needs 28 GB + minutes per 128^3 volume), so this arm runs the
oracle port of its mode-'3' algorithm (oracle/zernike_oracle.c) with all host threads on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIM, ORDER, BATCH, TILE = 128, 50, 128, 64
METRIC = "order-50 decompositions/sec (128^3 vol)"
UNIT = "decompositions/s"
SM_COUNT, FP32_LANES = 148, 128


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); smax = float(r[2]); power.append(float(r[3]))
            except ValueError:
                continue
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def algorithmic_flops(order_moments: int, dim: int) -> float:
    """W = 4 * N_mom * N^3 per volume: one real x complex MAC per moment per voxel."""
    return 4.0 * order_moments * dim ** 3


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_sample(threads=None, planes=8):
    """Oracle port of the reference's mode-'3' analysis on `planes` of the 128 planes of one volume."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    from zernike3d_b200.synthetic import volume

    oracle.build()
    if threads:
        oracle.set_threads(threads)
    cores = oracle.max_threads()
    v = volume(1, DIM)
    lo = DIM // 2 - planes // 2
    oracle.decompose(v, ORDER, oracle.FAITHFUL, rows=(lo, lo + 1))  # warm-up (page-in, thread start)
    t0 = time.perf_counter()
    oracle.decompose(v, ORDER, oracle.FAITHFUL, rows=(lo, lo + planes))
    dt = time.perf_counter() - t0
    per_volume = dt * DIM / planes
    return {"value": 1.0 / per_volume, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle/zernike_oracle.c (faithful restatement of zm mode '3', fused + streamed, pthreads) on "
                      f"{planes} of {DIM} planes of one synthetic 128^3 volume at order 50, {dt:.1f} s, extrapolated x{DIM // planes}; "
                      "the reference's own numpy run of this volume took 216 s basis + 95 s decomposition on 8 cores "
                      "(tests/golden/ref_c3_d128_o50.npz `timing`)"}, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    samples = []
    base = None
    for i in range(args.warmup + args.steps):
        base, dt = cpu_sample(planes=4)
        if i >= args.warmup:
            samples.append(dt)
    ms = 1e3 * float(np.mean(samples))
    value = 1.0 / (float(np.mean(samples)) * DIM / 4)
    base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "128^3 order-50 spharm decomposition; each step = 4 of 128 planes of one volume "
                                   "(bounded sample), value extrapolated to whole volumes"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------- GPU arm
def _time_ms(fn, reps, warm=2):
    """CUDA-event time of `fn` on torch's current stream: warm-ups, synchronize on both sides, mean over reps."""
    import torch

    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def other_configs(rank, world, local, maxreduce, oracle):
    """The BASELINE.json configurations that are not the headline batch, on this rank's GPU, each with a parity sample against the
    float64 oracle (rank 0; bounded CPU work).  C5a is the 1024-volume ensemble of 96^3 at order 60 split over the ranks."""
    import torch

    from zernike3d_b200.plan import Plan
    from zernike3d_b200.synthetic import bundle_3dva, volume

    out = {}
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    # C1 / C2: 64^3 order 50, one volume, decomposition and reconstruction (round trip), parity of both against the oracle
    P = Plan(50, 64, local)
    v = volume(7, 64)
    dv = dev(v)
    mom = P.decompose(dv)
    c1 = _time_ms(lambda: P.decompose(dv), 20)
    c2 = _time_ms(lambda: P.reconstruct(mom), 20)
    entry = {"c1_decompose_64_o50_ms": c1, "c2_reconstruct_64_o50_ms": c2}
    if rank == 0:
        ex = oracle.decompose(v, 50, oracle.EXACT)
        entry["c1_rel_l2_vs_f64_oracle"] = _rel(mom.cpu().numpy(), ex)
        rec = P.reconstruct(mom).cpu().numpy()
        want = oracle.reconstruct(mom.cpu().numpy(), 64, 50, oracle.EXACT, rows=(20, 24))
        entry["c2_rel_l2_vs_f64_oracle_planes_20_23"] = _rel(rec[20:24], want)
    out["c1_c2"] = entry
    del P
    # C4: 3DVA bundle, base map + 5 components at 128^3 order 50 in ONE call
    P = Plan(50, 128, local)
    _, cons, comps = bundle_3dva(128, components=5, particles=8, seed=3)
    bundle = dev(np.stack([cons, *comps]))
    mb = P.decompose(bundle)
    entry = {"bundle_6x128_o50_ms": _time_ms(lambda: P.decompose(bundle), 10), "volumes": 6}
    if rank == 0:
        entry["rel_l2_vs_f64_oracle_component_5"] = _rel(mb[5].cpu().numpy(), oracle.decompose(comps[4], 50, oracle.EXACT))
    out["c4_3dva_bundle"] = entry
    del P, bundle
    # C5a: 1024 x 96^3 at order 60 over the ranks (independent volumes, no collective)
    total = 1024
    share = total // world
    P = Plan(60, 96, local)
    base = np.stack([volume(5000 + 100 * rank + s, 96) for s in range(8)])
    vols = dev(np.concatenate([base] * (share // 8)))
    o = torch.empty((share, P.num_moments), dtype=torch.complex64, device="cuda")
    ms = maxreduce(_time_ms(lambda: P.decompose(vols, out=o), 2, warm=1))
    info = P.info()
    entry = {"volumes_total": total, "volumes_per_gpu": share, "ms": ms, "decompositions_per_s": total / (ms * 1e-3),
             "scaling": "strong", "kernel": "k_decompose_gen", "column_sets": info["gen_sets"], "chunk_splits": info["gen_splits"]}
    if rank == 0:
        entry["rel_l2_vs_f64_oracle_volume_3"] = _rel(o[3].cpu().numpy(), oracle.decompose(base[3], 60, oracle.EXACT))
    out["c5a_ensemble_1024x96_o60"] = entry
    del P, vols, o
    # C5b: 256^3 order 50 single-volume stress case: latency, and two plane pairs against oracle rows
    if rank == 0:
        P = Plan(50, 256, local)
        v = volume(8, 256)
        dv = dev(v)
        entry = {"decompose_256_o50_ms": _time_ms(lambda: P.decompose(dv), 5)}
        errs = []
        for lo in (3, 127):
            got = P.decompose(dv, pairs=(lo, lo + 1)).cpu().numpy()
            want = (oracle.decompose(v, 50, oracle.EXACT, rows=(lo, lo + 1), return_f64=True)
                    + oracle.decompose(v, 50, oracle.EXACT, rows=(256 - lo - 1, 256 - lo), return_f64=True))
            errs.append(_rel(got, want))
        entry["rel_l2_vs_f64_oracle_plane_pairs_3_127"] = errs
        out["c5b_256_stress"] = entry
        del P, dv
    torch.cuda.empty_cache()
    return out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from zernike3d_b200 import dist as zdist
    from zernike3d_b200.plan import Plan, fma_peak_tflops, launch_count, tf32_peak_tflops
    from zernike3d_b200.synthetic import volume

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - zernike3d_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa_node = zdist.bind_to_gpu_numa(local)   # before any pinned allocation (no-op where the platform hides the topology)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxreduce(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    oracle = None
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle  # the checker of the parity samples below and the cpu_baseline leg; never on the timed path

        oracle.build()

    plan = Plan(ORDER, DIM, local)
    plan.profile(True)
    nmom = plan.num_moments
    info = plan.info()
    if not (info["gen"] == 1 and BATCH >= info["gen_min"] and info["tables_level"] == 0):
        raise SystemExit("bench.py: the table-free tensor-core path is unavailable for this plan")
    # distinct synthetic volumes per rank: 8 generated, tiled to the batch (contents do not change the work)
    base = np.stack([volume(1000 * rank + s, DIM) for s in range(8)])
    host = torch.from_numpy(np.concatenate([base] * (BATCH // 8))).pin_memory()
    dvol = host.cuda()
    out = torch.empty((BATCH, nmom), dtype=torch.complex64, device="cuda")

    warm = max(3, args.warmup)
    for _ in range(warm):
        plan.decompose(dvol, out=out)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        plan.decompose(dvol, out=out)
    e1.record()
    barrier()
    elapsed_ms = maxreduce(e0.elapsed_time(e1))
    launches = launch_count() - l0
    value = world * BATCH * args.steps / (elapsed_ms * 1e-3)
    # per-launch time of the dominant kernel, live (CUDA events on the launching stream around every k_decompose_gen launch)
    kernel_ms = []
    for _ in range(min(5, args.steps)):
        plan.decompose(dvol, out=out)
        kernel_ms.append(plan.last_kernel_ms())
    gen_moments = out[:8].cpu().numpy()

    # ---- end to end through the host-buffer C ABI call: pinned host volumes in, host moments out, per step
    hmom = torch.empty((BATCH, nmom), dtype=torch.complex64).pin_memory()
    for _ in range(2):
        plan.decompose_host(host, out=hmom)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        plan.decompose_host(host, out=hmom)
    torch.cuda.synchronize()
    e2e_rank_s = time.perf_counter() - t0
    e2e_s = maxreduce(e2e_rank_s)
    e2e_value = world * BATCH * args.steps / e2e_s
    # the same through the streaming form of the call (z3d_decompose_host_submit / z3d_host_wait): every step still copies its inputs in
    # and its results out inside the timed region, but step i + 1's copies overlap step i's last tile - reported beside e2e, not as e2e
    hmom2 = [hmom, torch.empty_like(hmom).pin_memory()]
    plan.decompose_host_submit(host, hmom2[0]); plan.host_wait()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        plan.decompose_host_submit(host, hmom2[i & 1])
    plan.host_wait()
    e2e_pipe_s = maxreduce(time.perf_counter() - t0)
    e2e_pipe_value = world * BATCH * args.steps / e2e_pipe_s
    pipe_ok = bool(torch.equal(torch.view_as_real(hmom2[0]), torch.view_as_real(hmom2[1])))
    h2d_step = BATCH * DIM ** 3 * 4
    rank_gbs = torch.tensor([h2d_step * args.steps / e2e_rank_s / 1e9], device="cuda", dtype=torch.float64)
    if world > 1:
        allg = [torch.empty_like(rank_gbs) for _ in range(world)]
        dist.all_gather(allg, rank_gbs)
        rank_gbs = torch.cat(allg)
    clocks = sampler.stop() if rank == 0 else None

    # ---- the opt-in table-fed path of round 1 beside it (z3d_plan_prepare_tables level 2: 13.7 GB T = R Q table streamed by TMA)
    table = None
    ptab = Plan(ORDER, DIM, local)
    ptab.profile(True)
    if ptab.prepare_tables(2) == 2:
        tms = _time_ms(lambda: ptab.decompose(dvol, out=out), max(2, min(5, args.steps)))
        tms = maxreduce(tms)
        ptab.decompose(dvol, out=out)
        ti = ptab.info()
        table = {"value": world * BATCH / (tms * 1e-3), "unit": UNIT, "ms_per_step": tms, "kernel": "z3d::k_decompose_stream",
                 "kernel_ms": ptab.last_kernel_ms() / (BATCH // TILE), "t_table_mib": ti["stream_table_mib"],
                 "rel_l2_vs_table_free_path": _rel(out[:8].cpu().numpy(), gen_moments),
                 "how": "Plan.prepare_tables(2): k_build_basis + k_build_T once, then k_fold_tile_kind + k_decompose_stream + k_reduce_stream "
                        "per tile; NOT the default - it materialises a folded basis table, which the north star rules out"}
    del ptab
    torch.cuda.empty_cache()

    # ---- one volume spread over the ranks (strong scaling of the single-volume latency): fused reduce + one-shot all-reduce over
    #      NVLink peer memory (own kernel), and compute + NCCL all-reduce; both checked against the float64 oracle
    fused_ms = fused_ok = fused_got = None
    onev = dvol[:1].contiguous()
    if world > 1:
        from zernike3d_b200.plan import PeerComm

        dist.broadcast(onev, src=0)   # the work-sharded path spreads ONE volume: every rank holds rank 0's volume 0 (the ranks' batches differ)

        comm = PeerComm(plan, batch_max=1)
        ref_m = zdist.decompose_work_sharded(plan.decompose, onev, deterministic=True)
        for _ in range(3):
            got = plan.decompose_allreduce(onev, comm)
        barrier()
        fused_ok = bool(torch.equal(torch.view_as_real(got), torch.view_as_real(ref_m))) and comm.error() == 0
        fused_got = got[0].cpu().numpy()
        barrier()   # the ranks enter the timed loop together: every call is a rendezvous, a late rank would be charged to the others
        fused_ms = maxreduce(_time_ms(lambda: plan.decompose_allreduce(onev, comm), 50, warm=3))
    sharded = zdist.decompose_work_sharded(plan.decompose, onev)
    barrier()
    sharded_ms = maxreduce(_time_ms(lambda: zdist.decompose_work_sharded(plan.decompose, onev), 50, warm=3))
    single_ms = _time_ms(lambda: plan.decompose(onev), 20)
    plan.decompose(onev)
    single_kernel_ms = plan.last_kernel_ms()

    # ---- reconstruction (same fused structure)
    moms = out[:8].contiguous()
    rec_ms = _time_ms(lambda: plan.reconstruct(moms), 3) / 8
    plan.reconstruct(moms[:1])
    rec_kernel_ms = plan.last_kernel_ms()

    # ---- batched reconstruction on the tensor cores (z3d_reconstruct_batched): one tile of 64 moment sets
    recb = None
    if info.get("synth"):
        m64 = out[:TILE].contiguous()
        rbuf = plan.reconstruct(m64, batched=True)
        recb_ms = maxreduce(_time_ms(lambda: plan.reconstruct(m64, batched=True, out=rbuf), 3, warm=1))
        import ctypes

        from zernike3d_b200 import _lib as zlib

        lay = (ctypes.c_int * 11)()
        zlib.lib().z3d_gen_layout_tiles(ORDER, info["sms"], 2, lay, 11)     # host-only planner query: columns over all sets of the layout
        # executed tensor work per tile: per K-step of 8 moment rows and PAIR of chunks one N = 32 and one N = 16 MMA (M 128, K 8)
        synth_flops = (lay[3] / 8.0) * (2.0 * 128 * 8 * (32 + 16)) * (info["chunks"] / 2.0)
        recb = {"ms_per_64_maps": recb_ms, "maps_per_s_per_gpu": TILE / (recb_ms * 1e-3), "kernel": "z3d::k_synth_gen + z3d::k_unfold_tile",
                "executed_tensor_flops_per_tile": synth_flops, "executed_tflops": synth_flops / (recb_ms * 1e-3) / 1e12,
                "partial_fold_bytes_per_tile": float(info["synth_partial_slots"]) * info["chunks"] * 2 * 128 * 8 * 4,
                "speedup_over_per_map_kernel": rec_ms * TILE / recb_ms,
                "rel_l2_vs_per_map_kernel_map_5": _rel(rbuf[5].cpu().numpy(), plan.reconstruct(m64[5:6], batched=False)[0].cpu().numpy())}
        if rank == 0:
            want = oracle.reconstruct(m64[9].cpu().numpy(), DIM, ORDER, oracle.EXACT, rows=(40, 42))
            recb["rel_l2_vs_f64_oracle_map_9_planes_40_41"] = _rel(rbuf[9, 40:42].cpu().numpy(), want)
        del rbuf
        torch.cuda.empty_cache()

    configs = other_configs(rank, world, local, maxreduce, oracle)

    if rank == 0:
        peaks = measured_peaks()
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        fp32_nominal = SM_COUNT * FP32_LANES * 2 * sm_max * 1e6 / 1e12     # 74.45 TFLOP/s at 1965 MHz
        fma_live = [fma_peak_tflops(local, v, 4096) for v in (0, 1)]
        tf32_live = tf32_peak_tflops(local, 4096)
        bf16_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 2250.0)))
        hbm_peak = float(peaks.get("hbm_gbs", 6553.0))
        # z3d_decompose runs the batch in tiles of 64 volumes: k_fold_tile_gen, k_decompose_gen, k_reduce_stream per tile;
        # last_kernel_ms() is the sum over the contraction launches of one call; the roofline unit is one launch = one tile
        n_launch = -(-BATCH // TILE)
        k_ms = float(np.mean(kernel_ms)) / n_launch
        chunks, cols = info["chunks"], info["gen_cols"]
        # executed tensor work of one launch: 3 TF32 MMAs (M = 128 = re/im x 64 volumes, K = 8) per accumulator column (moment rows
        # padded to 16 per (m, z parity) piece) and 8-voxel chunk of the 16-fold folded volume
        executed = 3.0 * 2 * 128 * 8 * cols * chunks
        W = algorithmic_flops(nmom, DIM) * TILE
        achieved = executed / (k_ms * 1e-3) / 1e12
        traffic = tsrc = None
        tpath = os.path.join(ROOT, "profiles", "traffic_bench.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic, tsrc = tj.get("dram_bytes_per_launch"), tj.get("source")
        alg_bytes = TILE * (DIM ** 3 * 4 + nmom * 8)
        folded = chunks * 8                                              # folded voxels (1/16 of the box, rows padded to 8)
        cpu, _ = cpu_sample(planes=16)
        exact = oracle.decompose(base[0], ORDER, oracle.EXACT)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"128^3 volumes, order 50 ({nmom} complex moments), batch of {BATCH} volumes per GPU per step "
                                   f"in tiles of {TILE}; inputs {BATCH * DIM ** 3 * 4 / 2 ** 20:.0f} MiB per GPU > 126 MB L2 (no flush needed)",
                       "dim": DIM, "order": ORDER, "batch_per_gpu": BATCH, "sharding": "independent volumes per GPU, no collective",
                       "tile_volumes": TILE, "column_sets": info["gen_sets"], "chunk_splits": info["gen_splits"],
                       "accumulator_columns": cols, "basis_tables": "none (table-free default path)",
                       "arithmetic": "3xTF32 tcgen05.mma (M 128 = re/im x 64 volumes, N = moment rows of one (m, z parity), K 8), float32 "
                                     "accumulate (TMEM); operand T = R_nl * Q_lm generated in registers by the reference's recurrences"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": BATCH * nmom * 8,
                    "per_rank_h2d_gbs": [round(float(x), 2) for x in rank_gbs.cpu().tolist()],
                    "rank0_numa_node": numa_node,
                    "streaming_calls_value": e2e_pipe_value, "streaming_calls_results_identical": pipe_ok,
                    "streaming_how": "Plan.decompose_host_submit x steps + host_wait (z3d_decompose_host_submit / z3d_host_wait): the next batch's "
                                     "H2D overlaps this batch's last tile and D2H; every step's copies are inside the timed region",
                    "how": "Plan.decompose_host -> z3d_decompose_host: pinned host volumes -> H2D -> kernels -> D2H moments, "
                           "64-volume groups double buffered"},
            "gpu_launches": int(launches),
            "parity": {"rel_l2_vs_f64_oracle_volume_0": _rel(gen_moments[0], exact),
                       "sharded_single_volume_rel_l2_vs_f64_oracle": _rel(sharded[0].cpu().numpy(), exact),
                       "fused_p2p_single_volume_rel_l2_vs_f64_oracle": _rel(fused_got, exact) if fused_got is not None else None},
            "roofline": {"bound": "tensor", "kernel": "z3d::k_decompose_gen", "achieved": achieved, "peak": tf32_live, "unit": "TFLOP/s",
                         "frac": achieved / tf32_live, "traffic": traffic, "traffic_source": tsrc,
                         "kernel_ms": k_ms, "launches_per_step": n_launch, "volumes_per_launch": TILE,
                         "executed_tensor_flops_per_launch": executed,
                         "peak_source": "live z3d_tf32_peak: back-to-back tcgen05.mma kind::tf32 M128 N256 K8 on all SMs (MEASURED_PEAKS.json "
                                        "carries only HBM and dense bf16 figures; half of its sustained bf16 figure would be "
                                        f"{bf16_peak / 2:.0f} TFLOP/s)",
                         "frac_of_half_bf16_sustained": achieved / (bf16_peak / 2),
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "traffic_over_algorithmic_bytes": (traffic / alg_bytes) if traffic else None,
                         "hbm_frac": (traffic / (k_ms * 1e-3) / 1e9 / hbm_peak) if traffic else None,
                         "algorithmic_flops_per_launch": W, "algorithmic_tflops": W / (k_ms * 1e-3) / 1e12,
                         "algorithmic_frac_of_fp32_peak": W / (k_ms * 1e-3) / 1e12 / fp32_nominal,
                         "note": "achieved = EXECUTED TF32 flops (3 MMAs x 2*128*8 per accumulator column and chunk) / k_decompose_gen time, "
                                 "against the measured TF32 tensor peak.  The algorithmic figure of SURVEY 8d (4*N_mom*N^3 flops per volume "
                                 "against the FP32 peak) is kept beside it; it exceeds 1 because mirror/transpose folding executes 1/16 of it"},
            "roofline_single_volume": {
                "bound": "fp32", "kernel": "z3d::k_decompose", "kernel_ms": single_kernel_ms, "call_ms": single_ms,
                "achieved": 2.0 * info["padded_accumulators"] * folded / (single_kernel_ms * 1e-3) / 1e12, "peak": fma_live[1],
                "unit": "TFLOP/s", "frac": 2.0 * info["padded_accumulators"] * folded / (single_kernel_ms * 1e-3) / 1e12 / fma_live[1],
                "peak_source": "live z3d_fma_peak (packed FFMA2)", "fma_microbench_tflops": {"ffma": fma_live[0], "ffma2": fma_live[1]},
                "executed_contraction_flops": 2.0 * info["padded_accumulators"] * folded,
                "algorithmic_flops": algorithmic_flops(nmom, DIM),
                "algorithmic_frac_of_fp32_peak": algorithmic_flops(nmom, DIM) / (single_kernel_ms * 1e-3) / 1e12 / fp32_nominal,
                "note": "one 128^3 volume (the north star's literal target: < 50 ms at >= 0.6 of the FP32 roofline, algorithmic); achieved = "
                        "the contraction's executed FMAs only (padded 8x8 register tiles x folded voxels); basis generation in registers "
                        "is the rest of the kernel's instructions (profiles/)"},
            "roofline_reconstruct": {
                "bound": "fp32", "kernel": "z3d::k_reconstruct", "kernel_ms": rec_kernel_ms, "ms_per_map_batched": rec_ms,
                "achieved": 2.0 * info["padded_accumulators"] * folded / (rec_kernel_ms * 1e-3) / 1e12, "peak": fma_live[1], "unit": "TFLOP/s",
                "frac": 2.0 * info["padded_accumulators"] * folded / (rec_kernel_ms * 1e-3) / 1e12 / fma_live[1],
                "algorithmic_flops": algorithmic_flops(nmom, DIM),
                "algorithmic_frac_of_fp32_peak": algorithmic_flops(nmom, DIM) / (rec_kernel_ms * 1e-3) / 1e12 / fp32_nominal,
                "maps_per_s_per_gpu": 1e3 / rec_ms},
            "cpu_baseline": cpu,
            "table_fed_path": table,
            "other_configs": configs,
            "sharded_single_volume": {"ms": sharded_ms, "decompositions_per_s": 1e3 / sharded_ms, "gpus": world, "scaling": "strong",
                                      "single_gpu_call_ms": single_ms,
                                      "how": "one 128^3 volume replicated, the kernel's work list split evenly over the ranks, "
                                             "one all-reduce (NCCL) of the partial moments; max over ranks",
                                      "fused_p2p_ms": fused_ms,
                                      "fused_p2p_matches_rank_ordered_sum_bitwise": fused_ok,
                                      "fused_how": "z3d_decompose_allreduce: k_decompose share + ONE kernel that reduces the "
                                                   "partials, pushes them block by block into every rank's peer-mapped buffer "
                                                   "(posted NVLink stores, one flag per block and source rank) and sums the "
                                                   "rows that arrived locally in rank order"},
            "reconstructions_per_s_per_gpu": 1e3 / rec_ms,
            "reconstruct_batched_tensor_core": recb,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the process's original stdout; everything else (NCCL's version banner, library chatter) was
    diverted to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)   # native libraries (NCCL prints its version to stdout) and stray prints -> stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
