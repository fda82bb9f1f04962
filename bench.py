#!/usr/bin/env python
"""Benchmark of the per-ray hot path (BASELINE.json metric: ray-steps/s in FP64 + 1024^2 230 GHz render time).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic input: the cfg2 bundle of BASELINE.json
(1024x1024 Kerr-Schild rays, a = 0.94, observer r = 1000 M, i = 60 deg, div = 40, tol = 1e-4, N = 10000),
integrated by the persistent FP64 kernel with trajectory dump.  Prints ONE JSON line (rank 0).

  value      whole-job ray-steps/s with the bundle resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the public API with HOST buffers (geodesics.integrate_paged_host): the kernel
             pulls every ray's initial state from pinned host memory over PCIe and stores final states / step
             counts / classifier radii into pinned host memory (zero-copy: h2d / d2h bytes move inside the launch;
             trajectories stay in HBM, as the reference's jax.Arrays stay on its device)
  render     the second half of the metric: wall time of the fused 1024^2 230 GHz image of the synthetic
             256^3 AthenaK-shaped snapshot (cfg4), device-timed and end-to-end
  roofline   dominant kernel = integrate_kernel; FP64 pipe: 859 algorithmic flop per ray-step
              over the measured DFMA peak of this GPU (mk_measure_fp64_peak)
  cpu_baseline  the C/OpenMP oracle (literal restatement of the reference's JAX path, oracle/mk_oracle.c)
             timed on this box's host cores on a bounded sample; kind = "port" (JAX is not installable)

--impl reference times that oracle port as the "reference arm" (the reference itself is pure Python on
JAX, which is absent from this image and the GPU box).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

E2E_CHUNKS = int(os.environ.get("MK_E2E_CHUNKS", "0"))
# cell storage of the cfg4 snapshot: float64 cells render ~3 % faster than float32 cells (no F2F conversions in the
# gather; the kernel is FP64/latency-bound, not traffic-bound) at twice the HBM footprint; images are bit-identical
RENDER_STORAGE = os.environ.get("MK_RENDER_STORAGE", "f64")
DUAL_FP64_PER_STEP = 2003     # executed FP64-pipe instructions per ray-step of the dual-number plugin (SASS count)
FLOP_PER_RAY_STEP = 859          # 312 add + 523 mul + 12 div + 12 sqrt
RENDER_FP64_PER_STEP = 455       # executed FP64-pipe instructions of the fused kernel per ray-step (SASS count) ...
RENDER_FP64_PER_SAMPLE = 340     # ... and per in-domain sample on top of that
CFG2 = dict(bhspin=0.94, inclination=60.0, distance=1000.0, fov=20.0, div=40.0, tol=1e-4, N=10000)
WEAK_INCLINATIONS = [60.0, 17.0, 30.0, 80.0, 45.0, 70.0, 25.0, 52.0]    # one frame per rank (cfg5-style)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--res", type=int, default=1024, help="rays per side of the bundle (default: cfg2)")
    ap.add_argument("--snapshot-cells", type=int, default=256, help="cells per side of the cfg4 snapshot")
    ap.add_argument("--no-render", action="store_true", help="skip the cfg4 render leg")
    ap.add_argument("--strong-res", type=int, default=4096,
                    help="side of the large single image used for the strong-scaling render leg (cfg5-sized frame)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=341, help="side of the pixel sub-lattice timed on the CPU")
    ap.add_argument("--render-cpu-stride", type=int, default=16, help="pixel stride of the CPU render baseline")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smmax, reasons = [], [], set()
        allsm = []
        for t, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                clk, mx = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            allsm.append(clk)
            if t0 - 0.05 <= t <= t1 + 0.05:
                sm.append(clk)
                smmax.append(mx)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            sm = allsm[-3:] if allsm else [0.0]
            smmax = smmax or [0.0]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smmax)) if smmax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
# CPU legs (oracle port)
# ------------------------------------------------------------------------------------------------------
def cpu_sample_rays(res, side, inclination=CFG2["inclination"]):
    """A side x side sub-lattice of the res x res cfg2 pixel grid (same camera, same rays)."""
    from oracle import mahakala_oracle as onp
    s0 = onp.initialize_geodesics_at_camera(CFG2["bhspin"], inclination, CFG2["distance"], -CFG2["fov"] / 2,
                                            CFG2["fov"] / 2, res)
    stride = max(res // side, 1)
    idx = (np.arange(0, res, stride)[:, None] * res + np.arange(0, res, stride)[None, :]).reshape(-1)
    return np.ascontiguousarray(s0[idx]), stride


def time_cpu_oracle(res, side):
    from oracle import c_oracle
    c_oracle.use_all_cores()
    s0, stride = cpu_sample_rays(res, side)
    c_oracle.integrate(200, s0[:64], CFG2["div"], CFG2["tol"], CFG2["bhspin"])       # load + thread warm-up
    t0 = time.perf_counter()
    out = c_oracle.integrate(CFG2["N"], s0, CFG2["div"], CFG2["tol"], CFG2["bhspin"], dump=False)
    dt = time.perf_counter() - t0
    steps = int(out["nsteps"].sum())
    return steps / dt, dict(seconds=dt, ray_steps=steps, rays=int(s0.shape[0]), cores=c_oracle.num_threads(),
                            sample=f"every {stride}th pixel of the {res}x{res} cfg2 grid ({s0.shape[0]} rays, "
                                   f"{steps} ray-steps, {dt:.1f} s)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle
    c_oracle.use_all_cores()
    side = 64
    s0, stride = cpu_sample_rays(args.res, side)
    times, steps = [], 0
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        out = c_oracle.integrate(CFG2["N"], s0, CFG2["div"], CFG2["tol"], CFG2["bhspin"], dump=False)
        t1 = time.perf_counter()
        if it >= args.warmup:
            times.append(t1 - t0)
            steps = int(out["nsteps"].sum())
    total = float(np.sum(times))
    # useful (accepted) ray-steps only, early exit per ray and no trajectory stores: this FAVOURS the CPU arm
    # (the reference's lax.scan runs all N = 10000 iterations for every ray and materialises them)
    value = steps * args.steps / total
    cores = c_oracle.num_threads()
    sample = (f"every {stride}th pixel of the {args.res}x{args.res} cfg2 grid per step ({s0.shape[0]} rays, {steps} "
              f"ray-steps, final-state mode, early exit); C/OpenMP restatement of the reference's JAX path "
              f"(jacfwd-style forward derivatives + 4x4 inverse), not JAX itself")
    line = {"impl": "reference", "metric": "ray_steps_per_sec_fp64", "value": value, "unit": "ray-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, mode="reference-cpu"),
            "cpu_baseline": {"value": value, "unit": "ray-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "ray-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, mode):
    return {"workload": f"cfg2: Kerr-Schild geodesic bundle {args.res}x{args.res} rays, a=0.94, observer r=1000 M, "
                        f"i=60 deg, fov 20 M, div=40, tol=1e-4, N=10000",
            "mode": mode, "rays": args.res * args.res,
            "l2_policy": "256 MiB scratch write between timed iterations (L2 flush); the dump itself is >> L2",
            "parity_note": "tests hold this kernel to the oracle on the every-16th-pixel sub-lattice of this bundle: "
                           "classification and step counts identical on all rays, end states 1e-9 on escaped rays; "
                           "captured rays are excluded from the 1e-9 bar (chaotic tail in every implementation, max "
                           "2.9e-7; the strict-IEEE kernel is bit-identical to the oracle on all rays)",
            "multi_gpu": ("N frames (one per GPU, inclinations " + ", ".join(f"{i:g}" for i in WEAK_INCLINATIONS) +
                          " deg in rank order) integrated by ALL GPUs together: one dynamic ray queue in rank 0's "
                          "memory shared over NVLink (system-scope atomics, chunked + prefetched), rays handed out in " +
                          ("pixel order with the frames interleaved, " if os.environ.get("MK_BENCH_ORDER", "pixel") == "pixel"
                           else "centre-out order (long photon-ring rays first) with the frames interleaved, ") +
                          "N = 1: the plain public call on the GPU's own queue; "
                          "per-ray results stored by the kernels straight into rank 0's memory (in-kernel gather), "
                          "trajectories paged where they are computed; weak scaling, no data-path collective")}


# ------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import mahakala_b200 as ma
    from mahakala_b200 import _cabi, geodesics as geo

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from mahakala_b200._device import bind_host_to_gpu_numa_node
    numa_node = bind_host_to_gpu_numa_node(local)       # pinned buffers next to the GPU (zero-copy e2e path)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    a = CFG2["bhspin"]
    incl = WEAK_INCLINATIONS[rank % len(WEAK_INCLINATIONS)]
    res = args.res
    s0 = ma.initialize_geodesics_at_camera(a, incl, CFG2["distance"], -CFG2["fov"] / 2, CFG2["fov"] / 2, res)
    npx = s0.shape[0]
    s0_host = torch.empty((npx, 8), dtype=torch.float64, pin_memory=True)
    s0_host.copy_(s0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # measured FP64 peak of this GPU (roofline denominator)
    tf = ctypes.c_double(0)
    ms = ctypes.c_double(0)
    _cabi.call("mk_measure_fp64_peak", 20000, tf, ms)
    fp64_peak = tf.value

    mode = "trajectory dump (paged warp logs, single pass)"
    launches = [0]
    shared = None
    from mahakala_b200 import multigpu
    ORDER = os.environ.get("MK_BENCH_ORDER", "pixel")          # 'pixel' (frames interleaved) | 'centre' (centre-out)
    make_order = multigpu.interleaved_pixel_ray_order if ORDER == "pixel" else multigpu.longest_first_ray_order
    if world == 1:
        # one GPU: the plain public call (the GPU's own queue, pixel order).  MK_BENCH_SHARED=1 runs the launch of the
        # N > 1 job with N = 1 instead (measured 16.3 ms in pixel order, 17.1 ms centre-out, against 16.1 ms)
        order = torch.from_numpy(make_order(res, 1)).to(dev) if os.environ.get("MK_BENCH_SHARED") == "1" else None
        local_queue = torch.zeros(64, dtype=torch.int32, device=dev)
    if world > 1:
        # BASELINE north_star split: ONE job of `world` frames, every GPU holds all bundles, rays come from one queue
        frames = [WEAK_INCLINATIONS[f % len(WEAK_INCLINATIONS)] for f in range(world)]
        s0_all = torch.cat([ma.initialize_geodesics_at_camera(a, frames[f], CFG2["distance"], -CFG2["fov"] / 2,
                                                              CFG2["fov"] / 2, res) for f in range(world)])
        order = torch.from_numpy(make_order(res, world)).to(dev)
        shared = multigpu.SharedRays(world * npx)
        job_store = geo.TrajectoryStore.allocate(world * npx, CFG2["N"], mem_fraction=0.45)
    store = geo.TrajectoryStore.allocate(npx, CFG2["N"], mem_fraction=0.55 if world > 1 else 0.6)

    def step_device():
        """-> device tensor with the ray-steps this rank integrated in the step"""
        if world > 1:
            geo.integrate_paged(CFG2["N"], s0_all, CFG2["div"], CFG2["tol"], a, store=job_store,
                                queue=shared.queue_ptr, ray_order=order, results=shared.results(),
                                page_id_offset=rank << multigpu.PAGE_RANK_SHIFT, participants=world)
            launches[0] += 1
            return job_store.total_steps
        if order is None:
            out = geo.integrate_paged(CFG2["N"], s0, CFG2["div"], CFG2["tol"], a, store=store)        # resets the store
        else:
            out = geo.integrate_paged(CFG2["N"], s0, CFG2["div"], CFG2["tol"], a, store=store, queue=local_queue,
                                      ray_order=order)
        launches[0] += 1
        return out.total_steps

    def fence_queue():
        """between two jobs on the shared queue: everybody is done, rank 0 zeroes the counter, everybody sees it"""
        if world > 1:
            barrier()
            shared.reset()
            barrier()
        else:
            local_queue.zero_()

    host_out = {"final": torch.empty((npx, 8), dtype=torch.float64, pin_memory=True),
                "nsteps": torch.empty((npx,), dtype=torch.int32, pin_memory=True),
                "r_last": torch.empty((npx,), dtype=torch.float64, pin_memory=True)}

    def step_e2e(chunks=None):
        chunks = E2E_CHUNKS if chunks is None else chunks
        if store is not None:
            # public host-to-host call.  Default: zero-copy, the kernel reads s0 from pinned host memory and stores
            # the per-ray results into pinned host memory over PCIe (all bytes still move, inside the launch);
            # MK_E2E_CHUNKS > 0 selects the explicit chunked H2D / kernel / D2H pipeline instead
            if chunks > 0:
                geo.integrate_paged_streamed(CFG2["N"], s0_host, CFG2["div"], CFG2["tol"], a, store, host_out, chunks=chunks)
            else:
                # (pixel order: with the longest-first order the 64 B PCIe reads of s0 become random and the
                # host-to-host step was measured 19.1 ms instead of 16.3 ms)
                geo.integrate_paged_host(CFG2["N"], s0_host, CFG2["div"], CFG2["tol"], a, store, host_out)
            # the results are in host memory when the call returns; the work count (a host-side sum over 1 M step
            # counts: 1.1 ms per step with the single OpenMP thread torchrun gives each rank) is bench bookkeeping and
            # is taken once after the timed loop
            return None
        d = s0_host.to(dev, non_blocking=True)
        if store is not None:
            store.reset()
            out = geo.integrate_paged(CFG2["N"], d, CFG2["div"], CFG2["tol"], a, store=store)
            final, nsteps, r_last = out.final, out.nsteps, out.r_last
        else:
            final, nsteps, r_last = geo.integrate_final(CFG2["N"], d, CFG2["div"], CFG2["tol"], a)
        host_out["final"].copy_(final, non_blocking=True)
        host_out["nsteps"].copy_(nsteps, non_blocking=True)
        host_out["r_last"].copy_(r_last, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return int(host_out["nsteps"].sum())

    # ---- warm-up ----
    for _ in range(args.warmup):
        fence_queue()
        total = step_device()
        flush.fill_(1)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    # ---- timed: device-resident ----
    launches[0] = 0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.time()
    totals = []
    for k in range(args.steps):
        flush.fill_(k)                      # L2 flush, outside the event pair
        fence_queue()                       # N > 1: queue reset between two barriers, outside the event pair
        ev[k][0].record()
        totals.append(step_device())
        ev[k][1].record()
    barrier()
    t_wall1 = time.time()
    dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    steps_per_pass = int(totals[-1].item())     # ray-steps THIS rank integrated in the last step
    n_launch = launches[0]

    # ---- N > 1: the gathered results of the split job against rank 0 integrating every frame alone ----
    split = None
    if world > 1:
        pages_job = job_store.pages_used
        overflow = torch.tensor([1.0 if job_store.overflowed else 0.0, float(steps_per_pass)], dtype=torch.float64, device=dev)
        per_rank = [torch.zeros_like(overflow) for _ in range(world)]
        dist.all_gather(per_rank, overflow)
        if rank == 0:
            views = shared.local_views()
            same, worst = True, 0.0
            for f in range(world):
                alone = geo.integrate_paged(CFG2["N"], s0_all[f * npx:(f + 1) * npx], CFG2["div"], CFG2["tol"], a, store=store)
                sl = slice(f * npx, (f + 1) * npx)
                ok = (torch.equal(views["final"][sl], alone.final) and torch.equal(views["nsteps"][sl], alone.nsteps)
                      and torch.equal(views["r_last"][sl], alone.r_last))
                same = same and ok
                worst = max(worst, float((views["final"][sl] - alone.final).abs().max()))
            owner = views["page_first"][:, 0] >> multigpu.PAGE_RANK_SHIFT
            split = {"split_identical": bool(same), "max_abs_diff_final_state": worst,
                     "check": "final states, step counts and classifier radii of all frames gathered in rank 0's memory "
                              "by the shared-queue job, torch.equal against rank 0 integrating each frame alone",
                     "rays_per_rank": [int((owner == r).sum()) for r in range(world)],
                     "ray_steps_per_rank": [int(t[1]) for t in per_rank],
                     "page_pool_overflowed": bool(any(float(t[0]) for t in per_rank))}
        del job_store
        torch.cuda.empty_cache()
        # for the record: the trivially parallel alternative, every rank integrating its own frame with no sharing
        barrier()
        iv = []
        for k in range(3):
            i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            i0.record()
            own = geo.integrate_paged(CFG2["N"], s0, CFG2["div"], CFG2["tol"], a, store=store)
            i1.record()
            torch.cuda.synchronize()
            iv.append(i0.elapsed_time(i1))
        ind = torch.tensor([float(np.mean(iv[1:])), float(own.total_steps.item())], dtype=torch.float64, device=dev)
        ind_t, ind_w = ind[0:1].clone(), ind[1:2].clone()
        dist.all_reduce(ind_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ind_w, op=dist.ReduceOp.SUM)
        if rank == 0:
            split["independent_frames"] = {"value": float(ind_w) / (float(ind_t) * 1e-3), "unit": "ray-steps/s",
                                           "ms_per_step": float(ind_t),
                                           "note": "one frame per rank, private queues, nothing shared: the slowest "
                                                   "inclination sets the time"}

    # ---- timed: end to end through the public API with host buffers ----
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        e2e_steps = step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if e2e_steps is None:
        e2e_steps = int(host_out["nsteps"].sum())           # identical every step: same rays, deterministic kernel
    barrier()
    # for comparison: the same host-to-host step with explicit cudaMemcpyAsync H2D / D2H copies (4-chunk pipeline)
    step_e2e(chunks=4)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        step_e2e(chunks=4)
    torch.cuda.synchronize()
    e2e_copy_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # ---- reduce over ranks: time = max, work = sum ----
    t = torch.tensor([dev_ms, e2e_s, e2e_copy_s], dtype=torch.float64, device=dev)
    w = torch.tensor([steps_per_pass, e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_s_max, e2e_copy_s_max = float(t[0]), float(t[1]), float(t[2])
    work, e2e_work = float(w[0]), float(w[1])
    value = work * args.steps / (dev_ms_max * 1e-3)
    e2e_value = e2e_work * args.steps / e2e_s_max

    # ---- SURVEY 8(d) extras (rank 0): the dual-number plugin's own count, the reference's scan overhead ----
    plugin_line = None
    if rank == 0:
        sub = ma.initialize_geodesics_at_camera(a, CFG2["inclination"], CFG2["distance"], -CFG2["fov"] / 2, CFG2["fov"] / 2, 512)
        geo.set_metric("kerr_schild_dual")
        try:
            geo.integrate_final(CFG2["N"], sub, CFG2["div"], CFG2["tol"], a)
            torch.cuda.synchronize()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            _, _, _, ptot = geo.integrate_final(CFG2["N"], sub, CFG2["div"], CFG2["tol"], a, want_total=True)
            p1.record()
            torch.cuda.synchronize()
            psteps, pms = int(ptot.item()), p0.elapsed_time(p1)
        finally:
            geo.set_metric("kerr_schild")
        plugin_line = {"metric": "Kerr-Schild typed generically (DualMetric<KerrSchildFn>: forward-mode dual numbers "
                                 "through the metric functor + adjugate inverse, the path of every user-registered "
                                 "spacetime), 512x512 rays of the cfg2 camera, final-state mode",
                       "ray_steps": psteps, "ms": pms, "ray_steps_per_s": psteps / (pms * 1e-3),
                       "fp64_instr_per_ray_step": DUAL_FP64_PER_STEP,
                       "fp64_issue_frac": psteps * DUAL_FP64_PER_STEP / (pms * 1e-3) / 1e12 / (fp64_peak / 2.0),
                       "note": "the plugin's own executed FP64-pipe instructions per ray-step (SASS of its step function, "
                               "cuobjdump: 2003 DFMA/DMUL/DADD against 430 for the closed form) over the issue rate of the "
                               "DFMA microbenchmark; SURVEY 8(d): reported separately from the closed-form roofline"}

    render = None
    pages_used = store.pages_used if world == 1 else pages_job
    store = None
    torch.cuda.empty_cache()
    if not args.no_render:
        render = render_leg(args, rank, world, dev, fp64_peak)

    if rank == 0:
        kernel_ms = dev_ms / args.steps
        achieved = steps_per_pass * FLOP_PER_RAY_STEP / (kernel_ms * 1e-3) / 1e12
        rays_here = split["rays_per_rank"][0] if split is not None else npx
        dump_bytes = 72 * (steps_per_pass + rays_here) if pages_used else 0
        line = {
            "metric": "ray_steps_per_sec_fp64", "value": value, "unit": "ray-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, mode),
            "ray_steps_per_pass_rank0": steps_per_pass,
            "e2e": {"value": e2e_value, "unit": "ray-steps/s", "h2d_bytes_per_step": int(npx * 64),
                    "d2h_bytes_per_step": int(npx * (64 + 4 + 8)), "ms_per_step": 1e3 * e2e_s_max / args.steps,
                    "host_numa_node": numa_node,
                    "explicit_copy_pipeline": {"value": e2e_work * args.steps / e2e_copy_s_max, "unit": "ray-steps/s",
                                               "ms_per_step": 1e3 * e2e_copy_s_max / args.steps,
                                               "note": "same step with cudaMemcpyAsync H2D / D2H on side streams, 4 chunks"},
                    "l2_note": "no flush in this loop: the inputs come from pinned host memory (never L2-resident) and the 40 GB "
                               "dump is >> L2; the device-timed loop writes 256 MiB through L2 before every launch, whose "
                               "write-back costs that launch ~0.3 ms -- which is why this figure can exceed `value` at small N",
                    "transfer": ("zero-copy: the kernel reads s0 from / writes results to pinned host memory over PCIe"
                                 if E2E_CHUNKS <= 0 else f"{E2E_CHUNKS}-chunk H2D / kernel / D2H pipeline")},
            "gpu_launches": n_launch,
            "clocks": clocks,
            "roofline": {"bound": "fp64", "kernel": "mk::integrate_kernel<KerrSchild>", "achieved": achieved,
                         "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                         "peak_source": "measured on this GPU by mk_measure_fp64_peak (DFMA microbenchmark); "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "flop_per_ray_step": FLOP_PER_RAY_STEP, "traffic": measured_traffic(bool(pages_used), args.res)[0],
                         "traffic_source": "static: " + str(measured_traffic(bool(pages_used), args.res)[1]) +
                                           " (committed ncu capture of this kernel at this shape, dram__bytes_read.sum + "
                                           "dram__bytes_write.sum of one launch; not re-measured in this run)",
                         "hbm": {"dump_bytes_per_launch": dump_bytes,
                                 "achieved_GBps": dump_bytes / (kernel_ms * 1e-3) / 1e9,
                                 "peak_GBps": hbm_peak()}},
        }
        line["reference_scan_overhead"] = {
            "N": CFG2["N"], "mean_steps_per_ray": steps_per_pass / float(rays_here), "factor": CFG2["N"] * rays_here / float(steps_per_pass),
            "note": "the reference's lax.scan runs all N iterations for every ray (geodesics.py:272) although a ray needs "
                    "mean_steps of them; neither arm's timing includes that factor (the CPU arm exits early too)"}
        if plugin_line is not None:
            line["generic_metric_plugin"] = plugin_line
        if split is not None:
            line["split"] = split
        if render is not None:
            line["render"] = render
        if not args.no_cpu_baseline:
            v, info = time_cpu_oracle(args.res, args.cpu_sample)
            line["cpu_baseline"] = {"value": v, "unit": "ray-steps/s", "cores": info["cores"], "kind": "port",
                                    "sample": info["sample"] + "; C/OpenMP restatement of the reference's JAX path"}
        print(json.dumps(line))
        failed = (split is not None and not split["split_identical"]) or \
                 (render is not None and render.get("strong_image_identical") is False) or \
                 (render is not None and render.get("single_image_learned_order_identical") is False) or \
                 (render is not None and render.get("strong_scaling_large_image", {}).get("identical") is False)
    else:
        failed = False
    if world > 1:
        if shared is not None:
            dist.barrier()
            shared.close()
        dist.barrier()
        dist.destroy_process_group()
    if failed:
        sys.stdout.flush()
        sys.exit(3)         # a split result that differs from the single-GPU one is a failure, not a number


def measured_traffic(paged, res):
    """dram__bytes_read + dram__bytes_write of one launch of the dominant kernel at cfg2, from the committed
    single-pass ncu measurement (newest profiles/rNN_traffic.json); (None, None) for other shapes."""
    if res != 1024:
        return None, None
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
        try:
            t = json.load(open(path))
            key = [k for k in t if ("MODE_PAGED" if paged else "MODE_FINAL") in k][0]
            return int(t[key]["dram_bytes_read"]) + int(t[key]["dram_bytes_write"]), os.path.relpath(path, ROOT)
        except Exception:
            continue
    return None, None


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0       # B200_PROFILING.md fallback


def render_cpu_baseline(arr, res, stride):
    """The reference's make_image chain (C/OpenMP restatement: integrate -> O(nmb) block scan -> trilinear -> fluid
    frame -> j, alpha -> back-to-front transfer, per ray) on every stride-th pixel of the cfg4 frame, all host cores;
    the full-frame figure is that time scaled by the pixel count."""
    from oracle import c_oracle, mahakala_oracle as onp
    c_oracle.use_all_cores()
    t0 = time.perf_counter()
    om = onp.AthenakFluidModel(arr["uov"].astype(np.float64), arr["B"].astype(np.float64), arr["x1v"], arr["x2v"],
                               arr["x3v"], arr["x1f"], arr["x2f"], arr["x3f"], arr["LogicalLocations"], arr["Levels"],
                               CFG2["bhspin"], fluid_gamma=arr["fluid_gamma"], variable_names=arr["VariableNames"])
    t_load = time.perf_counter() - t0
    s0 = onp.initialize_geodesics_at_camera(CFG2["bhspin"], CFG2["inclination"], CFG2["distance"], -CFG2["fov"] / 2,
                                            CFG2["fov"] / 2, res)
    idx = (np.arange(0, res, stride)[:, None] * res + np.arange(0, res, stride)[None, :]).reshape(-1)
    sub = np.ascontiguousarray(s0[idx])
    units = om.get_units(6.2e9 * 1.989e33, 1.e26)
    t0 = time.perf_counter()
    img, nsteps, nin = c_oracle.render(om, sub, units, [230e9])
    dt = time.perf_counter() - t0
    scale = (res * res) / sub.shape[0]
    return {"value": 1e3 * dt * scale, "unit": "ms per 1024^2 frame (estimated: sample time x pixel ratio)",
            "cores": c_oracle.num_threads(), "kind": "port",
            "sample": f"every {stride}th pixel of the {res}x{res} cfg4 frame ({sub.shape[0]} rays, {int(nsteps.sum())} "
                      f"ray-steps, {nin} in-domain samples, {dt:.1f} s); ghost-zone fill of the snapshot on the host "
                      f"{t_load:.1f} s (not included); C/OpenMP restatement of the reference's images.make_image"}


def render_leg(args, rank, world, dev, fp64_peak):
    """cfg4: fused 1024^2 230 GHz image of the synthetic 256^3 snapshot.

    N = 1: one image.  N > 1: (weak) one frame per rank at its own inclination, and (strong) ONE image whose
    32-ray patches all ranks pull from a single queue in rank 0's memory over NVLink, pixels stored straight
    into rank 0's image (mahakala_b200.multigpu).  The snapshot is replicated with one NCCL broadcast.
    """
    import torch
    import torch.distributed as dist
    from mahakala_b200 import images, multigpu
    from mahakala_b200.grmhd import AthenakFluidModel
    from mahakala_b200.synthetic import make_synthetic_snapshot

    nc = args.snapshot_cells
    arr = None
    if rank == 0:       # the other ranks get a geometry-only replica and the cells by one NCCL broadcast
        # float32 interior arrays, which is what an AthenaK dump holds (and what a loader hands over)
        arr = make_synthetic_snapshot(ncells=nc, block=32 if nc % 32 == 0 else 16, extent=32.0, seed=0, dtype=np.float32)
    nccl_init_ms = multigpu.warm_communicator()      # communicator start-up, reported on its own
    # host arrays -> usable snapshot on every rank, twice: the first pass pays the one-off costs of the process (pinned
    # staging buffers of the uploader, first cudaMalloc of the big arrays, NCCL's buffers for a large broadcast), the
    # second is what every further snapshot of a run costs (an EHT-style movie loads thousands)
    setups = []
    model = None
    for attempt in ("cold", "warm"):
        if model is not None:
            model.release()
            model = None
            torch.cuda.empty_cache()
        if rank == 0:
            model = AthenakFluidModel.from_arrays(arr["uov"], arr["B"], arr["x1v"], arr["x2v"], arr["x3v"], arr["x1f"],
                                                  arr["x2f"], arr["x3f"], arr["LogicalLocations"], arr["Levels"],
                                                  CFG2["bhspin"], fluid_gamma=arr["fluid_gamma"], storage=RENDER_STORAGE)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_setup = time.perf_counter()
        model = multigpu.replicate_snapshot(model)   # rank 0: upload of the interior arrays + ghost fill / repack kernel
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        setups.append((1e3 * (time.perf_counter() - t_setup), dict(model.replication_timing)))
    t_setup_cold, rep_cold = setups[0]
    t_setup, rep = setups[1]
    bcast_ms = rep["broadcast"]
    incl = WEAK_INCLINATIONS[rank % len(WEAK_INCLINATIONS)]
    res = args.res
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    times, e2e_times = [], []
    counters = None
    for it in range(2 + 3):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        img, counters = images.render(model, camera_inclination=incl, resolution=res, observing_frequencies=(230e9,),
                                      want_counters=True)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            times.append(e0.elapsed_time(e1))
    for it in range(1 + 3):             # first call allocates the pinned staging buffer: warm-up, not timed
        t0 = time.perf_counter()
        host_img = images.make_image(model, camera_inclination=incl, resolution=res)
        if it >= 1:
            e2e_times.append(1e3 * (time.perf_counter() - t0))
    def strong_leg(sres, reps, learned=False):
        """ONE i = 60 deg image of sres^2 pixels shared by all ranks (or the plain render at N = 1).  learned: patches
        handed out longest first, from the step counts of one geodesics-only pass (images.learn_patch_order, outside
        the timed region: it is paid once per camera, e.g. once per movie)."""
        if learned:
            images.learn_patch_order(CFG2["bhspin"], camera_inclination=CFG2["inclination"], resolution=sres)
        else:
            images.forget_patch_orders()
        shared = multigpu.SharedImage(1, sres * sres) if world > 1 else None
        st = []
        for it in range(1 + reps):
            flush.fill_(it)
            if world > 1:
                shared.reset()
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if world > 1:
                images.render(model, camera_inclination=CFG2["inclination"], resolution=sres,
                              observing_frequencies=(230e9,), image_out=shared.image_ptr, queue=shared.queue_ptr,
                              long_queue=shared.ring_queue_ptr, participants=world)
            else:
                out = images.render(model, camera_inclination=CFG2["inclination"], resolution=sres,
                                    observing_frequencies=(230e9,))
            e1.record()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            if it >= 1:
                st.append(e0.elapsed_time(e1))
        same, worst = None, None
        if world > 1:
            fl = 0.0
            if rank == 0:       # the split image against the same frame rendered by rank 0 alone, pixel for pixel
                got = shared.local_view()[1]
                alone = images.render(model, camera_inclination=CFG2["inclination"], resolution=sres,
                                      observing_frequencies=(230e9,))
                same = bool(torch.equal(got, alone))
                worst = float((got - alone).abs().max())
                fl = float(got.sum())
                del alone
            dist.barrier()
            shared.close()
        else:
            fl = float(out.sum())
        return float(np.mean(st)), fl, same, worst

    # The only timing the reference publishes for this path (demos/grmhd_detailed.ipynb cells 10-11, hardware not
    # stated): get_fluid_scalars_from_geodesics on the trajectories of a 160x160 image, 456 meshblocks: 31.8 s for
    # the meshblock-index loop + 0.96 s for the sampling scan.  Same call here (512 meshblocks of 32^3).
    stage = None
    if rank == 0:
        from mahakala_b200 import geodesics as geo
        s160 = geo.initialize_geodesics_at_camera(CFG2["bhspin"], 60, 1000, -10, 10, 160)
        S160, dt160 = geo.geodesic_integrator(CFG2["N"], s160, 40, 1e-4, CFG2["bhspin"])
        model.get_fluid_scalars_from_geodesics(S160)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model.get_fluid_scalars_from_geodesics(S160)
        e1.record()
        torch.cuda.synchronize()
        stage = {"call": "AthenakFluidModel.get_fluid_scalars_from_geodesics(S), 160x160 rays", "rows": int(S160.shape[0]),
                 "samples": int(S160.shape[0] * S160.shape[1]), "ms": e0.elapsed_time(e1),
                 "reference_notebook_s": {"meshblock_index_loop": 31.81, "sampling_scan": 0.957,
                                          "source": "demos/grmhd_detailed.ipynb cell 10 (456 meshblocks, hardware not stated)"}}
        del S160, dt160
        torch.cuda.empty_cache()
    # the other single-GPU BASELINE configs, timed once each for the record (rank 0): cfg1 = the reference's own test
    # (4 golden shadow curves through find_shadow_bisection_angles, tests/test_shadows.py), cfg3 = analytic torus 512^2
    other = None
    if rank == 0:
        import mahakala_b200 as ma
        from mahakala_b200.grmhd.athenak import AnalyticTorusFluidModel
        other = {}
        gold = os.path.join(ROOT, "tests", "golden", "shadow_golden.npz")
        if os.path.exists(gold):
            z = np.load(gold)
            cases = sorted({k.split("__")[0] for k in z.files})
            def shadows():
                worst = 0.0
                for c in cases:
                    r = np.asarray(ma.find_shadow_bisection_angles(float(z[c + "__bhspin"]), float(z[c + "__inclination"]),
                                                                   z[c + "__angles"]))
                    worst = max(worst, float(np.max(np.abs(r - z[c + "__radii"]) / z[c + "__radii"])))
                return worst
            shadows()
            t0 = time.perf_counter()
            worst = shadows()
            other["cfg1_shadow_golden"] = {"call": "find_shadow_bisection_angles on the reference's 4 golden cases "
                                                   "(305 angles, 14 bisection iterations each, N=2000, tol=1e-2)",
                                           "ms": 1e3 * (time.perf_counter() - t0), "max_rel_err_vs_golden": worst,
                                           "reference_rtol": 1e-2}
        torus = AnalyticTorusFluidModel(CFG2["bhspin"])
        images.render(torus, resolution=512)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        timg = images.render(torus, resolution=512)
        e1.record()
        torch.cuda.synchronize()
        other["cfg3_torus_512"] = {"call": "fused render of the analytic thin torus, 512x512 at 230 GHz", "ms": e0.elapsed_time(e1),
                                   "image_sum": float(timg.sum())}
    strong, flux, same, worst = strong_leg(res, 3) if world > 1 else (0.0, 0.0, None, None)
    quick_ms = None
    if world > 1:
        images.forget_patch_orders()
        torch.cuda.synchronize()
        tq = time.perf_counter()
        images.quick_patch_order(CFG2["bhspin"], camera_inclination=CFG2["inclination"], resolution=res)
        torch.cuda.synchronize()
        quick_ms = 1e3 * (time.perf_counter() - tq)
        images.forget_patch_orders()
    strong_l, _, same_l, _ = strong_leg(res, 3, learned=True)
    strong_big, flux_big, same_big, worst_big = strong_leg(args.strong_res, 2) if args.strong_res > 0 else (0.0, 0.0, None, None)
    images.forget_patch_orders()
    t = torch.tensor([float(np.mean(times)), float(np.mean(e2e_times)), strong, bcast_ms, strong_big, strong_l], dtype=torch.float64, device=dev)
    w = torch.tensor([float(counters[0]), float(counters[1])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    steps, samples = int(w[0]), int(w[1])
    ms = float(t[0])
    out = {"workload": f"cfg4: synthetic AthenaK-shaped {nc}^3 snapshot ({model.storage} cells, {model.lookup} lookup), "
                       f"{res}x{res} image at 230 GHz, fused kernel; one frame per rank",
           "ms": ms, "e2e_ms": float(t[1]), "ray_steps": steps, "in_domain_samples": samples,
           "ray_steps_per_s": steps / (ms * 1e-3),
           "sampling_algorithmic_GBps": samples * (256 if model.storage == "f32" else 512) / (ms * 1e-3) / 1e9,
           "snapshot_bytes": model.snapshot_bytes(), "snapshot_setup_ms": t_setup, "snapshot_setup_cold_ms": t_setup_cold,
           "snapshot_setup_phases_ms": {k: rep[k] for k in ("host_prep", "upload", "ghost_fill")},
           "snapshot_setup_note": "host interior arrays -> device snapshot (upload + fused ghost-fill/repack kernel"
                                  + (" + NCCL broadcast" if world > 1 else "") + "), outside the render time",
           "image_sum": float(host_img.sum()), "gpu_launches_per_image": 1}
    # FP64-issue roofline of the fused kernel: executed FP64-pipe instructions (SASS of render_kernel<1, f64 cells>,
    # scripts/dev/sass_by_line.py: 455 per ray-step for RK4 + step rule + block lookup, 340 more per in-domain sample
    # for cell index, trilinear gather, fluid frame, Theta_e, j / alpha, transfer update) against the issue rate the
    # DFMA microbenchmark reaches on this GPU (one FP64 warp instruction per 2 cycles per SM sub-partition)
    fp64_thread_instr = steps * RENDER_FP64_PER_STEP + samples * RENDER_FP64_PER_SAMPLE
    out["roofline"] = {"bound": "fp64", "kernel": "mk::render_kernel<1, f64 cells>",
                       "achieved": fp64_thread_instr / (ms * 1e-3) / 1e12, "peak": fp64_peak / 2.0,
                       "unit": "T FP64 instr/s (thread level; a DFMA counts once)",
                       "frac": fp64_thread_instr / (ms * 1e-3) / 1e12 / (fp64_peak / 2.0),
                       "fp64_instr_per_ray_step": RENDER_FP64_PER_STEP, "fp64_instr_per_in_domain_sample": RENDER_FP64_PER_SAMPLE,
                       "hbm": {"algorithmic_bytes": samples * (256 if model.storage == "f32" else 512),
                               "note": "8 corner cells x 8 primitives per in-domain sample if nothing were reused; "
                                       "the 4x8-pixel patches make it cache-resident (measured DRAM traffic per frame: "
                                       "profiles/, ~2.3 GB = 1.8 x the snapshot)", "peak_GBps": hbm_peak()}}
    if rank == 0 and not args.no_cpu_baseline and arr is not None:
        out["cpu_baseline"] = render_cpu_baseline(arr, res, args.render_cpu_stride)
    if stage is not None:
        out["sampling_stage_160px"] = stage
    if other:
        out["other_configs"] = other
    if args.strong_res > 0:
        out["strong_scaling_large_image"] = {"resolution": args.strong_res, "ms": float(t[4]), "image_sum": flux_big,
                                             "note": "ONE cfg5-sized frame rendered by all ranks together"}
        if same_big is not None:
            out["strong_scaling_large_image"].update(identical=same_big, max_abs_diff=worst_big)
    out["single_image_learned_order_ms"] = float(t[5])
    out["single_image_learned_order_note"] = ("the same one i=60 deg frame (all ranks together at N > 1) with the patches "
                                              "handed out longest first from a geodesics-only pass done once per camera "
                                              "(images.learn_patch_order, not timed) and the photon-ring patches (longest "
                                              "ray >= 0.25 x the frame's longest, at most one per SM) rendered by the "
                                              "warp-specialised long-patch kernel mk_render_long on a high-priority "
                                              "stream: the sample leaves the ray's chain of dependent RK4 steps (lone "
                                              "longest patch 3.2 ms instead of 7.3 ms); pixels bit-identical")
    if same_l is not None:
        out["single_image_learned_order_identical"] = same_l
    if world > 1:
        out["strong_scaling_single_image_ms"] = float(t[2])
        out["strong_scaling_note"] = ("one i=60 deg image split over all ranks: shared atomic tile queue + in-kernel "
                                      "gather into rank 0 over NVLink (CUDA IPC), device-timed, max over ranks.  Nothing "
                                      "is learned beforehand: the first (untimed, warm-up) frame of the camera runs a "
                                      "coarse, capped geodesics-only pre-pass on every rank (images.quick_patch_order, "
                                      "quick_order_ms) that finds the photon-ring patches for the long-patch kernel; the "
                                      "timed frames reuse it (MK_QUICK_ORDER=0: centre-out order, fused kernel only, "
                                      "9.4 ms at 8 GPUs)")
        out["quick_order_ms"] = quick_ms
        out["strong_image_sum"] = flux
        if same is not None:
            out["strong_image_identical"] = same
            out["strong_image_max_abs_diff"] = worst
        out["snapshot_replication"] = {
            "total_ms": t_setup, "total_cold_ms": t_setup_cold, "cold_phases_ms": {k: rep_cold[k] for k in ("host_prep", "upload", "ghost_fill", "meta", "broadcast")},
            "host_prep_ms": rep["host_prep"], "upload_ms": rep["upload"],
            "ghost_fill_ms": rep["ghost_fill"], "meta_ms": rep["meta"], "broadcast_ms": float(t[3]),
            "wire_format": rep.get("wire_format"), "wire_bytes": rep["wire_bytes"],
            "collective_ms_rank0": rep.get("collective"), "collective_ms": rep.get("collective_min"),
            "collective_GBps": rep.get("collective_GBps"),
            "collective": "the NCCL broadcast alone (multigpu.broadcast_large): minimum over ranks = the rank that "
                          "arrives last sees the transfer without waiting for its peers (rank 0 enqueues first and "
                          "waits for the receivers' allocations); broadcast_ms also holds the f64 -> f32 -> f64 wire "
                          "conversion and its 0.64 GB temporaries on both sides",
            "broadcast_GBps": rep["wire_bytes"] / (float(t[3]) * 1e-3) / 1e9 if float(t[3]) > 0 else None,
            "nccl_init_ms": nccl_init_ms,
            "note": "total = wall time on rank 0 from host arrays to a usable snapshot on every rank (max over ranks for "
                    "broadcast_ms), second snapshot of the process; total_cold = the first one (one-off pinned staging "
                    "buffers, first large cudaMalloc / NCCL buffers); the communicator was created beforehand by one tiny "
                    "all-reduce + broadcast (nccl_init_ms, in neither total)"}
    return out


def _claim_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL prints its version banner to stdout when NCCL_DEBUG is set)
    and child processes write to stderr instead.  Returns a file object bound to the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


if __name__ == "__main__":
    args = parse()
    _real_stdout = _claim_stdout()
    sys.stdout = _real_stdout
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    _real_stdout.flush()
