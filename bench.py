#!/usr/bin/env python
"""bench.py -- grid-nodes/sec end-to-end (Steps 1-2 heat-kernel summation + Step 3 constrained solve + shift) of the
B200-native signed-heat grid solver, and the reference arm beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

One "step" = one complete computeDistance over the workload (BASELINE.json metric):
  value  : N_grid_nodes * K / device-event time of K steps, sources already resident in HBM (shm3d_solve_device),
           result left on the device as float32 -- max over ranks;
  e2e    : the same through the public, reference-facing call SignedHeatGridSolver.computeDistance(V, F, options)
           (-> shm3d_prepare_mesh -> shm3d_solve): the mesh-level host half (rows a4-a6), H2D of the sources and D2H of
           the double field are all inside the timed region;
  roofline     : the dominant kernel, k_sum (Steps 1-2; SFU-bound: 2 MUFU per evaluated pair at 16/clk/SM);
  roofline_pcg : the PCG's fused p-update + stencil-apply + dot kernel (HBM-bound; 16 B/node/launch algorithmic);
  pcg_whole    : the whole Step-3 iteration against the HBM roofline, by SURVEY 8(d)'s 44 B/node/iteration formula and
                 by the bytes the V-cycle-preconditioned iteration really has to move;
  cpu_baseline : the reference's own computeDistance (oracle/_ref: its sources compiled against a shim) on the
                 workload's sources at 16^3, single-threaded; cpu_baseline_port_all_threads: the oracle's C port of the
                 Step 1-2 loop with OpenMP on a bounded sample of the real grid.
N > 1: the grid is z-slab partitioned over the ranks (NCCL halo exchange + all-reduces).  Default workloads keep
~512^3 nodes per GPU (512^3, 640^3, 768^3, 1024^3 at 1, 2, 4, 8 GPUs -> "scaling": "weak"); --workload sphereN fixes
the grid for strong-scaling runs.  At N > 1 the line also carries `dist_parity` (the slab-partitioned solve of a 256^3
case against the single-GPU solve of rank 0: rel-L2, iteration counts) and `strong` (the same grid solved by rank 0
alone -> speed-up of the N-GPU run).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "signed-heat-3d_b200"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "grid_nodes_per_sec_end_to_end"
UNIT = "grid-nodes/s"

WORKLOADS = {
    # name: (generator, grid nodes per axis, description).  "sphereN": N^3 grid around the 1e5-triangle unit sphere
    # (N = 16*2^h reproduces the reference's hCoef grids: 512 = hCoef 5, 1024 = hCoef 6; other N keep the same box).
    "knot128": ("knot", 128, "data/knot.obj (30504 faces, from tests/golden/knot.npz), 128^3 grid (hCoef 3)"),
}
for _n in (128, 256, 512, 640, 768, 1024):
    WORKLOADS[f"sphere{_n}"] = ("sphere", _n, f"synthetic unit sphere, 100000 outward triangles (Fibonacci hull), {_n}^3 grid")

# default workload per GPU count: ~512^3 nodes per GPU (weak scaling); N = 1 is the configuration BASELINE.json's
# target is quoted on (512^3, 1e5 triangles, 1 x B200), N = 8 its 1024^3 / 8 x B200 configuration
DEFAULT_BY_GPUS = {1: "sphere512", 2: "sphere640", 4: "sphere768", 8: "sphere1024"}


def make_workload(name):
    gen, n, desc = WORKLOADS[name]
    if gen == "sphere":
        from synth import fibonacci_sphere
        V, F = fibonacci_sphere(100000)
    else:
        z = np.load(os.path.join(ROOT, "tests", "golden", "knot.npz"))
        fo, fv = z["face_offsets"], z["face_vertices"]
        V, F = z["V"], [fv[fo[i]:fo[i + 1]].tolist() for i in range(len(fo) - 1)]
    return V, F, n, desc


def prepare(name):
    """host half of computeDistance (rows a4-a6) for the workload: (Params, pos, nrm, area, desc)"""
    import shm3d
    V, F, n, desc = make_workload(name)
    hc = int(round(np.log2(n / 16.0)))
    p, pos, nrm, area, _ = shm3d.prepare_mesh(V, F, hCoef=max(hc, 0))
    if p.nx != n:  # same box c +- 2r, n nodes per axis (the reference itself only produces 16*2^h)
        side = p.cell * (p.nx - 1)
        p.nx = p.ny = p.nz = n
        p.cell = side / (n - 1)
    return p, pos, nrm, area, desc


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                       "200", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            t = [x.strip() for x in line.split(",")]
            if len(t) < 8:
                continue
            try:
                sm.append(float(t[1]))
                mx.append(float(t[2]))
            except ValueError:
                continue
            for n, v in zip(names, t[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            busy = [s for s in sm if s >= 0.5 * max(sm)] or sm  # samples under load
            out = {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# --------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference's Step 1-2 loop on the host cores
# --------------------------------------------------------------------------------------------------------
def cpu_sample_plan(p, M, threads, seconds=8.0):
    """rows of one mid-domain plane such that the sample is ~`seconds` of CPU work (7e7 pair evaluations/s/thread
    assumed; every node costs the same M evaluations in the reference's brute-force loop)."""
    pairs = 7e7 * threads * seconds
    rows = int(max(1, min(p.ny, round(pairs / (p.nx * float(M))))))
    return rows


def cpu_step12_sample(p, pos, nrm, area, rows, threads):
    from oracle import shm_oracle as o
    g = o.Grid(p.nx, p.ny, p.nz, np.array(p.bbox_min), p.cell)
    k = p.nz // 2
    j0 = max(0, p.ny // 2 - rows // 2)
    t0 = time.perf_counter()
    Y = o.step12_box(g, p.lambda_, pos, nrm, area, j0, j0 + rows, k, k + 1, threads=threads)
    dt = time.perf_counter() - t0
    assert np.isfinite(Y).all()
    return rows * p.nx, dt


def cpu_baseline_obj(p, M, pos, nrm, area, seconds=12.0):
    from oracle import shm_oracle as o
    threads = o.max_threads()
    rows = cpu_sample_plan(p, M, threads, seconds)
    nodes, dt = cpu_step12_sample(p, pos, nrm, area, rows, threads)
    return {"value": nodes / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": (f"Steps 1-2 only (fp64 oracle port of src/signed_heat_grid_solver.cpp:48-65, OpenMP over rows), "
                       f"{rows} rows of plane k={p.nz // 2} = {nodes} nodes x {M} sources in {dt:.2f} s; Step 3 "
                       f"excluded (the reference's sparse LU of the KKT system is infeasible beyond 64^3: 945 s at "
                       f"64^3) -> an UPPER bound on the CPU path's end-to-end nodes/s")}


def reference_build_sample(name):
    """One run of the REFERENCE'S OWN computeDistance (oracle/_ref/libshm_ref.so: its two translation units compiled
    unmodified against oracle/ref_shim; sparse LU through a scipy SuperLU callback) on the workload's sources at the
    reference's smallest grid, hCoef 0 = 16^3.  Returns (nodes, seconds, description) or None when the library is absent."""
    from oracle import reference_build as rb
    if not rb.available():
        return None
    V, F, n, desc = make_workload(name)
    faces = F.tolist() if hasattr(F, "tolist") else F
    t0 = time.perf_counter()
    phi = rb.compute_distance_mesh(V, faces, hCoef=0)
    dt = time.perf_counter() - t0
    assert np.isfinite(phi).all()
    text = (f"the reference's own SignedHeatGridSolver::computeDistance (src/signed_heat_grid_solver.cpp + "
            f"src/signed_heat_3d.cpp compiled unmodified against oracle/ref_shim; Eigen's SparseLU replaced by a scipy "
            f"SuperLU callback), single-threaded like the reference, on the workload's {len(faces)} source faces at its "
            f"smallest grid hCoef 0 = 16^3: {phi.size} nodes, whole path incl. Step 3, in {dt:.2f} s.  Step 1-2 cost per "
            f"node does not depend on the grid and the KKT LU grows super-linearly, so nodes/s at 16^3 is an UPPER bound "
            f"for the {n}^3 workload")
    return phi.size, dt, text


def run_reference(args):
    """The reference's own CPU implementation of the path (oracle/_ref: its translation units compiled unmodified), on
    this box's host cores.  Nothing of the product is imported or loaded here."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import shm_oracle as o
    V, F, n_grid, desc = make_workload(args.workload)
    M = len(F)
    scaling = "weak" if args.workload == DEFAULT_BY_GPUS.get(args.gpus, "sphere512") else "strong"
    first = reference_build_sample(args.workload)  # also serves as warm-up
    if first is not None:
        t_tot, n_tot, text = 0.0, 0, first[2]
        for _ in range(args.steps):
            n, dt, text = reference_build_sample(args.workload)
            n_tot += n
            t_tot += dt
        v = n_tot / t_tot
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(1, args.steps), "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "same_config": False, "extrapolated": True,
                "config": {"workload": f"16^3 PROXY of: {desc} -- the workload's {M} sources on the reference's smallest grid "
                                       f"(hCoef 0 = 16^3); its sparse LU of the KKT system is infeasible beyond 64^3, so "
                                       f"the {n_grid}^3 grid itself cannot be run; nodes/s at 16^3 is an upper bound for it",
                           "grid": [16, 16, 16], "grid_of_the_gpu_arm": [n_grid] * 3, "sources": M},
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "per step: " + text},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    # oracle/_ref absent: the oracle's port of the Step 1-2 loop on a bounded sample of the real grid
    s_ = o.mesh_sources(V, F.tolist() if hasattr(F, "tolist") else F)
    hc = max(int(round(np.log2(n_grid / 16.0))), 0)
    g = o.make_grid(s_["centroid"], s_["radius"], hc)
    lam = o.lambda_from_h(s_["h"])
    threads = o.max_threads()
    rows = int(max(1, min(g.ny, round(7e7 * threads * 6.0 / (g.nx * float(M))))))

    def sample(nrows):
        k, j0 = g.nz // 2, max(0, g.ny // 2 - nrows // 2)
        t0 = time.perf_counter()
        Y = o.step12_box(g, lam, s_["pos"], s_["nrm"], s_["area"], j0, j0 + nrows, k, k + 1, threads=threads)
        dt = time.perf_counter() - t0
        assert np.isfinite(Y).all()
        return nrows * g.nx, dt
    for _ in range(args.warmup):
        sample(max(1, rows // 8))
    t_tot, n_tot = 0.0, 0
    for _ in range(args.steps):
        n, dt = sample(rows)
        n_tot += n
        t_tot += dt
    v = n_tot / t_tot
    text = (f"per step: Steps 1-2 (fp64 oracle port of the reference loop, OpenMP, {threads} threads) on {rows} rows "
            f"of one plane = {rows * g.nx} nodes x {M} sources; Step 3 excluded (reference sparse LU infeasible "
            f"beyond 64^3) -> upper bound on the CPU path's nodes/s; oracle/_ref (the reference's own source) is absent")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(1, args.steps), "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "same_config": False, "extrapolated": True,
            "config": {"workload": "Steps 1-2 ONLY on a bounded sample of: " + desc, "grid": [g.nx, g.ny, g.nz], "sources": M},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": text},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import shm3d

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000),
                   os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- workload (host) and contexts
    p, pos, nrm, area, desc = prepare(args.workload)
    M = len(area)
    N = p.N
    if world > 1:
        ids = [shm3d.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx = shm3d.Context(local, rank, world, ids[0])
    else:
        ctx = shm3d.Context(local)
    n_local = ctx.local_n(p)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    d_pos = torch.from_numpy(pos).to(dev)
    d_nrm = torch.from_numpy(nrm).to(dev)
    d_area = torch.from_numpy(area).to(dev)
    d_phi = torch.empty(n_local, dtype=torch.float32, device=dev)

    def step_device(flags=0):
        q = shm3d.Params.from_buffer_copy(p)
        q.flags |= flags | args.extra_flags
        return ctx.solve_device(q, d_pos.data_ptr(), d_nrm.data_ptr(), d_area.data_ptr(), d_phi.data_ptr(), M)

    # ---- N > 1: the slab-partitioned solve against the single-GPU solve of the same (256^3) problem, before timing
    dist_parity = None
    if world > 1 and not args.no_parity:
        dist_parity = dist_parity_check(ctx, dist, dev, rank, world, local)

    # ---- device-resident arm
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    torch.cuda.synchronize()
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    launches = 0
    stats = None
    prof_steps = []
    for _ in range(args.steps):
        # FLAG_PROFILE: CUDA event pairs on the solver's stream around the roofline kernels, recorded asynchronously and
        # resolved after the solve -- the per-kernel durations below are measured live inside the timed region
        stats = step_device(shm3d.FLAG_PROFILE)
        prof_steps.append(stats)
        launches += stats.kernel_launches
    e1.record(stream)
    torch.cuda.synchronize()
    stream.synchronize()
    wall = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = e0.elapsed_time(e1)
    t = torch.tensor([dev_ms, wall * 1e3, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, wall_ms, launches = float(tmax[0]), float(tmax[1]), int(tsum[2])
    else:
        wall_ms = wall * 1e3
    value = N * args.steps / (dev_ms * 1e-3)

    # ---- per-kernel roofline numbers from the timed steps (CUDA events on the solver's stream)
    class _Agg:
        pass
    prof = _Agg()
    for k in ("pcg_stencil_launches", "ms_pcg_stencil", "pairs_evaluated", "pairs_bruteforce", "ms_sum", "ms_pcg_vcycle",
              "ms_pcg_projector", "ms_pcg_update", "pcg_vcycles", "pcg_projector_applies", "ms_pcg", "cg_iters",
              "graph_replays"):
        setattr(prof, k, sum(getattr(st, k) for st in prof_steps))
    hbm_peak, sm_max, peak_src = peaks()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath) and world == 1:  # the capture is of the single-GPU launch
        try:
            traffic = json.load(open(tpath)).get(args.workload, {}).get("pcg_update_p_stencil_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline_pcg = None
    if prof.pcg_stencil_launches > 0 and prof.ms_pcg_stencil > 0:
        # fused PCG kernel: p <- (z - mean) + beta p ; q = K p ; p.q  -- reads z, p and writes p, q: 4 words = 16 B per node
        # (SURVEY.md section 8(d) counts the unfused pair as 2 + 3 words)
        bytes_per_launch = 16.0 * n_local
        ach = bytes_per_launch * prof.pcg_stencil_launches / (prof.ms_pcg_stencil * 1e-3) / 1e9
        roofline_pcg = {"kernel": "k_march<OpUpdateP> (PCG: p update + q = K p + p.q in one TMA-staged pass)", "bound": "hbm",
                        "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": traffic,
                        "peak_source": peak_src, "launches": int(prof.pcg_stencil_launches),
                        "avg_launch_us": 1e3 * prof.ms_pcg_stencil / prof.pcg_stencil_launches,
                        "algorithmic_bytes_per_launch": bytes_per_launch,
                        "note": "CUDA-event pairs around the launches of the first PCG iterations of every timed step (the "
                                "rest of each solve replays the same kernels from a CUDA graph)"}
    sm_mhz = (clocks or {}).get("sm_mhz") or sm_max
    sfu_peak = 16.0 * 148 * sm_mhz * 1e6  # MUFU ops/s at the clock observed under load
    mufu = 2.0 * prof.pairs_evaluated / (prof.ms_sum * 1e-3)
    roofline = {"kernel": "k_sum (Steps 1-2: heat-kernel summation + normalisation) -- the dominant kernel of the step",
                "bound": "sfu", "achieved": mufu / 1e9, "peak": sfu_peak / 1e9, "unit": "GMUFU-op/s", "frac": mufu / sfu_peak,
                "traffic": None, "share_of_step": prof.ms_sum / dev_ms if world == 1 else None,
                "peak_source": "16 MUFU/clk/SM x 148 SMs x SM clock sampled under load (B300_MICROARCH.md pipe table); "
                               "not HBM- or tensor-bound: 12 B/node written, no contraction",
                "algorithmic_ops_per_launch": 2.0 * prof.pairs_evaluated / args.steps,
                "pairs_evaluated": int(prof.pairs_evaluated // args.steps),
                "pairs_bruteforce": int(prof.pairs_bruteforce // args.steps),
                "bruteforce_equivalent_pairs_per_s": prof.pairs_bruteforce / (prof.ms_sum * 1e-3),
                "fp32_flops_per_s": 17.0 * prof.pairs_evaluated / (prof.ms_sum * 1e-3), "ms": prof.ms_sum / args.steps,
                "note": "rank 0 slab" if world > 1 else ""}
    # the whole Step 3 against the HBM roofline: (a) SURVEY 8(d)'s textbook count, 44 B per node and CG iteration;
    # (b) what one V(2,2)-preconditioned iteration has to move: fine level 8+12+4.5+8.5+12+12 (V-cycle) + 16 (p/q) + 24
    # (x/r) = 97 B/node, x 8/7 for the coarser levels' share of the V-cycle part
    its = max(1, prof.cg_iters)
    pcg_s = prof.ms_pcg * 1e-3
    b_vc = (57.0 * 8.0 / 7.0 + 40.0) * n_local
    pcg_whole = {"ms_per_iteration": prof.ms_pcg / its, "iterations": int(prof.cg_iters // args.steps),
                 "survey_44B": {"GBps": 44.0 * n_local * its / pcg_s / 1e9, "frac": 44.0 * n_local * its / pcg_s / 1e9 / hbm_peak},
                 "vcycle_inclusive": {"bytes_per_node_iteration": b_vc / n_local, "GBps": b_vc * its / pcg_s / 1e9,
                                      "frac": b_vc * its / pcg_s / 1e9 / hbm_peak},
                 "graph_replays_per_step": prof.graph_replays / args.steps}

    # ---- end-to-end arm: the public, reference-facing call with HOST inputs -- computeDistance(V, F, options) runs the
    # mesh-level host half (rows a4-a6: areas, normals, barycentres, mean edge length, grid), uploads the sources, solves,
    # and downloads the double field into the solver-owned page-locked buffer -- all inside the timed region
    solver = shm3d.SignedHeatGridSolver(context=ctx)
    Vw, Fw, n_w, _ = make_workload(args.workload)
    hc = int(round(np.log2(n_w / 16.0)))
    opts = shm3d.SignedHeat3DOptions(hCoef=max(hc, 0))
    producible = (16 << max(hc, 0)) == n_w   # grids the reference itself produces: 16 * 2^h

    def step_host():
        if producible:
            return solver.computeDistance(Vw, Fw, opts)
        # 640^3 / 768^3 (weak-scaling fillers) are not 16 * 2^h: same host half, then the box resampled to n nodes per axis
        q, ps, nr, ar, _ = shm3d.prepare_mesh(Vw, Fw, hCoef=max(hc, 0))
        side = q.cell * (q.nx - 1)
        q.nx = q.ny = q.nz = n_w
        q.cell = side / (n_w - 1)
        return solver._finish(q, ps, nr, ar, opts)

    for _ in range(min(args.warmup, 2)):
        step_host()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for _ in range(args.steps):
        phi = step_host()
        checksum += float(phi[::max(1, len(phi) // 1024)].sum())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = N * args.steps / float(te[0])
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(7 * 8 * M * world),
           "d2h_bytes_per_step": int(8 * N), "ms_per_step": 1e3 * float(te[0]) / args.steps,
           "api": "shm3d.SignedHeatGridSolver.computeDistance(V, F, options) -> shm3d_prepare_mesh + shm3d_solve: host "
                  "vertex/face arrays in, double field out; the a4-a6 host half (~40 ms at 1e5 faces), H2D and D2H are all "
                  "INSIDE the timed region" + ("" if producible else " (grid resampled to a non-16*2^h size via the class's "
                  "internal entry after the same host half)")}

    # ---- N > 1: the same grid on ONE GPU (rank 0 alone) -> strong-scaling speed-up of this run
    strong = None
    if world > 1 and not args.no_strong:
        strong = strong_block(p, d_pos, d_nrm, d_area, M, N, dev_ms / args.steps, local, rank, world, barrier)

    default_wl = args.workload == DEFAULT_BY_GPUS.get(world, "sphere512")
    scaling = "weak" if default_wl else "strong"
    scaling_note = ("default workloads hold ~512^3 nodes per GPU: 512^3 / 640^3 / 768^3 / 1024^3 at 1 / 2 / 4 / 8 GPUs "
                    "(1.00 / 0.98 / 0.84 / 1.00 x 512^3 per GPU), same 1e5-triangle sphere" if default_wl else
                    "fixed grid given by --workload, z-slabs split over the ranks")
    if rank == 0:
        cpu = cpu_port = None
        if world == 1 and not args.no_cpu:
            cpu_port = cpu_baseline_obj(p, M, pos, nrm, area)  # the oracle port, OpenMP over all host threads
            rs = reference_build_sample(args.workload)          # the reference's own source, single-threaded
            cpu = ({"value": rs[0] / rs[1], "unit": UNIT, "cores": 1, "kind": "reference", "sample": rs[2]}
                   if rs is not None else cpu_port)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc, "grid": [p.nx, p.ny, p.nz], "sources": M, "l2": "working set "
                           f"{4 * N / 1e6:.0f} MB per grid vector >> 126 MB L2 (no explicit flush needed)"
                           if 4 * N > 4 * 126e6 else "L2 flushed implicitly by the 3-component Y write of Steps 1-2",
                           "parallelism": f"z-slab x{world}" if world > 1 else "single GPU",
                           "nodes_per_gpu": N // world, "scaling_note": scaling_note,
                           "cull_tau": 10.0, "timing": "CUDA events on the solver's stream, max over ranks"},
                "wall_ms_per_step": wall_ms / args.steps, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "roofline": roofline, "roofline_pcg": roofline_pcg, "pcg_whole": pcg_whole,
                "dist_parity": dist_parity, "strong": strong, "cpu_baseline": cpu,
                "cpu_baseline_port_all_threads": cpu_port,
                "stages_ms": {"h2d+cluster": stats.ms_h2d, "sum(step1-2)": stats.ms_sum, "rhs": stats.ms_rhs,
                              "constraints+factor(host, overlapped with sum)": stats.ms_constraints,
                              "pcg": stats.ms_pcg, "shift": stats.ms_shift, "total": stats.ms_total},
                "pcg": {"iters": int(stats.cg_iters), "rel_residual": stats.cg_rel_residual,
                        "m_constraints": int(stats.m_constraints),
                        "tail_ops": int(stats.tail_ops), "graph_replays": int(stats.graph_replays),
                        "profiled_ms_per_iteration": {
                            "p_update+stencil": prof.ms_pcg_stencil / max(1, prof.pcg_stencil_launches),
                            "vcycle": prof.ms_pcg_vcycle / max(1, prof.pcg_vcycles),
                            "projector(2 applications)": 2 * prof.ms_pcg_projector / max(1, prof.pcg_projector_applies),
                            "x/r update": prof.ms_pcg_update / max(1, prof.pcg_stencil_launches),
                            "whole iteration (pcg ms / iterations, graph replay)": prof.ms_pcg / its}},
                "checksum": checksum}
        if world == 1 and not args.no_cpu:
            line["consumer_n3"] = consumer_leg(min(p.nx, 512))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


def dist_parity_check(ctx, dist, dev, rank, world, local):
    """sphere256 solved on the `world` z-slabs and by rank 0 alone: the slabs are gathered on the ranks and compared."""
    import torch
    import shm3d
    if 256 % world:
        return {"error": f"256 planes do not split evenly over {world} ranks"}
    p, pos, nrm, area, _ = prepare("sphere256")
    phi_slab, st = ctx.solve(p, pos, nrm, area)
    t = torch.from_numpy(np.ascontiguousarray(phi_slab, dtype=np.float64)).to(dev)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    out = None
    if rank == 0:
        solo = shm3d.Context(local)
        full, st1 = solo.solve(p, pos, nrm, area)
        solo.close()
        joined = torch.cat(parts).cpu().numpy()
        err = float(np.linalg.norm(joined - full) / np.linalg.norm(full))
        out = {"workload": "sphere256 (1e5-triangle sphere, 256^3)", "rel_l2_vs_single_gpu": err,
               "max_abs_diff": float(np.abs(joined - full).max()), "iters_dist": int(st.cg_iters),
               "iters_single": int(st1.cg_iters), "ranks": world}
    dist.barrier()
    return out


def strong_block(p, d_pos, d_nrm, d_area, M, N, ms_per_step_dist, local, rank, world, barrier):
    """The SAME grid solved by rank 0 alone (one warm-up + one timed step, CUDA events on its stream)."""
    import torch
    import shm3d
    out = None
    if rank == 0:
        try:
            solo = shm3d.Context(local)
            d_phi = torch.empty(N, dtype=torch.float32, device=d_pos.device)
            st = None
            ms = []
            for i in range(2):
                stream = torch.cuda.ExternalStream(solo.stream, device=d_pos.device)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                st = solo.solve_device(shm3d.Params.from_buffer_copy(p), d_pos.data_ptr(), d_nrm.data_ptr(),
                                       d_area.data_ptr(), d_phi.data_ptr(), M)
                e1.record(stream)
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            solo.close()
            out = {"grid": [p.nx, p.ny, p.nz], "ms_per_step_1gpu": ms[-1], "ms_per_step_this_run": ms_per_step_dist,
                   "speedup": ms[-1] / ms_per_step_dist, "gpus": world, "iters_1gpu": int(st.cg_iters),
                   "stages_ms_1gpu": {"sum": st.ms_sum, "pcg": st.ms_pcg, "constraints(host)": st.ms_constraints}}
        except Exception as e:  # e.g. the grid does not fit one GPU
            out = {"error": repr(e)[:300]}
    barrier()
    return out


def consumer_leg(n):
    """Row N3 (isosurface of the device-resident field) measured in a child process after the solver's own numbers are
    in: a separate context, and nothing it does can take the main line down.  Not part of `value` / `e2e`."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_consumer.py"), "--n", str(n)],
                           capture_output=True, text=True, timeout=180)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": (r.stderr or r.stdout or "no output")[-400:]}
    except Exception as e:
        return {"error": repr(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: ~512^3 nodes per GPU (sphere512 / 640 / 768 / 1024 at 1 / 2 / 4 / 8 GPUs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--extra-flags", type=int, default=0, help="diagnostics: SHM3D_FLAG_* bits OR-ed into every solve of the "
                    "device-resident arm (e.g. 2048 = no round-robin z-chunks for Steps 1-2 on slab contexts)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the dist_parity block")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling block")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = DEFAULT_BY_GPUS.get(args.gpus, "sphere512")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
