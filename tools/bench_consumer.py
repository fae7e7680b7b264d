"""Row N3 measurement (run as a child of bench.py, or by hand on a GPU box): isosurface extraction of a device-resident
float32 field -- the hand-off shm3d_solve_device leaves in HBM -- next to the reference consumer's own marching cubes
(polyscope's vendored MC.h, oracle/_ref/libshm_mc_ref.so) timed on one host core over the same field.

    python tools/bench_consumer.py --n 512 --steps 5 --warmup 2        -> one JSON line

The field is the signed distance of the unit sphere on the bench workload's box [-2,2]^3 (what the 512^3 sphere solve
converges to), synthesised on the device.  Roofline: the count pass reads every node once (4 B/node, algorithmic);
the emit passes only revisit the columns the surface crosses.  ms_device is the CUDA-event time of all four launches
including the host round trip for the two totals."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "signed-heat-3d_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--isoval", type=float, default=0.0)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    import numpy as np
    import torch
    import shm3d
    n = a.n
    dev = torch.device("cuda", 0)
    g = torch.linspace(-2.0, 2.0, n, device=dev, dtype=torch.float32)
    phi = torch.sqrt(g[None, None, :] ** 2 + g[None, :, None] ** 2 + g[:, None, None] ** 2) - 1.0   # [k, j, i]
    phi = phi.contiguous().view(-1)
    torch.cuda.synchronize()
    p = shm3d.Params()
    p.nx = p.ny = p.nz = n
    for ax in range(3):
        p.bbox_min[ax] = -2.0
    p.cell = 4.0 / (n - 1)
    ctx = shm3d.Context(0)
    bmin, bmax = (-2.0, -2.0, -2.0), (2.0, 2.0, 2.0)
    for _ in range(a.warmup):
        st = ctx.isosurface(p, phi.data_ptr(), a.isoval, bmin, bmax, fetch=False)
    ms = []
    for _ in range(a.steps):
        st = ctx.isosurface(p, phi.data_ptr(), a.isoval, bmin, bmax, fetch=False)
        ms.append(st.ms_device)
    t0 = time.perf_counter()
    V, T, st = ctx.isosurface(p, phi.data_ptr(), a.isoval, bmin, bmax)
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    N = n ** 3
    ms_med = float(np.median(ms))
    out = {"row": "N3 isosurface", "grid": [n, n, n], "isoval": a.isoval, "n_vertices": int(st.n_vertices),
           "n_triangles": int(st.n_triangles), "gpu_launches": int(st.gpu_launches), "ms_device": ms_med,
           "ms_device_all": ms, "nodes_per_s": N / (ms_med * 1e-3), "algorithmic_bytes": 4 * N,
           "achieved_GBps_algorithmic": 4 * N / (ms_med * 1e-3) / 1e9,
           "ms_with_fetch_to_host": e2e_ms, "d2h_bytes": int(V.nbytes + T.nbytes),
           "closed_manifold": bool(len(T) == 2 * len(V) - 4)}
    try:  # the same denominator bench.py uses for its roofline
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    out["roofline"] = {"bound": "hbm", "achieved": out["achieved_GBps_algorithmic"], "peak": peak, "unit": "GB/s",
                       "frac": out["achieved_GBps_algorithmic"] / peak, "traffic": None, "peak_source": src,
                       "note": "all four launches incl. the host round trip for the totals over 4 B/node (the count pass's "
                               "algorithmic bytes); the emit passes revisit only columns the surface crosses"}
    if not a.no_cpu:
        try:
            from oracle import reference_build as rb
            if rb.mc_available():
                h = phi.cpu().numpy()
                t0 = time.perf_counter()
                Vr, Tr = rb.isosurface(h, a.isoval, (n, n, n), bmin, bmax)
                cpu_s = time.perf_counter() - t0
                out["cpu_reference"] = {"kind": "reference", "what": "polyscope's vendored MC::marching_cube + vertex transform, "
                                        "compiled from the reference tree (oracle/_ref/libshm_mc_ref.so)", "cores": 1,
                                        "ms": 1e3 * cpu_s, "nodes_per_s": N / cpu_s}
                out["identical_to_reference"] = bool(np.array_equal(V, Vr) and np.array_equal(T, Tr))
            else:
                out["cpu_reference"] = None
        except Exception as e:  # the checker must not take the measurement down
            out["cpu_reference"] = {"error": repr(e)}
    ctx.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
