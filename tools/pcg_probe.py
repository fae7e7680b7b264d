"""Step-3 timing probe: one workload solved with the launch-reduction switches on and off.
    python tools/pcg_probe.py [sphere512] [reps]
Prints per variant: PCG ms, iterations, ms/iteration, kernel launches, tail ops, graph replays, and (profiled run) the
V-cycle / projector / stencil / update split of the kernel-by-kernel iterations."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "signed-heat-3d_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402
import shm3d  # noqa: E402
if os.environ.get("SHM3D_LIB_PATH"):  # experiment builds (e.g. -DSHM3D_TUNING_KNOBS)
    shm3d.LIB_PATH = os.path.abspath(os.environ["SHM3D_LIB_PATH"])
import bench  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "sphere512"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    p, pos, nrm, area, desc = bench.prepare(wl)
    ctx = shm3d.Context(0)
    dev = torch.device("cuda", 0)
    d_pos, d_nrm, d_area = (torch.from_numpy(a).to(dev) for a in (pos, nrm, area))
    d_phi = torch.empty(p.N, dtype=torch.float32, device=dev)
    variants = [("default (graph replay + PDL sweeps)", 0), ("no graph", shm3d.FLAG_NO_GRAPH),
                ("+ tail program", shm3d.FLAG_TAIL_PROGRAM), ("no PDL", shm3d.FLAG_NO_PDL),
                ("none (round-1 launch structure)", shm3d.FLAG_NO_GRAPH | shm3d.FLAG_NO_PDL),
                ("default + profile", shm3d.FLAG_PROFILE)]
    if os.environ.get("PCG_PROBE_QUICK"):
        variants = variants[:1]
    for name, fl in variants:
        best = None
        for _ in range(reps):
            q = shm3d.Params.from_buffer_copy(p)
            q.flags |= fl
            st = ctx.solve_device(q, d_pos.data_ptr(), d_nrm.data_ptr(), d_area.data_ptr(), d_phi.data_ptr(), len(area))
            if best is None or st.ms_pcg < best.ms_pcg:
                best = st
        st = best
        line = {"variant": name, "workload": wl, "ms_pcg": round(st.ms_pcg, 2), "iters": st.cg_iters,
                "ms_per_iter": round(st.ms_pcg / max(1, st.cg_iters), 3), "launches": st.kernel_launches,
                "tail_ops": st.tail_ops, "graph_replays": st.graph_replays, "ms_sum": round(st.ms_sum, 1),
                "ms_total": round(st.ms_total, 1), "ms_constraints_host": round(st.ms_constraints, 1)}
        if fl & shm3d.FLAG_PROFILE:
            line["eager_ms"] = {"vcycle": round(st.ms_pcg_vcycle / max(1, st.pcg_vcycles), 3),
                                "projector_x2": round(2 * st.ms_pcg_projector / max(1, st.pcg_projector_applies), 3),
                                "stencil": round(st.ms_pcg_stencil / max(1, st.pcg_stencil_launches), 3),
                                "update": round(st.ms_pcg_update / max(1, st.pcg_stencil_launches), 3)}
        print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
