"""Multi-GPU parity probe (run under torchrun, one rank per GPU): every rank solves the same problem twice -- as one
z-slab of a world-size partition (NCCL halo exchange / all-reduces) and alone on its own GPU -- and compares its slab.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_dist.py sphere:3 bunny_small:1
Prints one line per case from rank 0; exit code 1 if any slab differs by more than 1e-4 relative L2."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "signed-heat-3d_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import torch, torch.distributed as dist
import shm3d
from synth import fibonacci_sphere

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ids = [shm3d.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
dctx = shm3d.Context(local, rank, world, ids[0])
sctx = shm3d.Context(local)
bad = 0
for c in (sys.argv[1:] or ["sphere:3", "bunny_small:1", "bunny_small:0"]):
    fast = c.endswith(":fast")
    if fast:
        c = c[:-5]
    name, hc = c.split(":"); hc = int(hc)
    if name == "sphere":
        V, F = fibonacci_sphere(100000)
    else:
        z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")); fo = z["face_offsets"]; fv = z["face_vertices"]
        V = z["V"]; F = [fv[fo[i]:fo[i + 1]].tolist() for i in range(len(fo) - 1)]
    p, pos, nrm, area, h = shm3d.prepare_mesh(V, F, hCoef=hc)
    if fast:
        p.flags |= shm3d.FLAG_FAST  # fastIntegration: the z prefix sums chain over the ranks
    phi1, st1 = sctx.solve(p, pos, nrm, area)
    dist.barrier()
    t = time.time(); phid, std = dctx.solve(p, pos, nrm, area); dt = time.time() - t
    dist.barrier()
    t = time.time(); phid, std = dctx.solve(p, pos, nrm, area); dt = time.time() - t
    k0, k1 = dctx.slab(p.nz)
    ref = phi1.reshape(p.nz, p.ny, p.nx)[k0:k1].ravel()
    err = np.linalg.norm(phid - ref) / np.linalg.norm(ref)
    e = torch.tensor([err, dt], dtype=torch.float64, device="cuda")
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{name}{' fast' if fast else ''} {p.nx}^3 world={world}: max slab rel-L2 vs single GPU {float(e[0]):.3e}; its dist {std.cg_iters} / single {st1.cg_iters}; "
              f"wall dist {float(e[1])*1e3:.0f} ms (sum {std.ms_sum:.0f} constr {std.ms_constraints:.0f} pcg {std.ms_pcg:.0f}) / single {st1.ms_total:.0f} ms "
              f"(sum {st1.ms_sum:.0f} pcg {st1.ms_pcg:.0f})", flush=True)
    if float(e[0]) > 1e-4:
        bad = 1
# row N3 on slab contexts: every rank hands its device-resident float32 slab to shm3d_isosurface; the slabs are gathered on
# rank 0 over NVLink and contoured there -- the mesh must be the one a single GPU extracts from the whole field
V, F = fibonacci_sphere(100000)
p, pos, nrm, area, h = shm3d.prepare_mesh(V, F, hCoef=3)
dev = torch.device("cuda", local)
d_in = [torch.from_numpy(a).to(dev) for a in (pos, nrm, area)]
k0, k1 = dctx.slab(p.nz)
d_slab = torch.empty((k1 - k0) * p.ny * p.nx, dtype=torch.float32, device=dev)
dctx.solve_device(p, d_in[0].data_ptr(), d_in[1].data_ptr(), d_in[2].data_ptr(), d_slab.data_ptr(), len(area))
Vd, Td, std_ = dctx.isosurface(p, d_slab.data_ptr())
sl = dctx.slice(p, d_slab.data_ptr(), [p.bbox_min[0], p.bbox_min[1], p.bbox_min[2] + 0.37 * p.cell * (p.nz - 1)],
                [p.cell * 1.7, 0, 0], [0, p.cell * 1.3, 0], 40, 30)
parts = [torch.empty_like(d_slab) for _ in range(world)] if (p.nz % world == 0) else None
if parts is not None:
    dist.all_gather(parts, d_slab)
    if rank == 0:
        full = torch.cat(parts)
        Vs, Ts, sts = sctx.isosurface(p, full.data_ptr())
        sl1 = sctx.slice(p, full.data_ptr(), [p.bbox_min[0], p.bbox_min[1], p.bbox_min[2] + 0.37 * p.cell * (p.nz - 1)],
                         [p.cell * 1.7, 0, 0], [0, p.cell * 1.3, 0], 40, 30)
        same = np.array_equal(Vd.view(np.uint32), Vs.view(np.uint32)) and np.array_equal(Td, Ts)
        same_slice = np.array_equal(sl.view(np.uint32), sl1.view(np.uint32))
        print(f"isosurface on {world} slabs (gathered on rank 0): {len(Vd)} vertices, {len(Td)} triangles, identical to the "
              f"single-GPU mesh: {same}; slice identical: {same_slice}", flush=True)
        if not (same and same_slice and len(Vd) > 0):
            bad = 1
    else:
        assert len(Vd) == 0 and len(Td) == 0
dist.barrier()
dctx.close()
sctx.close()
dist.destroy_process_group()
sys.exit(bad)
