"""How evenly do z-slabs share Steps 1-2?  The 8 slabs of the 1024^3 bench grid (and the 4 of 768^3, 2 of 640^3) evaluated
one after the other on ONE GPU as stand-alone sub-grids (same nodes, same sources): k_sum time and pairs per slab."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "signed-heat-3d_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, shm3d, bench
ctx = shm3d.Context(0)
for wl, world in (("sphere1024", 8), ("sphere768", 4), ("sphere640", 2)):
    p, pos, nrm, area, _ = bench.prepare(wl)
    nz = p.nz
    out = []
    for r in range(world):
        k0, k1 = shm3d.slab_range(r, world, nz)
        q = shm3d.Params.from_buffer_copy(p)
        q.nz = k1 - k0
        q.bbox_min[2] = p.bbox_min[2] + k0 * p.cell
        best = None
        for _ in range(2):
            Y, st = ctx.step12(q, pos, nrm, area)
            best = st.ms_sum if best is None else min(best, st.ms_sum)
        del Y
        out.append((round(best, 1), round(st.pairs_evaluated / 1e9, 1)))
    ms = [o[0] for o in out]
    print(json.dumps({"workload": wl, "slabs": world, "ms_sum_per_slab": ms, "Gpairs_per_slab": [o[1] for o in out],
                      "max_over_mean": round(max(ms) / (sum(ms) / len(ms)), 3)}), flush=True)
ctx.close()
