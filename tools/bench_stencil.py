"""Device time per launch of the PCG / V-cycle stencil operations: row-streaming kernels vs TMA-staged marching kernels,
against the HBM roofline (algorithmic bytes per node: p-update+stencil 16, sweep 12, residual 12, fused first sweeps 8).
    python tools/bench_stencil.py [n ...]        # default 512 256
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "signed-heat-3d_b200"))
import shm3d  # noqa: E402

OPS = {0: ("update_p_stencil", 16), 1: ("smooth", 12), 2: ("smooth_dot", 12), 3: ("residual", 12), 4: ("smooth01", 8)}


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [512, 256]
    peak = 6458.7
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except OSError:
        pass
    ctx = shm3d.Context(0)
    rng = np.random.default_rng(0)
    for n in sizes:
        shape = (n + 2, n, n)
        a = rng.standard_normal(shape, dtype=np.float32)
        b = rng.standard_normal(shape, dtype=np.float32)
        w = rng.standard_normal(shape, dtype=np.float32)
        for op, (name, bpn) in OPS.items():
            scal = {0: [0.013, 0.37], 1: [0.013, 0.81], 2: [0.013, 0.81], 3: [0.013], 4: [0.013, 0.55, 1.7]}[op]
            line = {"op": name, "n": n, "algorithmic_bytes": bpn * n ** 3}
            for tma in (0, 1):
                *_, ms = ctx.debug_stencil_op(op, (n, n, n), 0, n, a, b, w, scal, use_tma=bool(tma), reps=20)
                gbs = bpn * n ** 3 / (ms * 1e-3) / 1e9
                line["tma" if tma else "rows"] = {"us": round(ms * 1e3, 1), "GBps": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak, 3)}
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
