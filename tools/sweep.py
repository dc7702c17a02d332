#!/usr/bin/env python
"""Kernel tuning sweep (needs the TUNE=1 build: make -C latticeboltzmann.jl_b200/csrc TUNE=1).

For every <lattice, collision, dtype, arith> times the fused pull kernel for each register-budget
variant (variant v = __launch_bounds__(256, v); v + 100 = same with 128-thread CTAs) and prints one
JSON line per measurement: MLUPS, GB/s (B_alg = 2 Q sizeof(T)), fraction of the measured HBM peak.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "latticeboltzmann.jl_b200"))
import lbm  # noqa: E402
from lbm import _abi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lattices", default="D2Q9,D2Q13,D2Q17,D2Q21,D2Q37")
    ap.add_argument("--models", default="SRT,TRT,MRT")
    ap.add_argument("--dtypes", default="f64,f32")
    ap.add_argument("--ariths", default="fast,exact")
    ap.add_argument("--variants", default="0,1,2,3,4,5,6,102,103,104")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--n", type=int, default=0, help="grid edge (0: 4096 for Q <= 13, 2048 otherwise)")
    a = ap.parse_args()
    peak = 6548.5
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    for lat in a.lattices.split(","):
        q = getattr(lbm.Quadratures, lat)
        n = a.n or (4096 if q.Q <= 13 else 2048)
        x = (np.arange(n) + 0.5) * (2 * np.pi / n)
        pert = 1e-3 * np.sin(x)[:, None] * np.cos(x)[None, :]
        f0 = np.empty((n, n, q.Q), order="F")
        for i in range(q.Q):
            f0[:, :, i] = q.weights[i] * (1 + pert * (1 + 0.1 * i))
        for dtype in a.dtypes.split(","):
            es = 8 if dtype == "f64" else 4
            for model in a.models.split(","):
                code = {"SRT": _abi.SRT, "TRT": _abi.TRT, "MRT": _abi.MRT}[model]
                taus = {"SRT": [0.8], "TRT": [0.8, 1.3333333333333333], "MRT": [0.8, 0.8, 0.8, 0.8]}[model]
                for arith in a.ariths.split(","):
                    with _abi.Context(n, n, lat, code, taus, [], dtype=_abi.F64 if dtype == "f64" else _abi.F32,
                                      arith=_abi.ARITH_FAST if arith == "fast" else _abi.ARITH_EXACT) as c:
                        c.upload_f(f0)
                        c.step(0, 2)
                        for v in [int(s) for s in a.variants.split(",")]:
                            c.set_option("variant", v)
                            c.step(0, 3)
                            c.sync()
                            c.timer_start()
                            c.step(0, a.steps)
                            ms = c.timer_stop()
                            mlups = n * n * a.steps / (ms * 1e-3) / 1e6
                            gbs = mlups * 1e6 * 2 * q.Q * es / 1e9
                            print(json.dumps(dict(lattice=lat, model=model, dtype=dtype, arith=arith, variant=v, n=n,
                                                  ms_per_step=ms / a.steps, mlups=round(mlups, 1), gbs=round(gbs, 1),
                                                  frac=round(gbs / peak, 4))), flush=True)


if __name__ == "__main__":
    main()
