#!/usr/bin/env python
"""Array-level operators on HOST arrays (the reference's collide!(cm, q, f_in, f_out) called on plain arrays,
src/collision_models.jl:19) and the population copies behind them: pageable numpy arrays (staged through the library's
page-locked chunk pipeline on several host threads) against page-locked arrays (lbm.pinned_empty) and the PCIe floor.

    python tools/bench_array_ops.py [--sizes 512,2048,4096] [--reps 5]

One JSON line per size."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "latticeboltzmann.jl_b200"))
import lbm  # noqa: E402
from lbm import _abi  # noqa: E402


def best(fn, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="512,2048,4096")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--pcie-gbs", type=float, default=55.0, help="one-direction PCIe rate used for the floor")
    a = ap.parse_args()
    q = lbm.D2Q9()
    cm = lbm.TRT(0.8, 1.1)
    for n in [int(v) for v in a.sizes.split(",")]:
        nbytes = n * n * q.Q * 8
        rng = np.random.default_rng(0)
        f_page = np.asfortranarray(np.broadcast_to(q.weights, (n, n, q.Q)) * (1 + 1e-3 * rng.standard_normal((n, n, 1))))
        out_page = np.empty_like(f_page, order="F")
        f_pin, out_pin = _abi.pinned_empty((n, n, q.Q)), _abi.pinned_empty((n, n, q.Q))
        f_pin[...] = f_page
        out_page[...] = 0.0  # touch the pages
        row = {"n": n, "bytes_each_way": nbytes, "pcie_floor_ms": round(2 * nbytes / (a.pcie_gbs * 1e9) * 1e3, 3)}
        lbm.collide_(cm, q, f_page, out_page)
        ref = out_page.copy()
        row["collide_pageable_ms"] = round(best(lambda: lbm.collide_(cm, q, f_page, out_page), a.reps) * 1e3, 3)
        lbm.collide_(cm, q, f_pin, out_pin)
        row["collide_pinned_ms"] = round(best(lambda: lbm.collide_(cm, q, f_pin, out_pin), a.reps) * 1e3, 3)
        assert np.array_equal(out_pin, ref)
        lbm.stream_(q, f_page, out_page)
        row["stream_pageable_ms"] = round(best(lambda: lbm.stream_(q, f_page, out_page), a.reps) * 1e3, 3)
        with _abi.Context(n, n, "D2Q9", _abi.TRT, [0.8, 1.1]) as c:
            c.upload_f(f_page); c.download_f(out_page)
            row["upload_pageable_gbs"] = round(nbytes / best(lambda: c.upload_f(f_page), a.reps) / 1e9, 1)
            row["download_pageable_gbs"] = round(nbytes / best(lambda: c.download_f(out_page), a.reps) / 1e9, 1)
            row["upload_pinned_gbs"] = round(nbytes / best(lambda: c.upload_f(f_pin), a.reps) / 1e9, 1)
            row["download_pinned_gbs"] = round(nbytes / best(lambda: c.download_f(out_pin), a.reps) / 1e9, 1)
            assert np.array_equal(out_pin, out_page)
        with _abi.Context(n, n, "D2Q9", _abi.TRT, [0.8, 1.1], dtype=_abi.F32, arith=_abi.ARITH_FAST) as c:
            c.upload_f(f_page); c.download_f(out_page)
            row["f32_upload_pageable_gbs"] = round(nbytes / best(lambda: c.upload_f(f_page), a.reps) / 1e9, 1)
            row["f32_download_pageable_gbs"] = round(nbytes / best(lambda: c.download_f(out_page), a.reps) / 1e9, 1)
        row["collide_pageable_over_floor"] = round(row["collide_pageable_ms"] / row["pcie_floor_ms"], 2)
        row["collide_pinned_over_floor"] = round(row["collide_pinned_ms"] / row["pcie_floor_ms"], 2)
        row["copy_threads"] = os.environ.get("LBM_COPY_THREADS", "default")
        print(json.dumps(row), flush=True)
        del f_pin, out_pin
    lbm.clear_scratch_contexts()


if __name__ == "__main__":
    main()
