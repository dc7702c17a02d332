"""Pins the oracle to the reference's only numeric golden vectors: the printed result table of
examples/notebooks/trt_magic_parameter.ipynb (cell 3; 41 visible rows of the 902,500-row
DataFrame), extracted by tests/golden/extract_trt_magic_parameter.py.

This one table fixes D2Q9 equilibrium, TRT with distinct relaxation times, the tau_a force shift,
pull streaming, North+South bounce-back, the `simulate` step-count convention (0:n_steps, next!(t+1)),
the TrackHydrodynamicErrors u/p/sigma formulas, unit scaling and VelocityConvergenceStoppingCriteria.
"""
import json
import os

import numpy as np
import pytest

import oracle.lbm_oracle as O

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "trt_magic_parameter.json")))
ROWS = G["rows"]


def solve(tau_s, tau_a):
    q = O.L.D2Q9()
    problem = O.PoiseuilleFlow((tau_s - 0.5) / q.css, 1)
    assert (problem.NX, problem.NY) == (3, 5)
    n_steps = round(100.0 / problem.delta_t())
    assert n_steps == 5000
    pm = O.TrackHydrodynamicErrors(problem, False, n_steps, O.VelocityConvergenceStoppingCriteria(1e-7, problem))
    cm = O.TRT(tau_s, tau_a, O._problem_force(problem))
    m = O.simulate(problem, q, pm=pm, should_process=False, strategy="ZeroVelocityInitialCondition", t_end=100.0,
                   collision=cm)
    return m.pm.df[-1]


def close_to_printed(value, printed, digits=6):
    """`printed` shows `digits` significant digits of `value`."""
    if np.isinf(printed):
        return np.isinf(value)
    ulp = 10.0 ** (np.floor(np.log10(abs(printed))) - (digits - 1))
    return abs(value - printed) <= 0.51 * ulp


@pytest.mark.parametrize("row", ROWS, ids=[f"row{r['row']}" for r in ROWS])
def test_golden_row(row):
    e = solve(row["tau_s"], row["tau_a"])
    assert close_to_printed(e["error_u"], row["error_u"]), (e["error_u"], row["error_u"])
    assert close_to_printed(e["error_p"], row["error_p"]), (e["error_p"], row["error_p"])
    if "error_sxx" in row:
        assert close_to_printed(e["error_sxx"], row["error_sxx"])
        assert close_to_printed(e["error_sxy"], row["error_sxy"]), (e["error_sxy"], row["error_sxy"])


def test_stop_criterion_fires_for_large_tau():
    """tau = 10 rows converge long before n_steps: the recorded row is the stop-criterion one."""
    e = solve(10.0, 10.0)
    assert e["timestep"] < 5000 and e["timestep"] % 100 == 0
