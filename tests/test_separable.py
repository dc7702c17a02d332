"""The skeleton decomposition that turns sampled analytic fields into the separable form lbm_reduce_errors takes
(lbm/separable.py; the Julia binding uses the same algorithm because it only has the package's pointwise functions to
call): exact to round-off on every shipped problem's fields, and it refuses fields of rank > 2."""
import numpy as np
import pytest

import lbm
from lbm.separable import cross_decompose, separable_from_fields

Q = lbm.D2Q9()
PROBLEMS = {
    "tgv": lambda: lbm.TGV(Q, 0.8, 2),
    "tgv-rect": lambda: lbm.TGV(Q, 0.8, 1, 24, 10),
    "shear-static": lambda: lbm.DecayingShearFlow(1 / 6, 2, static=True),
    "shear-decaying": lambda: lbm.DecayingShearFlow(1 / 6, 2, static=False),
    "poiseuille": lambda: lbm.PoiseuilleFlow(1 / 6, 2),
    "couette": lambda: lbm.CouetteFlow(1 / 6, 2),
    "thermal-diffusion": lambda: lbm.LinearizedThermalDiffusion(0.1, 0.2, 2),
    "transverse-shear": lambda: lbm.LinearizedTransverseShearWave(0.1, 0.2, 2),
}


def _fields(pr, t):
    X, Y = pr.grid(0, pr.NY)
    ux, uy = pr.velocity(X, Y, t)
    (sxx, sxy), (syx, syy) = pr.deviatoric_tensor(Q, X, Y, t)
    return [np.asarray(a, dtype=np.float64) * np.ones_like(X) for a in
            (pr.density(Q, X, Y, t), ux, uy, pr.pressure(Q, X, Y, t), sxx, sxy, syx, syy)]


@pytest.mark.parametrize("name", list(PROBLEMS))
@pytest.mark.parametrize("t", [0.0, 0.37])
def test_every_shipped_problem_decomposes_exactly(name, t):
    pr = PROBLEMS[name]()
    fields = _fields(pr, t)
    sep = separable_from_fields(fields)
    assert sep is not None, name
    for E, (c0, terms) in zip(fields, sep):
        assert len(terms) <= 2
        R = c0 + sum(a * np.outer(X, Y) for a, X, Y in terms) if terms else np.zeros_like(E) + c0
        assert np.abs(R - E).max() <= 1e-13 * max(np.abs(E).max(), 1e-300), name
    # ... and agrees with the problem's own analytic separable form
    own = pr.expected_separable(Q, t, 0, pr.NY)
    if own is not None:
        for E, (c0, terms) in zip(fields, own):
            R = np.zeros_like(E) + c0
            for a, X, Y in terms:
                R = R + a * np.outer(np.ones(pr.NX) if X is None else X, np.ones(pr.NY) if Y is None else Y)
            assert np.abs(R - E).max() <= 1e-12 * max(np.abs(E).max(), 1e-300), name


def test_rank_three_fields_are_refused():
    x = np.linspace(0, 1, 12)[:, None]
    y = np.linspace(0, 1, 9)[None, :]
    assert cross_decompose(np.sin(3 * x) * y + np.cos(x) * y ** 2 + x ** 2 * np.exp(y)) is None
    assert cross_decompose(np.zeros((4, 5))) == []
    one = cross_decompose(np.full((4, 5), 2.5))
    assert len(one) == 1
