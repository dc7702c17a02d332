"""The C restatement (oracle/lbm_oracle.c: CPU baseline + large-grid checker) must agree
bit-for-bit with the numpy restatement that the golden table pins."""
import numpy as np
import pytest

import oracle.lbm_oracle as O
from oracle.c_oracle import COracle, num_threads
from oracle.lattices import ALL
from conftest import random_populations


@pytest.mark.parametrize("mk", ALL, ids=[m.__name__ for m in ALL])
def test_c_oracle_is_bit_identical(mk):
    q = mk()
    ny, nx = 7, 9
    f = random_populations(q, nx, ny)
    rng = np.random.default_rng(3)
    field = (1e-5 * rng.standard_normal((ny, nx)), 1e-5 * rng.standard_normal((ny, nx)))
    cms = [O.SRT(0.8), O.SRT(0.8, (1e-5, -2e-5)), O.TRT(0.8, 1.1, (1e-5, 2e-5)), O.TRT(0.6, 0.9, field),
           O.MRT(q, [0.8, 0.9, 1.1, 1.3][:max(q.N, 2)], (1e-5, 2e-5)), O.MRT(q, 0.7)]
    bc_sets = [[], [O.BounceBack("S", (1, nx), (1, ny)), O.MovingWall("N", (1, nx), (1, ny), [0.01, 0.001])],
               [O.BounceBack("E", (1, nx), (2, ny)), O.BounceBack("W", (1, nx), (1, ny - 1)), O.BounceBack("N", (2, nx), (1, ny))]]
    for cm in cms:
        for bcs in bc_sets:
            a = f.copy()
            for _ in range(4):
                a, ac = O.step(cm, q, bcs, a)
            b, bc_ = COracle(q, cm, bcs).steps(f, 4)
            assert np.array_equal(a, b) and np.array_equal(ac, bc_)


def test_c_oracle_tiny_grids_and_threads():
    q = O.L.D2Q37()
    for shape in [(1, 1), (2, 3), (3, 1)]:
        f = random_populations(q, shape[1], shape[0])
        assert np.array_equal(O.stream(q, f), COracle(q, O.SRT(1.0)).stream(f))
    assert num_threads() >= 1
    f = random_populations(O.L.D2Q9(), 64, 64)
    one, _ = COracle(O.L.D2Q9(), O.TRT(0.8, 1.0), threads=1).steps(f, 3)
    many, _ = COracle(O.L.D2Q9(), O.TRT(0.8, 1.0), threads=4).steps(f, 3)
    assert np.array_equal(one, many)
