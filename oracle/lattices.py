"""ORACLE (test infrastructure, NOT product code) -- lattice constant tables.

CPU restatement of the seven velocity sets of LatticeBoltzmann.jl.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import anything under ``oracle/``.

Each table follows the reference file cited next to it (paths relative to
/root/reference).  Population order == column order of ``abscissae`` there;
indices here are 0-based (Julia index - 1).
"""
import math

import numpy as np


class Lattice:
    """abscissae/weights/speed_of_sound_squared of one quadrature.

    ``css`` is the reference's ``q.speed_of_sound_squared`` (= 1/c_s^2).
    ``order`` follows ``order(q)``; ``N = order // 2`` Hermite orders
    (src/collision_models/mrt.jl:68, velocity_distribution_function/hermite.jl:11).
    ``eq_order`` is the truncation order of the collision equilibrium
    (maxwell_boltzmann_equilibrium.jl:43-66 and
    velocity_distribution_function/quadratures.jl).
    """

    def __init__(self, name, cx, cy, w, css, order, eq_order, opposite_rule):
        self.name = name
        self.cx = np.asarray(cx, dtype=np.int64)
        self.cy = np.asarray(cy, dtype=np.int64)
        self.w = np.asarray(w, dtype=np.float64)
        self.css = float(css)
        self.order = int(order)
        self.N = self.order // 2
        self.eq_order = int(eq_order)
        self.Q = len(self.w)
        assert len(self.cx) == self.Q and len(self.cy) == self.Q
        self.opp = np.array([opposite_rule(i + 1) - 1 for i in range(self.Q)], dtype=np.int64)
        self.h = int(max(np.abs(self.cx).max(), np.abs(self.cy).max()))

    def __repr__(self):
        return self.name


def _opposite_generic(idx):
    # src/quadratures.jl:11-19 (1-based)
    if idx == 1:
        return 1
    if idx % 2 == 0:
        return idx + 1
    return idx - 1


def _opposite_d2q4(idx):
    # src/quadratures/D2Q4.jl:26-31
    return idx + 2 if idx <= 2 else idx - 2


def _opposite_d2q5(idx):
    # src/quadratures/D2Q5.jl:34-42
    if idx == 1:
        return 1
    return idx + 2 if idx <= 3 else idx - 2


def _opposite_d2q9(idx):
    # src/quadratures/D2Q9.jl:30-38
    if idx == 1:
        return 1
    return idx + 4 if idx <= 5 else idx - 4


def D2Q4():
    # src/quadratures/D2Q4.jl:16-25
    return Lattice("D2Q4", [1, 0, -1, 0], [0, 1, 0, -1], [1 / 4] * 4, 2.0, 3, 1, _opposite_d2q4)


def D2Q5():
    # src/quadratures/D2Q5.jl:17-33
    return Lattice("D2Q5", [0, 1, 0, -1, 0], [0, 0, 1, 0, -1],
                   [4 / 6, 1 / 12, 1 / 12, 1 / 12, 1 / 12], 6.0, 3, 1, _opposite_d2q5)


def D2Q9():
    # src/quadratures/D2Q9.jl:20-29
    return Lattice("D2Q9", [0, -1, -1, -1, 0, 1, 1, 1, 0], [0, 1, 0, -1, -1, -1, 0, 1, 1],
                   [4 / 9, 1 / 36, 1 / 9, 1 / 36, 1 / 9, 1 / 36, 1 / 9, 1 / 36, 1 / 9],
                   3.0, 5, 2, _opposite_d2q9)


def D2Q13():
    # src/quadratures/D2Q13.jl:10-29
    w0, w1, w2, w3 = 3 / 8, 1 / 12, 1 / 16, 1 / 96
    return Lattice("D2Q13",
                   [0, 1, -1, 0, 0, 1, -1, 1, -1, 2, -2, 0, 0],
                   [0, 0, 0, 1, -1, 1, -1, -1, 1, 0, 0, -2, 2],
                   [w0] + [w1] * 4 + [w2] * 4 + [w3] * 4, 2.0, 5, 2, _opposite_generic)


def D2Q17():
    # src/quadratures/D2Q17.jl:21-57
    sq = math.sqrt(193)
    w0 = (575 + 193 * sq) / 8100
    w1 = (3355 - 91 * sq) / 18000
    w2 = (655 + 17 * sq) / 27000
    w3 = (685 - 49 * sq) / 54000
    w4 = (1445 - 101 * sq) / 162000
    return Lattice("D2Q17",
                   [0, 1, -1, 0, 0, 1, -1, 1, -1, 2, -2, 2, -2, 3, -3, 0, 0],
                   [0, 0, 0, 1, -1, 1, -1, -1, 1, 2, -2, -2, 2, 0, 0, 3, -3],
                   [w0] + [w1] * 4 + [w2] * 4 + [w3] * 4 + [w4] * 4,
                   (125 + 5 * math.sqrt(193)) / 72, 7, 3, _opposite_generic)


def D2Q21():
    # src/quadratures/D2Q21.jl:15-71 -- 25 stored populations, the last four
    # ((+-3,+-3)) carry weight 0 (second assignment block, :26-33).
    w0, w1, w2, w3, w4, w5, w6 = 91 / 324, 1 / 12, 2 / 27, 7 / 360, 1 / 432, 1 / 1620, 0.0
    return Lattice("D2Q21",
                   [0, 1, -1, 0, 0, 1, -1, 1, -1, 2, -2, 0, 0, 2, -2, 2, -2, 3, -3, 0, 0, 3, -3, -3, 3],
                   [0, 0, 0, 1, -1, 1, -1, -1, 1, 0, 0, 2, -2, 2, -2, -2, 2, 0, 0, 3, -3, 3, -3, 3, -3],
                   [w0] + [w1] * 4 + [w2] * 4 + [w3] * 4 + [w4] * 4 + [w5] * 4 + [w6] * 4,
                   3 / 2, 7, 3, _opposite_generic)


def D2Q37():
    # src/quadratures/D2Q37.jl:11-74
    g1 = 0.23315066913235250228650
    g2 = 0.10730609154221900241246
    g3 = 0.05766785988879488203006
    g4 = 0.01420821615845075026469
    g5 = 0.00535304900051377523273
    g6 = 0.00101193759267357547541
    g7 = 0.00024530102775771734547
    g8 = 0.00028341425299419821740
    r = 1.19697977039307435897239
    return Lattice("D2Q37",
                   [0, 1, -1, 0, 0, 1, -1, 1, -1, 2, -2, 0, 0, 2, -2, -2, 2, 1, -1, 1, -1,
                    2, -2, 2, -2, 3, -3, 0, 0, 3, -3, 3, -3, 1, -1, -1, 1],
                   [0, 0, 0, 1, -1, 1, -1, -1, 1, 0, 0, 2, -2, 1, -1, 1, -1, 2, -2, -2, 2,
                    2, -2, -2, 2, 0, 0, 3, -3, 1, -1, -1, 1, 3, -3, 3, -3],
                   [g1] + [g2] * 4 + [g3] * 4 + [g4] * 4 + [g5] * 8 + [g6] * 4 + [g7] * 4 + [g8] * 8,
                   r * r, 9, 4, _opposite_generic)


ALL = (D2Q4, D2Q5, D2Q9, D2Q13, D2Q17, D2Q21, D2Q37)
BY_NAME = {f.__name__: f for f in ALL}
