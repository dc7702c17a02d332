"""Quadrature types of the Julia API (src/quadratures.jl, src/quadratures/D2Q*.jl).

Field names follow the reference structs: `abscissae` (2 x Q integers), `weights`,
`speed_of_sound_squared`.  `opposite(q, idx)` is 1-based like the Julia function; `q.opp` is
the same table 0-based.  The constant tables are checked against the library's built-in
ones on first use (`lbm_lattice_info`).
"""
import math

import numpy as np

from . import _abi


class Quadrature:
    name = ""
    _order = 0

    def __init__(self, abscissae, weights, speed_of_sound_squared):
        self.abscissae = np.array(abscissae, dtype=np.int64)
        self.weights = np.array(weights, dtype=np.float64)
        self.speed_of_sound_squared = float(speed_of_sound_squared)
        self.Q = len(self.weights)
        self.opp = np.array([self._opposite(i + 1) - 1 for i in range(self.Q)])
        self.lattice_id = _abi.LATTICE_IDS[self.name]
        self._checked = False

    # src/quadratures.jl:11-19
    def _opposite(self, idx):
        if idx == 1:
            return 1
        return idx + 1 if idx % 2 == 0 else idx - 1

    def check_against_library(self):
        """The kernels use compile-time tables; refuse to run if the host's differ."""
        if self._checked:
            return
        info = _abi.lattice_info(self.lattice_id)
        ok = (info["Q"] == self.Q and np.array_equal(info["cx"], self.abscissae[0])
              and np.array_equal(info["cy"], self.abscissae[1]) and np.array_equal(info["w"], self.weights)
              and info["css"] == self.speed_of_sound_squared and np.array_equal(info["opposite"], self.opp))
        if not ok:
            raise _abi.LbmError(-1, f"{self.name}: host tables differ from the library's built-in quadrature")
        self._checked = True

    def __repr__(self):
        return f'"{self.name}"'

    def __str__(self):
        return self.name


class D2Q4(Quadrature):  # src/quadratures/D2Q4.jl
    name, _order = "D2Q4", 3

    def __init__(self):
        super().__init__([[1, 0, -1, 0], [0, 1, 0, -1]], [1 / 4] * 4, 2.0)

    def _opposite(self, idx):
        return idx + 2 if idx <= 2 else idx - 2


class D2Q5(Quadrature):  # src/quadratures/D2Q5.jl
    name, _order = "D2Q5", 3

    def __init__(self):
        super().__init__([[0, 1, 0, -1, 0], [0, 0, 1, 0, -1]], [4 / 6, 1 / 12, 1 / 12, 1 / 12, 1 / 12], 6.0)

    def _opposite(self, idx):
        if idx == 1:
            return 1
        return idx + 2 if idx <= 3 else idx - 2


class D2Q9(Quadrature):  # src/quadratures/D2Q9.jl
    name, _order = "D2Q9", 5

    def __init__(self):
        super().__init__([[0, -1, -1, -1, 0, 1, 1, 1, 0], [0, 1, 0, -1, -1, -1, 0, 1, 1]],
                         [4 / 9, 1 / 36, 1 / 9, 1 / 36, 1 / 9, 1 / 36, 1 / 9, 1 / 36, 1 / 9], 3.0)

    def _opposite(self, idx):
        if idx == 1:
            return 1
        return idx + 4 if idx <= 5 else idx - 4


class D2Q13(Quadrature):  # src/quadratures/D2Q13.jl
    name, _order = "D2Q13", 5

    def __init__(self):
        super().__init__([[0, 1, -1, 0, 0, 1, -1, 1, -1, 2, -2, 0, 0], [0, 0, 0, 1, -1, 1, -1, -1, 1, 0, 0, -2, 2]],
                         [3 / 8] + [1 / 12] * 4 + [1 / 16] * 4 + [1 / 96] * 4, 2.0)


class D2Q17(Quadrature):  # src/quadratures/D2Q17.jl
    name, _order = "D2Q17", 7

    def __init__(self):
        sq = math.sqrt(193)
        w = ([(575 + 193 * sq) / 8100] + [(3355 - 91 * sq) / 18000] * 4 + [(655 + 17 * sq) / 27000] * 4
             + [(685 - 49 * sq) / 54000] * 4 + [(1445 - 101 * sq) / 162000] * 4)
        super().__init__([[0, 1, -1, 0, 0, 1, -1, 1, -1, 2, -2, 2, -2, 3, -3, 0, 0],
                          [0, 0, 0, 1, -1, 1, -1, -1, 1, 2, -2, -2, 2, 0, 0, 3, -3]], w,
                         (125 + 5 * math.sqrt(193)) / 72)


class D2Q21(Quadrature):  # src/quadratures/D2Q21.jl (25 stored populations, 4 with weight 0)
    name, _order = "D2Q21", 7

    def __init__(self):
        w = ([91 / 324] + [1 / 12] * 4 + [2 / 27] * 4 + [7 / 360] * 4 + [1 / 432] * 4 + [1 / 1620] * 4 + [0.0] * 4)
        super().__init__([[0, 1, -1, 0, 0, 1, -1, 1, -1, 2, -2, 0, 0, 2, -2, 2, -2, 3, -3, 0, 0, 3, -3, -3, 3],
                          [0, 0, 0, 1, -1, 1, -1, -1, 1, 0, 0, 2, -2, 2, -2, -2, 2, 0, 0, 3, -3, 3, -3, 3, -3]],
                         w, 3 / 2)


class D2Q37(Quadrature):  # src/quadratures/D2Q37.jl
    name, _order = "D2Q37", 9

    def __init__(self):
        w = ([0.23315066913235250228650] + [0.10730609154221900241246] * 4 + [0.05766785988879488203006] * 4
             + [0.01420821615845075026469] * 4 + [0.00535304900051377523273] * 8
             + [0.00101193759267357547541] * 4 + [0.00024530102775771734547] * 4
             + [0.00028341425299419821740] * 8)
        r = 1.19697977039307435897239
        super().__init__([[0, 1, -1, 0, 0, 1, -1, 1, -1, 2, -2, 0, 0, 2, -2, -2, 2, 1, -1, 1, -1,
                           2, -2, 2, -2, 3, -3, 0, 0, 3, -3, 3, -3, 1, -1, -1, 1],
                          [0, 0, 0, 1, -1, 1, -1, -1, 1, 0, 0, 2, -2, 1, -1, 1, -1, 2, -2, -2, 2,
                           2, -2, -2, 2, 0, 0, 3, -3, 1, -1, -1, 1, 3, -3, 3, -3]], w, r * r)


def order(q):
    return q._order


def dimension(q):
    return 2


def opposite(q, idx):
    """opposite(q, idx) with Julia's 1-based population index."""
    return q._opposite(idx)


class _Quadratures:
    """The `Quadratures` named tuple (src/quadratures.jl:30-38)."""

    def __init__(self):
        self._cache = {}

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        if name not in self._cache:
            self._cache[name] = {"D2Q4": D2Q4, "D2Q5": D2Q5, "D2Q9": D2Q9, "D2Q13": D2Q13, "D2Q17": D2Q17,
                                 "D2Q21": D2Q21, "D2Q37": D2Q37}[name]()
        return self._cache[name]

    def __iter__(self):
        return iter(getattr(self, n) for n in ("D2Q4", "D2Q5", "D2Q9", "D2Q13", "D2Q17", "D2Q21", "D2Q37"))


Quadratures = _Quadratures()
