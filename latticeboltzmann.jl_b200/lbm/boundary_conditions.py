"""Boundary-condition types of the Julia API (src/boundary_conditions*.jl).

`xs`, `ys` are 1-based inclusive ranges given as `(lo, hi)` tuples or Python `range`
objects with Julia meaning (`range(1, NX + 1)` == `1:NX`).
"""
from . import _abi


class Direction:
    code = -1


class North(Direction):
    code = _abi.NORTH


class East(Direction):
    code = _abi.EAST


class South(Direction):
    code = _abi.SOUTH


class West(Direction):
    code = _abi.WEST


def _bounds(r):
    if isinstance(r, range):
        if r.step != 1:
            raise ValueError("only unit-step ranges are supported")
        return r.start, r.stop - 1
    lo, hi = r
    return int(lo), int(hi)


class BoundaryCondition:
    pass


class BounceBack(BoundaryCondition):
    """BounceBack(direction, xs, ys): half-way bounce-back (bounce_back.jl:2-6)."""

    def __init__(self, direction, xs, ys):
        self.direction = direction() if isinstance(direction, type) else direction
        self.xs, self.ys = _bounds(xs), _bounds(ys)

    def to_abi(self):
        b = _abi.lbm_bc()
        b.kind, b.direction = _abi.BC_BOUNCE_BACK, self.direction.code
        (b.x0, b.x1), (b.y0, b.y1) = self.xs, self.ys
        b.rho = b.T = 1.0
        return b


class MovingWall(BoundaryCondition):
    """MovingWall(direction, xs, ys, u[, rho, T]) (moving_wall.jl:5-15).  Only North has an
    `apply!` method in the reference (moving_wall.jl:17); the library rejects the others."""

    def __init__(self, direction, xs, ys, u, rho=1.0, T=1.0):
        self.direction = direction() if isinstance(direction, type) else direction
        self.xs, self.ys = _bounds(xs), _bounds(ys)
        self.u = (float(u[0]), float(u[1]))
        self.rho, self.T = float(rho), float(T)

    def to_abi(self):
        b = _abi.lbm_bc()
        b.kind, b.direction = _abi.BC_MOVING_WALL, self.direction.code
        (b.x0, b.x1), (b.y0, b.y1) = self.xs, self.ys
        b.u[0], b.u[1] = self.u
        b.rho, b.T = self.rho, self.T
        return b
