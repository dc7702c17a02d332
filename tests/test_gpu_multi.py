"""y-slab multi-GPU parity: two ranks (one process per GPU, halos exchanged by liblbm_b200.so over
NCCL) must reproduce the single-domain oracle bit for bit."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _worker(rank, world, idq, out, lattice, model, nx, ny, nsteps, walls, dtype, overlap):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "latticeboltzmann.jl_b200"))
    import lbm
    from lbm import _abi
    import oracle.lbm_oracle as O
    try:
        if rank == 0:
            nid = _abi.nccl_unique_id()
            for _ in range(world - 1):
                idq.put(nid)
        else:
            nid = idq.get(timeout=60)
        qo = O.L.BY_NAME[lattice]()
        rng = np.random.default_rng(5)
        f = np.stack([qo.w[i] * (1 + 0.01 * rng.uniform(-1, 1, (ny, nx))) for i in range(qo.Q)])
        code = {"SRT": _abi.SRT, "TRT": _abi.TRT, "MRT": _abi.MRT}[model]
        taus = {"SRT": [0.8], "TRT": [0.8, 1.1], "MRT": [0.8, 0.9, 1.1, 1.3]}[model]
        bcs = []
        if walls:
            bcs = [lbm.BounceBack(lbm.South(), (1, nx), (1, ny)).to_abi(),
                   lbm.MovingWall(lbm.North(), (1, nx), (1, ny), [0.01, 0.0]).to_abi()]
        c = _abi.Context(nx, ny, lattice, code, taus, bcs, dtype=dtype, device=rank, rank=rank, world=world, nccl_id=nid)
        c.set_option("overlap", overlap)
        assert (c.y0, c.ny_local) == lbm.slab_rows(ny, rank, world)
        c.set_force_uniform(1e-6, 2e-6)
        slab = np.asfortranarray(np.transpose(f[:, c.y0:c.y0 + c.ny_local], (2, 1, 0)))
        c.upload_f(slab)
        c.step(0, 3)
        c.step(3, nsteps - 3)
        got = np.transpose(c.download_f(), (2, 1, 0))
        red = c.reduce(_abi.REDUCE_CONSERVED)
        c.step(nsteps, 2)  # resume path after a download
        got2 = np.transpose(c.download_f(), (2, 1, 0))
        out.put((rank, c.y0, got, got2, red, None))
        c.close()
    except Exception as e:  # pragma: no cover
        out.put((rank, -1, None, None, None, repr(e)))


@pytest.mark.parametrize("lattice,model,walls,overlap", [
    ("D2Q9", "TRT", False, 1), ("D2Q9", "TRT", True, 1), ("D2Q9", "SRT", True, 0),
    ("D2Q13", "TRT", False, 1), ("D2Q37", "TRT", True, 1), ("D2Q37", "MRT", False, 1), ("D2Q17", "SRT", False, 1),
])
def test_two_slabs_equal_single_domain(lattice, model, walls, overlap):
    import torch.multiprocessing as mp
    import oracle.lbm_oracle as O
    nx, ny, nsteps, world = 40, 37, 9, 2
    qo = O.L.BY_NAME[lattice]()
    rng = np.random.default_rng(5)
    f = np.stack([qo.w[i] * (1 + 0.01 * rng.uniform(-1, 1, (ny, nx))) for i in range(qo.Q)])
    force = (1e-6, 2e-6)
    cm = {"SRT": O.SRT(0.8, force), "TRT": O.TRT(0.8, 1.1, force), "MRT": O.MRT(qo, [0.8, 0.9, 1.1, 1.3], force)}[model]
    bcs = [O.BounceBack("S", (1, nx), (1, ny)), O.MovingWall("N", (1, nx), (1, ny), [0.01, 0.0])] if walls else []
    want = f
    for _ in range(nsteps):
        want, _ = O.step(cm, qo, bcs, want)
    want2 = want
    for _ in range(2):
        want2, _ = O.step(cm, qo, bcs, want2)
    ctx = mp.get_context("spawn")
    idq, out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, idq, out, lattice, model, nx, ny, nsteps, walls, 0, overlap))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    mass = 0.0
    for rank, y0, got, got2, red, err in res:
        assert err is None, err
        nyl = got.shape[1]
        if model == "MRT":
            assert np.abs(got - want[:, y0:y0 + nyl]).max() < 1e-13
            assert np.abs(got2 - want2[:, y0:y0 + nyl]).max() < 1e-13
        else:
            assert np.array_equal(got, want[:, y0:y0 + nyl]), f"rank {rank}"
            assert np.array_equal(got2, want2[:, y0:y0 + nyl]), f"rank {rank} (resume)"
        mass += red[0]
    assert np.isclose(mass, O.density(qo, [want[i] for i in range(qo.Q)]).sum(), rtol=1e-13)
