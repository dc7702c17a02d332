"""y-slab multi-GPU parity: the ranks (one process per GPU; halos pushed into the neighbours' ghost rows through
peer memory by the boundary-row launch, or exchanged with NCCL send/recv) must reproduce the single-domain oracle
bit for bit."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _worker(rank, world, idq, out, lattice, model, nx, ny, nsteps, walls, dtype, overlap, p2p=1, single_steps=False, arith=0,
            persistent=2):
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
        if p2p < 0 and rank == world - 1:
            os.environ["LBM_P2P"] = "0"  # ONE rank cannot (here: will not) map its neighbours -> all must agree on NCCL
        c = _abi.Context(nx, ny, lattice, code, taus, bcs, dtype=dtype, arith=arith, device=rank, rank=rank, world=world, nccl_id=nid)
        c.set_option("overlap", overlap)
        c.set_option("p2p", 1 if p2p else 0)
        c.set_option("persistent", persistent)  # 0: plain launches / graphs, 1: persistent kernel wherever possible, 2: automatic
        path = c.halo_path
        assert (c.y0, c.ny_local) == lbm.slab_rows(ny, rank, world)
        c.set_force_uniform(1e-6, 2e-6)
        slab = np.asfortranarray(np.transpose(f[:, c.y0:c.y0 + c.ny_local], (2, 1, 0)))
        c.upload_f(slab)
        if single_steps:  # one batch per step: exercises the batch barrier / final-epoch wait every step
            for t in range(nsteps):
                c.step(t, 1)
        else:
            c.step(0, 3)
            c.step(3, nsteps - 3)
        got = np.transpose(c.download_f(), (2, 1, 0))
        red = c.reduce(_abi.REDUCE_CONSERVED)
        c.step(nsteps, 2)  # resume path after a download
        got2 = np.transpose(c.download_f(), (2, 1, 0))
        out.put((rank, c.y0, got, got2, red, None, path))
        c.close()
    except Exception as e:  # pragma: no cover
        out.put((rank, -1, None, None, None, repr(e), -1))


def _gpus():
    import torch
    return torch.cuda.device_count()


def _run_slabs(lattice, model, walls, overlap, p2p, world=2, nx=40, ny=37, nsteps=9, single_steps=False, persistent=2):
    import torch.multiprocessing as mp
    import oracle.lbm_oracle as O
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
    procs = [ctx.Process(target=_worker, args=(r, world, idq, out, lattice, model, nx, ny, nsteps, walls, 0, overlap, p2p,
                                               single_steps, 0, persistent))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    mass = 0.0
    paths = set()
    for rank, y0, got, got2, red, err, path in res:
        assert err is None, err
        paths.add(path)
        nyl = got.shape[1]
        if model == "MRT":
            assert np.abs(got - want[:, y0:y0 + nyl]).max() < 1e-13
            assert np.abs(got2 - want2[:, y0:y0 + nyl]).max() < 1e-13
        else:
            assert np.array_equal(got, want[:, y0:y0 + nyl]), f"rank {rank}"
            assert np.array_equal(got2, want2[:, y0:y0 + nyl]), f"rank {rank} (resume)"
        mass += red[0]
    assert np.isclose(mass, O.density(qo, [want[i] for i in range(qo.Q)]).sum(), rtol=1e-13)
    assert len(paths) == 1, f"ranks disagree on the halo path: {paths}"
    if p2p <= 0:
        assert paths == {1}
    return paths.pop()


@pytest.mark.parametrize("lattice,model,walls,overlap,p2p", [
    ("D2Q9", "TRT", False, 1, 1), ("D2Q9", "TRT", True, 1, 1), ("D2Q9", "SRT", True, 0, 1),
    ("D2Q13", "TRT", False, 1, 1), ("D2Q37", "TRT", True, 1, 1), ("D2Q37", "MRT", False, 1, 1), ("D2Q17", "SRT", False, 1, 1),
    # NCCL send/recv path (LBM_P2P=0 / boxes without peer access)
    ("D2Q9", "TRT", True, 1, 0), ("D2Q9", "SRT", True, 0, 0), ("D2Q37", "TRT", True, 1, 0),
])
@pytest.mark.parametrize("persistent", [0, 1])
def test_two_slabs_equal_single_domain(lattice, model, walls, overlap, p2p, persistent):
    """persistent = 0: boundary-row + interior launches per step (CUDA graphs for long batches); 1: the persistent
    multi-step kernel whose edge CTAs push the boundary rows and hand-shake with the neighbours mid-step."""
    if persistent and not p2p:
        pytest.skip("the persistent kernel exchanges halos through peer memory only")
    _run_slabs(lattice, model, walls, overlap, p2p, persistent=persistent)


def _peer_access(a=0, b=1):
    import torch
    return torch.cuda.device_count() > max(a, b) and torch.cuda.can_device_access_peer(a, b) and torch.cuda.can_device_access_peer(b, a)


def test_peer_memory_path_is_taken_on_nvlink_boxes():
    """Where the GPUs can map each other (the B200 boxes: NVSwitch) the default halo path must be peer memory."""
    path = _run_slabs("D2Q9", "TRT", False, 1, 1)
    if not _peer_access():
        pytest.skip("GPUs 0 and 1 cannot access each other's memory on this box: NCCL path (checked above)")
    assert path == 2


@pytest.mark.parametrize("lattice,model,walls,nx,ny,nsteps,single_steps", [
    ("D2Q9", "TRT", False, 1024, 203, 60, False),   # many CTAs per boundary launch, uneven slabs
    ("D2Q9", "SRT", True, 300, 64, 40, True),       # one batch per step
    ("D2Q37", "TRT", True, 515, 47, 24, False),     # halo 3, odd width
    ("D2Q21", "TRT", False, 64, 13, 12, True),      # thin slabs (6/7 rows <= 2H+1 on one rank): single-launch path
])
@pytest.mark.parametrize("persistent", [0, 1])
def test_peer_memory_stress(lattice, model, walls, nx, ny, nsteps, single_steps, persistent):
    _run_slabs(lattice, model, walls, 1, 1, nx=nx, ny=ny, nsteps=nsteps, single_steps=single_steps, persistent=persistent)


@pytest.mark.parametrize("lattice,model,walls,nx,ny,nsteps", [
    ("D2Q9", "TRT", False, 1024, 2048, 40),    # 1024 x 1024 per GPU: the launch-bound case the persistent kernel is for
    ("D2Q9", "SRT", True, 1024, 512, 101),     # C3-like channel, odd step count
    ("D2Q37", "TRT", True, 2048, 64, 21),      # halo 3: the boundary rows of an edge span several CTAs
    ("D2Q17", "MRT", False, 256, 300, 33),
])
def test_persistent_kernel_slabs_at_production_sizes(lattice, model, walls, nx, ny, nsteps):
    """The persistent multi-step kernel across two GPUs at sizes where every SM holds several CTAs of the cooperative
    grid: bit-identical (SRT/TRT) to the single-domain oracle."""
    _run_slabs(lattice, model, walls, 1, 1, nx=nx, ny=ny, nsteps=nsteps, persistent=1)


@pytest.mark.parametrize("lattice,model,walls,nx,ny,nsteps,single_steps", [
    ("D2Q9", "TRT", False, 40, 37, 9, False), ("D2Q9", "SRT", True, 40, 37, 9, False),
    ("D2Q9", "TRT", False, 1024, 203, 60, False),   # many CTAs per row, uneven slabs, graph replays
    ("D2Q9", "SRT", True, 1024, 512, 101, False),   # C3-like channel, odd step count
    ("D2Q37", "TRT", True, 515, 47, 24, False),     # halo 3, odd width: the edge rows span several row groups
    ("D2Q13", "MRT", False, 300, 64, 40, True),     # one batch per step
    ("D2Q21", "TRT", False, 64, 13, 12, True),      # thin slabs: falls back to the whole-slab launch
])
def test_merged_boundary_and_interior_launch(lattice, model, walls, nx, ny, nsteps, single_steps):
    """Option overlap = 2: ONE launch per step on a y-slab -- its first CTAs take the edge rows and the halo hand-shake, the
    rest the interior rows.  Bit-identical (SRT / TRT) to the single-domain oracle like the two-launch form."""
    _run_slabs(lattice, model, walls, 2, 1, nx=nx, ny=ny, nsteps=nsteps, single_steps=single_steps, persistent=0)


@pytest.mark.parametrize("world", [3, 4, 8])
@pytest.mark.parametrize("overlap", [1, 2])
@pytest.mark.parametrize("lattice,model,walls", [("D2Q9", "TRT", True), ("D2Q37", "TRT", False)])
def test_more_slabs(world, lattice, model, walls, overlap):
    """Rings of 3 / 4 / 8 slabs (distinct up and down neighbours), two-launch and merged-launch forms."""
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    _run_slabs(lattice, model, walls, overlap, 1, world=world, nx=136, ny=16 * world + 5, nsteps=15)


def test_ranks_agree_on_the_halo_path_when_one_cannot_map_its_neighbours():
    """lbm_create takes the min over ranks of "mapped my neighbours": a single rank that opts out (LBM_P2P=0 in its
    environment only) moves every rank to the NCCL path -- no mixed protocols, results unchanged."""
    assert _run_slabs("D2Q9", "TRT", True, 1, -1) == 1


def test_two_slabs_in_one_process():
    """Both contexts in ONE process (two host threads, one per GPU -- what a single Julia process driving several GPUs
    does): the neighbours' buffers are reached through plain peer access instead of cudaIpc; same bit-identical result."""
    import threading
    import oracle.lbm_oracle as O
    import lbm
    from lbm import _abi
    lattice, nx, ny, nsteps, world = "D2Q13", 72, 41, 40, 2
    qo = O.L.BY_NAME[lattice]()
    rng = np.random.default_rng(5)
    f = np.stack([qo.w[i] * (1 + 0.01 * rng.uniform(-1, 1, (ny, nx))) for i in range(qo.Q)])
    force = (1e-6, 2e-6)
    cm = O.TRT(0.8, 1.1, force)
    want = f
    for _ in range(nsteps):
        want, _ = O.step(cm, qo, [], want)
    nid = _abi.nccl_unique_id()
    res, errs = {}, []

    def work(rank):
        try:
            c = _abi.Context(nx, ny, lattice, _abi.TRT, [0.8, 1.1], [], device=rank, rank=rank, world=world, nccl_id=nid)
            c.set_force_uniform(*force)
            c.upload_f(np.asfortranarray(np.transpose(f[:, c.y0:c.y0 + c.ny_local], (2, 1, 0))))
            c.step(0, 3)
            c.step(3, nsteps - 3)
            res[rank] = (c.y0, np.transpose(c.download_f(), (2, 1, 0)), c.halo_path)
            c.close()
        except Exception as e:  # pragma: no cover
            errs.append(repr(e))

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=240)
    assert not errs, errs
    assert len(res) == world
    for rank, (y0, got, path) in res.items():
        assert path == (2 if _peer_access() else 1), "peer access is available on this box but the NCCL path was taken"
        assert np.array_equal(got, want[:, y0:y0 + got.shape[1]]), f"rank {rank}"
