"""Exhaustive model check of the peer-memory halo protocol (csrc/kernels_inst.cu p2p_wait / p2p_store /
p2p_signal).

Every rank runs, per step e = 1, 2, ..., one boundary launch B(e) made of four atomic actions

    WAIT   : proceed only when both neighbours have published epoch e - 1
    READ   : pull from my ghost rows of buffer (e - 1) % 2        -- they must hold the neighbours' rows of state e - 1
    PUSH   : store my rows of state e into the neighbours' ghost rows of buffer e % 2
             -- the same physical rows held state e - 2, which a neighbour reads in ITS B(e - 1): it must be done with it
    SIGNAL : publish epoch e to both neighbours

(interior launches never touch ghost rows and are left out).  The ranks are independent GPUs: any interleaving of their
actions can happen.  The test walks ALL reachable interleavings for small rings and checks that no READ sees stale or
future data, no PUSH overwrites rows a neighbour still needs, and nobody deadlocks -- and that weakened variants of the
protocol (waiting for e - 2 only, or signalling before the push) are caught by the same checker."""
import itertools

import pytest


def explore(n_ranks, n_epochs, wait_for=1, signal_before_push=False, final_wait=True):
    """DFS over all interleavings.  Returns None if the protocol is safe and live, else a description of the violation.
    After its last step every rank runs the end-of-batch hook (k_p2p_wait_epoch: WAIT for the last epoch) and then an
    operation outside the protocol that pulls from the ghost rows of the final state (lbm_download_f / lbm_moments)."""
    order = ("WAIT", "READ", "SIGNAL", "PUSH") if signal_before_push else ("WAIT", "READ", "PUSH", "SIGNAL")
    tail = ("WAIT", "READ") if final_wait else ("READ",)  # pseudo-step n_epochs + 1: wait for epoch n_epochs, read its halos
    up = lambda r: (r + 1) % n_ranks        # noqa: E731
    down = lambda r: (r - 1) % n_ranks      # noqa: E731
    # state: per rank (epoch being produced, index of the next action), flags[r] = (from_down, from_up),
    #        ghost[r][buf] = (tag of the rows in the bottom ghost rows, tag in the top ghost rows)
    init = (tuple((1, 0) for _ in range(n_ranks)),
            tuple((0, 0) for _ in range(n_ranks)),
            tuple(((0, 0), (-1, -1)) for _ in range(n_ranks)))  # buffer 0 holds state 0 everywhere (set up before the batch)
    seen, stack = {init}, [init]
    while stack:
        pcs, flags, ghost = stack.pop()
        enabled = False
        for r in range(n_ranks):
            e, a = pcs[r]
            if e > n_epochs + 1:
                continue
            acts = order if e <= n_epochs else tail
            act = acts[a]
            nflags, nghost = flags, ghost
            if act == "WAIT":
                if flags[r][0] < e - wait_for or flags[r][1] < e - wait_for:
                    continue  # guard not satisfied: this rank spins
            elif act == "READ":
                have = ghost[r][(e - 1) % 2]
                if have != (e - 1, e - 1):
                    return f"rank {r} step {e}: ghost rows hold states {have}, expected {e - 1}"
            elif act == "PUSH":
                g = [list(map(list, gr)) for gr in ghost]
                for nb, side in ((up(r), 0), (down(r), 1)):  # my top rows -> up's bottom ghosts; my bottom rows -> down's top ghosts
                    ne, na = pcs[nb]
                    # the rows being overwritten held state e - 2; nb reads them in its B(e - 1)
                    nb_done_reading = ne > e - 1 or (ne == e - 1 and na > order.index("READ"))  # (e - 1 <= n_epochs here)
                    if e >= 2 and not nb_done_reading:
                        return f"rank {r} step {e}: overwrites ghost rows rank {nb} has not read yet (it is at step {ne}, action {na})"
                    g[nb][e % 2][side] = e
                nghost = tuple(tuple(tuple(x) for x in gr) for gr in g)
            else:  # SIGNAL
                f = [list(x) for x in flags]
                f[up(r)][0] = e      # I am up's `down`
                f[down(r)][1] = e    # I am down's `up`
                nflags = tuple(tuple(x) for x in f)
            enabled = True
            npcs = list(pcs)
            npcs[r] = (e, a + 1) if a + 1 < len(acts) else (e + 1, 0)
            nxt = (tuple(npcs), nflags, nghost)
            if nxt not in seen:
                seen.add(nxt)
                stack.append(nxt)
        if not enabled and any(e <= n_epochs + 1 for e, _ in pcs):
            return f"deadlock at {pcs} with flags {flags}"
    return None


@pytest.mark.parametrize("n_ranks,n_epochs", [(2, 6), (3, 6), (4, 4), (5, 3), (6, 2)])
def test_protocol_is_safe_and_live_under_every_interleaving(n_ranks, n_epochs):
    assert explore(n_ranks, n_epochs) is None


@pytest.mark.parametrize("n_ranks", [2, 3])
def test_checker_catches_a_wait_that_is_one_epoch_short(n_ranks):
    """waiting for e - 2 instead of e - 1 lets a rank read ghost rows its neighbour has not written yet"""
    msg = explore(n_ranks, 4, wait_for=2)
    assert msg is not None and ("ghost rows hold" in msg or "overwrites" in msg), msg


@pytest.mark.parametrize("n_ranks", [2, 3])
def test_checker_catches_signalling_before_the_push(n_ranks):
    """publishing the epoch before the rows have been stored lets the neighbour read stale ghost rows"""
    msg = explore(n_ranks, 3, signal_before_push=True)
    assert msg is not None and "ghost rows hold" in msg, msg


@pytest.mark.parametrize("n_ranks", [2, 3])
def test_checker_catches_a_missing_end_of_batch_wait(n_ranks):
    """without k_p2p_wait_epoch a download right after the batch can pull halos that have not landed yet"""
    msg = explore(n_ranks, 2, final_wait=False)
    assert msg is not None and "ghost rows hold" in msg, msg


def test_epochs_of_replayed_graph_launches():
    """lbm_step replays graphs of 16 steps whose P2P launches carry epoch = flags[EPOCH_BASE] + k (k = 1..16); the host sets
    the base to its epoch counter before each replay and advances the counter by 16: the sequence of absolute epochs is
    gap-free across plain launches and replays in any mix."""
    host_epoch, issued = 0, []
    for chunk in (1, 1, 16, 16, 1, 1, 1, 16, 1):  # plain launches (1) and graph replays (16), as do_steps mixes them
        if chunk == 1:
            host_epoch += 1
            issued.append(host_epoch)
        else:
            base = host_epoch            # k_p2p_set_base(flags, c->epoch)
            issued.extend(base + k for k in range(1, 17))
            host_epoch += 16
    assert issued == list(range(1, len(issued) + 1))
    assert list(itertools.accumulate([1] * len(issued))) == issued
