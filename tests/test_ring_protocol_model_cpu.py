"""Discrete-event model of the chunk ring of psmf_stream.cuh (producer thread, pass warps, `full` / `done`
mbarriers), used to pin down the parity-aliasing bug fixed by the per-slot lap counter.

mbarrier semantics modelled: a barrier counts completed phases; `try_wait.parity(P)` succeeds iff P is the parity
of the immediately PRECEDING phase (i.e. differs from the parity of the phase in progress).  Bulk loads are issued
in order but complete in any order.  A pass warp handles tiles wp, wp + NPW, ...: it jumps NPW / TS chunks per tile,
so with a shallow ring it can reach a slot whose previous load is still in flight; the parity test then succeeds on
the phase two laps back and the warp reads a stale chunk.  With the guard (`gen[slot] >= lap` before the parity
wait) this cannot happen, whatever the schedule."""

import random

import pytest


def simulate(nslot, nchunks_total, npw, ts, guard, seed):
    rng = random.Random(seed)
    full_phase = [0] * nslot            # completed phases of full[s]
    content = [None] * nslot            # chunk id whose data the slot holds (None while a load is in flight)
    gen = [0] * nslot                   # lap of the load last requested for the slot
    done_cnt = {}                       # chunk id -> tiles processed
    inflight = []                       # (slot, chunk) loads issued, not yet completed
    # producer state
    next_store = 0                      # next chunk whose `done` the producer waits for
    for c in range(min(nslot, nchunks_total)):
        inflight.append((c, c))
    issued = min(nslot, nchunks_total)
    # warps: list of tile indices still to process
    ntiles = nchunks_total * ts
    todo = [list(range(w, ntiles, npw)) for w in range(npw)]
    stale_reads = 0
    steps = 0
    while any(todo) or next_store < nchunks_total:
        steps += 1
        assert steps < 200000, "model deadlock"
        actors = []
        if inflight:
            actors.append("complete")
        if next_store < nchunks_total and done_cnt.get(next_store, 0) == ts:
            actors.append("producer")
        for w in range(npw):
            if todo[w]:
                kk = todo[w][0] // ts
                slot, lap = kk % nslot, kk // nslot
                if guard and gen[slot] < lap:
                    continue
                if (full_phase[slot] & 1) != (lap & 1):         # try_wait.parity(lap & 1) succeeds
                    actors.append(w)
        assert actors, "model deadlock"
        a = rng.choice(actors)
        if a == "complete":                                      # any in-flight load may land next
            slot, chunk = inflight.pop(rng.randrange(len(inflight)))
            content[slot] = chunk
            full_phase[slot] += 1
        elif a == "producer":                                    # store chunk next_store, reload its slot
            slot = next_store % nslot
            if issued < nchunks_total:
                content[slot] = None
                gen[slot] = issued // nslot
                inflight.append((slot, issued))
                issued += 1
            next_store += 1
        else:
            tile = todo[a].pop(0)
            kk = tile // ts
            if content[kk % nslot] != kk:
                stale_reads += 1                                 # processed a slot that does not hold its chunk
            done_cnt[kk] = done_cnt.get(kk, 0) + 1
    return stale_reads


@pytest.mark.parametrize("nslot", [5, 7, 9])
def test_guarded_ring_never_reads_a_stale_slot(nslot):
    for seed in range(40):
        assert simulate(nslot, 60, 12, 4, guard=True, seed=seed) == 0


def test_unguarded_shallow_ring_can_alias_parities():
    """Without the lap counter a 5-slot ring with 12 pass warps (3 chunks per jump) does read stale slots under
    some schedules -- the failure seen on the GPU at d = 2M (launch failure) and 1.5M (stale rows of C)."""
    hits = 0
    for seed in range(60):
        try:
            hits += simulate(5, 60, 12, 4, guard=False, seed=seed) > 0
        except AssertionError:
            hits += 1                                            # the corrupted protocol may also deadlock
    assert hits > 0
