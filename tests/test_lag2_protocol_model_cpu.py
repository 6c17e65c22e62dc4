"""Discrete-event model of the grid-level protocol of psmf_stream.cuh: data CTAs, their reduce warps and the
control CTA exchange three kinds of double-buffered tagged cells and never meet at a barrier.

    pass p of CTA c      needs parameter set p-1 (cell parity (p-1)&1, tag p)       writes partial[p&1][c], tag p+1
    reduce warp c, step t needs partial[t&1][*] with tag t+1                          writes total[t&1][c], tag t+1
    control CTA, step t   needs total[t&1][*] with tag t+1                            writes params[(t+1)&1], tag t+2

A value is overwritten two steps later; the protocol is correct iff nobody can still be waiting for the old tag at
that time (the waiter would spin forever).  The model runs the actors under random schedules and checks that every
run completes, i.e. that two steps in flight never become three."""

import random

import pytest


def run_model(nctas, nsteps, seed):
    rng = random.Random(seed)
    params = [1, 0]                       # set 0 (tag 1) published before the loop
    partial = [[0] * nctas, [0] * nctas]
    total = [[0] * nctas, [0] * nctas]
    pass_of = [0] * nctas                 # next pass of each data CTA (nsteps = flush pass, needs sets n-1 and n)
    red_of = [0] * nctas                  # next step of each reduce warp
    ctl = 0                               # next step of the control CTA
    max_lead = 0
    for _ in range(100000):
        moves = []
        for c in range(nctas):
            p = pass_of[c]
            if p < nsteps and (p == 0 or params[(p - 1) & 1] == p):
                moves.append(("pass", c))
            elif p == nsteps and params[(p - 1) & 1] == p and params[p & 1] == p + 1:
                moves.append(("flush", c))
            t = red_of[c]
            if t < nsteps and all(v == t + 1 for v in partial[t & 1]):
                moves.append(("reduce", c))
        if ctl < nsteps and all(v == ctl + 1 for v in total[ctl & 1]):
            moves.append(("control", 0))
        if not moves:
            break
        kind, c = rng.choice(moves)
        if kind == "pass":
            p = pass_of[c]
            partial[p & 1][c] = p + 1
            pass_of[c] = p + 1
        elif kind == "flush":
            pass_of[c] = nsteps + 1
        elif kind == "reduce":
            t = red_of[c]
            total[t & 1][c] = t + 1
            red_of[c] = t + 1
        else:
            params[(ctl + 1) & 1] = ctl + 2
            ctl += 1
        max_lead = max(max_lead, max(pass_of) - min(pass_of))
    finished = ctl == nsteps and all(p == nsteps + 1 for p in pass_of) and all(t == nsteps for t in red_of)
    return finished, max_lead


@pytest.mark.parametrize("nctas", [1, 2, 5, 16])
def test_protocol_completes_under_random_schedules(nctas):
    for seed in range(50):
        finished, lead = run_model(nctas, 25, seed)
        assert finished, "a waiter lost its value (seed %d)" % seed
        assert lead <= 2                    # data CTAs are never more than two passes apart
