"""A host model of the synchronisation protocol of the fused persistent sweep kernel (csrc/acquire_fused.cu).

The kernel decouples two item streams with per-slot counters in global memory:
    ready[slot]  build items finished     a consumer waits for (generation + 1) * nJ before it loads a tile's digits
    done[slot]   row blocks finished      the arrival that completes (generation + 1) * nI finalises the tile
    freed[slot]  tiles finalised          a builder waits for `generation` before it overwrites the slot
Work items are drawn in order from two global counters; consumers go through groups of G candidate tiles (heaviest row
block first inside a group), the ring holds R = 2 G tiles.  The claims checked here, under random interleavings of an
arbitrary number of resident CTAs (including fewer CTAs than a group has items, and CTAs that arrive late):
  * progress: every item is processed, whatever the schedule (no deadlock);
  * safety: a consumer only ever reads a slot that holds the digits of ITS tile, complete, and a builder never overwrites
    a slot whose tile has not been finalised;
  * every tile is finalised exactly once, after all of its row blocks.
This is test infrastructure for the host-visible logic of the kernel (the decode of work items is restated from the
device code); the arithmetic is tested on the GPU in tests/test_gpu_fused.py."""
import random

import pytest


def decode(w, nct, nI, G):
    """acquire_fused.cu, producer warp: work item -> (candidate tile, row block), None for padding items, 'end' when exhausted"""
    per_group = nI * G
    total = -(-nct // G) * per_group
    if w >= total:
        return "end"
    grp, rem = divmod(w, per_group)
    gc = min(G, nct - grp * G)
    ib = nI - 1 - rem // gc
    ct = grp * G + rem % gc
    return None if ib < 0 else (ct, ib)


def simulate(nct, nI, nJ, G, n_cta, seed, late=0):
    rng = random.Random(seed)
    R = 2 * G
    ready, done, freed = [0] * R, [0] * R, [0] * R
    slot_tile = [None] * R            # which tile's digits a slot holds (set when its first build item starts)
    slot_built = [0] * R              # build items written for that tile
    finalised = [0] * nct
    blocks_done = [0] * nct
    nxt = {"work": 0, "build": 0}
    # every CTA runs two independent agents; each agent is a little state machine
    agents = []
    for c in range(n_cta):
        agents.append({"kind": "consumer", "state": "fetch", "start": rng.randrange(0, late + 1)})
        agents.append({"kind": "builder", "state": "fetch", "start": rng.randrange(0, late + 1)})
    alive = len(agents)
    steps = 0
    while alive:
        steps += 1
        assert steps < 200 * (nct * (nI + nJ) + len(agents)) + 10000, "no progress: deadlock"
        a = rng.choice(agents)
        if a["state"] == "exit" or steps < a["start"]:
            continue
        if a["kind"] == "consumer":
            if a["state"] == "fetch":
                item = decode(nxt["work"], nct, nI, G); nxt["work"] += 1
                if item == "end":
                    a["state"] = "exit"; alive -= 1
                elif item is not None:
                    a["item"], a["state"] = item, "wait_ready"
            elif a["state"] == "wait_ready":
                ct, ib = a["item"]
                if ready[ct % R] >= (ct // R + 1) * nJ:
                    a["state"] = "compute"
            elif a["state"] == "compute":
                ct, ib = a["item"]
                s = ct % R
                assert slot_tile[s] == ct and slot_built[s] == nJ, "consumer reads a slot that does not hold its complete tile"
                blocks_done[ct] += 1
                done[s] += 1
                if done[s] == (ct // R + 1) * nI:             # the arrival that completes the tile finalises it
                    assert blocks_done[ct] == nI
                    finalised[ct] += 1
                    freed[s] += 1
                a["state"] = "fetch"
        else:
            if a["state"] == "fetch":
                b = nxt["build"]; nxt["build"] += 1
                if b >= nct * nJ:
                    a["state"] = "exit"; alive -= 1
                else:
                    a["item"], a["state"] = divmod(b, nJ), "wait_freed"
            elif a["state"] == "wait_freed":
                ct, jb = a["item"]
                if freed[ct % R] >= ct // R:
                    a["state"] = "build"
            elif a["state"] == "build":
                ct, jb = a["item"]
                s = ct % R
                if slot_tile[s] != ct:                        # first block of a new generation: the old tile must be finalised
                    assert slot_tile[s] is None or finalised[slot_tile[s]] == 1, "builder overwrites a tile that is still in use"
                    slot_tile[s], slot_built[s] = ct, 0
                slot_built[s] += 1
                ready[s] += 1
                a["state"] = "fetch"
    assert all(f == 1 for f in finalised), "every tile is finalised exactly once"
    assert all(b == nI for b in blocks_done)


@pytest.mark.parametrize("nct,nI,nJ,G,n_cta", [
    (37, 4, 2, 3, 5),          # fewer CTAs than a group has items
    (64, 8, 4, 4, 40),         # more CTAs than a group has items: several groups in flight
    (10, 32, 16, 17, 148),     # the benchmark's geometry (N = 4096), a sweep shorter than one group
    (5, 3, 2, 1, 2),           # G = 1: ring of two tiles
    (100, 2, 1, 8, 7),
])
def test_fused_protocol_makes_progress_and_is_safe_under_random_schedules(nct, nI, nJ, G, n_cta):
    for seed in range(6):
        simulate(nct, nI, nJ, G, n_cta, seed)
        simulate(nct, nI, nJ, G, n_cta, 100 + seed, late=2000)      # CTAs that become resident late


def test_work_item_decode_covers_every_tile_and_row_block_once():
    for nct, nI, G in [(37, 4, 3), (64, 8, 4), (1, 32, 17), (35, 32, 17)]:
        seen = {}
        w = 0
        while True:
            it = decode(w, nct, nI, G); w += 1
            if it == "end":
                break
            if it is not None:
                seen[it] = seen.get(it, 0) + 1
        assert len(seen) == nct * nI and set(seen.values()) == {1}
        # inside a group the heaviest row block comes first
        first = decode(0, nct, nI, G)
        assert first == (0, nI - 1)
