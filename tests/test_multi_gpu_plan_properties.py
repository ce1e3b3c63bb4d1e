"""Property tests (CPU) of the multi-GPU plans over random volumes, chunk pitches and
world sizes: slabs partition the volume, every chunk row has one owner, the balanced
dealing is optimal among contiguous dealings, plane and box transfers deliver exactly the
planes the workers lack, and lending never raises the largest load."""
import itertools

import numpy as np
from hypothesis import given, settings, strategies as st

from magellanmapper_b200 import multi_gpu as mg


def _rows(n_planes, pitch, overlap):
    n = -(-n_planes // pitch)
    return [(k * pitch, min(k * pitch + pitch + overlap, n_planes)) for k in range(n)]


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 5000), st.integers(1, 9), st.integers(1, 40))
def test_slab_bounds_partition(n_planes, world, align):
    b = mg.slab_bounds(n_planes, world, align)
    assert len(b) == world and b[0][0] == 0 and b[-1][1] == n_planes
    for (a0, a1), (b0, b1) in zip(b[:-1], b[1:]):
        assert a0 <= a1 == b0 <= b1
        assert a1 % align == 0 or a1 == n_planes
    for z in (0, n_planes // 2, n_planes - 1):
        r = mg.owner_of(z, b)
        assert b[r][0] <= z < b[r][1]


@settings(max_examples=120, deadline=None)
@given(st.integers(20, 3000), st.integers(10, 600), st.integers(0, 9), st.integers(1, 8))
def test_balanced_rows_are_contiguous_complete_and_optimal(n_planes, pitch, overlap, world):
    zb = _rows(n_planes, pitch, overlap)
    rows = mg.assign_chunk_rows_balanced(zb, world)
    assert len(rows) == world
    flat = [k for r in rows for k in r]
    assert flat == list(range(len(zb)))                       # complete, contiguous, in order
    cost = lambda run: sum(zb[k][1] - zb[k][0] for k in run)
    worst = max(cost(r) for r in rows)
    # brute force over every contiguous dealing when that is cheap
    n = len(zb)
    if n <= 9:
        best = min(max(cost(range(a, b)) for a, b in zip((0,) + cuts, cuts + (n,)))
                   for cuts in itertools.combinations_with_replacement(range(n + 1), world - 1))
        assert worst == best
    else:
        assert worst >= sum(cost([k]) for k in range(n)) / world


@settings(max_examples=120, deadline=None)
@given(st.integers(20, 3000), st.integers(10, 600), st.integers(0, 9), st.integers(1, 8))
def test_plane_transfers_deliver_exactly_what_is_missing(n_planes, pitch, overlap, world):
    zb = _rows(n_planes, pitch, overlap)
    held = mg.slab_bounds(n_planes, world)
    rows = mg.assign_chunk_rows_balanced(zb, world)
    wanted = []
    for r in range(world):
        wanted.append((min(zb[k][0] for k in rows[r]), max(zb[k][1] for k in rows[r]))
                      if rows[r] else (held[r][0], held[r][0]))
    plan = mg.transfer_plan(held, wanted)
    for r in range(world):
        have = np.zeros(n_planes, dtype=np.int32)
        have[held[r][0]:held[r][1]] += 1
        for src, dst, z0, z1 in plan:
            assert src != dst and held[src][0] <= z0 < z1 <= held[src][1]
            if dst == r:
                have[z0:z1] += 1
        w0, w1 = wanted[r]
        assert np.all(have[w0:w1] == 1)                       # every wanted plane, once
        assert mg.wanted_range(rows[r], zb, held[r])[0] <= w0 or not rows[r]


@settings(max_examples=80, deadline=None)
@given(st.integers(100, 3000), st.integers(50, 600), st.integers(1, 8), st.integers(1, 6),
       st.booleans())
def test_lending_never_raises_the_largest_load_and_boxes_cover_the_units(n_planes, pitch, world,
                                                                          n_cols, with_x):
    zb = _rows(n_planes, pitch, 5)
    yb = _rows(n_cols * 300 - 17, 300, 5)
    xb = _rows(800, 300, 5) if with_x else None
    held = mg.slab_bounds(n_planes, world)
    rows = mg.assign_chunk_rows_balanced(zb, world)
    loans = mg.loan_units(rows, zb, yb, xb)
    assert loans == sorted(loans) and len({(k, j) for k, j, _, _ in loans}) == len(loans)

    def weight(k, j):
        area = (zb[k][1] - zb[k][0]) * (yb[j][1] - yb[j][0])
        if xb is None:
            return float(area)
        return sum(area * (b - a) + mg.CHUNK_OVERHEAD_VOXELS for a, b in xb)
    load = [sum(weight(k, j) for k in rows[r] for j in range(len(yb))) for r in range(world)]
    before = max(load)
    for k, j, owner, worker in loans:
        assert k in rows[owner] and owner != worker and 0 <= worker < world
        load[owner] -= weight(k, j)
        load[worker] += weight(k, j)
    assert max(load) <= before + 1e-6
    plan = mg.box_transfer_plan(loans, zb, yb, held)
    for k, j, _, worker in loans:
        pieces = sorted((z0, z1) for s, d, kk, jj, z0, z1 in plan if (kk, jj, d) == (k, j, worker))
        assert pieces and pieces[0][0] == zb[k][0] and pieces[-1][1] == zb[k][1]
        assert all(a[1] == b[0] for a, b in zip(pieces[:-1], pieces[1:]))
    for src, dst, k, j, z0, z1 in plan:
        assert held[src][0] <= z0 < z1 <= held[src][1]


@settings(max_examples=150, deadline=None)
@given(st.integers(30, 5000), st.integers(1, 8), st.integers(1, 40), st.integers(1, 40))
def test_seamless_plan_halos(n_planes, world, block_depth, halo):
    own, ext = mg.seamless_plan(n_planes, world, block_depth, halo)
    assert own[0][0] == 0 and own[-1][1] == n_planes
    for (a, b), (ea, eb) in zip(own, ext):
        if a == b:
            continue
        assert ea <= a and eb >= b and 0 <= ea and eb <= n_planes
        assert ea % block_depth == 0 and (eb % block_depth == 0 or eb == n_planes)
        assert a - ea >= halo or ea == 0
        assert eb - b >= halo or eb == n_planes
