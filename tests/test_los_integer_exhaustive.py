"""Independent check of the line-of-sight / collision predicate (N2 / N4).

Kernel and oracle state the predicate the same way — four edge tests per rectangle, each two pairs of fp32
orientation signs with the half-open rule — so a shared spec bug would go unnoticed.  Here the SAME question
("does the centre-to-centre segment count this obstruction?") is answered by a different formulation in exact
integer arithmetic and compared exhaustively:

  * coordinates doubled: cell centres become odd integers, rectangle sides even integers, so no endpoint ever lies
    on a side's line and every orientation determinant is an exact integer;
  * separating-axis test of segment vs rectangle: the obstruction is crossed iff the x-extents overlap, the
    y-extents overlap, the rectangle's corners are not all on one side of the segment's line — and the endpoints
    are not BOTH inside (a segment inside the rectangle crosses no side);
  * a corner exactly ON the segment's line is put on the right of p->q: the half-open rule of the spec ("> 0"
    on one side) is exactly what an infinitesimal shift of the segment to its left produces
    (orient(p + eps n, q + eps n, c) = orient(p, q, c) - eps |pq|^2), so grazing a corner counts or not depending
    on which side the rectangle lies — asserted below on hand cases, too.

No per-edge test, no fp32, no shared code with oracle/rt_oracle.c or csrc/rt_device.cuh.
"""
import numpy as np
import pytest

from oracle import oracle as O


def sat_blocked(px, py, qx, qy, x0, y0, x1, y1):
    """Integer separating-axis form.  p, q: int64 arrays of DOUBLED centre coordinates (odd); rectangle sides doubled
    (even).  Returns bool array."""
    overlap_x = (np.minimum(px, qx) < x1) & (np.maximum(px, qx) > x0)
    overlap_y = (np.minimum(py, qy) < y1) & (np.maximum(py, qy) > y0)
    dx, dy = qx - px, qy - py
    left = np.zeros(px.shape, np.int64)
    for cx, cy in ((x0, y0), (x1, y0), (x1, y1), (x0, y1)):
        left += (dx * (cy - py) - dy * (cx - px)) > 0  # corner strictly to the left of p->q (on-line counts as right)
    straddles = (left > 0) & (left < 4)
    p_in = (px > x0) & (px < x1) & (py > y0) & (py < y1)
    q_in = (qx > x0) & (qx < x1) & (qy > y0) & (qy < y1)
    return overlap_x & overlap_y & straddles & ~(p_in & q_in)


def all_pairs(G):
    c = np.arange(G * G)
    p, q = np.meshgrid(c, c, indexing="ij")
    keep = p != q
    p, q = p[keep], q[keep]
    return p % G, p // G, q % G, q // G


def check_grid(G, max_side=None):
    pxc, pyc, qxc, qyc = all_pairs(G)
    segs = np.stack([pxc + 0.5, pyc + 0.5, qxc + 0.5, qyc + 0.5], 1).astype(np.float32)
    px, py, qx, qy = (2 * v.astype(np.int64) + 1 for v in (pxc, pyc, qxc, qyc))
    n_rect = n_cases = n_blocked = n_graze = 0
    top = G if max_side is None else min(G, max_side)
    for w in range(1, top + 1):
        for h in range(1, top + 1):
            for x0 in range(0, G - w + 1):
                for y0 in range(0, G - h + 1):
                    want = sat_blocked(px, py, qx, qy, 2 * x0, 2 * y0, 2 * (x0 + w), 2 * (y0 + h))
                    got = O.rect_blocks(segs, [x0, y0, x0 + w, y0 + h]).astype(bool)
                    if not np.array_equal(got, want):
                        i = int(np.flatnonzero(got != want)[0])
                        raise AssertionError(f"G={G} rect=({x0},{y0},{x0 + w},{y0 + h}) seg={segs[i].tolist()}: "
                                             f"edge tests say {bool(got[i])}, separating-axis says {bool(want[i])}")
                    n_rect += 1
                    n_cases += want.size
                    n_blocked += int(want.sum())
    return n_rect, n_cases, n_blocked


@pytest.mark.parametrize("G", [2, 3, 4, 5, 6, 7, 8])
def test_every_agent_cell_source_cell_rectangle_small_grids(G):
    n_rect, n_cases, n_blocked = check_grid(G)
    assert n_rect == (G * (G + 1) // 2) ** 2 and n_cases == n_rect * G * G * (G * G - 1)
    assert 0 < n_blocked < n_cases


def test_every_agent_cell_source_cell_rectangle_grid_12():
    # every rectangle that fits a 12 x 12 grid: 6,084 rectangles x 20,592 ordered cell pairs = 125 M obstruction decisions
    n_rect, n_cases, n_blocked = check_grid(12)
    assert n_rect == 78 * 78 and n_cases == n_rect * 144 * 143 and 0 < n_blocked < n_cases


def test_corner_grazing_follows_the_half_open_rule():
    # rectangle [2,4] x [2,4]; the diagonal through its corner (2,4) only touches it
    seg_a = [[0.5, 5.5, 3.5, 2.5]]  # heading down-right along x + y = 6
    seg_b = [[3.5, 2.5, 0.5, 5.5]]  # the same segment, reversed
    rect = [2, 2, 4, 4]
    a = bool(O.rect_blocks(seg_a, rect)[0])
    b = bool(O.rect_blocks(seg_b, rect)[0])
    # (0.5,5.5)->(3.5,2.5) passes (2,4)?  x + y = 6 at every point: yes — but it ALSO cuts through the interior
    # ((3,3) lies on it), so both directions must report a crossing
    assert a and b
    # a true graze: the line x + y = 8 touches only the corner (4,4) of [2,4] x [2,4]
    g1 = [[2.5, 5.5, 5.5, 2.5]]
    g2 = [[5.5, 2.5, 2.5, 5.5]]
    r1, r2 = bool(O.rect_blocks(g1, rect)[0]), bool(O.rect_blocks(g2, rect)[0])
    # heading down-right the rectangle lies to the RIGHT of p->q: the on-line corner joins the right side, no straddle
    # -> not counted; reversed, the rectangle lies to the LEFT: the on-line corner is on the right, the others left
    # -> counted.  (Physically a measure-zero event; the point is that both implementations agree on it.)
    assert (r1, r2) == (False, True)
    for seg, want in ((g1, r1), (g2, r2)):
        s = np.asarray(seg[0])
        px, py, qx, qy = (np.asarray([int(round(2 * v))], np.int64) for v in s)
        assert bool(sat_blocked(px, py, qx, qy, 4, 4, 8, 8)[0]) == want


def test_segment_inside_or_leaving_a_rectangle():
    rect = [1, 1, 5, 5]
    assert not O.rect_blocks([[1.5, 1.5, 4.5, 3.5]], rect)[0]   # both centres inside: no side crossed
    assert O.rect_blocks([[1.5, 1.5, 6.5, 3.5]], rect)[0]       # leaves through one side
    assert O.rect_blocks([[0.5, 2.5, 5.5, 2.5]], rect)[0]       # straight through
    assert not O.rect_blocks([[0.5, 0.5, 5.5, 0.5]], rect)[0]   # passes below


@pytest.mark.gpu
def test_device_line_of_sight_and_refused_moves_match_the_integer_predicate():
    """The DEVICE against the integer separating-axis predicate (no oracle in between): fixed scenarios on a 12 x 12
    grid — every source cell x a spread of 1..5 rectangles — with 8 agents per env on assorted cells.  After every
    step LAST_LOS must equal the number of rectangles the integer predicate counts on agent -> source, and a move
    must have been refused exactly when the predicate says the centre-to-centre step crosses a rectangle."""
    import torch
    from rad_team_b200.envs import BatchedRadSearchEnv

    G, A = 12, 8
    rng = np.random.default_rng(2026)
    n_src = G * G
    E = n_src * 3                                   # every source cell, three rectangle sets each
    src = np.stack([np.arange(E) % n_src % G, np.arange(E) % n_src // G], 1).astype(np.int32)
    n_poly = rng.integers(1, 6, E).astype(np.int32)
    rects = np.zeros((E, 5, 4), np.int64)           # x0, y0, x1, y1 (cell units)
    edges = np.zeros((E, 20, 4), np.float32)
    for e in range(E):
        for p in range(n_poly[e]):
            while True:
                w, h = rng.integers(1, 6, 2)
                x0, y0 = rng.integers(0, G - w + 1), rng.integers(0, G - h + 1)
                if not (x0 <= src[e, 0] < x0 + w and y0 <= src[e, 1] < y0 + h):
                    break
            rects[e, p] = (x0, y0, x0 + w, y0 + h)
            X0, Y0, X1, Y1 = map(float, rects[e, p])
            edges[e, 4 * p:4 * p + 4] = [(X0, Y0, X1, Y0), (X1, Y0, X1, Y1), (X1, Y1, X0, Y1), (X0, Y1, X0, Y0)]
    start = np.zeros((E, A, 2), np.int32)
    for e in range(E):                              # starts outside every rectangle of the env
        k = 0
        while k < A:
            c = rng.integers(0, G, 2)
            if not any(r[0] <= c[0] < r[2] and r[1] <= c[1] < r[3] for r in rects[e, :n_poly[e]]):
                start[e, k] = c
                k += 1
    env = BatchedRadSearchEnv(n_envs=E, n_agents=A, grid=G, max_steps=40, fixed_scenario=1, poly_min=0, poly_max=5,
                              transmission=0.5, seed=3)
    env.set_scenario(src_xy=src, src_int=np.full(E, 1e6), bkg=np.full(E, 20.0), start_xy=start, n_poly=n_poly,
                     edges=edges)
    env.reset()

    def crossed(p, q):                              # [E, A] rectangles counted on p -> q, integer arithmetic
        px, py, qx, qy = (2 * v.astype(np.int64) + 1 for v in (p[..., 0], p[..., 1], q[..., 0], q[..., 1]))
        n = np.zeros(px.shape, np.int64)
        for k in range(5):
            r = 2 * rects[:, k][:, None, :]         # doubled, broadcast over agents
            hit = sat_blocked(px, py, qx, qy, r[..., 0], r[..., 1], r[..., 2], r[..., 3])
            n += hit & (k < n_poly)[:, None] & ((px != qx) | (py != qy))
        return n

    pos = start.copy()
    d = {1: (0, 1), 2: (1, 0), 3: (-1, 0), 4: (0, -1)}
    refused = moved = 0
    for t in range(30):
        act = rng.integers(1, 5, (E, A)).astype(np.int32)
        xy, _, _, _, _ = env.step(torch.as_tensor(act, device=env.device))
        step = np.array([d[a] for a in act.reshape(-1)], np.int32).reshape(E, A, 2)
        cand = np.clip(pos + step, 0, G - 1)
        block = crossed(pos, cand) > 0
        want = np.where(block[..., None], pos, cand)
        got = xy.cpu().numpy()
        assert np.array_equal(got, want), f"step {t}: positions differ from the integer predicate's"
        refused += int((block & (cand != pos).any(-1)).sum())
        moved += int((~block & (cand != pos).any(-1)).sum())
        pos = want
        srcb = np.broadcast_to(src[:, None, :], pos.shape)
        los = env.buffer("LAST_LOS").cpu().numpy()
        assert np.array_equal(los, crossed(pos, srcb).astype(np.uint8)), f"step {t}: LAST_LOS"
    assert refused > 1000 and moved > 10000 and env.errors() == 0


def test_unit_moves_cross_a_rectangle_iff_exactly_one_cell_is_inside():
    """The collision rule (N4) on the generated scenarios: for a move between neighbouring cells the edge tests reduce
    to "exactly one of the two cells lies inside the rectangle" — every rectangle and every unit move of a 12 x 12 grid."""
    G = 12
    moves = []
    for y in range(G):
        for x in range(G):
            for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
                if 0 <= x + dx < G and 0 <= y + dy < G:
                    moves.append((x, y, x + dx, y + dy))
    m = np.asarray(moves, np.int64)
    segs = (m + 0.5).astype(np.float32)
    for w in range(1, G + 1):
        for h in range(1, G + 1):
            for x0 in range(0, G - w + 1):
                for y0 in range(0, G - h + 1):
                    p_in = (m[:, 0] >= x0) & (m[:, 0] < x0 + w) & (m[:, 1] >= y0) & (m[:, 1] < y0 + h)
                    q_in = (m[:, 2] >= x0) & (m[:, 2] < x0 + w) & (m[:, 3] >= y0) & (m[:, 3] < y0 + h)
                    got = O.rect_blocks(segs, [x0, y0, x0 + w, y0 + h]).astype(bool)
                    assert np.array_equal(got, p_in ^ q_in), (x0, y0, w, h)
