"""Host-side work decomposition of the self-match (so_dso_place_recognition_b200/csrc/sc_match_tc.cu,
`launch_sc_match_tc_self`): the launcher's closed-form item count against an explicit enumeration of the lower block
triangle, the coverage argument of the transposed stores, and a model of the cost-weighted split over the CTA pairs.
No GPU: `sodso_debug_sc_self_items` is host code of the library."""
import numpy as np
import pytest

from so_dso_place_recognition_b200 import _native as N

QG, TILE = 4, 256


def _items(n, q0, q1):
    """(group, tile) pairs with tile_start <= group_end for the query groups of [q0, q1), x 2 channels"""
    tiles = (n + TILE - 1) // TILE
    cnt = 0
    for g in range(q0 // QG, (q1 + QG - 1) // QG):
        cnt += min((QG * g + QG - 1) // TILE + 1, tiles)
    return 2 * cnt


@pytest.mark.parametrize("n", [1, 3, 255, 256, 257, 1000, 1301, 5000, 12345, 50000])
def test_item_count_closed_form(n):
    f = N.lib().sodso_debug_sc_self_items
    assert f(n, 0, n) == _items(n, 0, n)
    for q0 in range(0, n, TILE * max(1, n // (4 * TILE))):          # chunk starts of the streamed path
        q1 = min(n, q0 + 512)
        assert f(n, q0, q1) == _items(n, q0, q1)
    assert f(n, 1, n) == -1 and f(n, 0, n + 1) == -1 and f(0, 0, 0) == -1


def test_streamed_chunks_partition_the_triangle():
    """chunks of queries (multiples of 256) run disjoint item sets whose union is the whole triangle"""
    n, f = 5000, N.lib().sodso_debug_sc_self_items
    bounds = [0, 256, 512] + list(range(1024, n, 512)) + [n]
    assert sum(f(n, a, b) for a, b in zip(bounds[:-1], bounds[1:])) == f(n, 0, n)


@pytest.mark.parametrize("n", [7, 300, 777, 1301])
def test_direct_and_transposed_entries_cover_the_matrix_once(n):
    """entry (i, j) is written directly iff tile_start(j) <= group_end(i); otherwise exactly the transposed store of
    (j, i) writes it -- the epilogue's rule (q0 / 256) * 256 > (row | 3)."""
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")           # i: query, j: DB row
    direct = (j // TILE) * TILE <= (i | (QG - 1))
    # a directly computed (query a, row b) also stores (b, a) when tile_start(a) > group_end(b)
    mirrored_from = direct.T & (((j // TILE) * TILE) > (i | (QG - 1)))       # position (i, j) written from (j, i)
    writes = direct.astype(int) + mirrored_from.astype(int)
    assert (writes == 1).all()


def test_cost_weighted_split_model():
    """model of item_at_cost: boundaries are monotone, start at 0, end at W, and balance cost within one item"""
    n, npairs, w0, w1 = 5000, 74, 11, 2
    groups = (n + QG - 1) // QG
    tiles = (n + TILE - 1) // TILE
    nt = np.array([min((QG * g + QG - 1) // TILE + 1, tiles) for g in range(groups)])
    T = np.concatenate([[0], np.cumsum(nt)])                                 # (group, tile) pairs before group g
    W = 2 * int(T[-1])
    assert W == N.lib().sodso_debug_sc_self_items(n, 0, n)

    def item_at(num):
        if num >= npairs:
            return W
        x = int(T[-1]) * (w0 + w1) * num // npairs
        g = int(np.searchsorted(T, x // (w0 + w1), side="right") - 1)
        rem = x - int(T[g]) * (w0 + w1)
        base = 2 * int(T[g])
        return base + rem // w0 if rem < nt[g] * w0 else base + int(nt[g]) + (rem - int(nt[g]) * w0) // w1

    b = [item_at(p) for p in range(npairs + 1)]
    assert b[0] == 0 and b[-1] == W and all(x <= y for x, y in zip(b[:-1], b[1:]))
    # cost of an item by its position: [ch0: nt items of weight w0][ch1: nt items of weight w1] per group
    cost = np.concatenate([np.r_[np.full(k, w0), np.full(k, w1)] for k in nt])
    share = np.array([cost[x:y].sum() for x, y in zip(b[:-1], b[1:])])
    assert share.max() - share.min() <= 2 * w0
