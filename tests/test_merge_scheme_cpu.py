"""CPU design check of the merge schedule's decomposition (csrc/kb_merge.cuh): tiles of T
consecutive nonzeros, the per-tile bookkeeping, group sizes, short tails finished by the tile
a row starts in, the carries of long rows and the second launch that finishes them.  A NumPy walk through the same index arithmetic must finish every row exactly once
and reproduce SciPy's product (bit for bit where every row is summed by one lane)."""
import numpy as np
import pytest
import scipy.sparse as sp


def tile_meta(rowptr, nnz, n_rows, T, TAIL):
    """kb_merge_tiles_kernel: per tile {first row it sums, last row, tail it finishes}."""
    def owner(key):
        if key <= 0:
            return 0
        if key >= nnz:
            return n_rows
        return int(np.searchsorted(rowptr, key, side="right")) - 1

    n_tiles = (nnz + T - 1) // T
    meta = []
    for t in range(n_tiles):
        a, b = t * T, min(t * T + T, nnz)
        r_lo, r_nx = owner(a), owner(b)
        r_last = r_nx - 1
        if r_nx < n_rows and rowptr[r_nx] < b:
            r_last = r_nx
        r_first = r_lo
        if rowptr[r_lo] < a:
            b0 = (rowptr[r_lo] // T + 1) * T
            if rowptr[r_lo + 1] - b0 <= TAIL:
                r_first = r_lo + 1
        tail = 0
        if (r_last >= r_first and rowptr[r_last + 1] > b and rowptr[r_last] >= a
                and rowptr[r_last + 1] - b <= TAIL):
            tail = rowptr[r_last + 1] - b
        maxlen = 0
        for r in range(r_first, r_last + 1):
            maxlen = max(maxlen, min(rowptr[r + 1], b + tail) - max(rowptr[r], a))
        nr = r_last - r_first + 1
        lg = min(range(6), key=lambda g: (-(-nr // (256 >> g)) * 125 + -(-int(maxlen) // (1 << g)) * 30
                                           + g * 20, g))  # kb_merge_lg's cost model
        meta.append((r_first, r_last, int(tail), lg))
    return meta


def merge_spmv(A, x, T, TAIL=None, nthreads=256, order=None):
    TAIL = T // 8 if TAIL is None else TAIL
    rowptr, cols, vals = A.indptr.astype(np.int64), A.indices, A.data
    n_rows, nnz = A.shape[0], A.nnz
    n_tiles = (nnz + T - 1) // T
    meta = tile_meta(rowptr, nnz, n_rows, T, TAIL)
    carry = np.full((n_tiles, 2), np.nan)
    y = np.full(n_rows, np.nan)
    done = np.zeros(n_rows, dtype=int)
    exact = np.ones(n_rows, dtype=bool)
    fix = []
    for tile in (order if order is not None else range(n_tiles)):
        a = tile * T
        cnt = min(T, nnz - a)
        b = a + cnt
        r_first, r_last, tail, lg = meta[tile]
        bt = b + tail
        assert tail == 0 or cnt == T
        prod = vals[a:bt] * x[cols[a:bt]]
        for row in range(r_first, r_last + 1):
            lo, hi = rowptr[row], rowptr[row + 1]
            jb, je = max(lo, a) - a, min(hi, bt) - a
            G = 1 << lg
            lane = np.zeros(G)
            for gl in range(G):
                for j in range(jb + gl, je, G):
                    lane[gl] += prod[j]
            if G > 1:
                exact[row] = False
            o = G >> 1
            while o:
                lane = lane + lane[np.arange(G) ^ o]
                o >>= 1
            s = lane[0]
            if lo < a:  # a long row's middle or last piece
                assert row == r_first and tail == 0 or row != r_last
                carry[tile, 0] = s
                if hi <= b:  # kb_merge_tiles_kernel's ends_long -> kb_merge_fixlist_kernel
                    fix.append((row, lo // T, tile))
            elif hi > bt:  # a long row's first piece
                assert row == r_last and tail == 0
                carry[tile, 1] = s
            else:
                y[row] = s
                done[row] += 1
    for row, t0, t1 in fix:  # kb_merge_fix_kernel: 32 strided partial sums + butterfly
        assert t1 > t0
        lane = np.zeros(32)
        lane[0] = carry[t0, 1]
        for l in range(32):
            for q in range(t0 + 1 + l, t1 + 1, 32):
                lane[l] += carry[q, 0]
        o = 16
        while o:
            lane = lane + lane[np.arange(32) ^ o]
            o >>= 1
        y[row] = lane[0]
        done[row] += 1
        exact[row] = False
    assert np.all(done == 1), "every row is finished exactly once"
    assert not np.any(np.isnan(y))
    return y, exact


def _cases():
    rng = np.random.default_rng(5)
    out = []
    out.append(("poisson1d", sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(300, 300)).tocsr()))
    out.append(("random_100", sp.random(400, 400, density=0.25, random_state=1, format="csr")))
    lens = np.minimum((rng.pareto(1.1, 500) * 3).astype(int) + 1, 450)
    rows = np.repeat(np.arange(500), lens)
    cols = np.concatenate([rng.choice(500, l, replace=False) for l in lens])
    out.append(("powerlaw", sp.csr_matrix((rng.standard_normal(len(rows)), (rows, cols)),
                                          shape=(500, 500))))
    # empty rows at both ends and in the middle, one dense row
    M = sp.lil_matrix((200, 200))
    M[5, :] = rng.standard_normal(200)
    for i in range(20, 120, 3):
        M[i, rng.choice(200, 4, replace=False)] = 1.5
    out.append(("empty_rows_dense_row", M.tocsr()))
    out.append(("single_dense_row", sp.csr_matrix(rng.standard_normal((1, 700)))))
    return out


@pytest.mark.parametrize("T", [16, 64, 256])
@pytest.mark.parametrize("name,A", _cases(), ids=[c[0] for c in _cases()])
def test_merge_decomposition(name, A, T):
    A = A.tocsr()
    A.sort_indices()
    x = np.random.default_rng(2).standard_normal(A.shape[1])
    y, exact = merge_spmv(A, x, T)
    ref = A @ x
    scale = abs(A) @ abs(x) + 1e-300
    assert np.all(np.abs(y - ref) <= 1e-13 * scale)
    # rows summed by a single lane (also across a tile boundary, through the tail) follow
    # csr_matvec's order exactly
    assert np.array_equal(y[exact], ref[exact])


def test_merge_any_tile_order():
    """No tile depends on another one (long rows are finished by the second launch), so the
    result is the same for every order in which CTAs happen to take the tiles."""
    A = _cases()[2][1].tocsr()
    A.sort_indices()
    x = np.random.default_rng(3).standard_normal(A.shape[1])
    n_tiles = (A.nnz + 31) // 32
    y0, _ = merge_spmv(A, x, 32)
    y1, _ = merge_spmv(A, x, 32, order=list(np.random.default_rng(0).permutation(n_tiles)))
    assert np.array_equal(y0, y1)
