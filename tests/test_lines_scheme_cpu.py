"""CPU design check of the line-marching SpMM's addressing scheme (krylov_b200/csrc/kb_lines.cuh).

The kernel itself is tested on the GPU (tests/test_gpu_kernels.py::test_line_marching_spmm_*).  This
file restates, in NumPy and one CTA at a time, what the producer copies into a ring slot and which
shared-memory entries a consumer thread reads for each diagonal -- with NaN-poisoned slots, so a read
of data that was never loaded is caught -- and checks the result bit for bit against SciPy's
``csr_matvecs`` for both work-item orders, lines shorter and longer than a chunk, a truncated last
line and one line per item.  It pins the index arithmetic (slot layout, halo, item decomposition,
short last line) independently of any hardware."""
import numpy as np
import pytest

from krylov_b200 import stencils as st

NS = 4


def _items(nlines, ncol, ch, lpp):
    """kb_lines_item: (chunk, first line, end line) of every non-empty work item, in launch order"""
    if lpp == 0:
        for it in range(ncol * ((nlines + ch - 1) // ch)):
            r0 = (it // ncol) * ch
            yield it % ncol, r0, min(r0 + ch, nlines)
        return
    nplanes = (nlines + lpp - 1) // lpp
    gpp = (lpp + ch - 1) // ch
    for it in range(gpp * ncol * nplanes):
        zp, gc = it % nplanes, it // nplanes
        r0 = zp * lpp + (gc // ncol) * ch
        r1 = min(r0 + ch, (zp + 1) * lpp, nlines)
        if r1 > r0:
            yield gc % ncol, r0, r1


def _pattern(A):
    A = A.tocsr()
    n = A.shape[0]
    rows = np.repeat(np.arange(n), np.diff(A.indptr))
    offs = sorted(set((A.indices - rows).tolist()))
    assert len(offs) == 7
    masks = np.zeros(n, dtype=np.int64)
    cv = [None] * 7
    for j in range(A.nnz):
        d = offs.index(A.indices[j] - rows[j])
        masks[rows[j]] |= 1 << d
        assert cv[d] is None or cv[d] == A.data[j]  # constant diagonals
        cv[d] = A.data[j]
    return offs, masks, cv


def _simulate(A, k, X, TR, ch, planes_fastest):
    offs, masks, cv = _pattern(A)
    n_rows, n_cols = A.shape
    inner = offs[4] * k
    H = (inner + 1) & ~1
    L, Pz, N, Nx = offs[5] * k, offs[6] * k, n_rows * k, n_cols * k
    nlines, ncol = (N + L - 1) // L, (L + TR - 1) // TR
    ch = min(ch, nlines)
    tail = N - (nlines - 1) * L
    lpp = Pz // L if planes_fastest and Pz % L == 0 else 0
    slotlen = 3 * TR + 2 * H
    x = X.reshape(-1)
    y = np.full(N, np.nan)
    ring = np.full((NS, slotlen), np.nan)
    tid = np.arange(256)
    cnt = 0
    seen = np.zeros(N, dtype=np.int64)
    for c, r0, r1 in _items(nlines, ncol, ch, lpp):
        nload = r1 - r0 + 2
        for l in range(nload):
            slot = cnt % NS
            # producer: window with halo | chunk one plane below | chunk one plane above
            ring[slot, :] = np.nan
            e0 = (r0 - 1 + l) * L + c * TR
            inside = 1 <= l <= nload - 2
            for s0, ln, base in ((e0 - H, TR + 2 * H, 0), (e0 - Pz, TR if inside else 0, TR + 2 * H),
                                 (e0 + Pz, TR if inside else 0, 2 * TR + 2 * H)):
                lo, hi = max(s0, 0), min(s0 + ln, Nx)
                if hi > lo:
                    ring[slot, base + lo - s0: base + hi - s0] = x[lo:hi]
            # consumers: line rc = r0 + l - 2 from the previous / current / next slot
            if l >= 2:
                rc = r0 + l - 2
                limc = tail if rc == nlines - 1 else L
                wprev, wcur, wnext = ring[(cnt - 2) % NS], ring[(cnt - 1) % NS], ring[slot]
                for q in range(TR // 256):
                    pos = c * TR + tid + q * 256
                    ok = pos < limc
                    e = rc * L + pos
                    own = H + tid + q * 256
                    vals = [wcur[TR + 2 * H + tid + q * 256], wprev[own], wcur[own - inner], wcur[own],
                            wcur[own + inner], wnext[own], wcur[2 * TR + 2 * H + tid + q * 256]]
                    m = masks[np.where(ok, e // k, 0)]
                    s = np.zeros(256)
                    for d in range(7):
                        use = ok & (((m >> d) & 1) == 1)
                        assert not np.isnan(vals[d][use]).any(), "read of an entry that was never loaded"
                        s = np.where(use, s + cv[d] * np.where(use, vals[d], 0.0), s)
                    y[e[ok]] = s[ok]
                    seen[e[ok]] += 1
            cnt += 1
    assert np.all(seen == 1), "every entry is computed exactly once"
    return y.reshape(n_rows, k)


@pytest.mark.parametrize("planes_fastest", [True, False])
@pytest.mark.parametrize("grid,k,TR,ch,cut", [
    ((70, 5, 4), 16, 1024, 3, 0),    # line (1120 entries) longer than a chunk, second chunk mostly empty
    ((33, 4, 5), 32, 1024, 32, 7),   # truncated last plane / short last line
    ((40, 3, 3), 2, 512, 1, 0),      # line shorter than a chunk, one line per item
    ((64, 6, 3), 16, 512, 4, 0),     # two full 512-entry chunks per line
])
def test_addressing_scheme_reproduces_csr_matvecs(grid, k, TR, ch, cut, planes_fastest):
    A = st.to_scipy(st.stencil7_csr(*grid, coeffs=st.convdiff_coeffs()))
    if cut:
        A = A[:-cut, :-cut].tocsr()
    X = np.random.default_rng(0).standard_normal((A.shape[1], k))
    Y = _simulate(A, k, X, TR, ch, planes_fastest)
    np.testing.assert_array_equal(Y, A @ X)
