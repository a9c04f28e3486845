"""CPU model of the 'sandwich' certificate for reads whose 8-bit pass is the final answer (44-84 bp at match 3, DESIGN.md §8
known gaps).  Not product code: it estimates how many pairs a fast path could certify and checks the claim against the oracle.

  U = plain Gotoh (what k_fast computes).  L = Gotoh with vertical gaps (F) switched off in HOT columns, hot = column whose
  U-maximum reaches 128 + go (only there can a lazy-F value fall into the window the signed exit test of ssw.c:311 mis-reads).
  L <= H(8-bit pass) <= U cell by cell; if the column maxima of L and U agree in every column and the smallest best row in the
  best column agrees, every forward output of the 8-bit pass equals Gotoh's.  Same on the reverse problem.
Usage: python tools/research/short_read_sandwich.py [n_pairs] [read_len_lo] [read_len_hi]"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np
import swbtest as T

TIGHT = os.environ.get("SANDWICH", "tight") == "tight"      # SANDWICH=hot: the column-flag variant


def sweep(read, ref, mat, go, ge, hot=None):
    """column-by-column Gotoh, ssw orientation (columns = window).  Returns (colmax, bestrow per column)."""
    m = len(read)
    H = np.zeros(m, dtype=np.int64); E = np.zeros(m, dtype=np.int64)
    idx = np.arange(m, dtype=np.int64)
    colmax = np.zeros(len(ref), dtype=np.int64); brow = np.zeros(len(ref), dtype=np.int64)
    sc = mat[:, read].astype(np.int64)                     # [n, m]
    for c, rb in enumerate(ref):
        diag = np.concatenate(([0], H[:-1])) + sc[rb]
        Hn = np.maximum(np.maximum(diag, E), 0)            # without the vertical gap
        if hot is None or not hot[c]:
            # F(i) = max_{k<i} Hn(k) - go - (i-1-k) ge   (a gap opened from a cell that itself came from F is dominated: go >= ge)
            t = np.maximum.accumulate(Hn + idx * ge)
            F = np.concatenate(([-10**9], t[:-1])) - go - (idx - 1) * ge
            Hn = np.maximum(Hn, F)
        E = np.maximum(E - ge, Hn - go)
        H = Hn
        colmax[c] = H.max(); brow[c] = int(np.argmax(H))
    return colmax, brow


def sweep_tight(read, ref, mat, go, ge):
    """L_tight, computed side by side with Gotoh (U): a vertical-gap CONTINUATION is dropped whenever the chain's value in the 8-bit
    pass -- which lies between L's and U's value of that chain -- can be in W = [128, 127 + go - ge], the only values the signed exit
    test of the 8-bit lazy-F loop (ssw.c:309-311) can mis-read as 'no lane needs F any more'.  A chain whose value is outside W is
    always seen correctly, and opens (H - go) are applied unconditionally (main loop / first lazy iteration).  E opens from H without F
    (ssw.c computes E before the lazy correction).  By induction over the cells L <= H(8-bit) <= U; no column flags needed."""
    m = len(read)
    HL = [0] * m; EL = [0] * m; HU = [0] * m; EU = [0] * m
    colmax = np.zeros(len(ref), dtype=np.int64); brow = np.zeros(len(ref), dtype=np.int64)
    sc = mat[:, read]
    lo, hi = 128, 127 + go - ge
    NEG = -10**9
    for c, rb in enumerate(ref):
        srow = sc[rb]
        hdL = hdU = 0; fL = fU = NEG; hpL = hpU = 0
        best = -1; bi = 0
        for i in range(m):
            s_ = int(srow[i])
            hnfL = max(hdL + s_, EL[i], 0)
            hU = max(hdU + s_, EU[i], 0)
            if i > 0:
                contL = fL - ge; contU = fU - ge
                if contL <= hi and contU >= lo:            # [contL, contU] meets W
                    contL = NEG
                fL = max(contL, hpL - go)
                fU = max(contU, hpU - go)
                hL = max(hnfL, fL); hU = max(hU, fU)
            else:
                hL = hnfL
            hdL = HL[i]; HL[i] = hL; hpL = hL
            hdU = HU[i]; HU[i] = hU; hpU = hU
            EL[i] = max(EL[i] - ge, hnfL - go)
            EU[i] = max(EU[i] - ge, hU - go)
            if hL > best:
                best = hL; bi = i
        colmax[c] = best; brow[c] = bi
    return colmax, brow


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 84
    b = T.make_pairs(n, read_len=(lo, hi), win_len=300, seed=77, grid=True, max_indel=10)
    ro, ao = T.oracle_parallel(b, threads=min(16, os.cpu_count() or 1))
    mat = b.mat.reshape(b.n, b.n).astype(np.int64)
    stats = dict(pairs=0, in_zone=0, fwd_pass=0, both_pass=0, unsound=0, gotoh_differs=0)
    for p in range(n):
        go, ge = int(b.gap_open[p]), int(b.gap_ext[p])
        if not go > ge:
            continue
        r, w = int(b.pair_read[p]), int(b.pair_win[p])
        read = b.reads[b.read_off[r]: b.read_off[r] + b.read_len[r]].astype(np.int64)
        ref = b.windows[b.win_off[w]: b.win_off[w] + b.win_len[w]].astype(np.int64)
        stats["pairs"] += 1
        U, Ur = sweep(read, ref, mat, go, ge)
        s = int(U.max())
        if s >= 255 - 2 or s < 128 + go + ge:              # escalates to 16 bits / provably safe already: other paths
            continue
        stats["in_zone"] += 1
        if TIGHT:
            Lc, Lr = sweep_tight(read, ref, mat, go, ge)
        else:
            Lc, Lr = sweep(read, ref, mat, go, ge, U >= 128 + go)
        cbest = int(np.argmax(U))
        ok_f = bool(np.array_equal(Lc, U) and Lr[cbest] == Ur[cbest])
        gotoh_f = (s, cbest, int(Ur[cbest]))
        real_f = (int(ro["score1"][p]), int(ro["ref_end1"][p]), int(ro["read_end1"][p]))
        if gotoh_f != real_f:
            stats["gotoh_differs"] += 1
        if not ok_f:
            continue
        stats["fwd_pass"] += 1
        if gotoh_f != real_f:
            stats["unsound"] += 1; print("UNSOUND forward", p, gotoh_f, real_f); continue
        # reverse problem (ssw.c:875-886): reversed read prefix against the reversed window prefix, stop at the first column reaching s
        rr = read[: gotoh_f[2] + 1][::-1]; wr = ref[: cbest + 1][::-1]
        U2, U2r = sweep(rr, wr, mat, go, ge)
        L2, L2r = sweep_tight(rr, wr, mat, go, ge) if TIGHT else sweep(rr, wr, mat, go, ge, U2 >= 128 + go)
        c2 = int(np.argmax(U2 >= s)) if (U2 >= s).any() else -1
        ok_r = c2 >= 0 and bool(np.array_equal(L2[: c2 + 1], U2[: c2 + 1]) and L2r[c2] == U2r[c2])
        if not ok_r:
            continue
        stats["both_pass"] += 1
        gotoh_r = (cbest - c2, gotoh_f[2] - int(U2r[c2]))
        real_r = (int(ro["ref_begin1"][p]), int(ro["read_begin1"][p]))
        if gotoh_r != real_r:
            stats["unsound"] += 1; print("UNSOUND reverse", p, gotoh_r, real_r)
    print(stats)
    z = max(1, stats["in_zone"])
    print("forward certified %.1f %%, forward+reverse certified %.1f %% of the pairs in the unsafe zone" % (100 * stats["fwd_pass"] / z, 100 * stats["both_pass"] / z))


if __name__ == "__main__":
    main()
