"""Generate the golden fixtures under tests/golden/ from the reference's own code.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

Two fixture families:
  * pairs_*.npz   — encoded read x window pairs + every s_align field and CIGAR produced by the
                    UNMODIFIED reference ssw.c (oracle/_ref/libssw_ref.so, built by oracle/Makefile).
  * sswpy_api.json — ASCII-level cases run through the reference's own sswpy.SSW class (sswpy.pyx +
                    ssw.c cythonized in a scratch directory under /tmp; nothing is copied into the repo),
                    pinning DNA_BASE_LUT, the score matrix, start_idx/end_idx handling, mask_len and the
                    CIGAR string formatting of the Python boundary.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import swbtest as T  # noqa: E402

REFERENCE = os.environ.get("REFERENCE", "/root/reference")


def quiet_ref(batch):
    """run the compiled reference with its stderr warnings silenced"""
    sys.stderr.flush()
    saved = os.dup(2)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 2)
    try:
        return T.reference().align_batch(batch)
    finally:
        os.dup2(saved, 2)
        os.close(devnull)
        os.close(saved)


def save_pairs(name, batch):
    res, arena = quiet_ref(batch)
    # cross-check the restatement while we are here
    ro, ao = T.oracle().align_batch(batch)
    T.compare(ro, ao, res, arena, what=name)
    np.savez_compressed(
        os.path.join(HERE, f"pairs_{name}.npz"),
        reads=batch.reads, read_off=batch.read_off, read_len=batch.read_len,
        windows=batch.windows, win_off=batch.win_off, win_len=batch.win_len,
        pair_read=batch.pair_read, pair_win=batch.pair_win,
        gap_open=batch.gap_open, gap_ext=batch.gap_ext,
        ref_beg=batch.ref_beg if batch.ref_beg is not None else np.zeros(0, np.int32),
        ref_len=batch.ref_len if batch.ref_len is not None else np.zeros(0, np.int32),
        mask_len=batch.mask_len if batch.mask_len is not None else np.zeros(0, np.int32),
        mat=batch.mat, n=batch.n, score_size=batch.score_size, flag=batch.flag,
        filters=batch.filters, filterd=batch.filterd,
        results=res, cigars=arena,
    )
    print(f"pairs_{name}.npz: {batch.n_pairs} pairs, byte-mode {float((res['score1'] < 255).mean()):.2f}, "
          f"flag2 {int((res['flag'] == 2).sum())}, status!=0 {int((res['status'] != 0).sum())}")


def edge_cases():
    """hand-made corner cases (SURVEY.md §10.1 zero-score corner, N handling, tiny inputs, sub-ranges)"""
    A, Cc, G, Tt, N = 0, 1, 2, 3, 4
    rng = np.random.default_rng(77)
    reads, wins, pr, pw, go, ge, rb, rl = [], [], [], [], [], [], [], []

    def add(r, w, o=3, e=1, beg=0, ln=None):
        reads.append(np.asarray(r, dtype=np.int8)); wins.append(np.asarray(w, dtype=np.int8))
        pr.append(len(reads) - 1); pw.append(len(wins) - 1); go.append(o); ge.append(e)
        rb.append(beg); rl.append(len(w) - beg if ln is None else ln)

    add([A], [A]); add([A], [Cc]); add([A], [Cc] * 30)           # 1x1 match, all-mismatch (score 0 corner)
    add([A] * 20, [Cc] * 50); add([N] * 25, [A] * 40); add([A] * 25, [N] * 40)
    add([A, Cc, G, Tt] * 10, [A, Cc, G, Tt] * 25)                  # periodic: many ties
    add([A] * 60, [A] * 100); add([A] * 100, [A] * 60)             # homopolymers, read longer than window
    w = rng.integers(0, 4, 300).astype(np.int8)
    add(w[100:250], w); add(w[100:250], w, beg=50); add(w[100:250], w, beg=120, ln=100)   # sub-ranges (start_idx/end_idx)
    add(w[0:150], w); add(w[150:300], w)                          # alignments touching both window edges
    r = np.concatenate([w[40:100], w[130:220]]); add(r, w, 3, 1); add(r, w, 3, 0); add(r, w, 150, 1); add(r, w, 150, 150)
    r = np.concatenate([w[40:100], rng.integers(0, 4, 25).astype(np.int8), w[100:165]]); add(r, w, 3, 1); add(r, w, 5, 0); add(r, w, 0, 0); add(r, w, 1, 3)
    add(w[::-1][:120].copy(), w)                                   # unrelated read
    add(w[10:26], w); add(w[10:25], w); add(w[10:27], w)           # around one striped segment
    for L in (7, 8, 9, 15, 16, 17, 31, 32, 33, 84, 85, 86, 127, 128, 129):
        s = int(rng.integers(0, 300 - L)); add(w[s:s + L], w)
    b = T.batch_from_lists(reads, wins, pr, pw, go, ge, ref_beg=rb, ref_len=rl)
    return b


def sswpy_cases():
    """build the reference's sswpy in /tmp and record Alignment tuples for ASCII inputs"""
    tmp = tempfile.mkdtemp(prefix="sswpy_ref_")
    try:
        for f in ("sswpy.pyx", "ssw.c", "ssw.h", "sse2neon.h"):
            shutil.copy(os.path.join(REFERENCE, "indelpost", f), tmp)
        with open(os.path.join(tmp, "setup.py"), "w") as fh:
            fh.write(
                "from setuptools import setup, Extension\nfrom Cython.Build import cythonize\n"
                "setup(ext_modules=cythonize([Extension('sswpy', ['sswpy.pyx', 'ssw.c'], "
                "extra_compile_args=['-Wno-unused-function'])], language_level=3))\n"
            )
        subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=tmp, check=True, capture_output=True)
        sys.path.insert(0, tmp)
        import sswpy  # the reference's own module

        rng = np.random.default_rng(5)
        alpha = np.array(list("ACGT"))
        win = "".join(rng.choice(alpha, 400))
        cases = []

        def run(match, mismatch, ref, read, **kw):
            a = sswpy.SSW(match, mismatch)
            a.setReference(ref)
            a.setRead(read)
            try:
                out = list(a.align(**kw))
                err = None
            except ValueError as e:
                out, err = None, "ValueError"
            cases.append(dict(match=match, mismatch=mismatch, ref=ref, read=read, kw=kw, out=out, err=err))

        read = win[120:270]
        run(3, 2, win, read)
        run(2, 2, win, read)
        run(3, 2, win, read[:70] + read[71:])                       # 1-bp deletion (SURVEY §8c example shape)
        run(3, 2, win, read[:70] + "ACGTTGCA" + read[70:])          # 8-bp insertion
        run(3, 2, win, read.lower())
        run(3, 2, win, read.replace("T", "U"))
        run(3, 2, win, read[:50] + "NNRYK" + read[55:])             # N and IUPAC codes -> 4
        run(3, 2, win[:200] + "NNNN" + win[204:], read)
        run(3, 2, win, read, gap_open=5, gap_extension=0)
        run(3, 2, win, read[:60] + read[75:], gap_open=len(read) - 15, gap_extension=1)   # mut aligner style
        run(3, 2, win, read, gap_open=300, gap_extension=300)        # narrowed to uint8 (300 -> 44)
        run(3, 2, win, read, gap_open=256, gap_extension=1)          # 256 -> 0
        run(3, 2, win, read, start_idx=100)
        run(3, 2, win, read, start_idx=130, end_idx=260)
        run(3, 2, win, read, start_idx=0, end_idx=200)
        run(3, 2, win, read, start_idx=401)                          # ValueError
        run(3, 2, win, read, start_idx=-1)                           # ValueError
        run(3, 2, win, win[10:40])                                   # 30-bp read: byte mode, mask_len 15
        run(3, 2, win, win[10:30] + win[200:230])                    # chimeric: sub-optimal score
        run(3, 2, win.encode(), read.encode())                       # bytes inputs
        run(3, 2, "ACGT" * 30, "ACGT" * 8)
        run(3, 2, "A" * 50, "C" * 20)                                # score 0 corner
        with open(os.path.join(HERE, "sswpy_api.json"), "w") as fh:
            json.dump([{**c, "ref": c["ref"].decode() if isinstance(c["ref"], bytes) else c["ref"],
                        "read": c["read"].decode() if isinstance(c["read"], bytes) else c["read"],
                        "bytes_input": isinstance(c["ref"], bytes)} for c in cases], fh, indent=0)
        print(f"sswpy_api.json: {len(cases)} cases")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    assert T.have_ref(), "oracle/_ref/libssw_ref.so missing: run `make -C oracle` with /root/reference present"
    save_pairs("cfg2_150x400", T.make_pairs(400, 150, 400, seed=11))
    save_pairs("cfg2_shared_windows", T.make_pairs(400, 150, 400, seed=12, reads_per_window=50, n_rate=0.005))
    save_pairs("grid_mixed", T.make_pairs(500, (20, 150), (60, 400), seed=13, grid=True, n_rate=0.01, junk_tail=0.2, low_complexity=0.1))
    save_pairs("short_reads", T.make_pairs(500, (30, 100), 300, seed=14, grid=True, max_indel=20))
    save_pairs("cfg4_250x1000", T.make_pairs(120, 250, 1000, seed=15, grid=True, max_indel=40))
    save_pairs("cfg5_250x2000", T.make_pairs(60, 250, 2000, seed=16, max_indel=200))
    save_pairs("tiny", T.make_pairs(400, (1, 40), (1, 60), seed=17, grid=True, max_indel=3, win_n_rate=0.05, n_rate=0.05))
    save_pairs("ge0_byte_bug_zone", T.make_pairs(500, (60, 130), 300, seed=18, go=5, ge=0, max_indel=15))
    save_pairs("odd_gaps", T.make_pairs(300, (40, 200), (100, 500), seed=19, go=2, ge=2))
    save_pairs("other_matrix", T.make_pairs(200, (40, 200), (100, 500), seed=20, go=6, ge=2, match=1, mismatch=4))
    save_pairs("edge_cases", edge_cases())
    b = T.make_pairs(200, (60, 150), 400, seed=21)
    b.score_size = 0
    save_pairs("score_size0", b)
    b = T.make_pairs(200, (60, 150), 400, seed=22)
    b.score_size = 1
    save_pairs("score_size1", b)
    b = T.make_pairs(200, 150, 400, seed=23)
    b.flag = 0
    save_pairs("flag0", b)
    b = T.make_pairs(200, 150, 400, seed=24)
    b.flag = 8
    save_pairs("flag8", b)
    b = T.make_pairs(200, (40, 150), 400, seed=25)
    b.mask_len = np.full(200, 10, dtype=np.int32)
    save_pairs("mask_lt15", b)
    sswpy_cases()


if __name__ == "__main__":
    main()
