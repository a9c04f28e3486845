"""Test harness around the UNMODIFIED reference pipeline built by oracle/build_ref_pipeline.py (oracle/_ref_pipeline:
the reference's compiled modules + the stub pysam).  Test infrastructure: only tests/ and bench.py import it.

run_locus(locus, ssw_cls) runs `Variant` -> `VariantAlignment` -> `count_alleles` (all flag combinations) -> `phase`
exactly like docs/intro.rst does, with `ssw_cls` swapped in as `indelpost.localn.SSW` -- the one name through which
every caller builds its aligners (localn.pyx:464-467) -- and `random.seed(123)` first (the reference samples from the
global RNG, pileup.pyx:86-98)."""
from __future__ import annotations

import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PIPELINE = os.path.join(ROOT, "oracle", "_ref_pipeline")

_mods = None


def available() -> bool:
    return os.path.isdir(os.path.join(REF_PIPELINE, "indelpost"))


def load():
    """-> (indelpost, pysam, indelpost.localn, reference SSW class)"""
    global _mods
    if _mods is None:
        if not available():
            raise RuntimeError("oracle/_ref_pipeline is missing: run `python oracle/build_ref_pipeline.py` where /root/reference exists")
        if REF_PIPELINE not in sys.path:
            sys.path.insert(0, REF_PIPELINE)
        import indelpost
        import indelpost.localn as localn
        import pysam
        from indelpost.sswpy import SSW as RefSSW

        _mods = (indelpost, pysam, localn, RefSSW)
    return _mods


def recording(base_cls, calls, key=lambda s: s):
    """subclass of an SSW class that appends (ref, read, match, mismatch, go, ge, start_idx, end_idx, result tuple) per align()"""

    class RecordingSSW(base_cls):
        def __init__(self, match_score=2, mismatch_penalty=2):
            super().__init__(match_score, mismatch_penalty)
            self._rec_ms, self._rec_mm = match_score, mismatch_penalty
            self._rec_ref = self._rec_read = None

        def setReference(self, reference):
            self._rec_ref = key(reference)
            return super().setReference(reference)

        def setRead(self, read):
            self._rec_read = key(read)
            return super().setRead(read)

        def align(self, gap_open=3, gap_extension=1, start_idx=0, end_idx=0):
            out = super().align(gap_open=gap_open, gap_extension=gap_extension, start_idx=start_idx, end_idx=end_idx)
            calls.append((self._rec_ref, self._rec_read, self._rec_ms, self._rec_mm, int(gap_open), int(gap_extension), int(start_idx), int(end_idx), tuple(out)))
            return out

    return RecordingSSW


def open_locus(locus, bam_cls=None):
    indelpost, pysam, localn, _ = load()
    fa = pysam.FastaFile({locus["chrom"]: locus["genome"]})
    segs = [pysam.AlignedSegment(**r) for r in locus["reads"]]
    bam = (bam_cls or pysam.AlignmentFile)(segs, (locus["chrom"],))
    return fa, bam


def file_backed_bam(src, base_cls=None):
    """the stub's AlignmentFile TYPE (the compiled reference checks it) over a bamio.AlignmentFile's DATA: fetch / count go to
    the native reader, every record becomes a stub AlignedSegment carrying the attributes pysam would report"""
    pysam = load()[1]
    base = base_cls or pysam.AlignmentFile

    class FileBackedBam(base):
        def fetch(self, contig=None, start=None, stop=None, until_eof=False):
            for seg in src.fetch(contig, start, stop):
                yield pysam.AlignedSegment(**seg.as_dict())

        def count(self, contig=None, start=None, stop=None, read_callback="nofilter"):
            return src.count(contig, start, stop, read_callback=read_callback)

    return FileBackedBam


def open_locus_files(locus, src_bam, src_fa, bam_cls=None):
    """like open_locus, but the reads come out of a BAM file and the genome out of a FASTA file (indelpost_b200.bamio)"""
    pysam = load()[1]
    fa = pysam.FastaFile({locus["chrom"]: src_fa.fetch(locus["chrom"])}, filename=src_fa.filename)
    bam = (bam_cls or file_backed_bam(src_bam))([], src_bam.references)
    return fa, bam


def _variant_tuple(v):
    return None if v is None else (v.chrom, v.pos, v.ref, v.alt)


def analyse(locus, fa, bam):
    """the reference's public API on one locus -> comparable summary"""
    indelpost = load()[0]
    random.seed(123)
    v = indelpost.Variant(locus["chrom"], locus["pos"], locus["ref"], locus["alt"], fa)
    valn = indelpost.VariantAlignment(v, bam, **locus.get("kwargs", {}))
    out = {"counts": tuple(valn.count_alleles()),
           "counts_fwrv": tuple(valn.count_alleles(fwrv=True)),
           "counts_frag": tuple(valn.count_alleles(by_fragment=True)),
           "counts_qc": tuple(valn.count_alleles(three_class=True))}
    try:
        ph = valn.phase()
        out["phased"] = _variant_tuple(ph)
    except Exception as e:  # noqa: BLE001 - the reference may raise on degenerate loci; both arms must then raise alike
        out["phased"] = ("error", type(e).__name__)
    try:
        out["target_indel"] = _variant_tuple(valn.get_target_indel())
    except Exception as e:  # noqa: BLE001
        out["target_indel"] = ("error", type(e).__name__)
    try:
        contig = valn.get_contig()
        out["contig"] = None if contig is None or getattr(contig, "failed", False) else (contig.get_contig_seq(), contig.get_reference_seq())
    except Exception as e:  # noqa: BLE001
        out["contig"] = ("error", type(e).__name__)
    return out


import contextlib
import threading


@contextlib.contextmanager
def swapped(ssw_cls):
    """`indelpost.localn.SSW = ssw_cls` for the duration of the block (process-wide: set it ONCE around concurrent tasks)"""
    localn = load()[2]
    saved = localn.SSW
    localn.SSW = ssw_cls
    try:
        yield
    finally:
        localn.SSW = saved


class ThreadCalls:
    """list-like sink that keeps one call list per thread (for recording under the wave scheduler)"""

    def __init__(self):
        self._tl = threading.local()

    def start(self, dest):
        self._tl.dest = dest

    def append(self, x):
        self._tl.dest.append(x)


def run_locus(locus, ssw_cls=None, calls=None, swap=True, bam_cls=None, files=None):
    """-> summary dict; with `calls` (a list) every SW call is recorded into it.  swap=False: the caller has already
    installed the SSW class (refpipe.swapped) -- required when several loci run concurrently.  files=(bamio.AlignmentFile,
    bamio.FastaFile): read the locus from those files instead of the in-memory stub (bam_cls then wraps file_backed_bam)."""
    indelpost, pysam, localn, RefSSW = load()
    opener = (lambda: open_locus(locus, bam_cls)) if files is None else (lambda: open_locus_files(locus, files[0], files[1], bam_cls))
    if not swap:
        fa, bam = opener()
        return analyse(locus, fa, bam)
    cls = ssw_cls or RefSSW
    if calls is not None:
        cls = recording(cls, calls)
    with swapped(cls):
        fa, bam = opener()
        return analyse(locus, fa, bam)


def classify_call(call, locus):
    """which call site of SURVEY.md §3.2 issued a recorded call (by its arguments; compiled Cython leaves no Python frames)"""
    ref, read, ms, mm, go, ge, s0, e0, out = call
    L = len(read)
    if go == L and ge == L:
        return "is_perfect_match"
    if go == L:
        return "is_target_by_ssw.mut"
    g, pos = locus["genome"], locus["pos"]
    if locus.get("intron"):
        i0, i1 = locus["intron"]
        if len(ref) == 200 and ref == g[i0 - 100: i0] + g[i1: i1 + 100]:
            return "overhang.junction"
    if ref == g[max(0, pos - 100): pos + 100]:
        return "overhang.genome"
    if len(read) >= 200 and read not in locus.get("_read_set", ()):
        return "decompose_complex_variant"
    return "grid_or_localn.ref"
