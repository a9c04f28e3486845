"""Record the REAL indelPost call stream: run the unmodified reference pipeline (VariantAlignment + count_alleles +
phase) on synthetic loci and log every Smith-Waterman call it issues, with the result the reference's own
sswpy/ssw.c returned.  Output: tests/golden/pipeline_calls.json (config-1 style fixture).

Run in the build container only (needs /root/reference):
    python tests/golden/make_pipeline_golden.py

How: the ten reference .pyx modules are copied to a scratch directory under /tmp and cythonized there, unmodified,
against a minimal stub `pysam` package (pysam itself is absent and cannot be installed offline; the stub provides
the handful of attributes indelPost touches: FastaFile.fetch/get_reference_length/references/filename,
AlignmentFile.fetch/count/references, AlignedSegment fields).  Nothing from the reference is copied into the repo.
The recorder is a Python subclass of the reference's `SSW` class swapped into `indelpost.localn` (through which
every caller builds its aligners, localn.pyx:464-467).
"""
import array
import json
import os
import random
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("REFERENCE", "/root/reference")

STUB = {
    "pysam/__init__.py": '''
import os
def get_include():
    return [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
from pysam.libcfaidx import FastaFile
from pysam.libcalignedsegment import AlignedSegment
from pysam.libcalignmentfile import AlignmentFile
from pysam.libcbcf import VariantFile, VariantRecord, VariantRecordFilter
''',
    "pysam/__init__.pxd": "",
    "pysam/libcfaidx.pxd": "cdef class FastaFile:\n    cdef public dict seqs\n    cdef public object filename\n",
    "pysam/libcfaidx.pyx": '''
cdef class FastaFile:
    def __init__(self, seqs, filename="synthetic.fa"):
        self.seqs = dict(seqs)
        self.filename = filename
    def fetch(self, reference=None, start=None, end=None):
        s = self.seqs[reference]
        start = 0 if start is None else max(0, start)
        end = len(s) if end is None else min(len(s), end)
        return s[start:end]
    def get_reference_length(self, reference):
        return len(self.seqs[reference])
    @property
    def references(self):
        return tuple(self.seqs.keys())
''',
    "pysam/libcalignedsegment.pxd": "cdef class AlignedSegment:\n    cdef public object query_name, query_sequence, query_qualities, cigarstring, reference_start, reference_end, mapping_quality, is_reverse, is_duplicate, is_secondary, is_supplementary, query_alignment_sequence, reference_name\n",
    "pysam/libcalignedsegment.pyx": '''
cdef class AlignedSegment:
    def __init__(self, **kw):
        self.is_duplicate = False; self.is_secondary = False; self.is_supplementary = False
        for k, v in kw.items():
            setattr(self, k, v)
''',
    "pysam/libcalignmentfile.pxd": "cdef class AlignmentFile:\n    cdef public list reads\n    cdef public tuple refs\n",
    "pysam/libcalignmentfile.pyx": '''
cdef class AlignmentFile:
    def __init__(self, reads, refs):
        self.reads = list(reads)
        self.refs = tuple(refs)
    @property
    def references(self):
        return self.refs
    def fetch(self, contig=None, start=None, stop=None, until_eof=False):
        for r in self.reads:
            if r.reference_name == contig and r.reference_start < stop and r.reference_end > start:
                yield r
    def count(self, contig=None, start=None, stop=None, read_callback="nofilter"):
        n = 0
        for r in self.reads:
            if r.reference_name == contig and r.reference_start < stop and r.reference_end > start:
                if read_callback == "all" and (r.is_duplicate or r.is_secondary):
                    continue
                n += 1
        return n
''',
    "pysam/libcbcf.pxd": "cdef class VariantRecord:\n    pass\ncdef class VariantRecordFilter:\n    pass\ncdef class VariantFile:\n    pass\n",
    "pysam/libcbcf.pyx": "cdef class VariantRecord:\n    pass\ncdef class VariantRecordFilter:\n    pass\ncdef class VariantFile:\n    pass\n",
}

SETUP = '''
from setuptools import setup, Extension
from Cython.Build import cythonize
import glob, os
exts = []
for p in sorted(glob.glob("pysam/*.pyx")):
    exts.append(Extension("pysam." + os.path.basename(p)[:-4], [p]))
for p in sorted(glob.glob("indelpost/*.pyx")):
    name = os.path.basename(p)[:-4]
    src = [p] + (["indelpost/ssw.c"] if name == "sswpy" else [])
    exts.append(Extension("indelpost." + name, src, include_dirs=["."], extra_compile_args=["-Wno-unused-function", "-w"]))
setup(ext_modules=cythonize(exts, language_level=3, include_path=["."], quiet=True))
'''


def build_reference(tmp):
    shutil.copytree(os.path.join(REFERENCE, "indelpost"), os.path.join(tmp, "indelpost"))
    for rel, txt in STUB.items():
        path = os.path.join(tmp, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as fh:
            fh.write(txt)
    with open(os.path.join(tmp, "setup.py"), "w") as fh:
        fh.write(SETUP)
    r = subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=tmp, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-3000:] + r.stderr[-6000:])
        raise SystemExit("reference build failed")


BASES = "ACGT"


def make_locus(rng, kind, n_reads=200, read_len=150, genome_len=4000, pos=2000, ev_len=1, vaf=0.5):
    """synthetic locus: random genome, one planted indel at `pos` (1-based anchor base), reads as an aligner would
    report them (ALT reads carry the gap in their CIGAR; 15 % of short-flank ALT reads are soft-clipped instead)"""
    genome = "".join(rng.choice(BASES) for _ in range(genome_len))
    if kind == "del":
        ref_allele = genome[pos - 1 : pos + ev_len]
        alt_allele = genome[pos - 1]
    else:
        ins = "".join(rng.choice(BASES) for _ in range(ev_len))
        ref_allele = genome[pos - 1]
        alt_allele = genome[pos - 1] + ins
    reads = []
    for k in range(n_reads):
        is_alt = rng.random() < vaf
        left = rng.randint(10, read_len - 10 - (ev_len if kind == "ins" else 0))      # bases up to and including the anchor
        start0 = pos - left                                                              # 0-based reference start
        if is_alt and kind == "del":
            right = read_len - left
            seq = genome[start0:pos] + genome[pos + ev_len : pos + ev_len + right]
            cigar = f"{left}M{ev_len}D{right}M"
            ref_end = pos + ev_len + right
            short = min(left, right)
            if short < 20 and rng.random() < 0.15:
                if right < left:
                    cigar = f"{left}M{right}S"; ref_end = pos
                else:
                    cigar = f"{left}S{right}M"; start0 = pos + ev_len
        elif is_alt and kind == "ins":
            right = read_len - left - ev_len
            seq = genome[start0:pos] + alt_allele[1:] + genome[pos : pos + right]
            cigar = f"{left}M{ev_len}I{right}M"
            ref_end = pos + right
            if right < 20 and rng.random() < 0.15:
                cigar = f"{left}M{ev_len + right}S"; ref_end = pos
        else:
            seq = genome[start0 : start0 + read_len]
            cigar = f"{read_len}M"
            ref_end = start0 + read_len
        seq = list(seq)
        for x in range(len(seq)):
            if rng.random() < 0.005:
                seq[x] = rng.choice([b for b in BASES if b != seq[x]])
        seq = "".join(seq)
        quals = array.array("B", [rng.choice((30, 35, 37, 40)) for _ in seq])
        reads.append(dict(query_name=f"r{k}", query_sequence=seq, query_qualities=quals, cigarstring=cigar, reference_start=start0,
                          reference_end=ref_end, mapping_quality=60, is_reverse=rng.random() < 0.5, reference_name="chr1",
                          query_alignment_sequence=seq))
    return genome, pos, ref_allele, alt_allele, reads


def main():
    tmp = tempfile.mkdtemp(prefix="indelpost_ref_")
    try:
        build_reference(tmp)
        sys.path.insert(0, tmp)
        import pysam
        import indelpost
        import indelpost.localn as localn
        from indelpost.sswpy import SSW as RefSSW

        calls = []
        seq_table, seq_index = [], {}

        def sid(s):
            s = s.decode() if isinstance(s, bytes) else s
            if s not in seq_index:
                seq_index[s] = len(seq_table)
                seq_table.append(s)
            return seq_index[s]

        class RecordingSSW(RefSSW):
            def __init__(self, match_score=2, mismatch_penalty=2):
                super().__init__(match_score, mismatch_penalty)
                self._ms, self._mm = match_score, mismatch_penalty
                self._ref_id = self._read_id = None

            def setReference(self, reference):
                self._ref_id = sid(reference)
                return super().setReference(reference)

            def setRead(self, read):
                self._read_id = sid(read)
                return super().setRead(read)

            def align(self, gap_open=3, gap_extension=1, start_idx=0, end_idx=0):
                out = super().align(gap_open=gap_open, gap_extension=gap_extension, start_idx=start_idx, end_idx=end_idx)
                calls.append(dict(locus=len(loci), ref=self._ref_id, read=self._read_id, match=self._ms, mismatch=self._mm, go=int(gap_open),
                                  ge=int(gap_extension), start_idx=int(start_idx), end_idx=int(end_idx), out=list(out)))
                return out

        localn.SSW = RecordingSSW          # every aligner is built by localn.make_aligner (localn.pyx:464-467)

        loci = []
        rng = random.Random(20260101)
        specs = [("del", 1, 200), ("ins", 8, 200), ("del", 12, 120), ("ins", 3, 150), ("del", 4, 100)]
        for kind, ev, nreads in specs:
            genome, pos, ref_a, alt_a, reads = make_locus(rng, kind, n_reads=nreads, ev_len=ev)
            fa = pysam.FastaFile({"chr1": genome})
            bam = pysam.AlignmentFile([pysam.AlignedSegment(**r) for r in reads], ("chr1",))
            random.seed(123)                                   # the reference samples from the global RNG
            n0 = len(calls)
            v = indelpost.Variant("chr1", pos, ref_a, alt_a, fa)
            valn = indelpost.VariantAlignment(v, bam)
            counts = valn.count_alleles()
            phased = valn.phase()
            loci.append(dict(kind=kind, ev_len=ev, n_reads=nreads, pos=pos, ref=ref_a, alt=alt_a, count_alleles=list(counts),
                             phased=[phased.chrom, phased.pos, phased.ref, phased.alt], n_calls=len(calls) - n0))
            print(f"locus {kind}{ev}: {nreads} reads -> count_alleles={counts}, phased={phased.pos}:{phased.ref}>{phased.alt}, SW calls={len(calls) - n0}")
        with open(os.path.join(HERE, "pipeline_calls.json"), "w") as fh:
            json.dump(dict(note="every SW call issued by the unmodified reference pipeline on synthetic loci, with the reference's own results",
                           loci=loci, seqs=seq_table, calls=calls), fh)
        print(f"pipeline_calls.json: {len(calls)} calls, {len(seq_table)} distinct sequences")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
