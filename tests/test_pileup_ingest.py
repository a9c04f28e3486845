"""Native pileup ingestion -- CPU tests.

1. libswbbam's writer and reader against the independent pure-Python BAM / BAI parser of oracle/bam_oracle.py (pysam / htslib
   are absent): records, bins, region queries with and without the index, counts, FASTA slices.
2. indelpost_b200.pileup.make_pileup (one columnar fetch + swb_pileup_columns) against the REFERENCE's own make_pileup
   (pileup.pyx:51-113, reached through oracle/ref_pileup_shim.pyx) on every locus of tests/loci.py: the reference reads the
   in-memory stub pysam, ours reads the BAM / FASTA files written from the same locus.  Every key of every read dict must
   be equal (the `read` objects are compared attribute by attribute, the Variant objects by chrom / pos / ref / alt).
3. the unmodified reference PIPELINE fed from the files through the native reader gives the outputs it gives in memory.
"""
import os
import random
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))

import bam_oracle  # noqa: E402
import loci as L  # noqa: E402
import refpipe  # noqa: E402
from indelpost_b200 import bamio, pileup  # noqa: E402

needs_ref = pytest.mark.skipif(not refpipe.available() or not any(f.startswith("refshim") for f in os.listdir(refpipe.REF_PIPELINE)),
                               reason="oracle/_ref_pipeline (with refshim) is not built")


def write_locus(tmp, locus, extra_refs=()):
    bam, fa = os.path.join(tmp, "locus.bam"), os.path.join(tmp, "locus.fa")
    seqs = {locus["chrom"]: locus["genome"]}
    seqs.update(extra_refs)
    bamio.write_fasta(fa, seqs)
    bamio.write_bam(bam, [(k, len(v)) for k, v in seqs.items()], locus["reads"])
    return bam, fa


def test_exports_and_layout():
    lib = bamio.load()
    for s in bamio.EXPORTS:
        assert hasattr(lib, s), s
    assert lib.swb_pileup_read_size() == bamio.PILEUP_READ_DTYPE.itemsize
    hdr = open(os.path.join(os.path.dirname(HERE), "include", "swbbam.h")).read()
    for s in bamio.EXPORTS:
        assert s + "(" in hdr, f"{s} is not declared in include/swbbam.h"
    # ... and every function the header declares is exported by the library
    import re
    declared = set(re.findall(r"\b(swb_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)))
    assert declared == set(bamio.EXPORTS), declared ^ set(bamio.EXPORTS)
    for s in declared:
        assert hasattr(lib, s), s


@pytest.mark.parametrize("seed,kind", [(7, "del"), (8, "spliced"), (9, "complex")])
def test_writer_reader_against_python_oracle(tmp_path, seed, kind):
    locus = L.make_locus(seed, kind=kind, ev_len=4, n_reads=180, n_rate=0.01)
    # flags the reader must carry through, an unplaced read and a read on a second contig
    rng = random.Random(seed)
    for r in locus["reads"]:
        r["is_duplicate"] = rng.random() < 0.1
        r["is_secondary"] = rng.random() < 0.05
    r0 = locus["reads"][0]
    other = dict(r0, query_name="other", reference_name="chrB", reference_start=10, reference_end=10 + r0["reference_end"] - r0["reference_start"])
    locus["reads"].append(other)
    bam, fa = write_locus(str(tmp_path), locus, {"chrB": "ACGT" * 200})
    text, refs, recs = bam_oracle.read_bam(bam)
    assert refs == [("chr1", len(locus["genome"])), ("chrB", 800)] and "SO:coordinate" in text
    assert [(r["tid"], r["pos"]) for r in recs] == sorted((r["tid"], r["pos"]) for r in recs)
    by = {r["query_name"]: r for r in locus["reads"]}
    for r in recs:
        s = by[r["name"]]
        assert (r["seq"], r["cigarstring"], r["pos"], r["end"], r["mapq"]) == (s["query_sequence"], s["cigarstring"], s["reference_start"], s["reference_end"], s["mapping_quality"])
        assert r["qual"] == list(s["query_qualities"])
        assert r["bin"] == bam_oracle.reg2bin(r["pos"], r["pos"] + max(1, r["reflen"]))
        assert bool(r["flag"] & 16) == s["is_reverse"] and bool(r["flag"] & 1024) == s["is_duplicate"] and bool(r["flag"] & 256) == s["is_secondary"]
    assert bam_oracle.read_fasta(fa) == {"chr1": locus["genome"], "chrB": "ACGT" * 200}
    # the index: every record's chunk lies in its bin, the linear index never points past a record of its window
    bai = bam_oracle.read_bai(bam + ".bai")
    assert len(bai) == 2 and sum(len(c) for c in bai[0][0].values()) >= 1
    f = bamio.AlignmentFile(bam)
    assert f.references == ("chr1", "chrB") and f.lengths == (len(locus["genome"]), 800) and f.has_index()
    noidx = os.path.join(str(tmp_path), "noidx.bam")
    os.link(bam, noidx)
    g = bamio.AlignmentFile(noidx)
    assert not g.has_index()
    for (beg, end) in [(0, 1 << 29), (1950, 2051), (1999, 2000), (0, 10), (2100, 2101), (3990, 4000), (1800, 1801)] + [(x, x + rng.randint(1, 300)) for x in rng.sample(range(1500, 2500), 12)]:
        exp = bam_oracle.overlapping(recs, 0, beg, end)
        for src in (f, g):
            b = src.fetch_columns("chr1", beg, end)
            assert [b.name(i) for i in range(len(b))] == [r["name"] for r in exp], (beg, end)
            assert src.count("chr1", beg, end) == len(exp)
            assert src.count("chr1", beg, end, read_callback="all") == sum(1 for r in exp if not r["flag"] & (4 | 256 | 512 | 1024))
    assert [s.query_name for s in f.fetch("chrB", 0, 800)] == ["other"]
    assert len(f.fetch_columns()) == len(recs)
    with pytest.raises(ValueError):
        f.fetch_columns("nope", 0, 10)
    # segments: pysam's attribute names
    b = f.fetch_columns("chr1", 1950, 2051)
    for seg in b:
        s = by[seg.query_name]
        for k in ("query_sequence", "cigarstring", "reference_start", "reference_end", "mapping_quality", "is_reverse", "is_duplicate", "is_secondary", "reference_name"):
            assert getattr(seg, k) == s[k], k
        assert seg.query_qualities == s["query_qualities"]
        words = bamio.parse_cigar(s["cigarstring"])
        lead = (words[0] >> 4) if words[0] & 15 == 4 else 0
        trail = (words[-1] >> 4) if words[-1] & 15 == 4 else 0
        assert seg.query_alignment_sequence == s["query_sequence"][lead: len(s["query_sequence"]) - trail]
    # packed read table == DNA_BASE_LUT codes of the bases
    table, off, ln = b.pack4()
    lut = np.full(256, 4, "u1")
    for k, c in enumerate("ACGT"):
        lut[ord(c)] = k
    for i in range(len(b)):
        codes = lut[np.frombuffer(b.sequence(i).encode(), "u1")]
        n = int(ln[i]); t = table[int(off[i]): int(off[i]) + (n + 1) // 2]
        un = np.empty(2 * len(t), "u1"); un[0::2] = t & 15; un[1::2] = t >> 4
        assert np.array_equal(un[:n], codes)
    # FASTA slices, clamped like pysam's
    fa_h = bamio.FastaFile(fa)
    gseq = locus["genome"]
    for (a, e) in [(0, 60), (59, 61), (100, 1234), (3990, 5000), (2000, 2000), (-5, 10)]:
        assert fa_h.fetch("chr1", a, e) == gseq[max(0, a): e]
    assert fa_h.fetch("chrB") == "ACGT" * 200 and fa_h.get_reference_length("chr1") == len(gseq)
    os.remove(fa + ".fai")                          # index rebuilt in memory
    assert bamio.FastaFile(fa).fetch("chr1", 777, 1999) == gseq[777:1999]


def test_reader_rejects_damage(tmp_path):
    locus = L.make_locus(3, n_reads=30)
    bam, _ = write_locus(str(tmp_path), locus)
    raw = bytearray(open(bam, "rb").read())
    bad = os.path.join(str(tmp_path), "bad.bam")
    raw[len(raw) // 2] ^= 0x55
    open(bad, "wb").write(raw)
    with pytest.raises(OSError):
        bamio.AlignmentFile(bad).fetch_columns()
    open(bad, "wb").write(b"not a bam at all")
    with pytest.raises(OSError):
        bamio.AlignmentFile(bad)
    with pytest.raises(OSError):
        bamio.write_bam(os.path.join(str(tmp_path), "no", "such", "dir.bam"), [("chr1", 10)], [])


def test_large_file_many_blocks(tmp_path):
    """40 k reads over 2 Mb: hundreds of BGZF blocks, several index levels; region queries against the linear scan"""
    rng = random.Random(5)
    glen = 2_000_000
    reads = []
    for k in range(40000):
        p = rng.randrange(0, glen - 200)
        cig = rng.choice(["150M", "70M5D80M", "60M40000N90M", "20S130M", "75M3I72M"])
        s = "".join(rng.choice("ACGT") for _ in range(150))
        reads.append(dict(query_name=f"q{k}", query_sequence=s, query_qualities=None, cigarstring=cig, reference_name="chr1", reference_start=p, mapping_quality=60, is_reverse=False))
    bam = os.path.join(str(tmp_path), "big.bam")
    bamio.write_bam(bam, [("chr1", glen)], reads)
    _, _, recs = bam_oracle.read_bam(bam)
    assert len(recs) == 40000 and recs[0]["qual"] is None
    f = bamio.AlignmentFile(bam)
    for _ in range(60):
        beg = rng.randrange(0, glen); end = beg + rng.choice([1, 50, 1000, 20000, 300000])
        exp = [r["name"] for r in bam_oracle.overlapping(recs, 0, beg, end)]
        b = f.fetch_columns("chr1", beg, end)
        assert [b.name(i) for i in range(len(b))] == exp, (beg, end)


def _cmp_variant(a, b):
    return (a.chrom, a.pos, a.ref, a.alt) == (b.chrom, b.pos, b.ref, b.alt)


def _cmp_pileups(ref_p, our_p, what):
    assert len(ref_p) == len(our_p), what
    for x, y in zip(ref_p, our_p):
        assert set(x) == set(y), (what, set(x) ^ set(y))
        for k in x:
            if k == "read":
                for a in ("query_name", "query_sequence", "cigarstring", "reference_start", "reference_end", "mapping_quality", "is_reverse"):
                    assert getattr(x[k], a) == getattr(y[k], a), (what, a)
                assert x[k].query_qualities == y[k].query_qualities
            elif k in ("I", "D"):
                assert len(x[k]) == len(y[k]), (what, k)
                for u, v in zip(x[k], y[k]):
                    assert u[:-1] == v[:-1], (what, k, u[:-1], v[:-1])
                    assert _cmp_variant(u[-1], v[-1]), (what, k)
            else:
                assert x[k] == y[k] and type(x[k]) is type(y[k]), (what, x["read_name"], k, x[k], y[k])


def _ref_pileup(locus, kw):
    indelpost, pysam, _, _ = refpipe.load()
    import refshim
    from indelpost.local_reference import UnsplicedLocalReference as RefULR

    fa, bam = refpipe.open_locus(locus)
    v = indelpost.Variant(locus["chrom"], locus["pos"], locus["ref"], locus["alt"], fa)
    u = RefULR(v.chrom, v.pos, fa.get_reference_length(v.chrom), kw["window"], fa)
    random.seed(99)
    return refshim.ref_make_pileup(v, bam, u, kw["exclude_duplicates"], kw["window"], kw["downsamplethresh"], kw["basequalthresh"]), v, fa


def _our_pileup(tmp, locus, kw, v_ref, fa_stub):
    indelpost = refpipe.load()[0]
    bam_p, fa_p = write_locus(tmp, locus)
    bam, fa = bamio.AlignmentFile(bam_p), bamio.FastaFile(fa_p)

    class Target:           # the target's coordinates with OUR FastaFile as its reference
        chrom, pos, reference = v_ref.chrom, v_ref.pos, fa

        @staticmethod
        def generate_equivalents():
            return v_ref.generate_equivalents()

    u = pileup.UnsplicedLocalReference(v_ref.chrom, v_ref.pos, fa.get_reference_length(v_ref.chrom), kw["window"], fa)
    random.seed(99)
    # the reference's Variant wants the (stub) FastaFile type: build it with the stub, compare by value
    factory = lambda c, p, r, a, ref, skip_validation=True: indelpost.Variant(c, p, r, a, fa_stub, skip_validation=True)  # noqa: E731
    return pileup.make_pileup(Target, bam, u, kw["exclude_duplicates"], kw["window"], kw["downsamplethresh"], kw["basequalthresh"], variant_factory=factory)


@needs_ref
def test_make_pileup_equals_reference(tmp_path):
    specs = L.parity_specs()
    n_reads = n_indel = n_spliced = 0
    for k, spec in enumerate(specs):
        locus = L.make_locus(**spec)
        rng = random.Random(k)
        for r in locus["reads"]:
            r["is_duplicate"] = rng.random() < 0.05
            r["is_secondary"] = rng.random() < 0.03
        # BAM order: the reference's stub iterates the list as given, so give it coordinate order like a real BAM
        locus["reads"].sort(key=lambda r: r["reference_start"])
        for excl in (True, False):
            kw = dict(exclude_duplicates=excl, window=locus["kwargs"]["window"], downsamplethresh=1000 if k % 3 else 60, basequalthresh=20 if k % 2 else 32)
            (ref_p, ref_sf), v, fa_stub = _ref_pileup(locus, kw)
            our_p, our_sf = _our_pileup(str(tmp_path), locus, kw, v, fa_stub)
            assert ref_sf == our_sf, (spec, kw)
            _cmp_pileups(ref_p, our_p, (spec, kw))
            n_reads += len(ref_p); n_indel += sum(len(r["I"]) + len(r["D"]) for r in ref_p); n_spliced += sum(r["is_spliced"] for r in ref_p)
    assert n_reads > 8000 and n_indel > 1500 and n_spliced > 300, (n_reads, n_indel, n_spliced)


@needs_ref
def test_make_pileup_odd_cigars(tmp_path):
    """CIGAR shapes the locus generator never makes: clips on both ends, =/X, two introns, insertion next to a deletion,
    hard clips, an event at the first aligned base, low qualities at the ends"""
    rng = random.Random(11)
    genome = "".join(rng.choice("ACGT") for _ in range(6000))
    pos = 3000
    shapes = ["10S60M2I40M3D28M10S", "5H20S100M30S", "50=1X99=", "40M200N40M300N70M", "30M1I1D119M", "1M2D149M", "75M75S", "75S75M", "148M2S",
              "20M5I20M5D20M400N85M", "60M10D10I80M", "3S10M1D10M1I10M1D116M", "150M", "100M50N50M"]
    reads = []
    for k in range(400):
        cig = shapes[k % len(shapes)]
        words = bamio.parse_cigar(cig)
        qlen = sum(w >> 4 for w in words if "MIDNSHP=X"[w & 15] in "MIS=X")
        rlen = sum(w >> 4 for w in words if "MIDNSHP=X"[w & 15] in "MDN=X")
        start = pos - rng.randint(5, 140)
        seq = "".join(rng.choice("ACGT") for _ in range(qlen)) if k % 3 else genome[start: start + qlen]
        q = L._quals(rng, qlen, low_rate=0.3 if k % 4 == 0 else 0.02)
        reads.append(dict(query_name=f"o{k}", query_sequence=seq, query_qualities=q, cigarstring=cig, reference_start=start, reference_end=start + rlen,
                          mapping_quality=rng.choice((0, 20, 60)), is_reverse=rng.random() < 0.5, reference_name="chr1", query_alignment_sequence=seq))
    reads.sort(key=lambda r: r["reference_start"])
    locus = dict(genome=genome, chrom="chr1", pos=pos, ref=genome[pos - 1: pos + 3], alt=genome[pos - 1], reads=reads, kwargs=dict(window=50))
    for thresh, window in ((20, 50), (36, 30), (5, 167)):
        kw = dict(exclude_duplicates=True, window=window, downsamplethresh=1000, basequalthresh=thresh)
        (ref_p, ref_sf), v, fa_stub = _ref_pileup(locus, kw)
        our_p, our_sf = _our_pileup(str(tmp_path), locus, kw, v, fa_stub)
        assert ref_sf == our_sf
        _cmp_pileups(ref_p, our_p, kw)
        assert len(ref_p) > 100


@needs_ref
def test_make_pileup_batch_matches_dict_pileup(tmp_path):
    locus = L.make_locus(21, kind="spliced", ev_len=2, n_reads=150)
    locus["reads"].sort(key=lambda r: r["reference_start"])
    kw = dict(exclude_duplicates=True, window=50, downsamplethresh=1000, basequalthresh=20)
    (_, _), v, fa_stub = _ref_pileup(locus, kw)
    our_p, _ = _our_pileup(str(tmp_path), locus, kw, v, fa_stub)
    bam_p, fa_p = os.path.join(str(tmp_path), "locus.bam"), os.path.join(str(tmp_path), "locus.fa")
    bam, fa = bamio.AlignmentFile(bam_p), bamio.FastaFile(fa_p)

    class Target:
        chrom, pos, reference = v.chrom, v.pos, fa
        generate_equivalents = staticmethod(v.generate_equivalents)

    u = pileup.UnsplicedLocalReference(v.chrom, v.pos, fa.get_reference_length(v.chrom), 50, fa)
    pb = pileup.make_pileup_batch(Target, bam, u, True, 50, 1000, 20)
    assert [pb.batch.name(i) for i in pb.keep] == [r["read_name"] for r in our_p]
    table, off, ln, idx = pb.read_table()
    assert list(idx) == list(pb.keep) and len(off) == len(pb.batch)


@needs_ref
@pytest.mark.parametrize("spec_idx", [0, 9, 14, 22, 31, 35])
def test_reference_pipeline_from_files(tmp_path, spec_idx):
    """the unmodified reference pipeline (oracle SSW = its own ssw.c) fed from BAM / FASTA files through the native reader
    gives the outputs it gives from the in-memory stub"""
    indelpost, pysam, _, _ = refpipe.load()
    locus = L.make_locus(**L.parity_specs()[spec_idx])
    locus["reads"].sort(key=lambda r: r["reference_start"])
    mem = refpipe.run_locus(locus)
    bam_p, fa_p = write_locus(str(tmp_path), locus)
    files = (bamio.AlignmentFile(bam_p), bamio.FastaFile(fa_p))
    assert refpipe.run_locus(locus, files=files) == mem


def test_long_reads_unplaced_reads_and_foreign_blocks(tmp_path):
    """records longer than a BGZF block (they span blocks), unplaced reads after the last contig, records without qualities
    or CIGAR, and a file whose blocks were written by another gzip implementation with a different block size"""
    rng = random.Random(17)
    big = "".join(rng.choice("ACGTN") for _ in range(200_000))
    reads = [
        dict(query_name="long1", query_sequence=big, query_qualities=array_of(rng, len(big)), cigarstring="1000S198000M1000S", reference_name="chr1", reference_start=500, mapping_quality=7, is_reverse=True),
        dict(query_name="short", query_sequence="ACGTACGTAC", query_qualities=None, cigarstring="10M", reference_name="chr1", reference_start=100_000, mapping_quality=60, is_reverse=False),
        dict(query_name="long2", query_sequence=big[:70_000], query_qualities=None, cigarstring="70000M", reference_name="chr1", reference_start=150_000, mapping_quality=60, is_reverse=False),
        dict(query_name="nocigar", query_sequence="ACGT", query_qualities=None, cigarstring=None, reference_name="chr1", reference_start=160_000, mapping_quality=0, is_reverse=False, is_unmapped=True),
        dict(query_name="unplaced", query_sequence="TTTTGGGG", query_qualities=None, cigarstring=None, reference_name=None, reference_start=-1, mapping_quality=0, is_reverse=False, is_unmapped=True),
    ]
    bam = os.path.join(str(tmp_path), "long.bam")
    bamio.write_bam(bam, [("chr1", 400_000)], reads, level=1)
    _, _, recs = bam_oracle.read_bam(bam)
    assert [r["name"] for r in recs] == ["long1", "short", "long2", "nocigar", "unplaced"]
    assert recs[0]["seq"] == big and recs[0]["end"] == 500 + 198_000 and recs[4]["tid"] == -1 and recs[4]["pos"] == -1
    f = bamio.AlignmentFile(bam)
    b = f.fetch_columns()
    assert len(b) == 5 and b.sequence(0) == big and b.qualities(1) is None and b.qualities(0) == reads[0]["query_qualities"]
    assert b.segment(3).cigarstring is None and b.segment(3).reference_end is None and b.segment(4).reference_name is None
    assert [s.query_name for s in f.fetch("chr1", 100_005, 100_006)] == ["long1", "short"]
    assert [s.query_name for s in f.fetch("chr1", 198_499, 198_501)] == ["long1", "long2"]
    assert [s.query_name for s in f.fetch("chr1", 198_500, 198_501)] == ["long2"]
    assert [s.query_name for s in f.fetch("chr1", 160_000, 160_001)] == ["long1", "long2", "nocigar"]       # a record without CIGAR covers one base
    assert f.count("chr1", 160_000, 160_001, read_callback="all") == 2
    # the same uncompressed stream re-blocked by Python's zlib into 4 kB blocks with the BC subfield (still valid BGZF)
    import gzip
    import struct
    import zlib
    raw = gzip.open(bam, "rb").read()
    out = bytearray()
    for o in range(0, len(raw), 4096):
        chunk = raw[o: o + 4096]
        co = zlib.compressobj(9, zlib.DEFLATED, -15)
        body = co.compress(chunk) + co.flush()
        out += struct.pack("<4BI2BH2BHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, len(body) + 25) + body + struct.pack("<II", zlib.crc32(chunk), len(chunk))
    out += bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0])
    other = os.path.join(str(tmp_path), "reblocked.bam")
    open(other, "wb").write(out)
    g = bamio.AlignmentFile(other)                      # no index: scans
    b2 = g.fetch_columns("chr1", 198_499, 198_501)
    assert [b2.name(i) for i in range(len(b2))] == ["long1", "long2"] and b2.sequence(0) == big
    assert len(g.fetch_columns()) == 5


def array_of(rng, n):
    import array
    return array.array("B", [rng.choice((2, 20, 30, 40)) for _ in range(n)])


def test_fetch_pack4_many_regions_equals_per_region_fetches(tmp_path):
    """swb_bam_fetch_pack4 (many regions, host threads, own file handle per thread) == one fetch_columns + pack4 per region with
    fetch_reads' filter applied; swb_fai_fetch_many == one fetch per slice"""
    rng = random.Random(23)
    lcs, reads, seqs = [], [], {}
    for k in range(24):
        lc = L.make_locus(600 + k, kind=("del", "ins", "spliced")[k % 3], ev_len=1 + k % 7, n_reads=60 + 5 * k, read_len=(100, 150, 151)[k % 3], n_rate=0.01)
        name = f"c{k:02d}"
        for r in lc["reads"]:
            r["reference_name"] = name
            r["is_duplicate"] = rng.random() < 0.1
            r["is_secondary"] = rng.random() < 0.05
        lc["reads"][0]["reference_start"] = 0                     # dropped by `and read.reference_start`
        lc["reads"][0]["reference_end"] = lc["reads"][0]["reference_end"] - lc["reads"][0]["reference_start"]
        seqs[name] = lc["genome"]; reads.extend(lc["reads"]); lcs.append((name, lc["pos"]))
    bam_p, fa_p = os.path.join(str(tmp_path), "m.bam"), os.path.join(str(tmp_path), "m.fa")
    bamio.write_fasta(fa_p, seqs)
    bamio.write_bam(bam_p, [(k, len(v)) for k, v in seqs.items()], reads)
    bam, fa = bamio.AlignmentFile(bam_p), bamio.FastaFile(fa_p)
    regions = [(name, pos - 51, pos + 50) for name, pos in lcs] + [("c03", 0, 10), ("c05", 3990, 4000), ("c07", 100, 101)]
    for excl_dup in (True, False):
        for threads in (1, 3, 8):
            pk = bam.fetch_pack4(regions, exclude=bamio.FSECONDARY | (bamio.FDUP if excl_dup else 0), need_cigar=True, drop_pos0=excl_dup, threads=threads)
            first = [0]; tabs = []; lens = []; poss = []; flags = []
            for (name, a, e) in regions:
                b = bam.fetch_columns(name, a, e)
                keep = ((b.flag & bamio.FSECONDARY) == 0) & (b.n_cigar > 0)
                if excl_dup:
                    keep &= ((b.flag & bamio.FDUP) == 0) & (b.pos != 0)
                t, o, ln = b.pack4()
                for i in keep.nonzero()[0]:
                    tabs.append(t[int(o[i]): int(o[i]) + (int(ln[i]) + 1) // 2]); lens.append(int(ln[i])); poss.append(int(b.pos[i])); flags.append(int(b.flag[i]))
                first.append(len(lens))
            assert pk.n_reads == len(lens) and list(pk.region_first) == first
            assert list(pk.read_len) == lens and list(pk.pos) == poss and list(pk.flag) == flags
            assert np.array_equal(pk.table, np.concatenate(tabs) if tabs else np.zeros(0, "u1"))
            assert np.array_equal(pk.read_off, np.concatenate([[0], np.cumsum((np.array(lens) + 1) // 2)[:-1]]))
            assert np.array_equal(pk.region_of_read(), np.repeat(np.arange(len(regions)), np.diff(first)))
    empty = bam.fetch_pack4([])
    assert empty.n_reads == 0 and list(empty.region_first) == [0]
    with pytest.raises(ValueError):
        bam.fetch_pack4([("nope", 0, 10)])
    slices = [(name, pos - 150, pos + 150) for name, pos in lcs] + [("c00", 3900, 5000), ("c01", -5, 3), ("c02", 10, 10)]
    blob, off = fa.fetch_many(slices)
    assert blob.tobytes() == b"".join(fa.fetch_bytes(*s) for s in slices)
    assert list(np.diff(off)) == [len(fa.fetch_bytes(*s)) for s in slices]


def test_bai_with_samtools_metadata(tmp_path):
    """an index as samtools writes it: the metadata pseudo-bin 37450 inside every reference and the trailing n_no_coor count
    (SAMv1 5.2) must not change any query"""
    import struct
    locus = L.make_locus(31, kind="ins", ev_len=5, n_reads=300)
    bam, _ = write_locus(str(tmp_path), locus)
    want = [s.query_name for s in bamio.AlignmentFile(bam).fetch("chr1", 1990, 2010)]
    raw = open(bam + ".bai", "rb").read()
    (n_ref,) = struct.unpack_from("<i", raw, 4)
    out = bytearray(raw[:8]); o = 8
    for _ in range(n_ref):
        (n_bin,) = struct.unpack_from("<i", raw, o); o += 4
        out += struct.pack("<i", n_bin + 1)
        for _ in range(n_bin):
            b, n_chunk = struct.unpack_from("<Ii", raw, o)
            out += raw[o: o + 8 + 16 * n_chunk]; o += 8 + 16 * n_chunk
        out += struct.pack("<IiQQQQ", 37450, 2, 0, 1 << 40, 300, 0)            # ref_beg, ref_end / n_mapped, n_unmapped
        (n_intv,) = struct.unpack_from("<i", raw, o)
        out += raw[o: o + 4 + 8 * n_intv]; o += 4 + 8 * n_intv
    out += struct.pack("<Q", 7)                                                 # n_no_coor
    open(bam + ".bai", "wb").write(out)
    f = bamio.AlignmentFile(bam)
    assert f.has_index() and [s.query_name for s in f.fetch("chr1", 1990, 2010)] == want
    # a truncated index is refused (the reader then scans): same answer
    open(bam + ".bai", "wb").write(out[: len(out) // 2])
    g = bamio.AlignmentFile(bam)
    assert not g.has_index() and [s.query_name for s in g.fetch("chr1", 1990, 2010)] == want


def test_random_records_all_bin_levels(tmp_path):
    """seeded fuzz: records with every CIGAR operation, IUPAC bases, arbitrary flags, on contigs up to 400 Mb (all six levels of the
    binning scheme; introns that push a record into a coarse bin) -- writer vs the Python parser, indexed region queries vs the
    linear scan rule"""
    rng = random.Random(2026)
    refs = [("big", 400_000_000), ("mid", 40_000_000), ("small", 20_000)]
    reads = []
    for k in range(6000):
        name, ln = refs[rng.choice((0, 0, 0, 1, 1, 2))]
        ops = []
        if rng.random() < 0.2:
            ops.append((rng.randint(1, 30), "S"))
        ops.append((rng.randint(1, 120), "M"))
        for _ in range(rng.randint(0, 4)):
            ops.append((rng.randint(1, 12) if rng.random() < 0.8 else rng.choice((5_000, 200_000, 3_000_000, 40_000_000)), rng.choice("IDN=XP" if rng.random() < 0.5 else "ID")))
            if ops[-1][1] not in "DN" and ops[-1][0] > 100:            # only reference-skipping operations get the long lengths
                ops[-1] = (rng.randint(1, 20), ops[-1][1])
            ops.append((rng.randint(1, 100), rng.choice("M=X")))
        if rng.random() < 0.2:
            ops.append((rng.randint(1, 30), "S"))
        if rng.random() < 0.05:
            ops = [(rng.randint(1, 5), "H")] + ops
        qlen = sum(n for n, o in ops if o in "MIS=X")
        rlen = sum(n for n, o in ops if o in "MDN=X")
        if rlen >= ln - 2:
            continue
        pos = rng.randrange(0, ln - rlen) if rng.random() < 0.7 else rng.choice((0, (1 << 14) - 1, 1 << 14, (1 << 17) - 3, (1 << 20) - 1, (1 << 23) - 50, (1 << 26) - 10)) % max(1, ln - rlen)
        seq = "".join(rng.choice("ACGTNRYKMSWBDHV=" if rng.random() < 0.05 else "ACGT") for _ in range(qlen))
        flag = rng.choice((0, 16, 1024, 256, 512, 2048, 99, 147, 1 | 64 | 32))
        reads.append(dict(query_name=f"f{k}", query_sequence=seq, query_qualities=None if k % 7 == 0 else array_of(rng, qlen), cigarstring="".join(f"{n}{o}" for n, o in ops),
                          reference_name=name, reference_start=pos, mapping_quality=rng.randrange(0, 61), flag=flag))
    bam = os.path.join(str(tmp_path), "fuzz.bam")
    n = bamio.write_bam(bam, refs, reads)
    _, prefs, recs = bam_oracle.read_bam(bam)
    assert prefs == refs and len(recs) == n
    by = {r["query_name"]: r for r in reads}
    levels = set()
    for r in recs:
        s = by[r["name"]]
        assert (r["seq"], r["cigarstring"], r["pos"], r["flag"], r["mapq"]) == (s["query_sequence"].upper(), s["cigarstring"], s["reference_start"], s["flag"], s["mapping_quality"])
        b = bam_oracle.reg2bin(r["pos"], r["pos"] + max(1, r["reflen"]))
        assert r["bin"] == b
        levels.add(sum(b >= x for x in (1, 9, 73, 585, 4681)))
    assert levels == {0, 1, 2, 3, 4, 5}, levels
    f = bamio.AlignmentFile(bam)
    whole = f.fetch_columns()
    assert [whole.name(i) for i in range(len(whole))] == [r["name"] for r in recs]
    assert [whole.cigarstring(i) for i in range(0, len(whole), 37)] == [recs[i]["cigarstring"] for i in range(0, len(recs), 37)]
    for q in range(300):
        t = rng.randrange(3)
        ln = refs[t][1]
        beg = rng.randrange(0, ln)
        end = min(ln, beg + rng.choice((1, 100, 20_000, 1 << 14, 1 << 20, 50_000_000)))
        if q % 10 == 0 and recs:
            r = recs[rng.randrange(len(recs))]; t, beg = r["tid"], max(0, r["pos"] + rng.randint(-3, 3)); end = beg + rng.randint(1, 5)
        exp = [r["name"] for r in bam_oracle.overlapping(recs, t, beg, end)]
        b = f.fetch_columns(refs[t][0], beg, end)
        assert [b.name(i) for i in range(len(b))] == exp, (refs[t][0], beg, end)
        assert f.count(refs[t][0], beg, end) == len(exp)


def test_records_with_aux_tags(tmp_path):
    """real BAMs carry optional fields after the qualities (NM, MD, RG, ...): the reader must step over them by block_size"""
    import gzip
    import struct
    import zlib
    locus = L.make_locus(41, kind="del", ev_len=3, n_reads=80)
    bam, _ = write_locus(str(tmp_path), locus)
    raw = gzip.open(bam, "rb").read()
    (l_text,) = struct.unpack_from("<i", raw, 4)
    o = 8 + l_text
    (n_ref,) = struct.unpack_from("<i", raw, o); o += 4
    for _ in range(n_ref):
        (ln,) = struct.unpack_from("<i", raw, o); o += 4 + ln + 4
    out = bytearray(raw[:o])
    k = 0
    while o < len(raw):
        (bs,) = struct.unpack_from("<i", raw, o)
        aux = b"NMC" + bytes([k % 7]) + b"MDZ" + (b"%dA%d" % (k, 150 - k)) + b"\0" + b"RGZgrp1\0" + b"XSi" + struct.pack("<i", -k) + b"ZBBs\x02\x00\x00\x00\x01\x00\x02\x00"
        out += struct.pack("<i", bs + len(aux)) + raw[o + 4: o + 4 + bs] + aux
        o += 4 + bs; k += 1
    blocks = bytearray()
    for p in range(0, len(out), 30000):
        chunk = bytes(out[p: p + 30000])
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = co.compress(chunk) + co.flush()
        blocks += struct.pack("<4BI2BH2BHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, len(body) + 25) + body + struct.pack("<II", zlib.crc32(chunk), len(chunk))
    blocks += bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0])
    tagged = os.path.join(str(tmp_path), "tagged.bam")
    open(tagged, "wb").write(blocks)
    a, b = bamio.AlignmentFile(bam).fetch_columns(), bamio.AlignmentFile(tagged).fetch_columns()
    assert len(a) == len(b) == k == 80
    for i in range(len(a)):
        assert (a.name(i), a.sequence(i), a.cigarstring(i), int(a.pos[i]), int(a.end[i]), int(a.flag[i])) == (b.name(i), b.sequence(i), b.cigarstring(i), int(b.pos[i]), int(b.end[i]), int(b.flag[i]))
        assert a.qualities(i) == b.qualities(i)
    _, _, recs = bam_oracle.read_bam(tagged)
    assert recs[5]["aux"].startswith(b"NMC") and len(recs) == 80
    sub = bamio.AlignmentFile(tagged).fetch_columns("chr1", 1999, 2000)
    assert len(sub) == len(bamio.AlignmentFile(bam).fetch_columns("chr1", 1999, 2000))


def test_make_pileup_reproduces_golden(tmp_path):
    """tests/golden/pileup_dicts.json.gz: 447 read dicts the REFERENCE's own make_pileup produced (tests/golden/make_pileup_golden.py);
    the native ingest reproduces every key of every dict from BAM / FASTA files, without the reference build being present"""
    import gzip
    import json

    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_pileup_golden as MG

    doc = json.load(gzip.open(os.path.join(HERE, "golden", "pileup_dicts.json.gz"), "rt"))
    assert len(doc) == len(MG.CASES)
    n = 0
    for entry, case in zip(doc, MG.CASES):
        assert entry["case"] == json.loads(json.dumps(case))
        locus = MG.prepare(case)
        bam_p, fa_p = write_locus(str(tmp_path), locus)
        bam, fa = bamio.AlignmentFile(bam_p), bamio.FastaFile(fa_p)
        rpos = entry["rpos"]

        class Target:
            chrom, pos, reference = locus["chrom"], locus["pos"], fa

            @staticmethod
            def generate_equivalents(p=rpos):
                return [type("V", (), {"pos": p})]

        u = pileup.UnsplicedLocalReference(locus["chrom"], locus["pos"], len(locus["genome"]), entry["window"], fa)
        random.seed(99)
        got, sf = pileup.make_pileup(Target, bam, u, case["excl"], entry["window"], case["down"], case["thresh"])
        assert sf == entry["sample_factor"], case
        got = json.loads(json.dumps([MG.plain(d) for d in got]))
        assert len(got) == len(entry["pileup"]), case
        for a, b in zip(got, entry["pileup"]):
            assert a == b, (case, a["read_name"], [k for k in b if a.get(k) != b[k]])
        n += len(got)
    assert n == 447
