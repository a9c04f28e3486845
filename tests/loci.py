"""Synthetic indel loci for the pipeline-level parity tests and the loci/s bench leg.

A locus is a random genome, one planted event at `pos` (1-based position of the anchor base, VCF style) and a pileup
of reads as a mapper would report them: ALT reads carry the gap in their CIGAR, or are soft-clipped / aligned through
the event with mismatches (the reads the realignment stage exists for).  Kinds cover every Smith-Waterman call site
of the reference pipeline:

  del / ins        simple events (retarget grid, update_read_info, is_target_by_ssw)
  long_ins         insertions >= 20 bp: (3,0) first in the grid, window/3 recursion of retarget (pileup.pyx:715-732)
  complex          deletion + insertion at one position given as ONE complex variant: decompose_complex_variant
                   (variant.pyx:609) and the second-target path
  hidden           no read carries the event in its CIGAR (all ALT reads clipped / mismatched): the contig fails and the
                   whole pileup goes through grid_search (varaln.pyx:251-364)
  read_end         the only gapped evidence sits at read ends: read_end_evidence_only -> is_perfect_match
                   (varaln.pyx:451-470, 1228-1234)
  spliced          RNA-style locus next to an exon junction: spliced reads (N in the CIGAR) and unspliced overhang reads:
                   spliced windows (utilities.pyx:528-575)
  spliced_hidden   the same with the event hidden from the CIGARs: the contig fails next to an intron ->
                   check_overhangs / filter_spurious_overhangs (pileup.pyx:435-574, two 200-bp windows per overhang read)

Everything is seeded; the same (seed, spec) gives the same locus on every machine.  Reads are plain dicts (the keyword
arguments of the stub pysam's AlignedSegment, oracle/pysam_stub)."""
from __future__ import annotations

import array
import random

BASES = "ACGT"


def _rand_seq(rng, n):
    return "".join(rng.choice(BASES) for _ in range(n))


def _mutate(rng, seq, sub_rate, n_rate=0.0):
    if sub_rate <= 0 and n_rate <= 0:
        return seq
    s = list(seq)
    for x in range(len(s)):
        u = rng.random()
        if u < sub_rate:
            s[x] = rng.choice([b for b in BASES if b != s[x]])
        elif u < sub_rate + n_rate:
            s[x] = "N"
    return "".join(s)


def _quals(rng, n, low_rate=0.0):
    return array.array("B", [(rng.choice((5, 10, 15)) if rng.random() < low_rate else rng.choice((30, 35, 37, 40))) for _ in range(n)])


def _read(rng, k, seq, cigar, start0, ref_end, sub_rate, n_rate=0.0, low_rate=0.0, mapq=60):
    seq = _mutate(rng, seq, sub_rate, n_rate)
    return dict(query_name=f"r{k}", query_sequence=seq, query_qualities=_quals(rng, len(seq), low_rate), cigarstring=cigar, reference_start=start0,
                reference_end=ref_end, mapping_quality=mapq, is_reverse=rng.random() < 0.5, reference_name="chr1", query_alignment_sequence=seq)


def make_locus(seed, kind="del", n_reads=200, read_len=150, ev_len=1, ins_len=None, vaf=0.5, genome_len=4000, pos=2000, clip_frac=0.15, clip_flank=20,
               mismatch_frac=0.0, sub_rate=0.005, n_rate=0.0, low_qual_rate=0.0, window=50, intron_len=400, repeat_unit=None, low_mapq_frac=0.0):
    """-> dict(genome, chrom, pos, ref, alt, reads, window, kind, kwargs): `kwargs` are the VariantAlignment keyword arguments
    the locus is meant to be run with (window, downsample_threshold)."""
    rng = random.Random(seed)
    genome = _rand_seq(rng, genome_len)
    if repeat_unit:
        # put the event into a short tandem repeat: shiftable indels, equivalence handling, exact_match_for_shiftable
        rep = repeat_unit * (24 // len(repeat_unit))
        genome = genome[:pos] + rep + genome[pos + len(rep):]
    dl = il = 0
    if kind in ("del", "hidden_del", "read_end_del", "spliced", "spliced_hidden"):
        dl = ev_len
    elif kind in ("ins", "long_ins", "hidden_ins", "read_end_ins"):
        il = ev_len
    elif kind == "complex":
        dl, il = ev_len, (ins_len if ins_len is not None else max(1, ev_len // 2))
    else:
        raise ValueError(kind)
    if repeat_unit and il:
        ins = (repeat_unit * il)[:il]
    else:
        ins = _rand_seq(rng, il)
        while il and dl and ins[0] == genome[pos]:            # keep the complex event from collapsing into a shorter one
            ins = _rand_seq(rng, il)
    ref_allele = genome[pos - 1: pos + dl]
    alt_allele = genome[pos - 1] + ins
    hidden = kind.startswith("hidden") or kind == "spliced_hidden"
    spliced = kind.startswith("spliced")
    read_end = kind.startswith("read_end")

    # spliced loci: the event sits 2 bases before the end of the left exon; [istart0, iend0) is the intron (0-based)
    istart0 = pos + dl + 2 if spliced else None
    iend0 = istart0 + intron_len if spliced else None

    reads = []
    for k in range(n_reads):
        is_alt = rng.random() < vaf
        mapq = 0 if rng.random() < low_mapq_frac else 60
        left = rng.randint(10, read_len - 10 - il)                 # read bases up to and including the anchor
        if read_end and is_alt and rng.random() < 0.5:
            # gapped evidence only at read ends: flank shorter than 15 % of the read (gappedaln.pyx:121-124)
            short = rng.randint(6, max(7, int(0.12 * read_len)))
            left = short if rng.random() < 0.5 else read_len - il - short
        start0 = pos - left
        if spliced:
            u = rng.random()
            if u < 0.45:
                # spliced REF read across the junction: left exon | intron skipped | right exon
                a = rng.randint(20, read_len - 20)
                s0 = istart0 - a
                seq = genome[s0:istart0] + genome[iend0: iend0 + read_len - a]
                reads.append(_read(rng, k, seq, f"{a}M{intron_len}N{read_len - a}M", s0, iend0 + read_len - a, sub_rate, n_rate, low_qual_rate, mapq))
                continue
            if u < 0.65:
                # spliced read whose last bases were pushed into the intron by the mapper (the spurious overhang)
                a = rng.randint(read_len - 12, read_len - 3)
                s0 = istart0 - a
                seq = genome[s0:istart0] + genome[iend0: iend0 + read_len - a]
                reads.append(_read(rng, k, seq, f"{read_len}M", s0, s0 + read_len, sub_rate, n_rate, low_qual_rate, mapq))
                continue
            # else: unspliced (pre-mRNA / genomic) read, REF or ALT, handled below like a DNA read
        if is_alt:
            right = read_len - left - il
            seq = genome[start0:pos] + ins + genome[pos + dl: pos + dl + right]
            if dl and il:
                cigar = f"{left}M{il}I{dl}D{right}M"
            elif dl:
                cigar = f"{left}M{dl}D{right}M"
            else:
                cigar = f"{left}M{il}I{right}M"
            ref_end = pos + dl + right
            u = rng.random()
            short_flank = min(left, right) < clip_flank
            force_hide = hidden or (read_end and min(left, right) >= 0.15 * read_len)
            if force_hide or (short_flank and u < clip_frac):
                if u < mismatch_frac or (force_hide and rng.random() < 0.3 and il == 0):
                    # aligned straight through the event: mismatches after it, no gap in the CIGAR
                    cigar = f"{read_len}M"; ref_end = start0 + read_len
                elif right < left:
                    cigar = f"{left}M{il + right}S"; ref_end = pos
                else:
                    cigar = f"{left + il}S{right}M"; ref_start_new = pos + dl
                    start0_clipped = ref_start_new
                    reads.append(_read(rng, k, seq, cigar, start0_clipped, ref_end, sub_rate, n_rate, low_qual_rate, mapq))
                    continue
        else:
            seq = genome[start0: start0 + read_len]
            cigar = f"{read_len}M"
            ref_end = start0 + read_len
        reads.append(_read(rng, k, seq, cigar, start0, ref_end, sub_rate, n_rate, low_qual_rate, mapq))
    kwargs = dict(window=window)
    if n_reads > 1000:
        kwargs["downsample_threshold"] = n_reads
    return dict(genome=genome, chrom="chr1", pos=pos, ref=ref_allele, alt=alt_allele, reads=reads, kind=kind, seed=seed, kwargs=kwargs,
                n_reads=n_reads, read_len=read_len, intron=(istart0, iend0) if spliced else None)


# the parity suite: every call site, short and long events, 100-250 bp reads, shiftable indels, low-quality / N bases
def parity_specs():
    specs = []
    s = 1000
    def add(**kw):
        nonlocal s
        s += 1
        specs.append(dict(seed=s, **kw))
    for ev in (1, 2, 4, 7, 12, 19):
        add(kind="del", ev_len=ev, n_reads=120)
    for ev in (1, 3, 8, 15):
        add(kind="ins", ev_len=ev, n_reads=120)
    for ev in (20, 27, 40):
        add(kind="long_ins", ev_len=ev, n_reads=100)
    for ev, il in ((5, 3), (3, 6), (9, 2), (6, 6)):
        add(kind="complex", ev_len=ev, ins_len=il, n_reads=100)
    for ev in (1, 3, 6, 14):
        add(kind="hidden_del", ev_len=ev, n_reads=100)
    for ev in (2, 5, 11, 24):
        add(kind="hidden_ins", ev_len=ev, n_reads=100)
    for ev in (2, 5):
        add(kind="read_end_del", ev_len=ev, n_reads=120, vaf=0.6)
    for ev in (3, 6):
        add(kind="read_end_ins", ev_len=ev, n_reads=120, vaf=0.6)
    for ev in (1, 3, 8):
        add(kind="spliced", ev_len=ev, n_reads=150)
    for ev in (1, 2, 5, 9):
        add(kind="spliced_hidden", ev_len=ev, n_reads=150)
    add(kind="del", ev_len=2, n_reads=100, repeat_unit="AC")
    add(kind="ins", ev_len=3, n_reads=100, repeat_unit="CAG")
    add(kind="del", ev_len=1, n_reads=100, repeat_unit="A")
    add(kind="hidden_ins", ev_len=4, n_reads=100, repeat_unit="TG")
    add(kind="del", ev_len=3, n_reads=100, n_rate=0.01, low_qual_rate=0.05)
    add(kind="ins", ev_len=5, n_reads=100, n_rate=0.01, low_qual_rate=0.05, low_mapq_frac=0.1)
    add(kind="del", ev_len=6, n_reads=100, read_len=100)
    add(kind="ins", ev_len=4, n_reads=100, read_len=100)
    add(kind="hidden_del", ev_len=3, n_reads=80, read_len=75)
    add(kind="del", ev_len=10, n_reads=80, read_len=250, window=167, genome_len=6000, pos=3000)
    add(kind="ins", ev_len=12, n_reads=80, read_len=250, window=167, genome_len=6000, pos=3000)
    add(kind="complex", ev_len=12, ins_len=5, n_reads=80, read_len=250, window=167, genome_len=6000, pos=3000)
    add(kind="hidden_ins", ev_len=30, n_reads=80, read_len=250, window=167, genome_len=6000, pos=3000)
    add(kind="del", ev_len=1, n_reads=200, vaf=0.1)
    add(kind="ins", ev_len=2, n_reads=200, vaf=0.9)
    add(kind="del", ev_len=25, n_reads=100, clip_frac=0.6, clip_flank=40)
    add(kind="ins", ev_len=9, n_reads=100, clip_frac=0.6, clip_flank=40, mismatch_frac=0.3)
    return specs
