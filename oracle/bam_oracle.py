"""oracle/bam_oracle.py — independent pure-Python reader of BAM / BAI / FASTA (TEST INFRASTRUCTURE; parity unpinned by htslib).

pysam / htslib / samtools are absent from the image and cannot be installed offline, so the native reader and writer of
indelpost_b200/csrc/swbbam.c are checked against this second, deliberately different implementation of the SAM/BAM
specification (SAMv1 §4.1 BGZF = concatenated gzip members, §4.2 record layout, §5.2 BAI): Python's own `gzip` module
inflates the whole file (it knows nothing of BGZF's BC subfield), `struct` decodes the records one by one, and region
queries are a linear scan with the overlap rule spelled out.  Slow and simple on purpose.  Only tests/ import it.
"""
from __future__ import annotations

import gzip
import struct

NT16 = "=ACMGRSVTWYHKDBN"
OPS = "MIDNSHP=XB"


def read_bam(path):
    """-> (header_text, [(name, length)], [record dict ...]) in file order"""
    with gzip.open(path, "rb") as fh:
        data = fh.read()
    assert data[:4] == b"BAM\1", "not a BAM file"
    (l_text,) = struct.unpack_from("<i", data, 4)
    text = data[8: 8 + l_text].decode()
    o = 8 + l_text
    (n_ref,) = struct.unpack_from("<i", data, o); o += 4
    refs = []
    for _ in range(n_ref):
        (ln,) = struct.unpack_from("<i", data, o); o += 4
        name = data[o: o + ln - 1].decode(); o += ln
        (l_ref,) = struct.unpack_from("<i", data, o); o += 4
        refs.append((name, l_ref))
    recs = []
    while o < len(data):
        (bs,) = struct.unpack_from("<i", data, o); o += 4
        tid, pos, l_name, mapq, bin_, n_cig, flag, l_seq, ntid, npos, tlen = struct.unpack_from("<iiBBHHHiiii", data, o)
        p = o + 32
        name = data[p: p + l_name - 1].decode(); p += l_name
        cig = struct.unpack_from("<%dI" % n_cig, data, p); p += 4 * n_cig
        sq = data[p: p + (l_seq + 1) // 2]; p += (l_seq + 1) // 2
        seq = "".join(NT16[sq[k >> 1] >> 4] if k % 2 == 0 else NT16[sq[k >> 1] & 15] for k in range(l_seq))
        qual = data[p: p + l_seq]; p += l_seq
        reflen = sum(w >> 4 for w in cig if OPS[w & 15] in "MDN=X")
        recs.append(dict(tid=tid, pos=pos, mapq=mapq, bin=bin_, flag=flag, name=name, cigar=list(cig), seq=seq,
                         qual=None if (l_seq and qual[0] == 0xFF) else list(qual), next_tid=ntid, next_pos=npos, tlen=tlen,
                         cigarstring="".join(f"{w >> 4}{OPS[w & 15]}" for w in cig), reflen=reflen,
                         end=(pos + reflen) if (n_cig and not flag & 4) else None, aux=data[p: o + bs]))
        o += bs
    return text, refs, recs


def overlapping(recs, tid, beg, end):
    """htslib's region rule: pos < end and pos + max(1, reflen) > beg"""
    return [r for r in recs if r["tid"] == tid and r["pos"] < end and r["pos"] + (r["reflen"] or 1) > beg]


def reg2bin(beg, end):
    """SAMv1 §5.3"""
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


def read_bai(path):
    """-> [ {bin: [(beg, end) ...]}, [ioffset ...] ] per reference"""
    data = open(path, "rb").read()
    assert data[:4] == b"BAI\1"
    (n_ref,) = struct.unpack_from("<i", data, 4)
    o = 8
    out = []
    for _ in range(n_ref):
        (n_bin,) = struct.unpack_from("<i", data, o); o += 4
        bins = {}
        for _ in range(n_bin):
            b, n_chunk = struct.unpack_from("<Ii", data, o); o += 8
            bins[b] = [struct.unpack_from("<QQ", data, o + 16 * k) for k in range(n_chunk)]; o += 16 * n_chunk
        (n_intv,) = struct.unpack_from("<i", data, o); o += 4
        ioff = list(struct.unpack_from("<%dQ" % n_intv, data, o)); o += 8 * n_intv
        out.append((bins, ioff))
    return out


def read_fasta(path):
    seqs, name = {}, None
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith(">"):
            name = line[1:].split()[0]; seqs[name] = []
        elif name is not None:
            seqs[name].append(line)
    return {k: "".join(v) for k, v in seqs.items()}
