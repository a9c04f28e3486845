"""Native pileup ingestion: ctypes host layer over libswbbam.so (include/swbbam.h).

What indelPost reaches through pysam in pileup.pyx:51-160 -- `AlignmentFile.fetch / count / references`, the
`AlignedSegment` attributes `dictize_read` reads, `FastaFile.fetch / get_reference_length / references / filename` -- under the
same names and argument meaning, on top of a C reader written from the SAM/BAM specification (pysam and htslib are not in
the image).  Beside the per-read objects the reference expects there is the columnar form a batched aligner wants:

    bam = AlignmentFile("sample.bam")                     # BAM + BAI
    batch = bam.fetch_columns("chr1", 1950, 2051)         # ReadBatch: one numpy array per field, arenas for bases / CIGARs
    table, off, length = batch.pack4()                    # SWB_SEQ_PACKED4 read table for swb_align_batch / BatchAligner
    for seg in bam.fetch("chr1", 1950, 2051): ...         # pysam-style AlignedSegment views of the same records

`write_bam` / `write_fasta` produce coordinate-sorted BAM + BAI and FASTA + FAI (the synthetic configs of BASELINE.json are
"written to BAM"; without pysam this is the writer).  No GPU code here; the library must be present (no Python fallback).
"""
from __future__ import annotations

import array
import ctypes as C
import os
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libswbbam.so")

FUNMAP, FREVERSE, FSECONDARY, FQCFAIL, FDUP, FSUPPLEMENTARY = 4, 16, 256, 512, 1024, 2048
CIGAR_OPS = "MIDNSHP=XB"
_OP_CODE = {c: i for i, c in enumerate(CIGAR_OPS)}

EXPORTS = (
    "swb_bam_last_error", "swb_bam_version", "swb_bam_open", "swb_bam_close", "swb_bam_n_ref", "swb_bam_ref_name", "swb_bam_ref_len",
    "swb_bam_tid", "swb_bam_header_text", "swb_bam_has_index", "swb_bam_fetch", "swb_bam_batch_free", "swb_bam_count",
    "swb_bam_batch_pack4", "swb_bam_batch_cigar_text", "swb_bam_fetch_pack4", "swb_bam_pack_free", "swb_fai_fetch_many", "swb_pileup_columns", "swb_pileup_cols_free", "swb_pileup_read_size",
    "swb_bam_create", "swb_bam_write", "swb_bam_writer_close", "swb_fasta_write", "swb_fai_open", "swb_fai_close", "swb_fai_n",
    "swb_fai_name", "swb_fai_len", "swb_fai_fetch",
)


class _CBatch(C.Structure):
    _fields_ = [
        ("n", C.c_int64),
        ("tid", C.c_void_p), ("pos", C.c_void_p), ("end", C.c_void_p), ("flag", C.c_void_p), ("mapq", C.c_void_p), ("l_seq", C.c_void_p),
        ("n_cigar", C.c_void_p), ("next_tid", C.c_void_p), ("next_pos", C.c_void_p), ("tlen", C.c_void_p),
        ("name_off", C.c_void_p), ("seq_off", C.c_void_p), ("cigar_off", C.c_void_p),
        ("names", C.c_void_p), ("names_len", C.c_int64),
        ("seq", C.c_void_p), ("seq_len", C.c_int64),
        ("qual", C.c_void_p),
        ("seq4", C.c_void_p), ("seq4_len", C.c_int64), ("seq4_off", C.c_void_p),
        ("cigar", C.c_void_p), ("cigar_len", C.c_int64),
    ]


class _CPack(C.Structure):
    _fields_ = [
        ("n_regions", C.c_int64), ("n_reads", C.c_int64), ("region_first", C.c_void_p), ("read_off", C.c_void_p), ("read_len", C.c_void_p),
        ("table", C.c_void_p), ("table_len", C.c_int64), ("pos", C.c_void_p), ("end", C.c_void_p), ("flag", C.c_void_p), ("mapq", C.c_void_p),
    ]


class _CCols(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("reads", C.c_void_p),
        ("subreads", C.c_void_p), ("n_subreads", C.c_int64),
        ("indels", C.c_void_p), ("n_indels", C.c_int64),
        ("ref_seq", C.c_void_p), ("ref_seq_len", C.c_int64), ("ref_seq_off", C.c_void_p),
    ]


# swb_pileup_read (include/swbbam.h)
PILEUP_READ_DTYPE = np.dtype(
    [
        ("aln_start", "<i4"), ("start_offset", "<i4"), ("read_start", "<i4"), ("aln_end", "<i4"), ("end_offset", "<i4"), ("read_end", "<i4"),
        ("low_qual_base_num", "<i4"),
        ("is_end_dirty", "u1"), ("is_dirty", "u1"), ("is_covering", "u1"), ("is_spliced", "u1"),
        ("covering_start", "<i4"), ("covering_end", "<i4"), ("intron_start", "<i4"), ("intron_end", "<i4"),
        ("n_subreads", "<i4"), ("subread_off", "<i8"), ("n_ins", "<i4"), ("n_del", "<i4"), ("indel_off", "<i8"),
        ("is_reference_seq", "u1"), ("n_count_gt1", "u1"), ("pad_", "u1", (2,)), ("splice_pos", "<i4"),
    ],
    align=True,
)
PILEUP_INDEL_DTYPE = np.dtype([("pos", "<i4"), ("len", "<i4"), ("read_split", "<i4"), ("ref_split", "<i4")])

_lib = None


def load():
    """the loaded libswbbam.so (raises when it is missing: there is no Python implementation behind it)"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C indelpost_b200/csrc`")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, cp = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_char_p
    sig = {
        "swb_bam_last_error": (cp, []), "swb_bam_version": (cp, []),
        "swb_bam_open": (vp, [cp]), "swb_bam_close": (None, [vp]), "swb_bam_n_ref": (i32, [vp]), "swb_bam_ref_name": (cp, [vp, i32]),
        "swb_bam_ref_len": (i64, [vp, i32]), "swb_bam_tid": (i32, [vp, cp]), "swb_bam_header_text": (vp, [vp, C.POINTER(i64)]),
        "swb_bam_has_index": (C.c_int, [vp]),
        "swb_bam_fetch": (C.POINTER(_CBatch), [vp, i32, i64, i64, u32, u32]), "swb_bam_batch_free": (None, [C.POINTER(_CBatch)]),
        "swb_bam_count": (i64, [vp, i32, i64, i64, u32, u32]),
        "swb_bam_batch_pack4": (i64, [C.POINTER(_CBatch), vp, vp]), "swb_bam_batch_cigar_text": (i64, [C.POINTER(_CBatch), vp, i64, vp]),
        "swb_bam_fetch_pack4": (C.POINTER(_CPack), [vp, i64, vp, vp, vp, u32, u32, C.c_int, C.c_int, C.c_int]), "swb_bam_pack_free": (None, [C.POINTER(_CPack)]),
        "swb_fai_fetch_many": (i64, [vp, i64, C.POINTER(cp), vp, vp, vp, i64, vp]),
        "swb_pileup_columns": (C.POINTER(_CCols), [C.POINTER(_CBatch), i32, i32, i32, vp, i64, i64, i64, i64]),
        "swb_pileup_cols_free": (None, [C.POINTER(_CCols)]), "swb_pileup_read_size": (i32, []),
        "swb_bam_create": (vp, [cp, cp, i32, C.POINTER(cp), vp, C.c_int]),
        "swb_bam_write": (C.c_int, [vp, i64] + [vp] * 16), "swb_bam_writer_close": (C.c_int, [vp, C.c_int]),
        "swb_fasta_write": (C.c_int, [cp, i32, C.POINTER(cp), C.POINTER(cp), vp, C.c_int]),
        "swb_fai_open": (vp, [cp]), "swb_fai_close": (None, [vp]), "swb_fai_n": (i32, [vp]), "swb_fai_name": (cp, [vp, i32]),
        "swb_fai_len": (i64, [vp, cp]), "swb_fai_fetch": (i64, [vp, cp, i64, i64, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.swb_pileup_read_size() != PILEUP_READ_DTYPE.itemsize:
        raise RuntimeError("swb_pileup_read layout differs between libswbbam.so and bamio.py")
    _lib = lib
    return lib


def _err(lib) -> str:
    return (lib.swb_bam_last_error() or b"").decode("utf-8", "replace")


def _view(ptr, n, dtype):
    """numpy copy of n items at a C pointer (the C memory stays with its owner)"""
    dt = np.dtype(dtype)
    if n == 0 or not ptr:
        return np.zeros(0, dt)
    buf = (C.c_char * (n * dt.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dt, count=n).copy()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------------------------------------- records
class AlignedSegment:
    """One record of a ReadBatch under pysam.AlignedSegment's attribute names (what pileup.pyx:138-200 and the callers of
    read["read"] touch).  A view: bases, qualities and the CIGAR are cut out of the batch arenas on access and cached."""

    __slots__ = ("_b", "_i", "_seq", "_qual", "_cig")

    def __init__(self, batch: "ReadBatch", i: int):
        self._b, self._i = batch, i
        self._seq = self._qual = self._cig = None

    @property
    def query_name(self) -> str:
        return self._b.name(self._i)

    @property
    def flag(self) -> int:
        return int(self._b.flag[self._i])

    @property
    def reference_id(self) -> int:
        return int(self._b.tid[self._i])

    @property
    def reference_name(self) -> Optional[str]:
        t = int(self._b.tid[self._i])
        return self._b.references[t] if 0 <= t < len(self._b.references) else None

    @property
    def reference_start(self) -> int:
        return int(self._b.pos[self._i])

    @property
    def reference_end(self) -> Optional[int]:
        e = int(self._b.end[self._i])
        return None if e < 0 else e

    @property
    def mapping_quality(self) -> int:
        return int(self._b.mapq[self._i])

    @property
    def is_reverse(self) -> bool:
        return bool(self._b.flag[self._i] & FREVERSE)

    @property
    def is_duplicate(self) -> bool:
        return bool(self._b.flag[self._i] & FDUP)

    @property
    def is_secondary(self) -> bool:
        return bool(self._b.flag[self._i] & FSECONDARY)

    @property
    def is_supplementary(self) -> bool:
        return bool(self._b.flag[self._i] & FSUPPLEMENTARY)

    @property
    def is_unmapped(self) -> bool:
        return bool(self._b.flag[self._i] & FUNMAP)

    @property
    def is_qcfail(self) -> bool:
        return bool(self._b.flag[self._i] & FQCFAIL)

    @property
    def query_length(self) -> int:
        return int(self._b.l_seq[self._i])

    @property
    def query_sequence(self) -> Optional[str]:
        if self._seq is None:
            self._seq = self._b.sequence(self._i)
        return self._seq or None

    @property
    def query_qualities(self):
        if self._qual is None:
            self._qual = self._b.qualities(self._i)
        return self._qual

    @property
    def cigartuples(self) -> Optional[List[Tuple[int, int]]]:
        c = self._b.cigar_words(self._i)
        return [(int(w) & 15, int(w) >> 4) for w in c] if len(c) else None

    @property
    def cigarstring(self) -> Optional[str]:
        if self._cig is None:
            self._cig = self._b.cigarstring(self._i)
        return self._cig or None

    @property
    def query_alignment_start(self) -> int:
        c = self._b.cigar_words(self._i)
        s = 0
        for w in c:
            if int(w) & 15 == 4:
                s += int(w) >> 4
            elif int(w) & 15 != 5:
                break
        return s

    @property
    def query_alignment_end(self) -> int:
        c = self._b.cigar_words(self._i)
        e = int(self._b.l_seq[self._i])
        for w in c[::-1]:
            if int(w) & 15 == 4:
                e -= int(w) >> 4
            elif int(w) & 15 != 5:
                break
        return e

    @property
    def query_alignment_sequence(self) -> Optional[str]:
        s = self.query_sequence
        return None if s is None else s[self.query_alignment_start: self.query_alignment_end]

    @property
    def next_reference_id(self) -> int:
        return int(self._b.next_tid[self._i])

    @property
    def next_reference_start(self) -> int:
        return int(self._b.next_pos[self._i])

    @property
    def template_length(self) -> int:
        return int(self._b.tlen[self._i])

    def as_dict(self) -> dict:
        """the keyword arguments of a stub / real pysam AlignedSegment carrying the same record (tests, interop)"""
        return dict(query_name=self.query_name, query_sequence=self.query_sequence, query_qualities=self.query_qualities,
                    cigarstring=self.cigarstring, reference_start=self.reference_start, reference_end=self.reference_end,
                    mapping_quality=self.mapping_quality, is_reverse=self.is_reverse, is_duplicate=self.is_duplicate,
                    is_secondary=self.is_secondary, is_supplementary=self.is_supplementary, reference_name=self.reference_name,
                    query_alignment_sequence=self.query_alignment_sequence)

    def __repr__(self):
        return f"<AlignedSegment {self.query_name} {self.reference_name}:{self.reference_start} {self.cigarstring}>"


class PackedReads:
    """swb_bam_fetch_pack4's output: the reads of many regions as ONE SWB_SEQ_PACKED4 table (`table`, `read_off`, `read_len`) in
    region order -- region r holds reads [region_first[r], region_first[r+1]) -- with pos / end / flag / mapq per read"""

    __slots__ = ("n_regions", "n_reads", "region_first", "read_off", "read_len", "table", "pos", "end", "flag", "mapq")

    def region_of_read(self):
        """int32[n_reads]: the region every read belongs to (pair_win of a locus-per-window batch)"""
        return np.repeat(np.arange(self.n_regions, dtype=np.int32), np.diff(self.region_first))


class PileupColumns:
    """swb_pileup_columns' output: `reads` (structured array, one row per record of the batch), `subreads` (n x 2),
    `indels` (insertions then deletions of each read) and the per-read ref_seq arena"""

    def __init__(self, reads, subreads, indels, ref_seq, ref_seq_off):
        self.reads, self.subreads, self.indels, self._ref, self.ref_seq_off = reads, subreads, indels, ref_seq, ref_seq_off

    def ref_seq(self, i: int) -> str:
        return self._ref[int(self.ref_seq_off[i]): int(self.ref_seq_off[i + 1])]


class ReadBatch:
    """The records of one region, columnar (swb_bam_batch).  Columns are numpy arrays owned by Python; the C batch stays
    alive with the object for pack4 / pileup_columns."""

    # columns are copied out of the C batch on first use: the columnar path touches four of them, the per-read path all
    _COLS = {"tid": "<i4", "pos": "<i4", "end": "<i4", "flag": "<u2", "mapq": "u1", "l_seq": "<i4", "n_cigar": "<i4", "next_tid": "<i4",
             "next_pos": "<i4", "tlen": "<i4", "name_off": "<i8", "seq_off": "<i8", "cigar_off": "<i8"}

    def __init__(self, lib, cb, references):
        self._lib, self._cb, self.references = lib, cb, references
        self.n = int(cb.contents.n)
        self._seq_str = None
        self._cig_text = None

    def __getattr__(self, name):
        # only reached for attributes not materialised yet
        cb = self.__dict__.get("_cb")
        if cb is None:
            raise AttributeError(name)
        b = cb.contents
        if name in ReadBatch._COLS:
            v = _view(getattr(b, name), self.n, ReadBatch._COLS[name])
        elif name == "names":
            v = C.string_at(b.names, int(b.names_len))
        elif name == "seq":
            v = C.string_at(b.seq, int(b.seq_len))                 # ASCII bases, all records back to back
        elif name == "qual":
            v = C.string_at(b.qual, int(b.seq_len))
        elif name == "cigar":
            v = _view(b.cigar, int(b.cigar_len), "<u4")
        else:
            raise AttributeError(name)
        self.__dict__[name] = v
        return v

    def __del__(self):
        cb, self._cb = getattr(self, "_cb", None), None
        if cb:
            self._lib.swb_bam_batch_free(cb)

    def __len__(self):
        return self.n

    def name(self, i):
        o = int(self.name_off[i])
        return self.names[o: self.names.index(b"\0", o)].decode("ascii")

    def sequence(self, i) -> str:
        if self._seq_str is None:
            self._seq_str = self.seq.decode("ascii")
        o = int(self.seq_off[i])
        return self._seq_str[o: o + int(self.l_seq[i])]

    def qualities(self, i):
        """array('B') like pysam's query_qualities; None when the record has none (0xff-filled in BAM)"""
        o, L = int(self.seq_off[i]), int(self.l_seq[i])
        if L and self.qual[o] == 0xFF:
            return None
        return array.array("B", self.qual[o: o + L])

    def cigar_words(self, i):
        o = int(self.cigar_off[i])
        return self.cigar[o: o + int(self.n_cigar[i])]

    def cigarstring(self, i) -> str:
        if self._cig_text is None:
            need = self._lib.swb_bam_batch_cigar_text(self._cb, None, 0, None)
            buf = np.zeros(max(1, need), "u1"); off = np.zeros(self.n + 1, "<i8")
            self._lib.swb_bam_batch_cigar_text(self._cb, _ptr(buf), need, _ptr(off))
            self._cig_text = (buf.tobytes().decode("ascii"), off)
        text, off = self._cig_text
        return text[int(off[i]): int(off[i + 1]) - 1]

    def count_overlapping(self, start: int, stop: int, exclude: int = 0) -> int:
        """records of the batch overlapping [start, stop) under htslib's rule (what AlignmentFile.count would report for a
        region inside the fetched one) without another pass over the file"""
        end = np.where(self.end < 0, self.pos + 1, np.maximum(self.end, self.pos + 1))
        m = (self.pos < stop) & (end > start)
        if exclude:
            m &= (self.flag & exclude) == 0
        return int(m.sum())

    def segment(self, i) -> AlignedSegment:
        return AlignedSegment(self, i)

    def __iter__(self) -> Iterator[AlignedSegment]:
        for i in range(self.n):
            yield AlignedSegment(self, i)

    def pack4(self):
        """-> (table uint8[], off int64[n], len int32[n]): the batch's reads as a SWB_SEQ_PACKED4 table (DNA_BASE_LUT codes,
        two per byte) straight from the BAM's own 4-bit bases: what BatchAligner / swb_align_batch take as the read table"""
        table = np.empty(max(1, int(self._cb.contents.seq4_len)), "u1"); off = np.empty(max(1, self.n), "<i8")
        w = self._lib.swb_bam_batch_pack4(self._cb, table.ctypes.data, off.ctypes.data)
        return table[:w], off[: self.n], self.l_seq

    def pileup_columns(self, pos, rpos, basequalthresh, contig: Optional[bytes] = None, contig_start=0, local_start=None, local_len=None) -> PileupColumns:
        """the integer core of dictize_read for every record (swb_pileup_columns); `contig` = reference bases from
        contig_start on (bytes) -> ref_seq / low_qual_base_num / is_reference_seq are filled too"""
        if contig is not None:
            cbuf = np.frombuffer(contig, "u1"); cptr, clen = _ptr(cbuf), len(contig)
        else:
            cbuf, cptr, clen = None, None, 0
        if local_start is None:
            local_start, local_len = contig_start, clen
        pc = self._lib.swb_pileup_columns(self._cb, int(pos), int(rpos), int(basequalthresh), cptr, int(contig_start), clen, int(local_start), int(local_len))
        if not pc:
            raise MemoryError(_err(self._lib))
        try:
            c = pc.contents
            reads = _view(c.reads, int(c.n), PILEUP_READ_DTYPE)
            sub = _view(c.subreads, 2 * int(c.n_subreads), "<i4").reshape(-1, 2)
            ind = _view(c.indels, int(c.n_indels), PILEUP_INDEL_DTYPE)
            ref = C.string_at(c.ref_seq, int(c.ref_seq_len)).decode("ascii", "replace")
            roff = _view(c.ref_seq_off, int(c.n) + 1, "<i8")
        finally:
            self._lib.swb_pileup_cols_free(pc)
        return PileupColumns(reads, sub, ind, ref, roff)


# ---------------------------------------------------------------------------------------------------------------- files
class AlignmentFile:
    """pysam.AlignmentFile's read side as indelPost uses it (pileup.pyx:71, 83, 134-136), BAM + BAI through libswbbam."""

    def __init__(self, path: str, mode: str = "rb"):
        if mode not in ("rb", "r"):
            raise ValueError("AlignmentFile reads BAM; use write_bam() to write")
        self._lib = load()
        self.filename = path
        self._h = self._lib.swb_bam_open(os.fsencode(path))
        if not self._h:
            raise OSError(_err(self._lib))
        n = self._lib.swb_bam_n_ref(self._h)
        self.references = tuple(self._lib.swb_bam_ref_name(self._h, i).decode() for i in range(n))
        self.lengths = tuple(int(self._lib.swb_bam_ref_len(self._h, i)) for i in range(n))
        self.nreferences = n
        self._tid = {r: i for i, r in enumerate(self.references)}

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.swb_bam_close(h)

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def header_text(self) -> str:
        n = C.c_int64(0)
        p = self._lib.swb_bam_header_text(self._h, C.byref(n))
        return C.string_at(p, n.value).decode("utf-8", "replace") if p else ""

    def has_index(self) -> bool:
        return bool(self._lib.swb_bam_has_index(self._h))

    def get_tid(self, reference: str) -> int:
        return self._tid.get(reference, -1)

    def get_reference_name(self, tid: int) -> str:
        return self.references[tid]

    def _region(self, contig, start, stop):
        if contig is None:
            return -1, 0, 1 << 29
        tid = self._tid.get(contig, -1)
        if tid < 0:
            raise ValueError(f"invalid contig `{contig}`")          # pysam's message
        start = 0 if start is None else int(start)
        stop = self.lengths[tid] if stop is None else int(stop)
        if start < 0 or stop < start:
            raise ValueError(f"invalid coordinates: start ({start}) > stop ({stop})" if stop < start else f"start out of range ({start})")
        return tid, start, stop

    def fetch_columns(self, contig=None, start=None, stop=None, require: int = 0, exclude: int = 0) -> ReadBatch:
        """every record overlapping [start, stop) as ONE columnar batch (no per-read Python objects)"""
        tid, start, stop = self._region(contig, start, stop)
        cb = self._lib.swb_bam_fetch(self._h, tid, start, stop, require, exclude)
        if not cb:
            raise OSError(_err(self._lib))
        return ReadBatch(self._lib, cb, self.references)

    def fetch_pack4(self, regions, require: int = 0, exclude: int = 0, need_cigar: bool = True, drop_pos0: bool = False, threads: int = 0) -> PackedReads:
        """the reads of MANY regions -- [(contig, start, stop), ...] -- in one call on host threads, as one SWB_SEQ_PACKED4 read
        table (swb_bam_fetch_pack4).  fetch_reads' filter (pileup.pyx:138-155) is exclude=FSECONDARY (| FDUP), need_cigar=True,
        drop_pos0=exclude_duplicates."""
        n = len(regions)
        tid = np.empty(n, "<i4"); beg = np.empty(n, "<i8"); end = np.empty(n, "<i8")
        for k, (contig, start, stop) in enumerate(regions):
            tid[k], beg[k], end[k] = self._region(contig, start, stop)
        pp = self._lib.swb_bam_fetch_pack4(self._h, n, _ptr(tid), _ptr(beg), _ptr(end), require, exclude, int(need_cigar), int(drop_pos0), int(threads))
        if not pp:
            raise OSError(_err(self._lib))
        try:
            c = pp.contents
            out = PackedReads()
            out.n_regions, out.n_reads = n, int(c.n_reads)
            m = out.n_reads
            out.region_first = _view(c.region_first, n + 1, "<i8")
            out.read_off = _view(c.read_off, m, "<i8"); out.read_len = _view(c.read_len, m, "<i4")
            out.table = _view(c.table, int(c.table_len), "u1")
            out.pos = _view(c.pos, m, "<i4"); out.end = _view(c.end, m, "<i4"); out.flag = _view(c.flag, m, "<u2"); out.mapq = _view(c.mapq, m, "u1")
        finally:
            self._lib.swb_bam_pack_free(pp)
        return out

    def fetch(self, contig=None, start=None, stop=None, until_eof: bool = False) -> Iterator[AlignedSegment]:
        """pysam semantics: with a region the records overlapping it in file order (`until_eof` only matters without one)"""
        return iter(self.fetch_columns(contig, start, stop))

    def count(self, contig=None, start=None, stop=None, read_callback="nofilter") -> int:
        """pysam's count(): "all" skips unmapped / secondary / qc-fail / duplicate records, "nofilter" counts everything"""
        tid, start, stop = self._region(contig, start, stop)
        if read_callback == "all":
            exclude = FUNMAP | FSECONDARY | FQCFAIL | FDUP
        elif read_callback == "nofilter":
            exclude = 0
        else:
            return sum(1 for r in self.fetch(contig, start, stop) if read_callback(r))
        n = self._lib.swb_bam_count(self._h, tid, start, stop, 0, exclude)
        if n < 0:
            raise OSError(_err(self._lib))
        return int(n)


class FastaFile:
    """pysam.FastaFile as indelPost uses it: fetch(reference, start, end) (0-based, half open, clamped), get_reference_length,
    references, filename (pileup.pyx:69, 290; local_reference.pyx:11; variant.pyx)"""

    def __init__(self, path: str):
        self._lib = load()
        self.filename = path
        self._h = self._lib.swb_fai_open(os.fsencode(path))
        if not self._h:
            raise OSError(_err(self._lib))
        n = self._lib.swb_fai_n(self._h)
        self.references = tuple(self._lib.swb_fai_name(self._h, i).decode() for i in range(n))
        self.lengths = tuple(int(self._lib.swb_fai_len(self._h, r.encode())) for r in self.references)
        self.nreferences = n
        self._len = dict(zip(self.references, self.lengths))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.swb_fai_close(h)

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def get_reference_length(self, reference: str) -> int:
        try:
            return self._len[reference]
        except KeyError:
            raise KeyError(f"sequence '{reference}' not present") from None

    def fetch_bytes(self, reference, start=None, end=None) -> bytes:
        L = self.get_reference_length(reference)
        start = 0 if start is None else max(0, int(start))
        end = L if end is None else min(L, int(end))
        if end <= start:
            return b""
        buf = C.create_string_buffer(end - start)
        n = self._lib.swb_fai_fetch(self._h, reference.encode(), start, end, buf)
        if n < 0:
            raise OSError(_err(self._lib))
        return buf.raw[:n]

    def fetch(self, reference=None, start=None, end=None) -> str:
        return self.fetch_bytes(reference, start, end).decode("ascii")

    def fetch_many(self, regions):
        """[(reference, start, end), ...] -> (uint8 blob, int64 off[n + 1]): the slices back to back (a window table)"""
        n = len(regions)
        names = (C.c_char_p * max(1, n))(*[r[0].encode() for r in regions])
        beg = np.array([r[1] for r in regions] or [0], "<i8"); end = np.array([r[2] for r in regions] or [0], "<i8")
        off = np.zeros(n + 1, "<i8")
        cap = int(np.maximum(end - np.maximum(beg, 0), 0).sum()) if n else 0
        buf = np.empty(max(1, cap), "u1")
        w = self._lib.swb_fai_fetch_many(self._h, n, names, _ptr(beg), _ptr(end), _ptr(buf), cap, _ptr(off))
        if w < 0:
            raise OSError(_err(self._lib) or "fetch_many failed")
        return buf[:w], off


# ---------------------------------------------------------------------------------------------------------------- writing
def parse_cigar(cigarstring: str) -> List[int]:
    """"70M1D80M" -> BAM words (len << 4 | op)"""
    out, n = [], 0
    for ch in cigarstring:
        if ch.isdigit():
            n = n * 10 + ord(ch) - 48
        else:
            out.append(n << 4 | _OP_CODE[ch]); n = 0
    return out


def write_bam(path: str, references: Sequence[Tuple[str, int]], reads: Iterable[dict], index: bool = True, level: int = 6, header_text: Optional[str] = None) -> int:
    """coordinate-sorted BAM (+ BAI) from pysam-style read dicts (the keyword arguments tests/loci.py builds: query_name,
    query_sequence, query_qualities, cigarstring, reference_name, reference_start, mapping_quality, is_reverse, is_duplicate,
    is_secondary, is_supplementary; optional `flag` overrides the booleans).  Returns the number of records written."""
    lib = load()
    names = [r for r, _ in references]
    tid_of = {r: i for i, r in enumerate(names)}
    rows = []
    for k, r in enumerate(reads):
        t = tid_of[r["reference_name"]] if r.get("reference_name") is not None else -1
        rows.append((t if t >= 0 else 1 << 30, int(r["reference_start"]), k, r, t))
    rows.sort(key=lambda x: x[:3])
    n = len(rows)
    tid = np.zeros(n, "<i4"); pos = np.zeros(n, "<i4"); flag = np.zeros(n, "<u2"); mapq = np.zeros(n, "u1"); l_seq = np.zeros(n, "<i4")
    n_cig = np.zeros(n, "<i4"); name_off = np.zeros(n, "<i8"); seq_off = np.zeros(n, "<i8"); cig_off = np.zeros(n, "<i8")
    nb, sb, qb, cw = bytearray(), bytearray(), bytearray(), []
    any_qual = False
    for i, (_, p, _, r, t) in enumerate(rows):
        tid[i], pos[i] = t, p
        f = r.get("flag")
        if f is None:
            f = (FREVERSE if r.get("is_reverse") else 0) | (FDUP if r.get("is_duplicate") else 0) | (FSECONDARY if r.get("is_secondary") else 0) \
                | (FSUPPLEMENTARY if r.get("is_supplementary") else 0) | (FUNMAP if r.get("is_unmapped") else 0)
        flag[i] = f
        mapq[i] = int(r.get("mapping_quality", 0))
        s = (r.get("query_sequence") or "").encode("ascii")
        l_seq[i] = len(s)
        seq_off[i] = len(sb); sb += s
        q = r.get("query_qualities")
        if q is None:
            qb += b"\xff" * len(s)
        else:
            any_qual = True
            qb += bytes(bytearray(q))
        words = parse_cigar(r.get("cigarstring") or "")
        n_cig[i] = len(words); cig_off[i] = len(cw); cw.extend(words)
        name_off[i] = len(nb); nb += r.get("query_name", f"r{i}").encode("ascii") + b"\0"
    cig = np.array(cw if cw else [0], "<u4")
    nbuf = np.frombuffer(bytes(nb) or b"\0", "u1"); sbuf = np.frombuffer(bytes(sb) or b"\0", "u1"); qbuf = np.frombuffer(bytes(qb) or b"\0", "u1")
    if header_text is None:
        header_text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{r}\tLN:{ln}\n" for r, ln in references)
    cnames = (C.c_char_p * max(1, len(names)))(*[x.encode() for x in names])
    lens = np.array([ln for _, ln in references] or [0], "<i8")
    w = lib.swb_bam_create(os.fsencode(path), header_text.encode(), len(names), cnames, _ptr(lens), level)
    if not w:
        raise OSError(_err(lib))
    rc = lib.swb_bam_write(w, n, _ptr(tid), _ptr(pos), _ptr(flag), _ptr(mapq), _ptr(l_seq), _ptr(n_cig), _ptr(name_off), _ptr(seq_off), _ptr(cig_off),
                           _ptr(nbuf), _ptr(sbuf), _ptr(qbuf) if any_qual else None, _ptr(cig), None, None, None)
    msg = _err(lib) if rc != 0 else ""
    rc2 = lib.swb_bam_writer_close(w, 1 if index else 0)
    if rc != 0 or rc2 != 0:
        raise OSError(msg or _err(lib) or "BAM write failed")
    return n


def write_bam_columns(path: str, references: Sequence[Tuple[str, int]], tid, pos, flag, mapq, l_seq, n_cigar, name_off, seq_off, cigar_off,
                      names: bytes, seq: bytes, qual: Optional[bytes], cigar, index: bool = True, level: int = 6, header_text: Optional[str] = None) -> int:
    """coordinate-sorted BAM (+ BAI) straight from columnar arrays (the layout of ReadBatch / swb_bam_write): no per-read Python
    objects.  `names`: NUL-terminated names back to back; `seq` / `qual`: ASCII bases / phred bytes at seq_off; `cigar`: BAM words."""
    lib = load()
    n = int(len(pos))
    arr = lambda a, dt: np.ascontiguousarray(a, dtype=dt)  # noqa: E731
    tid, pos, flag, mapq = arr(tid, "<i4"), arr(pos, "<i4"), arr(flag, "<u2"), arr(mapq, "u1")
    l_seq, n_cigar, name_off, seq_off, cigar_off = arr(l_seq, "<i4"), arr(n_cigar, "<i4"), arr(name_off, "<i8"), arr(seq_off, "<i8"), arr(cigar_off, "<i8")
    cigar = arr(cigar, "<u4")
    nbuf, sbuf = np.frombuffer(names or b"\0", "u1"), np.frombuffer(seq or b"\0", "u1")
    qbuf = np.frombuffer(qual, "u1") if qual else None
    if header_text is None:
        header_text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{r}\tLN:{ln}\n" for r, ln in references)
    cnames = (C.c_char_p * max(1, len(references)))(*[r.encode() for r, _ in references])
    lens = np.array([ln for _, ln in references] or [0], "<i8")
    w = lib.swb_bam_create(os.fsencode(path), header_text.encode(), len(references), cnames, _ptr(lens), level)
    if not w:
        raise OSError(_err(lib))
    rc = lib.swb_bam_write(w, n, _ptr(tid), _ptr(pos), _ptr(flag), _ptr(mapq), _ptr(l_seq), _ptr(n_cigar), _ptr(name_off), _ptr(seq_off), _ptr(cigar_off),
                           _ptr(nbuf), _ptr(sbuf), _ptr(qbuf) if qbuf is not None else None, _ptr(cigar), None, None, None)
    msg = _err(lib) if rc != 0 else ""
    rc2 = lib.swb_bam_writer_close(w, 1 if index else 0)
    if rc != 0 or rc2 != 0:
        raise OSError(msg or _err(lib) or "BAM write failed")
    return n


def write_fasta(path: str, seqs, line_width: int = 60) -> None:
    """FASTA + .fai from {name: sequence} (or a list of pairs)"""
    lib = load()
    items = list(seqs.items()) if hasattr(seqs, "items") else list(seqs)
    n = len(items)
    names = (C.c_char_p * max(1, n))(*[k.encode() for k, _ in items])
    data = [v.encode("ascii") if isinstance(v, str) else bytes(v) for _, v in items]
    ptrs = (C.c_char_p * max(1, n))(*data)
    lens = np.array([len(d) for d in data] or [0], "<i8")
    if lib.swb_fasta_write(os.fsencode(path), n, names, ptrs, _ptr(lens), line_width) != 0:
        raise OSError(_err(lib))
