"""ctypes binding of libswb200.so (include/swb200.h).  No CPU fallback: if the shared library or a
usable B200 is missing, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libswb200.so")

SWB_SEQ_CODES = 0
SWB_SEQ_ASCII = 1
SWB_SEQ_PACKED4 = 2
SWB_SEQ_PACKED2 = 3
SWB_OK = 0
SWB_ERR_BYTE_ONLY = 1
SWB_ERR_BAD_INPUT = 2

# swb_result (40 bytes)
RESULT_DTYPE = np.dtype(
    [
        ("score1", "<u2"), ("score2", "<u2"),
        ("ref_begin1", "<i4"), ("ref_end1", "<i4"), ("read_begin1", "<i4"), ("read_end1", "<i4"), ("ref_end2", "<i4"),
        ("cigar_len", "<i4"), ("flag", "<u2"), ("status", "<u2"), ("cigar_off", "<i8"),
    ],
    align=True,
)
assert RESULT_DTYPE.itemsize == 40

# swb_indel (20 bytes): one I / D token of an alignment (include/swb200.h, section 3)
INDEL_DTYPE = np.dtype([("pair", "<i4"), ("cigar_op", "<u4"), ("ref_idx", "<i4"), ("read_idx", "<i4"), ("pos_off", "<i4")])
assert INDEL_DTYPE.itemsize == 20


class SwbBatch(C.Structure):
    _fields_ = [
        ("n_pairs", C.c_int32), ("n_reads", C.c_int32), ("n_windows", C.c_int32), ("seq_encoding", C.c_int32),
        ("reads", C.c_void_p), ("read_off", C.c_void_p), ("read_len", C.c_void_p),
        ("windows", C.c_void_p), ("win_off", C.c_void_p), ("win_len", C.c_void_p),
        ("pair_read", C.c_void_p), ("pair_win", C.c_void_p), ("ref_beg", C.c_void_p), ("ref_len", C.c_void_p),
        ("gap_open", C.c_void_p), ("gap_ext", C.c_void_p), ("mask_len", C.c_void_p),
        ("mat", C.c_void_p), ("n", C.c_int32),
        ("score_size", C.c_int8), ("flag", C.c_uint8), ("filters", C.c_uint16), ("filterd", C.c_int32),
    ]


class SwbTiming(C.Structure):
    _fields_ = [
        ("ms_total", C.c_float), ("ms_prepare", C.c_float), ("ms_forward", C.c_float), ("ms_reverse", C.c_float),
        ("ms_traceback", C.c_float), ("ms_h2d", C.c_float), ("ms_d2h", C.c_float),
        ("cells_forward", C.c_int64), ("cells_reverse", C.c_int64), ("cells_band", C.c_int64),
        ("n_fast", C.c_int64), ("n_exact", C.c_int64), ("n_launches", C.c_int64),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("ms_band_round0", C.c_float), ("ms_band_rest", C.c_float), ("ms_certify", C.c_float), ("band_rounds", C.c_int32),
        ("n_sw_certified", C.c_int32), ("n_sw_rejected", C.c_int32), ("n_sw_verified", C.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SAlign(C.Structure):  # s_align, reference ssw.h:55-66
    _fields_ = [
        ("score1", C.c_uint16), ("score2", C.c_uint16),
        ("ref_begin1", C.c_int32), ("ref_end1", C.c_int32), ("read_begin1", C.c_int32), ("read_end1", C.c_int32),
        ("ref_end2", C.c_int32), ("cigar", C.POINTER(C.c_uint32)), ("cigarLen", C.c_int32), ("flag", C.c_uint16),
    ]


EXPORTS = (
    "ssw_init", "init_destroy", "ssw_align", "align_destroy",
    "swb_cigar_int_to_op", "swb_cigar_int_to_len", "swb_to_cigar_int",
    "swb_device_count", "swb_create", "swb_destroy", "swb_last_error",
    "swb_align_batch", "swb_upload", "swb_compute", "swb_download", "swb_get_timing", "swb_rebase_cigar_offsets", "swb_slice_table",
    "swb_host_alloc", "swb_host_free", "swb_encode_dna", "swb_pack_table", "swb_version",
    "swb_indels", "swb_indels_from_cigars",
)

_lib = None


class SwbError(RuntimeError):
    pass


def load():
    """Load libswb200.so (built in-tree by indelpost_b200/csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SwbError(f"{LIB_PATH} is missing: build it with `make -C indelpost_b200/csrc` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.swb_device_count.restype = C.c_int
    lib.swb_create.restype = C.c_void_p
    lib.swb_create.argtypes = [C.c_int]
    lib.swb_destroy.argtypes = [C.c_void_p]
    lib.swb_last_error.restype = C.c_char_p
    lib.swb_last_error.argtypes = [C.c_void_p]
    lib.swb_align_batch.restype = C.c_int
    lib.swb_align_batch.argtypes = [C.c_void_p, C.POINTER(SwbBatch), C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    lib.swb_upload.restype = C.c_int
    lib.swb_upload.argtypes = [C.c_void_p, C.POINTER(SwbBatch)]
    lib.swb_compute.restype = C.c_int
    lib.swb_compute.argtypes = [C.c_void_p]
    lib.swb_download.restype = C.c_int
    lib.swb_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    lib.swb_get_timing.restype = C.c_int
    lib.swb_get_timing.argtypes = [C.c_void_p, C.POINTER(SwbTiming)]
    lib.swb_rebase_cigar_offsets.restype = None
    lib.swb_rebase_cigar_offsets.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
    lib.swb_slice_table.restype = C.c_int
    lib.swb_slice_table.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    lib.swb_indels.restype = C.c_int
    lib.swb_indels.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    lib.swb_indels_from_cigars.restype = C.c_int
    lib.swb_indels_from_cigars.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    lib.swb_host_alloc.restype = C.c_void_p
    lib.swb_host_alloc.argtypes = [C.c_int64]
    lib.swb_host_free.argtypes = [C.c_void_p]
    lib.swb_encode_dna.argtypes = [C.c_char_p, C.c_void_p, C.c_int64]
    lib.swb_pack_table.restype = C.c_int64
    lib.swb_pack_table.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.swb_version.restype = C.c_char_p
    lib.ssw_init.restype = C.c_void_p
    lib.ssw_init.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int8]
    lib.init_destroy.argtypes = [C.c_void_p]
    lib.ssw_align.restype = C.POINTER(SAlign)
    lib.ssw_align.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint16, C.c_int32, C.c_int32]
    lib.align_destroy.argtypes = [C.POINTER(SAlign)]
    lib.swb_cigar_int_to_op.restype = C.c_char
    lib.swb_cigar_int_to_op.argtypes = [C.c_uint32]
    lib.swb_cigar_int_to_len.restype = C.c_uint32
    lib.swb_cigar_int_to_len.argtypes = [C.c_uint32]
    lib.swb_to_cigar_int.restype = C.c_uint32
    lib.swb_to_cigar_int.argtypes = [C.c_uint32, C.c_char]
    _lib = lib
    return lib


class PinnedBuffer:
    """numpy view over cudaMallocHost memory (swb_host_alloc) — the staging arrays of the batched path."""

    def __init__(self, nbytes: int):
        self._lib = load()
        self.nbytes = int(max(nbytes, 1))
        self.ptr = self._lib.swb_host_alloc(self.nbytes)
        if not self.ptr:
            raise SwbError("swb_host_alloc failed (no CUDA device?)")
        self._raw = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(self._raw, dtype=np.uint8)

    def view(self, dtype, count, offset=0):
        dt = np.dtype(dtype)
        return self.array[offset : offset + count * dt.itemsize].view(dt)

    def close(self):
        if self.ptr:
            self.array = None
            self._raw = None
            self._lib.swb_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
