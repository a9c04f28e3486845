"""CPU-side checks of the drop-in boundary (no GPU needed): the shared library loads, exports every
symbol include/swb200.h declares, structure layouts match the header, and the product fails loudly
(instead of falling back to a CPU path) when no GPU is usable."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from indelpost_b200 import _lib as L

    lib = L.load()
    hdr = open(os.path.join(ROOT, "include", "swb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b((?:swb|ssw|init|align)_[a-z_0-9]*)\s*\(", hdr))
    declared = {d for d in declared if d not in ("swb_ctx",)}
    assert declared, "no declarations parsed"
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"libswb200.so does not export {sym}"
    assert set(L.EXPORTS) <= declared | {"swb_version"}


def test_struct_layouts_match_header():
    from indelpost_b200 import _lib as L

    assert L.RESULT_DTYPE.itemsize == 40
    assert C.sizeof(L.SAlign) == 40            # reference ssw.h:55-66 on LP64
    assert L.RESULT_DTYPE.fields["cigar_off"][1] == 32
    assert C.sizeof(L.SwbBatch) % 8 == 0


def test_host_helpers_without_gpu():
    from indelpost_b200 import _lib as L
    from indelpost_b200.batch import dna_score_matrix

    lib = L.load()
    s = b"ACGTUacgtuNnRYK-*"
    out = np.zeros(len(s), dtype=np.int8)
    lib.swb_encode_dna(s, out.ctypes.data, len(s))
    assert out.tolist() == [0, 1, 2, 3, 0, 0, 1, 2, 3, 0, 4, 4, 4, 4, 4, 4, 4]
    assert lib.swb_to_cigar_int(70, b"M") == 70 << 4 and lib.swb_to_cigar_int(3, b"D") == (3 << 4 | 2)
    assert lib.swb_cigar_int_to_op(5 << 4 | 1) == b"I" and lib.swb_cigar_int_to_len(5 << 4 | 1) == 5
    m = dna_score_matrix(3, 2).reshape(5, 5)
    assert m[0, 0] == 3 and m[0, 1] == -2 and (m[4] == 0).all() and (m[:, 4] == 0).all()


def test_pack_table_layouts():
    """swb_pack_table: SWB_SEQ_PACKED4 = two codes per byte, low nibble first; SWB_SEQ_PACKED2 = four per byte, bits 0-1 first;
    every entry starts on a byte boundary; ASCII goes through DNA_BASE_LUT; 2 bits refuse N"""
    from indelpost_b200.batch import pack_table

    codes = np.array([0, 1, 2, 3, 4, 3, 2,   1, 0, 3], dtype=np.int8)
    off = np.array([0, 7], dtype=np.int64)
    ln = np.array([7, 3], dtype=np.int32)
    pk, po = pack_table(codes, off, ln, bits=4)
    assert po.tolist() == [0, 4]
    assert pk.tolist() == [0x10, 0x32, 0x34, 0x02, 0x01, 0x03]
    with pytest.raises(ValueError):
        pack_table(codes, off, ln, bits=2)                   # code 4 (N) does not fit two bits
    codes2 = np.array([0, 1, 2, 3, 3, 2,   1], dtype=np.int8)
    pk2, po2 = pack_table(codes2, np.array([0, 6], dtype=np.int64), np.array([6, 1], dtype=np.int32), bits=2)
    assert po2.tolist() == [0, 2]
    assert pk2.tolist() == [0b11100100, 0b1011, 0b01]
    pk3, _ = pack_table(np.frombuffer(b"ACGTNacgu", dtype=np.int8), np.array([0], dtype=np.int64), np.array([9], dtype=np.int32), bits=4, ascii=True)
    assert pk3.tolist() == [0x10, 0x32, 0x04, 0x21, 0x00]


def test_no_cpu_fallback_without_gpu():
    from indelpost_b200 import _lib as L

    lib = L.load()
    if lib.swb_device_count() > 0:
        pytest.skip("a GPU is present")
    from indelpost_b200 import SSW, BatchAligner

    with pytest.raises(L.SwbError):
        BatchAligner(0)
    a = SSW(3, 2)
    a.setReference("ACGTACGTACGTACGTACGT")
    a.setRead("ACGTACGTAC")
    with pytest.raises(ValueError):          # ssw_align returned NULL: same error path as sswpy.pyx:222-223
        a.align()


def test_sswpy_argument_checks_mirror_reference():
    from indelpost_b200 import SSW

    a = SSW(3, 2)
    a.setReference("ACGT" * 10)
    a.setRead("ACGTACGT")
    with pytest.raises(ValueError, match="negative indexing"):
        a.align(start_idx=-1)
    with pytest.raises(ValueError, match="can't be greater than ref_length"):
        a.align(start_idx=41)
    b = SSW(3, 2)
    b.setRead("ACGT")
    with pytest.raises(ValueError, match="call setReference first"):
        b.align()
