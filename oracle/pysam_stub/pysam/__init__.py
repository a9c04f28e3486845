"""Minimal in-memory stand-in for pysam (test infrastructure; pysam/htslib are absent from the image and cannot be installed
offline).  It provides exactly the attributes the reference pipeline touches: FastaFile.fetch /
get_reference_length / references / filename, AlignmentFile.fetch / count / references, the AlignedSegment fields that
pileup.pyx:160-266 reads, and empty VariantFile / VariantRecord / VariantRecordFilter classes for the cimports."""
import os


def get_include():
    return [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]


from pysam.libcfaidx import FastaFile  # noqa: E402,F401
from pysam.libcalignedsegment import AlignedSegment  # noqa: E402,F401
from pysam.libcalignmentfile import AlignmentFile  # noqa: E402,F401
from pysam.libcbcf import VariantFile, VariantRecord, VariantRecordFilter  # noqa: E402,F401
