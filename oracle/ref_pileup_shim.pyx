# oracle/ref_pileup_shim.pyx -- TEST INFRASTRUCTURE.  The reference's `make_pileup` (pileup.pyx:51-113) is a cdef function:
# this shim cimports it through the reference's own pileup.pxd and hands it to Python unchanged, so that the native ingest of
# indelpost_b200/pileup.py can be compared with the reference's own output (tests/test_pileup_ingest.py).
# Built by oracle/build_ref_pipeline.py into oracle/_ref_pipeline/refshim*.so next to the unmodified reference modules.
from indelpost.pileup cimport make_pileup
from indelpost.variant cimport Variant
from indelpost.local_reference cimport UnsplicedLocalReference
from pysam.libcalignmentfile cimport AlignmentFile


def ref_make_pileup(Variant target, AlignmentFile bam, UnsplicedLocalReference unspl_loc_ref, bint exclude_duplicates, int window,
                    int downsamplethresh, int basequalthresh):
    return make_pileup(target, bam, unspl_loc_ref, exclude_duplicates, window, downsamplethresh, basequalthresh)
