"""indelpost_b200 — B200-native drop-in for indelPost's Smith-Waterman realignment hot path
(reference indelpost/sswpy.pyx + ssw.c).  See DESIGN.md / INTEGRATION.md."""
from .sswpy import SSW, Alignment, align_batch, force_align, format_force_align, prefetch_alignments, clear_prefetched  # noqa: F401
from .batch import BatchAligner, dna_score_matrix  # noqa: F401

__all__ = ["SSW", "Alignment", "align_batch", "force_align", "format_force_align", "prefetch_alignments", "clear_prefetched", "BatchAligner", "dna_score_matrix"]
