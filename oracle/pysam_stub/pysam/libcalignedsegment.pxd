cdef class AlignedSegment:
    cdef public object query_name, query_sequence, query_qualities, cigarstring, reference_start, reference_end
    cdef public object mapping_quality, is_reverse, is_duplicate, is_secondary, is_supplementary
    cdef public object query_alignment_sequence, reference_name
