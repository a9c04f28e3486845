cdef class AlignmentFile:
    cdef public list reads
    cdef public tuple refs
