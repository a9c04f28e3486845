cdef class FastaFile:
    cdef public dict seqs
    cdef public object filename
