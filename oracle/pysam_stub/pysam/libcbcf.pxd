cdef class VariantRecord:
    pass
cdef class VariantRecordFilter:
    pass
cdef class VariantFile:
    pass
