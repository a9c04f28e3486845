cdef class AlignedSegment:
    def __init__(self, **kw):
        self.is_duplicate = False
        self.is_secondary = False
        self.is_supplementary = False
        for k, v in kw.items():
            setattr(self, k, v)
