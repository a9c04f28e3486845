# in-memory BAM: a list of AlignedSegment; fetch/count select by overlap like an indexed BAM would
cdef class AlignmentFile:
    def __init__(self, reads, refs):
        self.reads = list(reads)
        self.refs = tuple(refs)

    @property
    def references(self):
        return self.refs

    def fetch(self, contig=None, start=None, stop=None, until_eof=False):
        for r in self.reads:
            if r.reference_name == contig and r.reference_start < stop and r.reference_end > start:
                yield r

    def count(self, contig=None, start=None, stop=None, read_callback="nofilter"):
        n = 0
        for r in self.reads:
            if r.reference_name == contig and r.reference_start < stop and r.reference_end > start:
                if read_callback == "all" and (r.is_duplicate or r.is_secondary):
                    continue
                n += 1
        return n
