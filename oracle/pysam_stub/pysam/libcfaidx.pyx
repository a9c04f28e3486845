# in-memory FASTA: {contig: sequence}
cdef class FastaFile:
    def __init__(self, seqs, filename="synthetic.fa"):
        self.seqs = dict(seqs)
        self.filename = filename

    def fetch(self, reference=None, start=None, end=None):
        s = self.seqs[reference]
        start = 0 if start is None else max(0, start)
        end = len(s) if end is None else min(len(s), end)
        return s[start:end]

    def get_reference_length(self, reference):
        return len(self.seqs[reference])

    @property
    def references(self):
        return tuple(self.seqs.keys())
