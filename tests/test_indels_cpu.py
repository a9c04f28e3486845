"""CIGAR -> indel extraction: the CPU restatement (oracle/indels_oracle.py) against the golden
vectors recorded from the unmodified reference (tests/golden/indels.json, made by tests/golden/make_indel_golden.py)."""
import json
import os
import sys
from collections import namedtuple

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import indels_oracle as O  # noqa: E402

Aln = namedtuple("Aln", "CIGAR reference_start read_start")


def golden_cases():
    with open(os.path.join(ROOT, "tests", "golden", "indels.json")) as fh:
        return json.load(fh)["cases"]


def test_oracle_reproduces_reference_make_insertion_first():
    cases = golden_cases()
    assert len(cases) >= 200 and sum(1 for c in cases if c["insertion_first"] != c["cigar"]) >= 10
    for c in cases:
        assert O.make_insertion_first(c["cigar"]) == c["insertion_first"], c["cigar"]


def test_oracle_reproduces_reference_findall_indels():
    for c in golden_cases():
        aln = Aln(c["cigar"], c["reference_start"], c["read_start"])
        out = O.findall_indels(aln, c["genome_aln_pos"], c["ref_seq"], c["read_seq"], report_snvs=c["report_snvs"])
        indels, snvs = out if c["report_snvs"] else (out, None)
        assert indels == c["indels"], c["cigar"]
        assert snvs == c["snvs"], c["cigar"]


def test_trailing_gap_run_quirk_is_kept():
    # utilities.pyx:366-375: a gap run that reaches the end of the CIGAR leaves its last token unmerged
    assert O.merge_consecutive_gaps(["5M", "2D", "3I"]) == ["5M", "2D", "3I"]
    assert O.merge_consecutive_gaps(["5M", "2D", "3I", "1D"]) == ["5M", "2D3I", "1D"]
    assert O.make_insertion_first("5M2D3I1D") == "5M3I2D1D"
    assert O.make_insertion_first("5M2D3I4M") == "5M3I2D4M"
