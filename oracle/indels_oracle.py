"""CPU restatement of indelPost's CIGAR -> indel extraction (test infrastructure only: imported by tests/ and
__graft_entry__.smoke(), never by the product).

Follows the reference line by line:
  merge_consecutive_gaps   utilities.pyx:360-380
  make_insertion_first     utilities.pyx:383-401
  findall_indels           localn.pyx:542-621
Pinned by tests/golden/indels.json (outputs of the unmodified reference, see tests/golden/make_indel_golden.py).
"""
import re

cigar_ptrn = re.compile(r"[0-9]+[MIDNSHPX=]")  # localn.pyx:12


def merge_consecutive_gaps(cigar_lst):
    """utilities.pyx:360-380: a gap token swallows the gap tokens that follow it directly."""
    cigar_lst = list(cigar_lst)
    merged = []
    while cigar_lst:
        c = cigar_lst[0]
        cigar_lst = cigar_lst[1:]
        if "I" in c or "D" in c:
            i = 0
            is_gap = True
            while i < len(cigar_lst) and is_gap:
                tmp = cigar_lst[i]
                is_gap = "I" in tmp or "D" in tmp
                i += 1
            # NB (utilities.pyx:373-375): i - 1 tokens are merged even when the scan stopped because the list ended
            # on a gap token, i.e. a gap run at the very END of the CIGAR leaves its last token unmerged
            if i - 1:
                c += "".join(cigar_lst[: i - 1])
                cigar_lst = cigar_lst[i - 1:]
        merged.append(c)
    return merged


def make_insertion_first(cigarstring):
    """utilities.pyx:383-401: inside a merged gap run holding both kinds, a run that starts with a deletion is reversed."""
    out = []
    for c in merge_consecutive_gaps(cigar_ptrn.findall(cigarstring)):
        if "I" in c and "D" in c:
            toks = cigar_ptrn.findall(c)
            out.append("".join(toks[::-1]) if "D" in toks[0] else "".join(toks))
        else:
            out.append(c)
    return "".join(out)


def indel_records(cigarstring, reference_start, read_start):
    """The integer core of findall_indels (localn.pyx:544-613): one record per I/D token of the reordered CIGAR,
    (op, length, ref_idx, read_idx, pos_off) with pos_off relative to genome_aln_pos (starts at -1, localn.pyx:544).
    Also returns the read index after the last token (start of rt_clipped, localn.pyx:615)."""
    pos = -1
    ref_idx, read_idx = reference_start, read_start
    recs = []
    for tok in cigar_ptrn.findall(make_insertion_first(cigarstring)):
        ev, n = tok[-1], int(tok[:-1])
        if ev == "I":
            recs.append(("I", n, ref_idx, read_idx, pos))
            read_idx += n
        elif ev == "D":
            recs.append(("D", n, ref_idx, read_idx, pos))
            ref_idx += n
            pos += n
        else:
            ref_idx += n
            read_idx += n
            pos += n
    return recs, read_idx


def findall_indels(ref_aln, genome_aln_pos, ref_seq, read_seq, report_snvs=False, basequals=None):
    """localn.pyx:542-621 on top of indel_records (same dict keys and values as the reference)."""
    recs, read_end = indel_records(ref_aln.CIGAR, ref_aln.reference_start, ref_aln.read_start)
    lt_clipped = read_seq[: ref_aln.read_start]
    rt_clipped = read_seq[read_end:]
    indels = []
    for op, n, ref_idx, read_idx, pos_off in recs:
        d = {"pos": genome_aln_pos + pos_off, "lt_ref": ref_seq[:ref_idx], "lt_flank": read_seq[:read_idx]}
        if basequals:
            d["lt_qual"] = basequals[:read_idx]
        if op == "I":
            d.update(indel_type="I", indel_seq=read_seq[read_idx: read_idx + n], rt_ref=ref_seq[ref_idx:], rt_flank=read_seq[read_idx + n:],
                     ref_idx=ref_idx, read_idx=read_idx)
            if basequals:
                d["rt_qual"] = basequals[read_idx + n:]
        else:
            d.update(indel_type="D", indel_seq="", del_seq=ref_seq[ref_idx: ref_idx + n], rt_ref=ref_seq[ref_idx + n:], rt_flank=read_seq[read_idx:],
                     ref_idx=ref_idx, read_idx=read_idx)
            if basequals:
                d["rt_qual"] = basequals[read_idx:]
        d["lt_clipped"] = lt_clipped
        d["rt_clipped"] = rt_clipped
        indels.append(d)
    if not report_snvs:
        return indels
    # localn.pyx:594-606
    snvs = []
    pos = genome_aln_pos - 1
    ref_idx, read_idx = ref_aln.reference_start, ref_aln.read_start
    for tok in cigar_ptrn.findall(make_insertion_first(ref_aln.CIGAR)):
        ev, n = tok[-1], int(tok[:-1])
        if ev == "I":
            read_idx += n
        elif ev == "D":
            ref_idx += n
            pos += n
        else:
            for i in range(n):
                a, b = ref_seq[ref_idx + i: ref_idx + i + 1], read_seq[read_idx + i: read_idx + i + 1]
                if a != b:
                    snvs.append({"pos": pos + i + 1, "ref": a, "alt": b})
            ref_idx += n
            read_idx += n
            pos += n
    return indels, snvs
