"""CPU restatement of the reference's keyword filter (``/root/reference/filtering/main.cc``).

TEST INFRASTRUCTURE ONLY.  Pinned against the reference binary itself: ``oracle/build_ref.py``
compiles ``filtering/main.cc`` unmodified into ``oracle/_ref/adVNTR-Filtering`` and
``tests/golden/make_golden.py`` stores its stdout for a seeded input (``tests/golden/kfilter_*``).

The Aho-Corasick machine (main.cc:56-174) reports, at every read position, every keyword that
ends there; summed per locus that is "occurrences of the locus's keywords in the read", which is
what this restatement counts with a dictionary of keywords (any mix of lengths).  Characters
other than A, C, G, T are one and the same fifth symbol (main.cc:43-54).
"""
from __future__ import annotations


def _canon(s):
    return "".join(c if c in "ACGT" else "N" for c in s)


def parse_keywords(text):
    """main.cc:176-218: one line per locus, `id kw kw ...`, duplicates within a line collapse."""
    out = []
    for line in text.split("\n"):
        tok = line.split()
        if not tok:
            break
        out.append((int(tok[0]), sorted(set(tok[1:]))))
    return out


def filter_output(keywords_by_locus, names, seqs, min_matches=5, max_reads=2000):
    """The stdout of `adVNTR-Filtering reads.fa < keywords.txt` (main.cc:229-334)."""
    by_len = {}
    for vid, words in keywords_by_locus:
        for w in set(words):
            by_len.setdefault(len(w), {}).setdefault(_canon(w), []).append(vid)
    vntr_ids = [vid for vid, _ in keywords_by_locus]
    vntr_read_list = {}
    read_sequences = {}
    for name, seq in zip(names, seqs):
        cs = _canon(seq)
        counts = {}
        for k, table in by_len.items():
            for i in range(len(cs) - k + 1):
                for vid in table.get(cs[i:i + k], ()):
                    counts[vid] = counts.get(vid, 0) + 1
        for vid in sorted(counts):                     # std::map iteration order (main.cc:279)
            lst = vntr_read_list.setdefault(vid, {})
            if len(lst) > max_reads * 3:
                continue
            if counts[vid] >= min_matches:
                lst[name] = counts[vid]
                read_sequences[name] = seq
    lines, listed = [], set()
    for vid in vntr_ids:
        ranked = sorted(((c, n) for n, c in vntr_read_list.get(vid, {}).items()), reverse=True)
        row = [str(vid), str(min(len(ranked), max_reads))]
        for j, (_, n) in enumerate(ranked):
            listed.add(n)
            row.append(n)
            if j >= max_reads:
                break
        lines.append(" ".join(row))
    for n in sorted(listed):
        lines.append("%s %s" % (n, read_sequences[n]))
    return "\n".join(lines) + "\n" if lines else ""
