"""Keyword pre-filter of unmapped reads: which reads go to which locus's Viterbi batch.

Host side of the GPU replacement of the reference's ``adVNTR-Filtering`` binary
(``/root/reference/filtering/main.cc``), as driven by ``genome_analyzer.py:173-197``.  The scan
(occurrences of every locus's keywords in every read) runs on the device
(``advhmm_kfilter_scan``); what remains here is the reference's selection logic and output
format (``main.cc:275-332``): a read is kept for a locus with >= ``min_matches`` keyword
occurrences; a locus stops accepting reads once it holds more than 3 x 2000; reads are reported
per locus by descending (occurrences, name), at most 2000 counted but 2001 listed (the
reference's ``if (j >= max) break`` comes after the print), then every reported read once with
its sequence, sorted by name.
"""
from __future__ import annotations

import numpy as np

from . import engine

MIN_MATCHES = 5          # main.cc:17
MAX_READS_PER_LOCUS = 2000   # main.cc:18


def get_keywords_for_filtering(left_flank, right_flank, repeat_segments, pattern, short_reads=True,
                               keyword_size=21):
    """``vntr_finder.py:140-154``: every 5th (6th for 5 bp patterns) k-mer of the locus."""
    vntr = "".join(repeat_segments)
    if len(vntr) < keyword_size:
        vntr = str(vntr) * (int(keyword_size / len(vntr)) + 1)
    locus = left_flank[-15:] + vntr + right_flank[:15]
    step = 5 if len(pattern) != 5 else 6
    queries = [locus[i:i + keyword_size] for i in range(0, len(locus) - keyword_size + 1, step)]
    if not short_reads:
        queries = [left_flank[-80:], right_flank[:80]]
    return set(queries)


class KeywordFilter(object):
    def __init__(self, keywords_by_locus, ctx=None):
        """``keywords_by_locus``: list of ``(vntr_id, iterable of keyword strings)`` in the order of
        the keywords file (one line per locus, duplicates within a line collapse, main.cc:198-207)."""
        self.ctx = ctx or engine.Context.default()
        self.vntr_ids = [int(v) for v, _ in keywords_by_locus]
        words, ids = [], []
        for vid, line in keywords_by_locus:
            for w in sorted(set(line)):
                words.append(w)
                ids.append(int(vid))
        self.filter = engine.DeviceKeywordFilter(self.ctx, words, ids)

    def close(self):
        self.filter.close()

    def occurrences(self, seqs, min_matches=MIN_MATCHES):
        """{(read index, vntr_id): occurrences} for the pairs reaching ``min_matches``."""
        R = len(seqs)
        off = np.zeros(R + 1, dtype=np.int64)
        if R:
            np.cumsum(np.fromiter(map(len, seqs), dtype=np.int64, count=R), out=off[1:])
        flat = np.frombuffer("".join(seqs).encode("ascii", "replace") or b"\0", dtype=np.uint8)
        hr, hl, hc = self.filter.scan(flat, off, min_matches)
        return dict(zip(zip(hr.tolist(), hl.tolist()), hc.tolist()))

    def filter_reads(self, names, seqs, min_matches=MIN_MATCHES, max_reads=MAX_READS_PER_LOCUS):
        """-> (per_locus, reads): ``per_locus[vid] = (count, [names ...])`` in ``vntr_ids`` order as the
        binary prints them, ``reads`` = sorted list of (name, sequence) of every listed read."""
        occ = self.occurrences(seqs, min_matches)
        accepted = {vid: {} for vid in self.vntr_ids}          # vntr_read_list (main.cc:246)
        for (r, vid) in sorted(occ):                            # reads in file order, loci ascending
            if vid not in accepted:
                accepted[vid] = {}
            if len(accepted[vid]) > max_reads * 3:
                continue
            accepted[vid][names[r]] = occ[(r, vid)]
        seq_of = {}
        for (r, vid) in occ:
            seq_of[names[r]] = seqs[r]
        per_locus, listed = [], set()
        for vid in self.vntr_ids:
            ranked = sorted(((c, n) for n, c in accepted[vid].items()), reverse=True)
            shown = [n for _, n in ranked[:max_reads + 1]]
            listed.update(shown)
            per_locus.append((vid, min(len(ranked), max_reads), shown))
        reads = [(n, seq_of[n]) for n in sorted(listed)]
        return per_locus, reads

    def format_output(self, names, seqs, min_matches=MIN_MATCHES, max_reads=MAX_READS_PER_LOCUS):
        """The binary's stdout, byte for byte (parity tests)."""
        per_locus, reads = self.filter_reads(names, seqs, min_matches, max_reads)
        lines = []
        for vid, count, shown in per_locus:
            lines.append(" ".join([str(vid), str(count)] + shown))
        for n, s in reads:
            lines.append("%s %s" % (n, s))
        return "\n".join(lines) + "\n" if lines else ""


def read_fasta_pairs(path):
    """The binary reads strict two-line records (main.cc:254-257)."""
    names, seqs = [], []
    with open(path) as fh:
        while True:
            name = fh.readline()
            seq = fh.readline()
            if not name or not seq:
                break
            names.append(name.rstrip("\n")[1:])
            seqs.append(seq.rstrip("\n"))
    return names, seqs
