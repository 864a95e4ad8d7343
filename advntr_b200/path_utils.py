"""Consumers of a Viterbi state path (the contract downstream of the hot path).

Restates ``/root/reference/advntr/hmm_utils.py:70-91, 122-286`` and the recruitment /
spanning predicates of ``vntr_finder.py:179-190, 311-322``.  They read nothing but the state
NAMES on the path (``vpath[1:-1]``: the model's own start/end are dropped), which is why the
engine returns the reference's exact state indices.  Implemented as one pass over the path
that fills a ``PathSummary``; the reference-named helpers are thin views of it.
"""
from __future__ import annotations

MIN_BP_IN_REPEAT = 3     # hmm_utils.py:165


def is_match_state(name):
    return name.startswith("M")


def is_emitting_state(name):
    return name.startswith(("M", "I", "start_random_matches", "end_random_matches"))


def _names(vpath):
    return [state.name for _, state in vpath[1:-1]]


class PathSummary(object):
    """Everything adVNTR derives from one path, computed in a single walk."""

    __slots__ = ("n_emitted", "n_match", "repeat_bp", "left_bp", "right_bp", "unit_starts",
                 "unit_ends", "repeats", "unit_lengths")

    def __init__(self, names):
        emitted = [is_emitting_state(n) for n in names]
        total = sum(emitted)
        self.n_emitted = total
        self.n_match = sum(1 for n in names if is_match_state(n))
        self.repeat_bp = self.left_bp = self.right_bp = 0
        starts = ends = 0
        first_start = last_start = first_end = last_end = None
        lengths, open_at = [], None
        bp = 0
        for n, emits in zip(names, emitted):
            if emits:
                bp += 1
                if n.endswith("suffix"):
                    self.left_bp += 1
                elif n.endswith("prefix"):
                    self.right_bp += 1
                if not n.endswith("fix"):
                    self.repeat_bp += 1
            if n.startswith("unit_start"):
                if total - bp >= MIN_BP_IN_REPEAT:
                    if first_start is None:
                        first_start = bp
                    last_start = bp
                    starts += 1
            if n.startswith("unit_end"):
                if open_at is not None:
                    lengths.append(bp - open_at)
                if bp >= MIN_BP_IN_REPEAT:
                    if first_end is None:
                        first_end = bp
                    last_end = bp
                    ends += 1
            if n.startswith("unit_start"):
                open_at = bp
        bonus = 0
        if None not in (first_start, last_start, first_end, last_end):
            if first_end < first_start and last_start > last_end:
                bonus = 1       # partial unit at both ends of the read (hmm_utils.py:184-187)
        self.unit_starts, self.unit_ends = starts, ends
        self.repeats = max(starts, ends) + bonus
        self.unit_lengths = lengths


def summarize(vpath):
    return PathSummary(_names(vpath))


def get_number_of_repeats_in_vpath(vpath):
    return summarize(vpath).repeats


def get_number_of_matches_in_vpath(vpath):
    return summarize(vpath).n_match


def get_number_of_repeat_bp_matches_in_vpath(vpath):
    return summarize(vpath).repeat_bp


def get_left_flanking_region_size_in_vpath(vpath):
    return summarize(vpath).left_bp


def get_right_flanking_region_size_in_vpath(vpath):
    return summarize(vpath).right_bp


def get_repeating_pattern_lengths(visited_states):
    return PathSummary(list(visited_states)).unit_lengths


def get_repeat_segments_from_visited_states_and_region(visited_states, region):
    """``hmm_utils.py:144-152``: consecutive pieces of ``region`` with the repeat-unit lengths."""
    out, at = [], 0
    for n in get_repeating_pattern_lengths(visited_states):
        out.append(region[at:at + n])
        at += n
    return out


def flank_match_counts(vpath, sequence, left_flank, right_flank):
    """(hits, bases) per flank, keyed ``"prefix"`` (right flank) / ``"suffix"`` (left flank): flank match states
    whose read base equals the flank base, and read bases spent in the flank (``hmm_utils.py:209-262``).
    The on-device path reducers deliver the same four numbers (``advhmm_read_summary``)."""
    names = _names(vpath)
    deepest = -1                      # index of the last left-flank column the path used
    prev = names[0]
    for n in names:
        if "suffix_end_suffix" in n:
            deepest = int(prev.split("_")[0][1:])
            break
        prev = n
    hits = {"prefix": 0, "suffix": 0}
    bases = {"prefix": 0, "suffix": 0}
    pos = 0
    for n in names:
        if "start" in n or "end" in n:
            continue
        col = int(n.split("_")[0][1:])
        emits = is_emitting_state(n)
        for side in ("prefix", "suffix"):
            if n.endswith(side):
                if is_match_state(n):
                    want = right_flank[col - 1] if side == "prefix" else left_flank[-(deepest - col + 1)]
                    if sequence[pos] == want:
                        hits[side] += 1
                if emits:
                    bases[side] += 1
        if emits:
            pos += 1
    return hits, bases


def get_flanking_regions_matching_rate(vpath, sequence, left_flank, right_flank, accuracy_filter=False):
    """Fraction of flank match states whose read base equals the flank base; the smaller of the
    two flanks' rates (``hmm_utils.py:209-268``)."""
    hits, bases = flank_match_counts(vpath, sequence, left_flank, right_flank)
    empty = 0.00001 if accuracy_filter else 1
    right = float(hits["prefix"]) / bases["prefix"] if bases["prefix"] else empty
    left = float(hits["suffix"]) / bases["suffix"] if bases["suffix"] else empty
    return min(right, left)


def extract_repeating_segments_from_read(sequence, visited_states):
    """Read substrings (and their state runs) between unit_start and unit_end
    (``hmm_utils.py:70-91``)."""
    repeats, runs = [], []
    open_pos = open_idx = None
    pos = 0
    for i, n in enumerate(visited_states):
        if n.startswith("unit_end") and open_pos is not None:
            repeats.append(sequence[open_pos:pos])
            runs.append(list(visited_states[open_idx + 1:i]))
        if n.startswith("unit_start"):
            open_pos, open_idx = pos, i
        if is_emitting_state(n):
            pos += 1
    return repeats, runs


def get_multiple_alignment_of_viterbi_paths(repeats_sequences, repeats_visited_states):
    """Column-anchored multiple alignment of repeat segments from their state runs
    (``hmm_utils.py:23-67``): one column per match index, followed by as many insert columns as
    the most-inserting segment needs.

    The reference marks EVERY remaining occurrence of a column's state as consumed when it places
    the first one, so a segment that visits ``I_k`` twice places one base for it and every later
    base of that segment moves one column to the left (its last base is dropped).  Kept: the
    re-estimated profile, and with it the tables the reads are decoded against, depend on it."""
    width = {}                      # 'M7' / 'I3' / 'D2' -> most visits by one segment
    last_index = 0
    for run in repeats_visited_states:
        visits = {}
        for name in run:
            label = name.split("_")[0]
            visits[label] = visits.get(label, 0) + 1
        for label, n in visits.items():
            last_index = max(last_index, int(label[1:]))
            width[label] = max(width.get(label, n), n)
    columns = []
    for i in range(last_index + 1):
        for label in ("M%s" % i, "I%s" % i):
            columns.extend([label] * width.get(label, 0))
    rows = []
    for seq, run in zip(repeats_sequences, repeats_visited_states):
        pending = [name.split("_")[0] for name in run]
        row, used = [], 0
        for label in columns:
            if label in pending:
                pending = [None if x == label else x for x in pending]
                row.append(seq[used])
                used += 1
            else:
                row.append("-")
        rows.append("".join(row))
    return rows


def get_multiple_alignment_of_repeats_from_reads(sequence_vpath_list):
    """``hmm_utils.py:94-103``: the repeat segments of every (read, Viterbi path), aligned."""
    seqs, runs = [], []
    for sequence, vpath in sequence_vpath_list:
        repeats, states = extract_repeating_segments_from_read(sequence, _names(vpath))
        seqs += repeats
        runs += states
    return get_multiple_alignment_of_viterbi_paths(seqs, runs)


def recruit_read(logp, vpath, min_score_to_count_read, read_sequence, left_flank, right_flank):
    """``vntr_finder.py:179-190``: does this read belong to the locus?"""
    if min_score_to_count_read is not None and logp > min_score_to_count_read:
        return get_flanking_regions_matching_rate(vpath, read_sequence, left_flank, right_flank) >= 0.9
    length = len(read_sequence)
    if min_score_to_count_read is None and \
            get_number_of_matches_in_vpath(vpath) >= 0.9 * length and logp > -length:
        return get_flanking_regions_matching_rate(vpath, read_sequence, left_flank, right_flank) >= 0.9
    return False


def get_emitted_basepair_from_visited_states(state, visited_states, sequence):
    """Read base emitted at the first visit of ``state`` (``hmm_utils.py:105-112``)."""
    pos = 0
    for n in visited_states:
        if n == state:
            return sequence[pos]
        if is_emitting_state(n):
            pos += 1
    return None


def read_flanks_repeats_with_confidence(vpath, sequence, left_flank, right_flank, min_left=5, min_right=5):
    """Spanning-read predicate (``vntr_finder.py:311-322``)."""
    if get_flanking_regions_matching_rate(vpath, sequence, left_flank, right_flank) < 0.95:
        return False
    s = summarize(vpath)
    return s.left_bp > min_left and s.right_bp > min_right


def frameshift_mutations(selected, pattern_length):
    """Indel states inside repeat units whose emitted length is off by <= 2 bp, counted over the
    selected reads (the path-walking part of ``vntr_finder.py:265-296``).  ``selected`` is an
    iterable of objects with ``.vpath`` and ``.sequence``.  Returns ``(mutations, repeat_bp)``."""
    mutations, repeat_bp = {}, 0
    for read in selected:
        names = _names(read.vpath)
        lengths = get_repeating_pattern_lengths(names)
        repeat_bp += get_number_of_repeat_bp_matches_in_vpath(read.vpath)
        unit = None
        for n in names:
            if n.endswith("fix") or n.startswith("M"):
                continue
            if n.startswith("unit_start"):
                unit = 0 if unit is None else unit + 1
            if unit is None or unit >= len(lengths):
                continue
            if not n.startswith("I") and not n.startswith("D"):
                continue
            if lengths[unit] == pattern_length:
                continue
            label = n.split("_")[0]
            if label.startswith("I"):
                label += get_emitted_basepair_from_visited_states(n, names, read.sequence)
            if abs(lengths[unit] - pattern_length) <= 2:
                mutations[label] = mutations.get(label, 0) + 1
    return mutations, repeat_bp


def frameshift_state_tables(names):
    """(class byte, number in the name) of every state: what ``advhmm_frameshift_candidates`` reads instead of
    the state names (``"I7_2"`` -> kind insert, repeat part, 7)."""
    import numpy as np
    cls = state_classes(names)
    label = np.full(len(names), -1, dtype=np.int32)
    for i, n in enumerate(names):
        if (cls[i] & 7) in (2, 3) and ((cls[i] >> 3) & 3) == 3:
            label[i] = int(n.split("_")[0][1:])
    return cls, label


def frameshift_label(call):
    """The reference's mutation key (``'I7A'`` / ``'D12'``) of an ``advhmm_frameshift_call`` record, or None."""
    if call["kind"] == 0:
        return None
    if call["kind"] == 2:
        return "I%d%s" % (call["column"], "ACGT"[call["base"]])
    return "D%d" % call["column"]


def state_classes(names, emis=None):
    """Class byte of every state for the on-device reducers (``include/advhmm.h``): what the
    functions above parse out of the state names, plus the base a flank match state expects
    (the symbol its emission row favours)."""
    import numpy as np
    out = np.zeros(len(names), dtype=np.uint8)
    for i, n in enumerate(names):
        if n.startswith("M"):
            kind = 1
        elif n.startswith("I"):
            kind = 2
        elif n.startswith("D"):
            kind = 3
        elif n.startswith("unit_start"):
            kind = 4
        elif n.startswith("unit_end"):
            kind = 5
        else:
            kind = 0
        part = 0
        if kind in (1, 2, 3):
            part = 1 if n.endswith("suffix") else 2 if n.endswith("prefix") else 3
        base = 0
        if kind == 1 and part in (1, 2) and emis is not None:
            base = int(np.argmax(emis[i]))
        out[i] = kind | (part << 3) | (base << 5)
    return out


def rescore_paths(baked, codes, paths):
    """Fold ``(v + t) + e`` along state paths with the model's own baked tables, in the reference's
    operation order (``hmm.pyx:2035-2042, 2056-2083``).  For a path Viterbi returned this reproduces its
    log-probability bit for bit: a check of paths that needs no second decoder and works at any batch
    size.  Raises ``AssertionError`` if a path does not run from the start to the end state, uses a
    transition that is not in the model, or does not emit exactly its read.
    ``codes`` / ``paths``: per read, uint8 symbol codes / int state indices.  -> float64 scores."""
    import numpy as np
    m, S = baked["n_states"], baked["silent_start"]
    W = np.full((m, m), -np.inf)
    dst = np.repeat(np.arange(m), np.diff(baked["in_off"]))
    W[baked["in_src"], dst] = baked["in_logp"]
    emis = baked["emis"]
    count = len(paths)
    plen = np.fromiter((len(p) for p in paths), dtype=np.int64, count=count)
    lens = np.fromiter((len(c) for c in codes), dtype=np.int64, count=count)
    P = np.full((count, int(plen.max())), -1, dtype=np.int64)
    for i, p in enumerate(paths):
        P[i, :plen[i]] = p
    sym = np.zeros((count, int(lens.max()) + 1), dtype=np.int64)
    for i, c in enumerate(codes):
        sym[i, :lens[i]] = c
    rows = np.arange(count)
    assert (P[:, 0] == baked["start_index"]).all(), "a path does not begin in the start state"
    assert (P[rows, plen - 1] == baked["end_index"]).all(), "a path does not finish in the end state"
    v = np.zeros(count)
    pos = np.zeros(count, dtype=np.int64)
    for k in range(1, P.shape[1]):
        live = k < plen
        a = np.where(live, P[:, k - 1], 0)
        b = np.where(live, P[:, k], 0)
        t = W[a, b]
        assert np.isfinite(t[live]).all(), "a path uses a transition the model does not have"
        nv = v + np.where(live, t, 0.0)
        emitting = live & (b < S)
        e = emis[np.where(emitting, b, 0), sym[rows, np.minimum(pos, lens)]]
        nv = np.where(emitting, nv + e, nv)
        v = np.where(live, nv, v)
        pos = pos + emitting
    assert (pos == lens).all(), "a path does not emit exactly its read"
    return v
