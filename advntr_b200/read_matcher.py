"""Per-locus read-matcher HMM: left-flank matcher + unrolled repeat units + right-flank matcher.

Host-side restatement of the model builders in ``/root/reference/advntr/hmm_utils.py:289-595``
and the repeat-unit profile of ``/root/reference/advntr/profile_hmm.py:13-175`` on top of
``advntr_b200.pomegranate``.  The reference's own ``hmm_utils.py`` also runs unmodified on
that module (tests/test_builder_parity.py proves both give bit-identical baked tables); this
restatement exists so the package is self-contained where the reference tree is absent.

State names are the contract the path consumers parse (``M{i}_{copy}``, ``I..``, ``D..``,
``unit_start_k``, ``unit_end_k``, ``.._suffix``, ``.._prefix``) and the order in which states
and transitions are inserted is part of the parity contract too: it fixes the in-edge order
of every state after ``bake`` and therefore Viterbi tie-breaking.
"""
from __future__ import annotations

import numpy as np

from . import pomegranate as _default_backend

ALPHABET = "ACGT"
DEFAULT_MAX_ERROR_RATE = 0.05     # settings.py:28 (Illumina); 0.3 for PacBio/nanopore


# --------------------------------------------------------------------------------- profile
class _SparseRow(dict):
    """A transition row of the repeat profile: pairs the alignment never showed read as 0."""

    def __missing__(self, key):
        return 0


def repeat_profile(alignment, error_rate):
    """Transition / emission probabilities of one repeat unit from aligned repeat segments.

    ``profile_hmm.py:13-161``.  ``alignment`` is a list of equal-length strings over
    ``ACGT-``.  Returns ``(transition, emission)`` as dicts of dicts keyed by state labels
    ``unit_start, I0, M1, D1, I1, ..., unit_end`` (no copy suffix).
    """
    n_seq = len(alignment)
    width = len(alignment[0])
    pseudo = (n_seq / 4.0) * (error_rate / 10)
    gap_limit = 0.5 * n_seq
    insert_cols = set(j for j in range(width)
                      if sum(1.0 for row in alignment if row[j] == "-") >= gap_limit)
    R = width - len(insert_cols)            # match columns

    emission = {"unit_start": dict.fromkeys(ALPHABET, 0), "unit_end": dict.fromkeys(ALPHABET, 0),
                "I0": dict.fromkeys(ALPHABET, 0)}
    for i in range(1, R + 1):
        for kind in "IMD":
            emission["%s%d" % (kind, i)] = dict.fromkeys(ALPHABET, 0)

    # the state path every aligned segment takes through the unit
    walks = []
    for row in alignment:
        walk, col = [], 1
        for j, ch in enumerate(row):
            if j in insert_cols:
                if ch != "-":
                    label = "I%d" % (col - 1)
                    walk.append(label)
                    emission[label][ch] += 1
            else:
                if ch == "-":
                    walk.append("D%d" % col)
                else:
                    label = "M%d" % col
                    walk.append(label)
                    emission[label][ch] += 1
                col += 1
        walks.append(walk)

    for label, table in emission.items():
        if label in ("unit_start", "unit_end") or label.startswith("D"):
            continue
        seen = 0
        for ch in table:
            seen += table[ch]
        if seen > 0:
            norm = 0
            for ch in table:
                table[ch] = (1.0 * table[ch]) / seen + pseudo
                norm += 1.0 * table[ch]
            for ch in table:
                table[ch] = table[ch] / norm
        else:
            for ch in table:
                table[ch] = 1.0 / len(ALPHABET)

    transition = {"unit_start": {"I0": 0, "D1": 0, "M1": 0}}
    for walk in walks:
        transition["unit_start"][walk[0]] += 1
    transition["I0"] = {"I0": 0, "D1": 0, "M1": 0}
    for walk in walks:
        for a, b in zip(walk[:-1], walk[1:]):
            row = transition.setdefault(a, {})
            row[b] = row.get(b, 0) + 1
        row = transition.setdefault(walk[-1], {})
        row["unit_end"] = row.get("unit_end", 0) + 1
    for kind in "IDM":
        label = "%s%d" % (kind, R)
        if label not in transition:
            transition[label] = {"unit_end": 0}
    for i in range(1, R + 1):
        for kind in "IMD":
            transition.setdefault("%s%d" % (kind, i), {})

    last = str(R)
    for label, row in transition.items():
        if label == "unit_end":
            continue
        seen = 0
        for b in row:
            seen += row[b]
        if label not in ("unit_start", "I0"):
            idx = label[1:]
            if idx != last:
                for b in ("I" + idx, "D%d" % (int(idx) + 1), "M%d" % (int(idx) + 1)):
                    row.setdefault(b, 0)
            else:
                row.setdefault("I" + idx, 0)
                row.setdefault("unit_end", 0)
        for b in row:
            if seen > 0:
                row[b] = 1.0 * row[b] / seen
                row[b] = (row[b] + pseudo) / (1 + pseudo * len(row))
            elif len(row) == 3:
                row[b] = 1.0 / 3
            elif len(row) == 2:
                row[b] = 1.0 / 2

    # profile_hmm.py:150-160 fills in a 0 for every (state, state) pair that was never seen; rows that
    # answer 0 for a missing key say the same without the (3R+2)^2 dictionary writes per locus
    labels = ["unit_start", "I0"]
    for i in range(1, R + 1):
        labels += ["M%d" % i, "D%d" % i, "I%d" % i]
    labels.append("unit_end")
    return {a: _SparseRow(transition.get(a, ())) for a in labels}, emission


def align_repeat_segments(segments):
    """``profile_hmm.py:165-171`` shells out to MUSCLE for more than one segment.  MUSCLE is an
    external binary that this package does not ship; segments of equal length are taken as
    already aligned (identity alignment), anything else must be aligned by the caller."""
    if len(segments) > 1 and len(set(len(s) for s in segments)) != 1:
        raise ValueError("repeat segments of unequal length need a multiple alignment "
                         "(pass aligned segments with '-' gaps)")
    return list(segments)


# ------------------------------------------------------------------------------ flank models
def _flank_matcher(pattern, tag, error_rate, pom=None):
    """Suffix (left flank, ``tag='suffix'``) or prefix (right flank) matcher.

    ``hmm_utils.py:357-420`` / ``:290-353``.  They differ in three places only: the suffix
    matcher may be entered at any match column (a read can start inside the flank), the
    prefix matcher may be left from any match column with probability 0.01 (a read can end
    inside the flank), and the order in which the entry transitions are declared.
    """
    pom = pom or _default_backend
    State, DiscreteDistribution = pom.State, pom.DiscreteDistribution
    L = len(pattern)
    title = "Suffix Matcher HMM Model" if tag == "suffix" else "Prefix Matcher HMM Model"
    hmm = pom.HiddenMarkovModel(name=title)
    uniform = DiscreteDistribution(dict.fromkeys(ALPHABET, 0.25))
    ins = [State(uniform, name="I%s_%s" % (i, tag)) for i in range(L + 1)]
    mat = []
    for i, base in enumerate(pattern):
        probs = dict.fromkeys(ALPHABET, 0.01)
        probs[base] = 0.97
        mat.append(State(DiscreteDistribution(probs), name="M%s_%s" % (i + 1, tag)))
    dele = [State(None, name="D%s_%s" % (i + 1, tag)) for i in range(L)]
    gate_in = State(None, name="%s_start_%s" % (tag, tag))
    gate_out = State(None, name="%s_end_%s" % (tag, tag))
    hmm.add_states(ins + mat + dele + [gate_in, gate_out])

    p_ins = error_rate * 2 / 5
    p_del = error_rate * 1 / 5
    p_adv = 1 - p_ins - p_del
    T = hmm.add_transition
    T(hmm.start, gate_in, 1)
    T(gate_out, hmm.end, 1)
    if tag == "suffix":
        T(gate_in, dele[0], p_del)
        T(gate_in, ins[0], p_ins)
        for k in range(L):
            T(gate_in, mat[k], p_adv / L)
    else:
        T(gate_in, mat[0], p_adv)
        T(gate_in, dele[0], p_del)
        T(gate_in, ins[0], p_ins)
    T(ins[0], ins[0], p_ins)
    T(ins[0], dele[0], p_del)
    T(ins[0], mat[0], p_adv)

    z = L - 1
    T(dele[z], gate_out, 1 - p_ins)
    T(dele[z], ins[z + 1], p_ins)
    T(mat[z], gate_out, 1 - p_ins)
    T(mat[z], ins[z + 1], p_ins)
    T(ins[z + 1], ins[z + 1], p_ins)
    T(ins[z + 1], gate_out, 1 - p_ins)

    for k in range(L):
        T(mat[k], ins[k + 1], p_ins)
        T(dele[k], ins[k + 1], p_ins)
        T(ins[k + 1], ins[k + 1], p_ins)
        if k < z:
            T(ins[k + 1], mat[k + 1], p_adv)
            T(ins[k + 1], dele[k + 1], p_del)
            if tag == "suffix":
                T(mat[k], mat[k + 1], p_adv)
                T(mat[k], dele[k + 1], p_del)
            else:
                T(mat[k], mat[k + 1], p_adv - 0.01)
                T(mat[k], dele[k + 1], p_del)
                T(mat[k], gate_out, 0.01)
            T(dele[k], dele[k + 1], p_del)
            T(dele[k], mat[k + 1], p_adv)
    hmm.bake(merge=None)
    return hmm


def get_suffix_matcher_hmm(pattern, error_rate=DEFAULT_MAX_ERROR_RATE, pom=None):
    return _flank_matcher(pattern, "suffix", error_rate, pom)


def get_prefix_matcher_hmm(pattern, error_rate=DEFAULT_MAX_ERROR_RATE, pom=None):
    return _flank_matcher(pattern, "prefix", error_rate, pom)


# ----------------------------------------------------------------------------- repeat models
def get_constant_number_of_repeats_matcher_hmm(patterns, copies, error_rate=DEFAULT_MAX_ERROR_RATE,
                                               profile=None, pom=None):
    """``copies`` unrolled copies of the repeat-unit profile (``hmm_utils.py:424-497``)."""
    pom = pom or _default_backend
    State, DiscreteDistribution = pom.State, pom.DiscreteDistribution
    hmm = pom.HiddenMarkovModel(name="Repeating Pattern Matcher HMM Model")
    trans, emis = profile or repeat_profile(align_repeat_segments(patterns), error_rate)
    R = sum(1 for label in emis if label.startswith("M"))
    T = hmm.add_transition
    prev_out = None
    for k in range(copies):
        ins = [State(DiscreteDistribution(emis["I%s" % i]), name="I%s_%s" % (i, k)) for i in range(R + 1)]
        mat = [State(DiscreteDistribution(emis["M%s" % i]), name="M%s_%s" % (i, k)) for i in range(1, R + 1)]
        dele = [State(None, name="D%s_%s" % (i, k)) for i in range(1, R + 1)]
        gate_in = State(None, name="unit_start_%s" % k)
        gate_out = State(None, name="unit_end_%s" % k)
        hmm.add_states(ins + mat + dele + [gate_in, gate_out])
        T(prev_out if k else hmm.start, gate_in, 1)
        if k == copies - 1:
            T(gate_out, hmm.end, 1)

        T(gate_in, mat[0], trans["unit_start"]["M1"])
        T(gate_in, dele[0], trans["unit_start"]["D1"])
        T(gate_in, ins[0], trans["unit_start"]["I0"])
        T(ins[0], ins[0], trans["I0"]["I0"])
        T(ins[0], dele[0], trans["I0"]["D1"])
        T(ins[0], mat[0], trans["I0"]["M1"])

        dR, mR, iR = "D%s" % R, "M%s" % R, "I%s" % R
        T(dele[R - 1], gate_out, trans[dR]["unit_end"])
        T(dele[R - 1], ins[R], trans[dR][iR])
        T(mat[R - 1], gate_out, trans[mR]["unit_end"])
        T(mat[R - 1], ins[R], trans[mR][iR])
        T(ins[R], ins[R], trans[iR][iR])
        T(ins[R], gate_out, trans[iR]["unit_end"])

        for i in range(1, R + 1):
            m_i, d_i, i_i = "M%s" % i, "D%s" % i, "I%s" % i
            T(mat[i - 1], ins[i], trans[m_i][i_i])
            T(dele[i - 1], ins[i], trans[d_i][i_i])
            T(ins[i], ins[i], trans[i_i][i_i])
            if i < R:
                m_n, d_n = "M%s" % (i + 1), "D%s" % (i + 1)
                T(ins[i], mat[i], trans[i_i][m_n])
                T(ins[i], dele[i], trans[i_i][d_n])
                T(mat[i - 1], mat[i], trans[m_i][m_n])
                T(mat[i - 1], dele[i], trans[m_i][d_n])
                T(dele[i - 1], mat[i], trans[d_i][m_n])
                T(dele[i - 1], dele[i], trans[d_i][d_n])
        prev_out = gate_out
    hmm.bake(merge=None)
    return hmm


def _last_nonzero(row):
    nz = np.nonzero(row)[0]
    return int(nz[-1]) if len(nz) else None


def get_variable_number_of_repeats_matcher_hmm(patterns, copies=1, error_rate=DEFAULT_MAX_ERROR_RATE,
                                               profile=None, pom=None):
    """Let a read leave the repeat block after any unit (``hmm_utils.py:501-549``).

    Goes through the same dense-matrix round trip as the reference (exp of the stored logs,
    edit, ``from_matrix`` re-logs) because the resulting last-ulp values are part of parity.
    """
    pom = pom or _default_backend
    State = pom.State
    base = get_constant_number_of_repeats_matcher_hmm(patterns, copies, error_rate, profile, pom)
    if hasattr(base, "sparse_transition_matrix"):
        return _variable_repeats_sparse(base, pom)
    mat = base.dense_transition_matrix()
    m = len(mat)
    states = list(base.states)
    states.append(State(None, name="start_repeating_pattern_match"))
    states.append(State(None, name="end_repeating_pattern_match"))
    enter, leave = m, m + 1
    mat = np.c_[mat, np.zeros(m), np.zeros(m)]
    mat = np.r_[mat, [np.zeros(m + 2)], [np.zeros(m + 2)]]

    first_unit = _last_nonzero(mat[base.start_index])
    mat[base.start_index][first_unit] = 0.0
    mat[base.start_index][enter] = 1
    mat[enter][first_unit] = 1
    for i, st in enumerate(states):
        if st.name.startswith("unit_end"):
            nxt = _last_nonzero(mat[i])
            mat[i][nxt] = 0.5
            mat[i][leave] = 0.5
    mat[leave][base.end_index] = 1

    starts = np.zeros(m + 2)
    starts[base.start_index] = 1.0
    ends = np.zeros(m + 2)
    ends[base.end_index] = 1.0
    out = pom.HiddenMarkovModel.from_matrix(mat, [s.distribution for s in states], starts, ends,
                                        name="Repeat Matcher HMM Model",
                                        state_names=[s.name for s in states], merge=None)
    out.bake(merge=None)
    return out


def _last_in_row(cells, i):
    cols = [j for (r, j) in cells if r == i]
    return max(cols) if cols else None


def _variable_repeats_sparse(base, pom):
    """Same edits as the dense branch of get_variable_number_of_repeats_matcher_hmm on the
    non-zero cells only (O(E) instead of O(m^2); bit-identical result)."""
    cells = base.sparse_transition_matrix()
    m = len(base.states)
    states = list(base.states)
    states.append(pom.State(None, name="start_repeating_pattern_match"))
    states.append(pom.State(None, name="end_repeating_pattern_match"))
    enter, leave = m, m + 1
    by_row = {}
    for (i, j) in cells:
        if i not in by_row or j > by_row[i]:
            by_row[i] = j
    first_unit = by_row[base.start_index]
    del cells[(base.start_index, first_unit)]
    cells[(base.start_index, enter)] = 1
    cells[(enter, first_unit)] = 1
    for i, st in enumerate(states):
        if st.name.startswith("unit_end"):
            cells[(i, by_row[i])] = 0.5
            cells[(i, leave)] = 0.5
    cells[(leave, base.end_index)] = 1
    starts = np.zeros(m + 2)
    starts[base.start_index] = 1.0
    ends = np.zeros(m + 2)
    ends[base.end_index] = 1.0
    out = pom.HiddenMarkovModel.from_sparse(cells, [s.distribution for s in states], starts, ends,
                                            name="Repeat Matcher HMM Model",
                                            state_names=[s.name for s in states], merge=None)
    out.bake(merge=None)
    return out


def get_read_matcher_model(left_flanking_region, right_flanking_region, patterns, copies=1,
                           vpaths=None, error_rate=DEFAULT_MAX_ERROR_RATE, profile=None, pom=None):
    """The model every read of a locus is decoded against (``hmm_utils.py:553-595``)."""
    if vpaths:
        # --update (vntr_finder.py:667-698): the repeat-unit profile is re-estimated from the repeat
        # segments the current model found in the selected reads (hmm_utils.py:427-429)
        from . import path_utils
        alignment = path_utils.get_multiple_alignment_of_repeats_from_reads(vpaths)
        profile = repeat_profile(alignment, error_rate)
    pom = pom or _default_backend
    hmm = get_suffix_matcher_hmm(left_flanking_region, error_rate, pom)
    hmm.concatenate(get_variable_number_of_repeats_matcher_hmm(patterns, copies, error_rate, profile, pom))
    hmm.concatenate(get_prefix_matcher_hmm(right_flanking_region, error_rate, pom))
    hmm.bake(merge=None)

    names = [s.name for s in hmm.states]
    if hasattr(hmm, "sparse_transition_matrix"):
        return _read_matcher_sparse(hmm, names, pom)
    mat = hmm.dense_transition_matrix()
    first_copy, repeat_matches, entry = [], [], None
    for i, nm in enumerate(names):
        tail = nm.split("_")[-1]
        if nm[0] == "M" and tail == "0":
            first_copy.append(i)
        if nm[0] == "M" and tail not in ("prefix", "suffix"):
            repeat_matches.append(i)
        if nm == "suffix_start_suffix":
            entry = i
    # a read may start in the left flank (0.3) or at any column of the first repeat copy (0.7)
    mat[hmm.start_index][entry] = 0.3
    for i in first_copy:
        mat[hmm.start_index][i] = 0.7 / len(first_copy)
    # ... and may end at any repeat match state
    for i in repeat_matches:
        to_end = 0.7 / len(repeat_matches)
        scale = 1 + to_end
        nz = mat[i] != 0
        mat[i][nz] /= scale
        mat[i][hmm.end_index] = to_end / scale

    starts = np.zeros(len(names))
    starts[hmm.start_index] = 1.0
    ends = np.zeros(len(names))
    ends[hmm.end_index] = 1.0
    out = pom.HiddenMarkovModel.from_matrix(mat, [s.distribution for s in hmm.states], starts, ends,
                                        name="Read Matcher", state_names=names, merge=None)
    out.bake(merge=None)
    return out


def _read_matcher_sparse(hmm, names, pom):
    """The final edit of get_read_matcher_model (hmm_utils.py:561-594) on the non-zero cells."""
    cells = hmm.sparse_transition_matrix()
    first_copy, repeat_matches, entry = [], set(), None
    for i, nm in enumerate(names):
        tail = nm.split("_")[-1]
        if nm[0] == "M" and tail == "0":
            first_copy.append(i)
        if nm[0] == "M" and tail not in ("prefix", "suffix"):
            repeat_matches.add(i)
        if nm == "suffix_start_suffix":
            entry = i
    cells[(hmm.start_index, entry)] = 0.3
    for i in first_copy:
        cells[(hmm.start_index, i)] = 0.7 / len(first_copy)
    to_end = 0.7 / len(repeat_matches)
    scale = 1 + to_end
    for key in list(cells):
        if key[0] in repeat_matches:
            cells[key] = cells[key] / scale
    for i in repeat_matches:
        cells[(i, hmm.end_index)] = to_end / scale
    starts = np.zeros(len(names))
    starts[hmm.start_index] = 1.0
    ends = np.zeros(len(names))
    ends[hmm.end_index] = 1.0
    out = pom.HiddenMarkovModel.from_sparse(cells, [s.distribution for s in hmm.states], starts, ends,
                                            name="Read Matcher", state_names=names, merge=None)
    out.bake(merge=None)
    return out


def build_reference_repeat_finder_hmm(patterns, copies=1, pom=None):
    """Model that segments the REFERENCE copy of a VNTR into its repeat units when a locus is added
    (``hmm_utils.py:598-680``, used by ``reference_vntr.py:80-87``): ``copies`` unrolled copies of
    ``patterns[0]`` with fixed 0.98 / 0.01 / 0.01 transitions, random-sequence states before and
    after, baked with the default ``merge='All'``."""
    pom = pom or _default_backend
    State, DiscreteDistribution = pom.State, pom.DiscreteDistribution
    pattern = patterns[0]
    R = len(pattern)
    hmm = pom.HiddenMarkovModel(name="HMM Model")
    uniform = DiscreteDistribution({"A": 0.25, "C": 0.25, "G": 0.25, "T": 0.25})
    lead = State(uniform, name="start_random_matches")
    trail = State(uniform, name="end_random_matches")
    hmm.add_states([lead, trail])
    T = hmm.add_transition
    prev_out = None
    for k in range(copies):
        ins = [State(uniform, name="I%s_%s" % (i, k)) for i in range(R + 1)]
        mat = []
        for i in range(R):
            table = {"A": 0.01, "C": 0.01, "G": 0.01, "T": 0.01}
            table[pattern[i]] = 0.97
            mat.append(State(DiscreteDistribution(table), name="M%s_%s" % (i + 1, k)))
        dele = [State(None, name="D%s_%s" % (i + 1, k)) for i in range(R)]
        gate_in = State(None, name="unit_start_%s" % k)
        gate_out = State(None, name="unit_end_%s" % k)
        hmm.add_states(ins + mat + dele + [gate_in, gate_out])
        if k:
            T(prev_out, gate_in, 0.5)
        else:
            T(hmm.start, gate_in, 0.5)
            T(hmm.start, lead, 0.5)
            T(lead, gate_in, 0.5)
            T(lead, lead, 0.5)
        T(gate_out, trail, 0.5)
        if k == copies - 1:
            T(gate_out, hmm.end, 0.5)
            T(trail, trail, 0.5)
            T(trail, hmm.end, 0.5)
        T(gate_in, mat[0], 0.98)
        T(gate_in, dele[0], 0.01)
        T(gate_in, ins[0], 0.01)
        T(ins[0], ins[0], 0.01)
        T(ins[0], dele[0], 0.01)
        T(ins[0], mat[0], 0.98)
        T(dele[R - 1], gate_out, 0.99)
        T(dele[R - 1], ins[R], 0.01)
        T(mat[R - 1], gate_out, 0.99)
        T(mat[R - 1], ins[R], 0.01)
        T(ins[R], ins[R], 0.01)
        T(ins[R], gate_out, 0.99)
        for i in range(R):
            T(mat[i], ins[i + 1], 0.01)
            T(dele[i], ins[i + 1], 0.01)
            T(ins[i + 1], ins[i + 1], 0.01)
            if i < R - 1:
                T(ins[i + 1], mat[i + 1], 0.98)
                T(ins[i + 1], dele[i + 1], 0.01)
                T(mat[i], mat[i + 1], 0.98)
                T(mat[i], dele[i + 1], 0.01)
                T(dele[i], dele[i + 1], 0.01)
                T(dele[i], mat[i + 1], 0.98)
        prev_out = gate_out
    hmm.bake()
    return hmm


def find_repeat_segments(pattern, estimated_repeats, region_in_ref, pom=None):
    """``ReferenceVNTR.find_repeat_segments`` (``reference_vntr.py:80-87``): Viterbi-decode the
    reference region against the repeat finder and cut it at the unit boundaries."""
    from . import path_utils
    model = build_reference_repeat_finder_hmm([pattern], copies=estimated_repeats, pom=pom)
    logp, path = model.viterbi(region_in_ref)
    visited = [state.name for _, state in path[1:-1]]
    return path_utils.get_repeat_segments_from_visited_states_and_region(visited, region_in_ref)


def copies_for_read_length(read_length, pattern_length):
    """``vntr_finder.py:98-99``: unrolled copies needed to cover a read."""
    return int(round(float(read_length) / pattern_length + 0.5))


def build_vntr_matcher_hmm(left_flank, right_flank, repeat_segments, copies, flank_size=100,
                           error_rate=DEFAULT_MAX_ERROR_RATE, pom=None):
    """``vntr_finder.py:108-115``: trim the flanks and build the read matcher.  ``pom`` selects the
    engine the model is built on (default: this package's; tests and the CPU baseline pass the
    compiled reference pomegranate to get a reference-engine model with identical tables)."""
    return get_read_matcher_model(left_flank[-flank_size:], right_flank[:flank_size],
                                  repeat_segments, copies, error_rate=error_rate, pom=pom)
