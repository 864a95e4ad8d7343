"""O(E) compiler for the read-matcher HMM of a locus.

``read_matcher.get_read_matcher_model`` follows the reference literally: three sub-models, two
dense m x m matrix round trips and eight ``bake`` calls per locus (0.1 s here, 1 s in the
reference, O(m^2)).  But for a given *shape* -- flank lengths, repeat-unit columns R, unrolled
copies C -- every locus has the same states, the same edges in the same order; only the numbers
differ.  This module therefore

1. builds one **template** per shape from a literal build (state names, CSR in-edge lists) and
   labels every edge with the parameter it carries (a flank constant, a repeat-profile
   transition, a glue constant), and every emitting state with its emission row;
2. compiles further loci of that shape by evaluating those few hundred parameters through the
   SAME chain of float operations the literal path applies (``log`` -> ``numpy.exp`` -> ``log``
   ... , the match-row rescaling of ``hmm_utils.py:578-584``) and scattering them.

Exactness is not assumed, it is checked: when a template is created the fast tables of the
template locus are compared bit for bit with the literal build, and
``tests/test_fast_compile.py`` does the same across shapes and loci.
"""
from __future__ import annotations

import math
import re

import numpy as np

from . import read_matcher
from .pomegranate import HiddenMarkovModel, State, DiscreteDistribution, NEGINF

_NAME = re.compile(r"^([IMD])(\d+)_(.+)$")

# chains of float operations a parameter goes through between the builder and the final table
CH_CONST0, CH_LOG, CH_F1, CH_F2, CH_F2_MATCH, CH_TO_END = range(6)


def _classify(src, dst, shape):
    """(chain, key) of the edge src -> dst of a final read-matcher model, from the state names."""
    ms, md = _NAME.match(src), _NAME.match(dst)
    part_s = ms.group(3) if ms else None
    part_d = md.group(3) if md else None
    # --- edges created or overwritten by the last from_matrix edit (hmm_utils.py:574-584) ---
    if src == "Suffix Matcher HMM Model-start":
        if dst == "suffix_start_suffix":
            return CH_LOG, ("const", 0.3)
        return CH_LOG, ("first_copy",)
    if dst == "Prefix Matcher HMM Model-end" and ms and ms.group(1) == "M" and part_s not in ("suffix", "prefix"):
        return CH_TO_END, ("to_end",)
    # --- flank matchers ---------------------------------------------------------------------
    for tag in ("suffix", "prefix"):
        gate_in, gate_out = "%s_start_%s" % (tag, tag), "%s_end_%s" % (tag, tag)
        inside_s = part_s == tag or src in (gate_in, gate_out)
        inside_d = part_d == tag or dst in (gate_in, gate_out)
        if inside_s and inside_d:
            if md:
                kind = md.group(1)
                if kind == "I":
                    return CH_F1, ("flank", tag, "ie")
                if kind == "D":
                    return CH_F1, ("flank", tag, "de")
                if src == gate_in and tag == "suffix":
                    return CH_F1, ("flank", tag, "adv_over_L")
                if ms and ms.group(1) == "M" and tag == "prefix":
                    return CH_F1, ("flank", tag, "adv_m01")
                return CH_F1, ("flank", tag, "adv")
            if dst == gate_out:
                if ms and ms.group(1) == "M" and tag == "prefix" and int(ms.group(2)) < shape["L_" + tag]:
                    return CH_F1, ("const", 0.01)
                return CH_F1, ("flank", tag, "one_minus_ie")
    # --- repeat units -----------------------------------------------------------------------
    def label(name, m):
        if m and m.group(3).isdigit():
            return m.group(1) + m.group(2)
        if name.startswith("unit_start_"):
            return "unit_start"
        if name.startswith("unit_end_"):
            return "unit_end"
        return None
    ls, ld = label(src, ms), label(dst, md)
    if ls is not None and ld is not None and ls != "unit_end":
        chain = CH_F2_MATCH if ls[0] == "M" else CH_F2
        return chain, ("profile", ls, ld)
    if ls == "unit_end":
        return CH_F1, ("const", 0.5)          # hmm_utils.py:530-536
    # --- glue: probability-1 edges between the parts ------------------------------------------
    return CH_CONST0, ("one",)


class ShapeTemplate(object):
    """Structure of the read-matcher model for one (L_left, L_right, R, C) + parameter labels."""

    def __init__(self, left, right, segments, copies, error_rate):
        aligned = read_matcher.align_repeat_segments(segments)
        literal = read_matcher.get_read_matcher_model(left, right, segments, copies, error_rate=error_rate)
        b = literal.baked
        self.names = [s.name for s in literal.states]
        self.n_states, self.silent_start = b["n_states"], b["silent_start"]
        self.start_index, self.end_index, self.finite = b["start_index"], b["end_index"], b["finite"]
        self.in_off, self.in_src = b["in_off"], b["in_src"]
        self.L_left, self.L_right = len(left), len(right)
        trans, emis = read_matcher.repeat_profile(aligned, error_rate)
        self.R = sum(1 for k in emis if k.startswith("M"))
        self.C = copies
        shape = {"L_suffix": self.L_left, "L_prefix": self.L_right}
        # edges -> parameter slots
        keys, slot_of, chains = [], {}, []
        dst = np.repeat(np.arange(self.n_states), np.diff(self.in_off))
        pidx = np.empty(len(self.in_src), dtype=np.int32)
        for e, (s, d) in enumerate(zip(self.in_src, dst)):
            chain, key = _classify(self.names[s], self.names[d], shape)
            full = (chain,) + key
            if full not in slot_of:
                slot_of[full] = len(keys)
                keys.append(key)
                chains.append(chain)
            pidx[e] = slot_of[full]
        self.keys, self.edge_slot = keys, pidx
        self.chains = np.asarray(chains, dtype=np.int32)
        # emitting states -> emission rows: 0..3 flank match on A,C,G,T; 4 uniform insert;
        # 5.. repeat profile rows (M1..MR then I0..IR)
        rows = np.empty(self.silent_start, dtype=np.int32)
        self.suffix_match_pos = np.empty(self.L_left, dtype=np.int32)
        self.prefix_match_pos = np.empty(self.L_right, dtype=np.int32)
        for i, nm in enumerate(self.names[:self.silent_start]):
            kind, idx, part = _NAME.match(nm).groups()
            idx = int(idx)
            if part in ("suffix", "prefix"):
                if kind == "I":
                    rows[i] = 4
                else:
                    rows[i] = 0
                    (self.suffix_match_pos if part == "suffix" else self.prefix_match_pos)[idx - 1] = i
            else:
                rows[i] = 5 + (idx - 1 if kind == "M" else self.R + idx)
        self.emis_row = rows
        self._states = None
        # self-check: the fast tables of the template locus must equal the literal build
        fast = self.compile(left, right, aligned, error_rate)
        if not (np.array_equal(fast["in_logp"].view(np.int64), b["in_logp"].view(np.int64)) and
                np.array_equal(fast["emis"].view(np.int64), b["emis"].view(np.int64))):
            raise AssertionError("fast compiler disagrees with the literal build for shape %r"
                                 % ((self.L_left, self.L_right, self.R, self.C),))

    # -- per-locus evaluation ------------------------------------------------------------------
    def _parameters(self, trans, error_rate):
        p_ins = error_rate * 2 / 5
        p_del = error_rate * 1 / 5
        p_adv = 1 - p_ins - p_del
        flank = {"ie": p_ins, "de": p_del, "adv": p_adv, "adv_m01": p_adv - 0.01,
                 "one_minus_ie": 1 - p_ins}
        out = np.empty(len(self.keys), dtype=np.float64)
        for i, key in enumerate(self.keys):
            kind = key[0]
            if kind == "profile":
                out[i] = trans[key[1]][key[2]]
            elif kind == "flank":
                if key[2] == "adv_over_L":
                    out[i] = p_adv / (self.L_left if key[1] == "suffix" else self.L_right)
                else:
                    out[i] = flank[key[2]]
            elif kind == "const":
                out[i] = key[1]
            elif kind == "first_copy":
                out[i] = 0.7 / self.R
            elif kind == "to_end":
                out[i] = 0.7 / (self.C * self.R)
            else:
                out[i] = 1.0
        return out

    @staticmethod
    def _log(values):
        return np.fromiter((math.log(v) if v > 0 else NEGINF for v in values.tolist()),
                           dtype=np.float64, count=len(values))

    def compile(self, left, right, aligned_segments, error_rate, profile=None):
        if (len(left), len(right)) != (self.L_left, self.L_right):
            raise ValueError("locus does not have this template's shape")
        trans, emis = profile or read_matcher.repeat_profile(aligned_segments, error_rate)
        p = self._parameters(trans, error_rate)
        if not (p > 0).all():
            raise ValueError("zero-probability transition: structure may differ, use the literal builder")
        ch = self.chains
        w = np.zeros(len(p), dtype=np.float64)
        n_match = self.C * self.R
        total = 1 + 0.7 / n_match
        # every parameter is logged once when its edge is first added (hmm.pyx:433)
        l1 = self._log(p)
        w[ch == CH_LOG] = l1[ch == CH_LOG]
        # one dense round trip: exp (hmm.pyx:514) then log again in from_matrix
        sel = (ch == CH_F1) | (ch == CH_F2) | (ch == CH_F2_MATCH)
        e1 = np.exp(l1[sel])
        l2 = self._log(e1)
        tmp = np.zeros(len(p)); tmp[sel] = l2
        w[ch == CH_F1] = tmp[ch == CH_F1]
        # repeat-unit parameters go through a second round trip (variable-copy wrapper + read matcher)
        sel2 = (ch == CH_F2) | (ch == CH_F2_MATCH)
        e2 = np.exp(tmp[sel2])
        is_match = (ch[sel2] == CH_F2_MATCH)
        e2 = np.where(is_match, e2 / total, e2)          # hmm_utils.py:578-583
        tmp2 = np.zeros(len(p)); tmp2[sel2] = self._log(e2)
        w[sel2] = tmp2[sel2]
        to_end = 0.7 / n_match
        w[ch == CH_TO_END] = math.log(to_end / total)    # hmm_utils.py:584
        in_logp = w[self.edge_slot]

        lg = math.log
        table = np.empty((5 + 2 * self.R + 1, 4), dtype=np.float64)
        table[:4] = lg(0.01)
        table[np.arange(4), np.arange(4)] = lg(0.97)
        table[4] = lg(0.25)
        for i in range(1, self.R + 1):
            row = emis["M%d" % i]
            table[4 + i] = [lg(row[c]) if row[c] > 0 else NEGINF for c in "ACGT"]
        for i in range(self.R + 1):
            row = emis["I%d" % i]
            table[5 + self.R + i] = [lg(row[c]) if row[c] > 0 else NEGINF for c in "ACGT"]
        rows = self.emis_row.copy()
        rows[self.suffix_match_pos] = _codes(left)
        rows[self.prefix_match_pos] = _codes(right)
        return {"n_states": self.n_states, "silent_start": self.silent_start,
                "start_index": self.start_index, "end_index": self.end_index, "finite": self.finite,
                "in_off": self.in_off, "in_src": self.in_src, "in_logp": in_logp,
                "emis": table[rows], "alphabet": "ACGT"}

    def states(self):
        """State objects (names only matter downstream); shared by all loci of the shape."""
        if self._states is None:
            dummy = DiscreteDistribution({"A": 0.25, "C": 0.25, "G": 0.25, "T": 0.25})
            S = self.silent_start
            self._states = [State(dummy if i < S else None, name=nm) for i, nm in enumerate(self.names)]
        return self._states


_LUT = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    _LUT[ord(_c)] = _i


def _codes(seq):
    out = _LUT[np.frombuffer(seq.encode("ascii"), dtype=np.uint8)]
    if (out == 255).any():
        raise ValueError("flank contains a non-ACGT symbol")
    return out


class CompiledHMM(HiddenMarkovModel):
    """A baked read-matcher model produced by the fast compiler: same decoding surface."""

    def __init__(self, template, baked):
        HiddenMarkovModel.__init__(self, name="Read Matcher")
        self._template = template
        self._baked = baked
        self.d = 1
        self.discrete = 1
        self.n_states, self.n_edges = baked["n_states"], len(baked["in_src"])
        self.silent_start = baked["silent_start"]
        self.start_index, self.end_index, self.finite = baked["start_index"], baked["end_index"], baked["finite"]
        self.keymap = [{c: i for i, c in enumerate("ACGT")}]
        self.states = template.states()
        self.start, self.end = self.states[self.start_index], self.states[self.end_index]

    def bake(self, verbose=False, merge="All"):
        raise ValueError("a compiled model is already baked")


_templates = {}


def get_read_matcher_model(left, right, segments, copies, error_rate=read_matcher.DEFAULT_MAX_ERROR_RATE):
    """Drop-in for ``read_matcher.get_read_matcher_model`` (no ``vpaths``): same tables, O(E)."""
    aligned = read_matcher.align_repeat_segments(segments)
    width = len(aligned[0])
    trans_emis = read_matcher.repeat_profile(aligned, error_rate)
    R = sum(1 for k in trans_emis[1] if k.startswith("M"))
    key = (len(left), len(right), R, width, copies)
    tpl = _templates.get(key)
    if tpl is None:
        tpl = _templates[key] = ShapeTemplate(left, right, segments, copies, error_rate)
    try:
        baked = tpl.compile(left, right, aligned, error_rate, profile=trans_emis)
    except ValueError:
        return read_matcher.get_read_matcher_model(left, right, segments, copies, error_rate=error_rate)
    return CompiledHMM(tpl, baked)


def build_vntr_matcher_hmm(left_flank, right_flank, repeat_segments, copies, flank_size=100,
                           error_rate=read_matcher.DEFAULT_MAX_ERROR_RATE):
    return get_read_matcher_model(left_flank[-flank_size:], right_flank[:flank_size],
                                  repeat_segments, copies, error_rate=error_rate)
