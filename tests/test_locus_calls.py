"""The native step after the decode (``advhmm_genotypes_from_summaries`` / ``advhmm_genotypes_from_counts``,
csrc/locus_calls.hpp; host code, no GPU): recruitment, strand choice, spanning test and genotype call of
many loci from the per-read device results.  Checked against

* the reference's own outputs: the 614 count lists of tests/golden/downstream.json (genotype AND the
  ``max_prob`` float of ``find_genotype_based_on_observed_repeats``, vntr_finder.py:486-532) and the count
  lists of the reference's tests/test_genotyping.py;
* ``pipeline.genotypes_from_summaries`` (the literal numpy / Python form, itself held against the
  reference's VNTRFinder through LocusDecoder in test_pipeline.py) on seeded random per-read results:
  every field of every locus, ``maximum_likelihood`` as the same float."""
import json
import os

import numpy as np
import pytest

from advntr_b200 import engine, genotype, pipeline
from conftest import GOLDEN


def _same_float(a, b):
    return a == b or (np.isnan(a) and np.isnan(b))


def test_count_lists_of_the_reference():
    cases = json.load(open(os.path.join(GOLDEN, "downstream.json")))["genotype"]
    assert len(cases) > 600
    for haploid in (False, True):
        sub = [c for c in cases if bool(c["haploid"]) == haploid]
        calls = engine.genotypes_from_counts([c["observed"] for c in sub], is_haploid=haploid)
        for c, got in zip(sub, calls):
            want = None if c["genotype"] is None else tuple(c["genotype"])
            have = (int(got["c1"]), int(got["c2"])) if got["has_call"] else None
            assert have == want, c
            assert _same_float(float(got["max_prob"]), c["max_prob"]), c


def test_reference_genotyping_kats():
    """tests/test_genotyping.py of the reference."""
    f = lambda obs, **kw: engine.genotypes_from_counts([obs], **kw)[0]
    assert (f([3, 3, 3, 3, 3])["c1"], f([3, 3, 3, 3, 3])["c2"]) == (3, 3)
    g = f([2, 3, 3, 3, 3], is_haploid=True)
    assert (g["c1"], g["c2"]) == (3, 3)
    g = f([2, 2, 3, 3, 3])
    assert sorted((g["c1"], g["c2"])) == [2, 3]
    g = f([4, 5, 5, 5, 7, 8, 8, 8, 9])
    assert sorted((g["c1"], g["c2"])) == [5, 8]


def test_odd_count_lists_equal_the_python_form():
    rng = np.random.default_rng(11)
    lists = [[], [0], [0, 0], [0, 0, 3], [7], [0, 5, 5], [2, 0, 2, 0, 0]]
    for _ in range(400):
        n = int(rng.integers(1, 40))
        lists.append(rng.choice(rng.integers(0, 12, size=int(rng.integers(1, 7))), size=n).tolist())
    for haploid in (False, True):
        for acc in (False, True):
            calls = engine.genotypes_from_counts(lists, accuracy_filter=acc, is_haploid=haploid)
            for obs, got in zip(lists, calls):
                with np.errstate(all="ignore"):
                    want, prob = genotype.dominant_copy_numbers(obs, acc, haploid)
                have = (int(got["c1"]), int(got["c2"])) if got["has_call"] else None
                assert have == want and _same_float(float(got["max_prob"]), float(prob)), (obs, haploid, acc)


def _random_results(rng, n_loci, recruit_rate):
    layout, goff = [], [0]
    for _ in range(n_loci):
        m, u = int(rng.integers(0, 70)), int(rng.integers(0, 25))
        if rng.random() < 0.05:
            m = u = 0                                           # a locus without reads
        layout.append((m, u))
        goff.append(goff[-1] + m + 2 * u)
    R = goff[-1]
    lens = rng.integers(100, 151, R)
    off = np.zeros(R + 1, np.int64)
    off[1:] = np.cumsum(lens)
    S = np.zeros(R, engine.SUMMARY_DTYPE)
    per_locus = np.repeat(np.arange(n_loci), np.diff(goff))
    base = rng.integers(1, 9, n_loci)[per_locus]
    S["repeats"] = np.where(rng.random(R) < 0.5, base, base + rng.integers(-1, 4, R)).clip(0)
    S["n_match"] = (lens * rng.uniform(0.86, 1.0, R)).astype(np.int32)
    S["repeat_bp"] = rng.integers(0, 40, R)
    S["left_bp"] = np.where(rng.random(R) < 0.2, 0, rng.integers(0, 60, R))
    S["right_bp"] = np.where(rng.random(R) < 0.2, 0, rng.integers(0, 60, R))
    lo = 0.97 - 0.12 * (1 - recruit_rate)
    S["left_hits"] = np.ceil(S["left_bp"] * rng.uniform(lo, 1.0, R)).astype(np.int32)
    S["right_hits"] = np.ceil(S["right_bp"] * rng.uniform(lo, 1.0, R)).astype(np.int32)
    logp = -lens * rng.uniform(0.2, 0.6 + 0.6 * (1 - recruit_rate), R)
    logp[rng.random(R) < 0.02] = -np.inf
    plen = np.where(np.isinf(logp) | (rng.random(R) < 0.02), -1, lens + 10).astype(np.int32)
    scores = [None if rng.random() < 0.3 else float(-rng.uniform(0.5, 1.0) * 140) for _ in range(n_loci)]
    return logp, S, plen, off, np.array(goff, np.int64), layout, scores


@pytest.mark.parametrize("recruit_rate", [0.3, 0.9])
@pytest.mark.parametrize("accuracy_filter,is_haploid", [(False, False), (False, True), (True, False), (True, True)])
def test_native_calls_equal_the_python_form(recruit_rate, accuracy_filter, is_haploid):
    rng = np.random.default_rng(int(recruit_rate * 10) * 4 + 2 * accuracy_filter + is_haploid)
    logp, S, plen, off, goff, layout, scores = _random_results(rng, 600, recruit_rate)
    with np.errstate(all="ignore"):
        want = pipeline.genotypes_from_summaries(logp, S, plen, np.diff(off).astype(np.float64), goff, layout, scores,
                                                 accuracy_filter, is_haploid)
    got = pipeline.native_genotypes_from_summaries(logp, S, plen, off, goff, layout, scores, accuracy_filter, is_haploid,
                                                   threads=3)
    assert sum(1 for g in got if g["copy_numbers"] is not None) > (20 if accuracy_filter else 150)
    assert accuracy_filter or sum(g["flanking_reads_count"] for g in got) > 500
    for w, g in zip(want, got):
        w = dict(w)
        prob = float(w.pop("maximum_likelihood"))
        assert _same_float(prob, g.pop("maximum_likelihood")), (w, g)
        assert w == g


def test_native_calls_refuse_inconsistent_groups():
    S = np.zeros(4, engine.SUMMARY_DTYPE)
    args = (np.zeros(4), S, np.zeros(4, np.int32), np.arange(5, dtype=np.int64) * 10)
    with pytest.raises(engine.EngineError):
        engine.genotypes_from_summaries([0, 4], [1], [1], None, *args)          # 1 + 2 x 1 != 4
    calls, cls = engine.genotypes_from_summaries([0, 4], [2], [1], None, *args, want_read_class=True)
    assert len(calls) == 1 and len(cls) == 4


# ------------------------------------------------------------------ --frameshift mode
def _summary_records(decoder, reads, logp, paths):
    """The on-device reducers' records, stated on the host with path_utils (CPU test: no device)."""
    from advntr_b200 import path_utils
    st = decoder.model.states
    S = np.zeros(len(reads), engine.SUMMARY_DTYPE)
    vpaths = []
    for i, (r, p) in enumerate(zip(reads, paths)):
        vp = [(int(x), st[x]) for x in p]
        vpaths.append(vp)
        ps = path_utils.summarize(vp)
        hits, bases = path_utils.flank_match_counts(vp, r, decoder.left_flank, decoder.right_flank)
        S[i] = (ps.repeats, ps.n_match, ps.repeat_bp, ps.left_bp, ps.right_bp, hits["suffix"], hits["prefix"], 0)
        assert (bases["suffix"], bases["prefix"]) == (ps.left_bp, ps.right_bp)
    return S, vpaths


def test_frameshift_candidates_equal_the_python_consumers():
    """advhmm_frameshift_candidates vs recruit_read + find_frameshift_from_selected_reads' path walk
    (path_utils.frameshift_mutations, LocusDecoder.frameshift_candidate) on CPU-oracle paths of loci whose
    sample carries a 1 bp indel (bench_workloads.frameshift_locus_reads)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import oracle
    import bench_workloads
    from advntr_b200 import locus_batch, path_utils
    goff, tables, plen_pat, lp_all, S_all, pl_all, po_all, pa_all, codes_all, want = [0], [], [], [], [], [], [], [], [], []
    at = 0
    for lid in (3, 4, 7, 11, 19, 26):
        loc, reads = bench_workloads.frameshift_locus_reads(lid, coverage=14)
        dec = locus_batch.LocusDecoder(loc.left, loc.right, loc.segments, read_length=150, locus_id=lid)
        codes = [oracle.encode(r) for r in reads]
        lp, paths = oracle.OracleModel(dec.model.baked).viterbi(codes)
        S, vpaths = _summary_records(dec, reads, lp, paths)
        sel = [locus_batch.SelectedRead(r, float(lp[i]), vpaths[i]) for i, r in enumerate(reads)
               if path_utils.recruit_read(lp[i], vpaths[i], None, r, dec.left_flank, dec.right_flank)]
        (label, count), repeat_bp = dec.frameshift_candidate(sel)
        want.append((label, count, repeat_bp, len(sel)))
        tables.append(path_utils.frameshift_state_tables([s.name for s in dec.model.states]))
        plen_pat.append(len(dec.pattern))
        for i, p in enumerate(paths):
            lp_all.append(lp[i]); pl_all.append(len(p)); po_all.append(at); pa_all.append(np.asarray(p, np.int32)); at += len(p)
            codes_all.append(np.asarray(codes[i], np.uint8))
        S_all.append(S)
        goff.append(goff[-1] + len(reads))
    seqs, off = engine.pack_reads(codes_all)
    got = engine.frameshift_candidates(goff, plen_pat, None, tables, lp_all, np.concatenate(S_all), pl_all, po_all,
                                       np.concatenate(pa_all), seqs, off, threads=2)
    assert sum(1 for w in want if w[0] is not None and w[1] >= 3) >= 3          # the planted indels are seen
    for w, g in zip(want, got):
        assert (path_utils.frameshift_label(g), int(g["count"]), int(g["repeat_bp"]), int(g["selected"])) == w


def test_frameshift_candidate_of_the_reference_callsite_golden():
    """tests/golden/callsite_frameshift.json: the candidate, its count and the coverage the reference's own
    VNTRFinder.find_frameshift_from_selected_reads arrives at (make_golden.py), from the native consumer on
    CPU-oracle paths of the same reads."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import oracle
    from advntr_b200 import locus_batch, path_utils, read_matcher
    case = json.load(open(os.path.join(GOLDEN, "callsite_frameshift.json")))
    left, right, segs = case["left"], case["right"], case["segments"]
    dec = locus_batch.LocusDecoder(left, right, segs, read_length=150, flank_size=150)
    om = oracle.OracleModel(dec.model.baked)
    both = [s for r in case["unmapped"] for s in (r, locus_batch.reverse_complement(r))]
    lp_u, paths_u = om.viterbi([oracle.encode(s) for s in both])
    reads = list(case["mapped"])
    st = dec.model.states
    for j in range(len(case["unmapped"])):                      # select_illumina_reads: the better strand, > 2 repeat bp
        k = 2 * j + 1 if lp_u[2 * j] < lp_u[2 * j + 1] else 2 * j
        if path_utils.get_number_of_repeat_bp_matches_in_vpath([(int(x), st[x]) for x in paths_u[k]]) > 2:
            reads.append(both[k])
    codes = [oracle.encode(r) for r in reads]
    lp, paths = om.viterbi(codes)
    S, _ = _summary_records(dec, reads, lp, paths)
    seqs, off = engine.pack_reads([np.asarray(c, np.uint8) for c in codes])
    plen = [len(p) for p in paths]
    poff = np.concatenate([[0], np.cumsum(plen)[:-1]])
    got = engine.frameshift_candidates([0, len(reads)], [len(dec.pattern)], None,
                                       [path_utils.frameshift_state_tables([s.name for s in st])], lp, S, plen, poff,
                                       np.concatenate([np.asarray(p, np.int32) for p in paths]), seqs, off)[0]
    assert (path_utils.frameshift_label(got), int(got["count"])) == (case["frameshift_candidate"], case["frameshift_count"])
    assert int(got["selected"]) == len(case["selected_sequences"])
    assert float(got["repeat_bp"]) / (30 * len(segs)) / 2 == case["avg_bp_coverage"]


def test_frameshift_candidates_refuse_foreign_paths():
    """A path that is not a path of its read on its model (state index out of range, emitted length) is an error,
    not a silent walk over someone else's memory."""
    S = np.zeros(1, engine.SUMMARY_DTYPE)
    S["n_match"], S["left_bp"], S["left_hits"], S["right_bp"], S["right_hits"] = 1, 1, 1, 1, 1
    cls = np.array([0, 1 | (3 << 3), 0], np.uint8)                 # start, a repeat-unit match state, end
    tables = [(cls, np.full(3, -1, np.int32))]
    ok = dict(group_off=[0, 1], pattern_len=[5], min_score=[-100.0], state_tables=tables, logp=[-1.0], summaries=S,
              path_len=[3], path_off=[0], paths=[0, 1, 2], seqs=[2, 0], seq_off=[0, 1])
    rec = engine.frameshift_candidates(**ok)
    assert int(rec[0]["selected"]) == 1 and int(rec[0]["kind"]) == 0
    with pytest.raises(engine.EngineError):
        engine.frameshift_candidates(**dict(ok, paths=[0, 7, 2]))
    with pytest.raises(engine.EngineError):
        engine.frameshift_candidates(**dict(ok, seq_off=[0, 2]))
