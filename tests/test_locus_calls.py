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
