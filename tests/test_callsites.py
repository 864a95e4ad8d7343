"""The batched call sites (advntr_b200/locus_batch.py, path_utils.py) reproduce the decisions of
the reference's own VNTRFinder on config-4 style inputs (frameshift reads, both-strand unmapped
reads).  Golden values: tests/golden/callsite_frameshift.json, produced by running
/root/reference/advntr/vntr_finder.py on the compiled reference engine (make_golden.py).

CPU test: host logic on oracle paths.  GPU test: the same through LocusDecoder on the device."""
import json
import os

import numpy as np
import pytest

import oracle
from advntr_b200 import locus_batch, path_utils, read_matcher
from conftest import GOLDEN


@pytest.fixture(scope="module")
def case():
    return json.load(open(os.path.join(GOLDEN, "callsite_frameshift.json")))


class _S(object):
    def __init__(self, name):
        self.name = name


def _check_downstream(case, selected, pattern_len, vntr_len, left, right):
    assert [s.sequence for s in selected] == case["selected_sequences"]
    assert [s.logp for s in selected] == case["selected_logp"]
    covered, flanking = [], []
    for s in selected:
        n = path_utils.get_number_of_repeats_in_vpath(s.vpath)
        (covered if path_utils.read_flanks_repeats_with_confidence(s.vpath, s.sequence, left, right)
         else flanking).append(n)
    assert covered == case["covered_repeats"] and sorted(flanking) == case["flanking_repeats"]
    mutations, repeat_bp = path_utils.frameshift_mutations(selected, pattern_len)
    ranked = sorted(mutations.items(), key=lambda x: x[1])
    assert ranked[-1] == (case["frameshift_candidate"], case["frameshift_count"])
    assert float(repeat_bp) / vntr_len / 2 == case["avg_bp_coverage"]


def test_host_logic_on_oracle_paths(case):
    left, right, segs = case["left"], case["right"], case["segments"]
    model = read_matcher.build_vntr_matcher_hmm(left, right, segs, read_matcher.copies_for_read_length(150, 30),
                                                flank_size=150)
    names = [s.name for s in model.states]
    om = oracle.OracleModel(model.baked)

    def decode(seqs):
        lp, paths = om.viterbi([oracle.encode(s) for s in seqs])
        return lp, [[(int(k), _S(names[k])) for k in p] for p in paths]

    selected = []
    lp, vps = decode(case["mapped"])
    got = [bool(path_utils.recruit_read(lp[i], vps[i], None, r, left, right)) for i, r in enumerate(case["mapped"])]
    assert got == case["mapped_recruited"]
    for i, r in enumerate(case["mapped"]):
        if got[i]:
            selected.append(locus_batch.SelectedRead(r, float(lp[i]), vps[i]))
    assert len(selected) == case["n_mapped_selected"]
    both = [s for r in case["unmapped"] for s in (r, locus_batch.reverse_complement(r))]
    lp, vps = decode(both)
    for j, r in enumerate(case["unmapped"]):
        f, rv = 2 * j, 2 * j + 1
        k = rv if lp[f] < lp[rv] else f
        if path_utils.recruit_read(lp[k], vps[k], None, both[k], left, right) and \
                path_utils.get_number_of_repeat_bp_matches_in_vpath(vps[k]) > 2:
            selected.append(locus_batch.SelectedRead(both[k], float(lp[k]), vps[k], False))
    _check_downstream(case, selected, 30, 30 * len(segs), left, right)


@pytest.mark.gpu
def test_locus_decoder_on_device(case):
    left, right, segs = case["left"], case["right"], case["segments"]
    dec = locus_batch.LocusDecoder(left, right, segs, read_length=150, flank_size=150)
    selected = dec.select_reads(case["mapped"], case["unmapped"])
    assert sum(1 for s in selected if s.is_mapped) == case["n_mapped_selected"]
    _check_downstream(case, selected, 30, 30 * len(segs), left, right)
    covered, flanking = dec.observed_repeats(selected)
    assert covered == case["covered_repeats"] and flanking == case["flanking_repeats"]
    (cand, count), _ = dec.frameshift_candidate(selected)
    assert (cand, count) == (case["frameshift_candidate"], case["frameshift_count"])


@pytest.mark.gpu
def test_decode_many_loci(case):
    from advntr_b200 import synth
    decs, reads = [], []
    for lid in (3, 8, 21):
        loc = synth.config2_locus(lid)
        decs.append(locus_batch.LocusDecoder(loc.left, loc.right, loc.segments, 150, locus_id=lid))
        mapped, unmapped = synth.config2_reads(loc, coverage=5, decoys=4)
        reads.append(mapped + unmapped)
    res = locus_batch.decode_many(decs, reads)
    k = 0
    for d, rs in zip(decs, reads):
        lp, paths = oracle.OracleModel(d.model.baked).viterbi([oracle.encode(r) for r in rs])
        for i in range(len(rs)):
            assert res.logp[k] == lp[i] and np.array_equal(res.path(k), paths[i])
            k += 1
