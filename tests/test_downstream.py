"""What adVNTR does with the Viterbi paths: genotype statistics, frameshift test, the --update
re-estimation and the PacBio call site -- against values produced by the reference's own
``vntr_finder.py`` / ``hmm_utils.py`` on the compiled reference engine
(tests/golden/make_golden_downstream.py -> downstream.json, update_model.npz).

CPU: the host logic.  GPU: whole loci through ``LocusDecoder`` on the device, ending in the same
genotype tuples ("identical genotype calls on every test locus")."""
import json
import os

import numpy as np
import pytest

from advntr_b200 import genotype, locus_batch, path_utils, read_matcher
from advntr_b200 import pomegranate as pom
from conftest import GOLDEN, same_bits
from test_builder_parity import _exp_log_reproduce


@pytest.fixture(scope="module")
def down():
    return json.load(open(os.path.join(GOLDEN, "downstream.json")))


def _tuple(x):
    return None if x is None else tuple(x)


# ------------------------------------------------------------------------------------ CPU
def test_genotype_statistics_match_the_reference(down):
    assert len(down["genotype"]) > 600
    for case in down["genotype"]:
        got, prob = genotype.find_genotype_based_on_observed_repeats(case["observed"], case["haploid"])
        assert _tuple(got) == _tuple(case["genotype"]), case
        assert prob == case["max_prob"] or (np.isnan(prob) and np.isnan(case["max_prob"])), case


def test_reference_genotyping_kats():
    """The count lists of the reference's tests/test_genotyping.py (its asserts predate the
    (genotype, max_prob) return value; the genotypes are what they expect)."""
    f = genotype.find_genotype_based_on_observed_repeats
    assert f([3, 3, 3, 3, 3])[0] == (3, 3)
    assert f([2, 3, 3, 3, 3], is_haploid=True)[0] == (3, 3)
    assert tuple(sorted(f([2, 2, 3, 3, 3])[0])) == (2, 3)
    assert tuple(sorted(f([4, 5, 5, 5, 7, 8, 8, 8, 9])[0])) == (5, 8)


def test_frameshift_binomial_test_matches_the_reference(down):
    for case in down["frameshift"]:
        assert genotype.identify_frameshift(case["coverage"], case["observed"], case["expected"]) == case["result"], case


def test_multiple_alignment_of_viterbi_paths_reference_fixtures(down):
    fx = json.load(open(os.path.join(GOLDEN, "ref_tests_hmm_utils.json")))
    repeats, runs = path_utils.extract_repeating_segments_from_read(fx["sequence"], fx["visited_states"].split(","))
    assert path_utils.get_multiple_alignment_of_viterbi_paths(repeats, runs) == down["msa"]["real_data_alignment"]
    two = down["msa"]["two_sequences"]
    assert path_utils.get_multiple_alignment_of_viterbi_paths(two["repeats"], two["states"]) == two["alignment"]


class _S(object):
    def __init__(self, name):
        self.name = name


def _oracle_decoder(model):
    import oracle
    om = oracle.OracleModel(model.baked)
    names = [s.name for s in model.states]

    def decode(seqs):
        lp, paths = om.viterbi([oracle.encode(s) for s in seqs])
        return lp, [[(int(k), _S(names[k])) for k in p] for p in paths]
    return decode


def test_update_model_tables_bit_identical(down):
    """--update: the model rebuilt from the Viterbi paths of the selected reads has the same baked
    tables as the reference's get_read_matcher_model(..., vpaths)."""
    z = np.load(os.path.join(GOLDEN, "update_model.npz"))
    inp = json.loads(str(z["inputs"]))
    left, right, segs = inp["left"], inp["right"], inp["segments"]
    copies = inp["copies"]
    base = read_matcher.build_vntr_matcher_hmm(left, right, segs, copies, flank_size=150)
    decode = _oracle_decoder(base)
    sel = inp["selected_sequences"]
    _, vps = decode(sel + [s.upper() for s in segs])
    vpaths = list(zip(sel + [s.upper() for s in segs], vps))
    assert path_utils.get_multiple_alignment_of_repeats_from_reads(vpaths) == inp["alignment"]
    new = read_matcher.get_read_matcher_model(left[-150:], right[:150], None, copies, vpaths)
    b = new.baked
    assert [s.name for s in new.states] == str(z["names"]).split("\n")
    assert list(z["scalars"]) == [b["n_states"], b["silent_start"], b["start_index"], b["end_index"], b["finite"]]
    assert np.array_equal(b["in_off"], z["in_off"]) and np.array_equal(b["in_src"], z["in_src"])
    assert same_bits(b["in_logp"], z["in_logp"]) and same_bits(b["emis"], z["emis"])
    # the native compiler, given the alignment the paths mark out, builds the same model (what
    # LocusDecoder.updated_model does)
    from advntr_b200 import fast_compile
    nat = fast_compile.get_read_matcher_model(left[-150:], right[:150], inp["alignment"], copies)
    assert [s.name for s in nat.states] == str(z["names"]).split("\n")
    assert np.array_equal(nat.baked["in_src"], z["in_src"])
    assert same_bits(nat.baked["in_logp"], z["in_logp"]) and same_bits(nat.baked["emis"], z["emis"])
    # and the reads decoded on it give the reference's paths (oracle restatement here, device below)
    lp, vps = _oracle_decoder(new)(inp["reads"])
    assert same_bits(lp, z["logp"])
    off = z["path_off"]
    for i, vp in enumerate(vps):
        assert [k for k, _ in vp] == list(z["paths"][off[i]:off[i + 1]])


def test_illumina_loci_host_logic_on_oracle_paths(down):
    for case in down["loci"]:
        left, right, segs = case["left"], case["right"], case["segments"]
        copies = read_matcher.copies_for_read_length(150, len(segs[0]))
        model = read_matcher.build_vntr_matcher_hmm(left, right, segs, copies, flank_size=150)
        decode = _oracle_decoder(model)
        selected = []
        lp, vps = decode(case["mapped"])
        for i, r in enumerate(case["mapped"]):
            if path_utils.recruit_read(lp[i], vps[i], None, r, left, right):
                selected.append(locus_batch.SelectedRead(r, float(lp[i]), vps[i]))
        both = [s for r in case["unmapped"] for s in (r, locus_batch.reverse_complement(r))]
        lp, vps = decode(both)
        for j in range(len(case["unmapped"])):
            f, rv = 2 * j, 2 * j + 1
            k = rv if lp[f] < lp[rv] else f
            if path_utils.recruit_read(lp[k], vps[k], None, both[k], left, right) and \
                    path_utils.get_number_of_repeat_bp_matches_in_vpath(vps[k]) > 2:
                selected.append(locus_batch.SelectedRead(both[k], float(lp[k]), vps[k], False))
        assert [s.sequence for s in selected] == case["selected_sequences"]
        _check_locus_result(case, selected, left, right)


def _check_locus_result(case, selected, left, right):
    for mode, acc in (("plain", False), ("accuracy_filter", True)):
        want = case["result"][mode]
        covered, flanking = [], []
        for s in selected:
            n = path_utils.get_number_of_repeats_in_vpath(s.vpath)
            if path_utils.read_flanks_repeats_with_confidence(s.vpath, s.sequence, left, right):
                covered.append(n)
            elif not acc:
                flanking.append(n)
        assert covered == want["covered_repeats"] and sorted(flanking) == want["flanking_repeats"]
        cn, prob = genotype.genotype_from_illumina_counts(covered, flanking, acc)
        assert _tuple(cn) == _tuple(want["copy_numbers"]) and prob == want["max_prob"]


def test_reference_segmentation_model_tables(down):
    """build_reference_repeat_finder_hmm is baked with the default merge='All' (hmm_utils.py:674):
    same tables as the reference's, and the oracle decodes the region into the same repeat units."""
    z = np.load(os.path.join(GOLDEN, "segmentation.npz"))
    inp = json.loads(str(z["inputs"]))
    model = read_matcher.build_reference_repeat_finder_hmm([inp["pattern"]], copies=inp["copies"])
    b = model.baked
    assert [s.name for s in model.states] == str(z["names"]).split("\n")
    assert np.array_equal(b["in_off"], z["in_off"]) and np.array_equal(b["in_src"], z["in_src"])
    assert same_bits(b["in_logp"], z["in_logp"]) and same_bits(b["emis"], z["emis"])
    lp, vps = _oracle_decoder(model)([inp["region"]])
    assert same_bits(lp, z["logp"]) and [k for k, _ in vps[0]] == list(z["paths"])
    visited = [st.name for _, st in vps[0][1:-1]]
    assert path_utils.get_repeat_segments_from_visited_states_and_region(visited, inp["region"]) == inp["segments"]


def _fuzz_model(pom, seed, merge):
    """Small random HMM: silent / emitting states, forward edges (no silent cycles), self loops,
    probability-1 edges, rows that do not sum to 1, dangling states."""
    import random
    rng = random.Random(seed)
    hmm = pom.HiddenMarkovModel(name="fuzz")
    states = []
    for i in range(rng.randint(3, 9)):
        if rng.random() < 0.5 and i:
            states.append(pom.State(None, name="s%d" % i))
        else:
            p = [rng.random() for _ in range(4)]
            t = sum(p)
            states.append(pom.State(pom.DiscreteDistribution(dict(zip("ACGT", [x / t for x in p]))), name="e%d" % i))
    hmm.add_states(states)
    nodes = [hmm.start] + states + [hmm.end]
    for i, a in enumerate(nodes[:-1]):
        outs = [b for b in nodes[i + 1:] if rng.random() < 0.45]
        if not a.is_silent() and rng.random() < 0.4:
            outs.append(a)
        if not outs and rng.random() < 0.7:
            outs = [nodes[rng.randrange(i + 1, len(nodes))]]
        if not outs:
            continue
        if rng.random() < 0.35:
            outs, probs = outs[:1], [1.0]
        else:
            probs = [rng.random() for _ in outs]
            if rng.random() < 0.6:
                t = sum(probs)
                probs = [x / t for x in probs]
        for b, p in zip(outs, probs):
            if not (a is hmm.start and b is hmm.end):
                hmm.add_transition(a, b, p)
    hmm.bake(merge=merge)
    return hmm


@pytest.mark.parametrize("merge", ["All", "Partial", None])
def test_bake_merge_semantics_match_compiled_reference(merge):
    """Orphan removal (with the reference's never-reset degree counters), row normalisation and the
    merging of silent states with a probability-1 out-edge (hmm.pyx:720-838), model by model against
    bake() of the compiled reference engine."""
    import oracle
    refenv = pytest.importorskip("refenv")
    if not refenv.have_reference_engine():
        pytest.skip("oracle/_ref not built")
    from advntr_b200 import pomegranate as mine
    ref = refenv.reference_pomegranate()
    checked = 0
    for seed in range(160):
        try:
            want = oracle.baked_from_reference_model(_fuzz_model(ref, seed, merge))
        except UnboundLocalError:       # the reference's bake() trips over a model left without emitting states
            continue
        got = _fuzz_model(mine, seed, merge)
        b = got.baked
        assert [s.name for s in got.states] == want["names"], seed
        assert np.array_equal(b["in_off"], want["in_off"]) and np.array_equal(b["in_src"], want["in_src"]), seed
        assert same_bits(b["in_logp"], want["in_logp"]), seed
        assert (b["start_index"], b["end_index"], b["silent_start"]) == \
            (want["start_index"], want["end_index"], want["silent_start"]), seed
        checked += 1
    assert checked >= 120


# ------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_illumina_loci_genotypes_on_device(down):
    for case in down["loci"]:
        dec = locus_batch.LocusDecoder(case["left"], case["right"], case["segments"], read_length=150, flank_size=150)
        selected = dec.select_reads(case["mapped"], case["unmapped"])
        assert [s.sequence for s in selected] == case["selected_sequences"]
        for mode, acc in (("plain", False), ("accuracy_filter", True)):
            want = case["result"][mode]
            res = dec.genotype(selected, accuracy_filter=acc)
            assert _tuple(res["copy_numbers"]) == _tuple(want["copy_numbers"])
            assert res["maximum_likelihood"] == want["max_prob"]
            assert (res["recruited_reads_count"], res["spanning_reads_count"], res["flanking_reads_count"]) == \
                (want["recruited"], want["spanning"], want["flanking"])
        assert sorted(dec.genotype(selected)["copy_numbers"]) == sorted(case["alleles"])


@pytest.mark.gpu
def test_update_model_on_device(down):
    z = np.load(os.path.join(GOLDEN, "update_model.npz"))
    inp = json.loads(str(z["inputs"]))
    case = down["loci"][0]
    dec = locus_batch.LocusDecoder(case["left"], case["right"], case["segments"], read_length=150, flank_size=150)
    selected = dec.select_reads(case["mapped"], case["unmapped"])
    new = dec.updated_model(selected)
    assert same_bits(new.baked["in_logp"], z["in_logp"]) and same_bits(new.baked["emis"], z["emis"])
    res = new.viterbi_batch(inp["reads"])
    assert same_bits(res.logp, z["logp"])
    off = z["path_off"]
    for i in range(len(inp["reads"])):
        assert np.array_equal(res.path(i), z["paths"][off[i]:off[i + 1]])


@pytest.mark.gpu
def test_pacbio_dominant_copy_numbers_on_device(down):
    case = down["pacbio"]
    for mode, acc in (("plain", False), ("accuracy_filter", True)):
        cn, prob, observed = locus_batch.dominant_copy_numbers_from_spanning_reads(
            case["left"], case["right"], case["segments"], case["reads"], error_rate=case["error_rate"],
            accuracy_filter=acc)
        assert observed == case["observed_repeats"]
        assert _tuple(cn) == _tuple(case["result"][mode]["copy_numbers"]) and prob == case["result"][mode]["max_prob"]
    assert sorted(cn) == sorted(case["alleles"])


@pytest.mark.gpu
def test_find_repeat_segments_on_device():
    """reference_vntr.py:80-87 as a drop-in: the repeat finder model (generic kernel) on the device."""
    z = np.load(os.path.join(GOLDEN, "segmentation.npz"))
    inp = json.loads(str(z["inputs"]))
    model = read_matcher.build_reference_repeat_finder_hmm([inp["pattern"]], copies=inp["copies"])
    logp, path = model.viterbi(inp["region"])
    assert same_bits([logp], z["logp"]) and [k for k, _ in path] == list(z["paths"])
    assert read_matcher.find_repeat_segments(inp["pattern"], inp["copies"], inp["region"]) == inp["segments"]


def _stored_model():
    z = np.load(os.path.join(GOLDEN, "stored_model.npz"))
    import gzip
    return z, gzip.decompress(z["json_gz"].tobytes()).decode(), json.loads(str(z["inputs"]))


def test_stored_model_json_round_trip_equals_the_reference():
    """vntr_finder.py:116-137: ``to_json`` of a freshly built read matcher is the reference's text byte
    for byte, and ``from_json`` of it (baked with the default merge='All', hmm.pyx:3143) gives the
    reference's reloaded tables -- fewer silent states than the stored model -- and the same text again."""
    import hashlib
    z, text, inp = _stored_model()
    built = read_matcher.get_read_matcher_model(inp["left"], inp["right"], inp["segments"], inp["copies"])
    assert len(built.states) == int(z["n_states_stored"])
    assert built.to_json() == text
    again = pom.HiddenMarkovModel.from_json(text)
    b = again.baked
    assert [s.name for s in again.states] == str(z["names"]).split("\n")
    assert list(z["scalars"]) == [b["n_states"], b["silent_start"], b["start_index"], b["end_index"], b["finite"]]
    assert b["n_states"] < len(built.states)
    assert np.array_equal(b["in_off"], z["in_off"]) and np.array_equal(b["in_src"], z["in_src"])
    assert same_bits(b["in_logp"], z["in_logp"]) and same_bits(b["emis"], z["emis"])
    assert hashlib.sha256(again.to_json().encode()).hexdigest() == str(z["json_again_sha"])
    lp, vps = _oracle_decoder(again)(inp["reads"])
    assert same_bits(lp, z["logp"])
    off = z["path_off"]
    for i, vp in enumerate(vps):
        assert [k for k, _ in vp] == list(z["paths"][off[i]:off[i + 1]])


def test_stored_model_from_a_file_and_bad_input(tmp_path):
    z, text, inp = _stored_model()
    path = tmp_path / "7_150.json"
    path.write_text(text)
    again = pom.HiddenMarkovModel("unused").from_json(str(path))     # vntr_finder.py:127-128 calls it on an instance
    assert len(again.states) == int(z["scalars"][0])
    with pytest.raises(IOError):
        pom.HiddenMarkovModel.from_json("neither json nor a file")


@pytest.mark.gpu
def test_stored_model_decodes_on_device():
    z, text, inp = _stored_model()
    again = pom.HiddenMarkovModel.from_json(text)
    res = again.viterbi_batch(inp["reads"])
    assert same_bits(res.logp, z["logp"])
    off = z["path_off"]
    for i in range(len(inp["reads"])):
        assert np.array_equal(res.path(i), z["paths"][off[i]:off[i + 1]])
    assert again.viterbi(inp["reads"][0])[0] == z["logp"][0]


def test_locus_decoder_stores_and_reloads_its_model(tmp_path):
    """get_vntr_matcher_hmm with USE_TRAINED_HMMS (vntr_finder.py:116-137): the first run stores
    '<id>_<read_length>.json', later runs load it."""
    z, text, inp = _stored_model()
    args = (inp["left"], inp["right"], inp["segments"])
    first = locus_batch.LocusDecoder(*args, read_length=40, flank_size=60, locus_id=7, trained_hmms_dir=str(tmp_path))
    stored = tmp_path / "7_40.json"
    assert stored.is_file()
    fresh = read_matcher.build_vntr_matcher_hmm(*args, read_matcher.copies_for_read_length(40, 14), flank_size=60)
    assert stored.read_text() == fresh.to_json()
    assert [s.name for s in first.model.states] == [s.name for s in fresh.states]
    second = locus_batch.LocusDecoder(*args, read_length=40, flank_size=60, locus_id=7, trained_hmms_dir=str(tmp_path))
    reloaded = pom.HiddenMarkovModel.from_json(stored.read_text())
    assert [s.name for s in second.model.states] == [s.name for s in reloaded.states]
    assert len(second.model.states) < len(first.model.states)
