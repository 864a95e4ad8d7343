"""The oracle (oracle/hmm_oracle.c) is pinned against the reference: golden vectors made from the
compiled unmodified engine, the SURVEY 8c sanity constants, and -- where oracle/_ref is
present -- the engine itself on fresh random reads."""
import random

import numpy as np
import pytest

import oracle
from conftest import assert_paths_equal, same_bits


def test_oracle_matches_golden_viterbi(golden):
    om = oracle.OracleModel(golden.baked)
    logp, paths = om.viterbi(golden.codes())
    assert same_bits(logp, golden.logp)
    assert_paths_equal(paths, [golden.path(i) for i in range(len(golden.reads))], golden.name)


def test_oracle_matches_golden_forward(golden):
    om = oracle.OracleModel(golden.baked)
    fwd = om.log_probability(golden.codes())
    # same libm here as when the vectors were made: bit-equal; elsewhere within 1e-12 relative
    assert np.allclose(fwd, golden.forward, rtol=1e-12, atol=0)


def test_survey_sanity_constants(golden_config1):
    g = golden_config1
    b = g.baked
    assert (b["n_states"], b["silent_start"], b["start_index"], b["end_index"]) == (1176, 768, 768, 1175)
    assert len(b["in_src"]) == 3877
    assert g.logp[0] == -22.324033385541433 and len(g.path(0)) == 171 and g.ru_count[0] == 3
    assert g.forward[0] == -22.21091055930884
    assert g.logp[1] == -959.9433067568683          # viterbi('')
    assert g.logp[2] == -9.347753111308677          # viterbi('A')
    names = [g.names[i] for i in g.path(0)]
    assert names[:4] == ["Read Matcher-start", "Suffix Matcher HMM Model-start", "suffix_start_suffix", "M71_suffix"]
    assert names[-4:] == ["M30_prefix", "prefix_end_prefix", "Prefix Matcher HMM Model-end", "Read Matcher-end"]


def test_oracle_decodes_random_reads_to_well_formed_paths(golden_config1):
    """Smoke property of the oracle alone (the pin against the compiled reference engine is the next
    test and the golden vectors above): every random read gets a finite score and a path from the start
    state to the end state that emits exactly the read."""
    b = golden_config1.baked
    om = oracle.OracleModel(b)
    rng = random.Random(5)
    reads = ["".join(rng.choice("ACGT") for _ in range(rng.randint(0, 160))) for _ in range(8)]
    logp, paths = om.viterbi([oracle.encode(r) for r in reads])
    assert np.all(np.isfinite(logp))
    S = b["silent_start"]
    for r, p in zip(reads, paths):
        assert p[0] == b["start_index"] and p[-1] == b["end_index"]
        assert int((np.asarray(p) < S).sum()) == len(r)


def test_oracle_vs_reference_engine_fresh_models():
    """Build fresh models with the reference's own builders (needs /root/reference) and compare
    the oracle with the engine on random reads, both strands."""
    refenv = pytest.importorskip("refenv")
    if not (refenv.have_reference_engine() and refenv.have_reference_sources()):
        pytest.skip("reference engine / sources not present (GPU box)")
    pom = refenv.reference_pomegranate()
    hu = refenv.reference_hmm_utils(pom, "ref")
    rng = random.Random(99)
    rnd = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
    ru, left, right = rnd(9), rnd(25), rnd(30)
    model = hu.get_read_matcher_model(left, right, [ru] * 3, copies=4)
    b = oracle.baked_from_reference_model(model)
    om = oracle.OracleModel(b)
    locus = left + ru * 6 + right
    reads = [locus[s:s + 40] for s in range(0, len(locus) - 40, 7)] + [rnd(30) for _ in range(10)] + ["", "G"]
    logp, paths = om.viterbi([oracle.encode(r) for r in reads])
    fwd = om.log_probability([oracle.encode(r) for r in reads])
    for i, r in enumerate(reads):
        lp, vp = model.viterbi(r)
        assert lp == logp[i]
        assert [k for k, _ in vp] == list(paths[i])
        assert model.log_probability(r) == fwd[i]
