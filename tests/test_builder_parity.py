"""Host model compiler parity: the pomegranate-compatible surface + the restated builders give
the SAME baked tables (state order, in-edge order, bit-equal log weights) as the reference's
builders on the reference engine (golden vectors), and the reference's own hmm_utils.py runs
unmodified on the new surface (this container only)."""
import math
import os

import numpy as np
import pytest

from advntr_b200 import pomegranate as pom
from advntr_b200 import read_matcher
from conftest import GOLDEN


def _exp_log_reproduce():
    """The builders round-trip probabilities through numpy.exp and libm log exactly like the
    reference (hmm.pyx:514, utils.pyx:70); bit-equality with vectors made elsewhere needs both to
    reproduce on this machine."""
    z = np.load(os.path.join(GOLDEN, "fingerprint.npz"))
    e = np.exp(z["x"])
    return np.array_equal(e, z["exp"]) and np.array_equal(np.array([math.log(v) for v in e]), z["log"])


def _assert_same_tables(model, golden):
    b = model.baked
    assert [s.name for s in model.states] == golden.names
    for k in ("n_states", "silent_start", "start_index", "end_index", "finite"):
        assert b[k] == golden.baked[k], k
    assert np.array_equal(b["in_off"], golden.baked["in_off"])
    assert np.array_equal(b["in_src"], golden.baked["in_src"])
    if _exp_log_reproduce():
        assert np.array_equal(b["in_logp"], golden.baked["in_logp"])
        assert np.array_equal(b["emis"], golden.baked["emis"])
    else:   # different libm / numpy SIMD path than where the vectors were made
        assert np.allclose(b["in_logp"], golden.baked["in_logp"], rtol=1e-14, atol=1e-15)
        assert np.allclose(b["emis"], golden.baked["emis"], rtol=1e-14, atol=1e-15)


def test_restated_builders_match_golden(golden):
    i = golden.inputs
    model = read_matcher.get_read_matcher_model(i["left"], i["right"], i["segments"], i["copies"],
                                                error_rate=i["error_rate"])
    _assert_same_tables(model, golden)


def test_reference_hmm_utils_runs_unmodified_on_new_surface(golden):
    refenv = pytest.importorskip("refenv")
    if not refenv.have_reference_sources():
        pytest.skip("/root/reference not present (GPU box)")
    hu = refenv.reference_hmm_utils(pom, "mine")
    settings = refenv.reference_settings()
    i = golden.inputs
    settings.MAX_ERROR_RATE = i["error_rate"]
    try:
        model = hu.get_read_matcher_model(i["left"], i["right"], i["segments"], copies=i["copies"])
    finally:
        settings.MAX_ERROR_RATE = 0.05
    _assert_same_tables(model, golden)


def test_builder_error_behaviour():
    m = pom.HiddenMarkovModel("x")
    s = pom.State(pom.DiscreteDistribution({"A": 1.0}), name="dup")
    m.add_state(s)
    with pytest.raises(ValueError, match="already exists"):
        m.add_state(pom.State(None, name="dup"))
    with pytest.raises(ValueError, match="must bake model"):
        m.viterbi("A")
    with pytest.raises(ValueError, match="must bake model"):
        m.log_probability("A")
    m.bake(merge="All")       # default merge: implemented (tests/test_downstream.py checks its semantics)


def test_from_matrix_wires_last_state_to_end():
    """hmm.pyx:3231-3235: whatever index `ends` marks, the LAST listed state gets the edge."""
    d = pom.DiscreteDistribution({"A": 0.5, "C": 0.5})
    mat = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [0.0, 0.0, 0.0]])
    m = pom.HiddenMarkovModel.from_matrix(mat, [d, d, None], [1.0, 0, 0], [0.7, 0, 0], name="T",
                                          state_names=["a", "b", "z"], merge=None)
    b = m.baked
    names = [s.name for s in m.states]
    end = b["end_index"]
    srcs = [names[k] for k in b["in_src"][b["in_off"][end]:b["in_off"][end + 1]]]
    assert srcs == ["z"]
    assert b["in_logp"][b["in_off"][end]] == math.log(0.7)


def test_concatenate_and_silent_order():
    d = pom.DiscreteDistribution({"A": 0.25, "C": 0.25, "G": 0.25, "T": 0.25})
    a, b = pom.HiddenMarkovModel("A"), pom.HiddenMarkovModel("B")
    for m, tag in ((a, "a"), (b, "b")):
        e = pom.State(d, name="e_" + tag)
        s = pom.State(None, name="s_" + tag)
        m.add_states(e, s)
        m.add_transition(m.start, s, 1.0)
        m.add_transition(s, e, 1.0)
        m.add_transition(e, e, 0.5)
        m.add_transition(e, m.end, 0.5)
        m.bake(merge=None)
    a.concatenate(b)
    a.bake(merge=None)
    names = [s.name for s in a.states]
    assert names[:2] == ["e_a", "e_b"] and a.silent_start == 2
    # order produced by the compiled reference engine for the same construction (silent states
    # sorted by name, then networkx-1.11 DFS order: B-end has no silent successor and is seeded
    # before A-start, so it comes first)
    assert names[2:] == ["B-end", "A-start", "s_a", "A-end", "B-start", "s_b"]
    assert (a.silent_start, a.start_index, a.end_index) == (2, 3, 2)
    assert a.end.name == "B-end" and a.finite == 1
