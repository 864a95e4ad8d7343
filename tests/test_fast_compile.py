"""The native per-shape compiler (fast_compile.py -> advhmm_models_create_for_loci) gives
bit-identical tables to the literal (reference-shaped) builder, across shapes, loci and error rates;
and identical to the golden vectors.  tests/test_native_compile.py goes further (degenerate shapes,
gapped alignments, the kernel-side tables, batches)."""
import random

import numpy as np
import pytest

from advntr_b200 import fast_compile, read_matcher, synth


def _same(a, b):
    for k in ("n_states", "silent_start", "start_index", "end_index", "finite"):
        assert a[k] == b[k], k
    for k in ("in_off", "in_src"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("in_logp", "emis"):
        assert np.array_equal(np.asarray(a[k]).view(np.int64), np.asarray(b[k]).view(np.int64)), k


@pytest.mark.parametrize("locus_id", [1, 2, 5, 7, 11, 12])
def test_fast_equals_literal_config2(locus_id):
    loc = synth.config2_locus(locus_id)
    lit = read_matcher.build_vntr_matcher_hmm(loc.left, loc.right, loc.segments, loc.copies, flank_size=150)
    fast = fast_compile.build_vntr_matcher_hmm(loc.left, loc.right, loc.segments, loc.copies, flank_size=150)
    _same(fast.baked, lit.baked)
    assert [s.name for s in fast.states] == [s.name for s in lit.states]


def test_template_reuse_across_loci_of_one_shape():
    rng = random.Random(5)
    base = synth.config2_locus(9)
    R, n = len(base.pattern), len(base.segments)
    for i in range(4):
        ru = synth.rand_dna(rng, R)
        segs = [synth.substitute(rng, ru, 0.15) for _ in range(n)]
        left, right = synth.rand_dna(rng, 300), synth.rand_dna(rng, 300)
        for eps in (0.05,):
            lit = read_matcher.build_vntr_matcher_hmm(left, right, segs, base.copies, flank_size=150, error_rate=eps)
            fast = fast_compile.build_vntr_matcher_hmm(left, right, segs, base.copies, flank_size=150, error_rate=eps)
            _same(fast.baked, lit.baked)
        assert fast.baked["shape"] == (150, 150, R, base.copies)


def test_fast_equals_golden(golden):
    i = golden.inputs
    fast = fast_compile.get_read_matcher_model(i["left"], i["right"], i["segments"], i["copies"],
                                               error_rate=i["error_rate"])
    assert [s.name for s in fast.states] == golden.names
    assert np.array_equal(fast.baked["in_src"], golden.baked["in_src"])
    assert np.allclose(fast.baked["in_logp"], golden.baked["in_logp"], rtol=1e-14, atol=1e-15)
