"""GPU parity: the CUDA path, called through the C-ABI (ctypes), against the golden vectors
made from the reference engine and against the oracle on fresh seeded inputs.
Bit-exact bar: log-probabilities compared as 64-bit patterns, state paths as arrays."""
import os
import random

import numpy as np
import pytest

import oracle
from conftest import Golden, assert_paths_equal, same_bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from advntr_b200 import engine
    c = engine.Context(device=0)
    yield c
    c.close()


def _decode(ctx, baked, codes, **kw):
    from advntr_b200 import engine
    dm = engine.DeviceModel(ctx, baked)
    try:
        res = dm.viterbi(codes, **kw)
        return dm.kind, res
    finally:
        dm.close()


@pytest.mark.parametrize("force_generic", [False, True], ids=["banded", "generic"])
def test_viterbi_matches_golden(ctx, golden, force_generic):
    kind, res = _decode(ctx, golden.baked, golden.codes(), force_generic=force_generic)
    assert kind == "banded"
    assert same_bits(res.logp, golden.logp)
    assert_paths_equal([res.path(i) for i in range(len(res))],
                       [golden.path(i) for i in range(len(golden.reads))], golden.name)


def test_both_strands(ctx, golden_config1):
    g = golden_config1
    from advntr_b200 import synth
    reads = [r for r in g.reads if r][:40]
    codes = [oracle.encode(r) for r in reads]
    _, res = _decode(ctx, g.baked, codes, both_strands=True)
    om = oracle.OracleModel(g.baked)
    both = []
    for r in reads:
        both += [oracle.encode(r), oracle.encode(synth.revcomp(r))]
    lp, paths = om.viterbi(both)
    assert same_bits(res.logp, lp)
    assert_paths_equal([res.path(i) for i in range(len(res))], paths)


def test_score_only(ctx, golden_config1):
    g = golden_config1
    _, res = _decode(ctx, g.baked, g.codes(), want_path=False)
    assert same_bits(res.logp, g.logp)


def test_forward_log_probability(ctx, golden):
    from advntr_b200 import engine
    dm = engine.DeviceModel(ctx, golden.baked)
    fwd = dm.log_probability(golden.codes())                         # banded forward kernel
    gen = dm.log_probability(golden.codes(), force_generic=True)      # generic CSR forward kernel
    dm.close()
    # north-star tolerance for log-probabilities: 1e-9 relative (device exp/log vs glibc)
    assert np.allclose(fwd, golden.forward, rtol=1e-9, atol=0)
    assert np.allclose(gen, golden.forward, rtol=1e-9, atol=0)


def test_forward_long_reads_and_fresh_reads_vs_oracle(ctx, golden_config1):
    """Forward on 400 fresh config-1 reads (banded kernel) and on reads longer than the banded
    kernel's 320 bases (generic kernel), against the C restatement of hmm.pyx:1371-1484."""
    import random
    import oracle
    from advntr_b200 import engine, synth
    loc = synth.config1_locus()
    rng = random.Random(99)
    reads = loc.reads(rng, 400) + [synth.revcomp(r) for r in loc.reads(rng, 50)]
    reads += [loc.reads(rng, 1, length=L)[0] for L in (1, 2, 5, 33, 97, 160, 222, 319, 320)]
    long_reads = [loc.left + loc.pattern * 8 + loc.right, synth.rand_dna(rng, 321), synth.rand_dna(rng, 700)]
    om = oracle.OracleModel(golden_config1.baked)
    dm = engine.DeviceModel(ctx, golden_config1.baked)
    for batch in (reads, long_reads, reads[:20] + long_reads):
        codes = [oracle.encode(r) for r in batch]
        got = dm.log_probability(codes)
        want = om.log_probability(codes)
        assert np.allclose(got, want, rtol=1e-9, atol=0)
    dm.close()


def test_fresh_reads_vs_oracle_and_chunking(ctx, golden_config1, monkeypatch):
    """1,000 config-1 reads; a tiny workspace budget forces many chunks through the same buffers."""
    from advntr_b200 import engine, synth
    g = golden_config1
    reads = synth.config1_reads(1000)
    codes = [oracle.encode(r) for r in reads]
    om = oracle.OracleModel(g.baked)
    lp, paths = om.viterbi(codes)
    monkeypatch.setenv("ADVHMM_WORKSPACE_MB", "8")
    small = engine.Context(device=0)
    try:
        for force in (False, True):
            _, res = _decode(small, g.baked, codes, force_generic=force)
            assert same_bits(res.logp, lp)
            assert_paths_equal([res.path(i) for i in range(len(res))], paths)
    finally:
        small.close()


def test_many_loci_one_call(ctx):
    """advhmm_viterbi_multi: reads of several loci (different models) in one launch."""
    from advntr_b200 import engine
    cases = [Golden(n) for n in ("small_a", "small_b", "divergent", "config1")]
    models = [engine.DeviceModel(ctx, c.baked) for c in cases]
    groups = [c.codes() for c in cases]
    res = ctx.viterbi_multi(models, groups)
    want_lp = np.concatenate([c.logp for c in cases])
    want_paths = [c.path(i) for c in cases for i in range(len(c.reads))]
    assert same_bits(res.logp, want_lp)
    assert_paths_equal([res.path(i) for i in range(len(res))], want_paths)
    for m in models:
        m.close()


def test_reads_longer_than_one_stripe(ctx, golden_config1):
    """Reads longer than 320 positions leave the 150 bp kernel for the striped long-read kernel
    (several stripes of 256 positions, carried through global memory)."""
    g = golden_config1
    from advntr_b200 import synth
    loc = synth.config1_locus()
    rng = random.Random(4)
    reads = [loc.sequence[:400], loc.sequence[50:450], synth.rand_dna(rng, 333), loc.sequence,
             loc.left + loc.pattern * 25 + loc.right, synth.rand_dna(rng, 700), loc.sequence[:321],
             loc.sequence[:256 + 150], loc.sequence[:512]] + g.reads[:5]
    codes = [oracle.encode(r) for r in reads]
    lp, paths = oracle.OracleModel(g.baked).viterbi(codes)
    for force in (False, True):
        _, res = _decode(ctx, g.baked, codes, force_generic=force)
        assert same_bits(res.logp, lp)
        assert_paths_equal([res.path(i) for i in range(len(res))], paths)


def test_pacbio_like_model_and_reads(ctx):
    """BASELINE config 3, scaled down: 60 bp repeat unit x 16 unrolled copies, 100 bp flanks,
    error rate 0.3 (advntr_commands.py:66-69); the model's tables (3,600 states, 253 KB) do not
    fit in shared memory, reads are ~1 kb with CLR-like errors."""
    from advntr_b200 import engine, read_matcher, synth
    rng = random.Random(33)
    ru = synth.rand_dna(rng, 60)
    left, right = synth.rand_dna(rng, 100), synth.rand_dna(rng, 100)
    model = read_matcher.get_read_matcher_model(left, right, [ru], 16, error_rate=0.3)
    dm = engine.DeviceModel(ctx, model.baked)
    assert dm.kind == "banded" and dm.info.smem_bytes > 227 * 1024
    reads = []
    for copies in (3, 9, 16, 12, 7):
        true = left + ru * copies + right
        reads.append(synth.sequencing_errors(rng, true, 0.02, 0.05, 0.05))
    reads += [reads[0][:300], reads[1][:150], ""]
    codes = [oracle.encode(r) for r in reads]
    lp, paths = oracle.OracleModel(model.baked).viterbi(codes)
    res = dm.viterbi(codes)
    dm.close()
    assert same_bits(res.logp, lp)
    assert_paths_equal([res.path(i) for i in range(len(res))], paths)
    from advntr_b200 import path_utils
    st = model.states
    counts = [path_utils.get_number_of_repeats_in_vpath([(int(k), st[k]) for k in res.path(i)]) for i in range(5)]
    assert counts == [3, 9, 16, 12, 7]


def test_non_profile_model_on_device(ctx):
    in_off = np.array([0, 2, 4, 4, 4], dtype=np.int32)
    in_src = np.array([0, 2, 0, 1], dtype=np.int32)
    in_logp = np.log(np.array([0.6, 1.0, 0.4, 1.0]))
    emis = np.log(np.array([[0.7, 0.1, 0.1, 0.1], [0.1, 0.1, 0.1, 0.7]]))
    baked = {"n_states": 4, "silent_start": 2, "start_index": 2, "end_index": 3, "finite": 0,
             "in_off": in_off, "in_src": in_src, "in_logp": in_logp, "emis": emis}
    codes = [np.array(x, dtype=np.uint8) for x in ([0, 0, 3, 3], [3], [0, 1, 2, 3, 0], [])]
    lp, paths = oracle.OracleModel(baked).viterbi(codes)
    kind, res = _decode(ctx, baked, codes)
    assert kind == "generic"
    assert same_bits(res.logp, lp)
    assert_paths_equal([res.path(i) for i in range(len(res))], paths)


def test_bad_symbol_is_rejected(ctx, golden_config1):
    from advntr_b200 import engine
    dm = engine.DeviceModel(ctx, golden_config1.baked)
    with pytest.raises(engine.EngineError) as ei:
        dm.viterbi([np.array([0, 1, 7, 2], dtype=np.uint8)])
    assert ei.value.code == engine.ESYMBOL
    dm.close()


def test_drop_in_surface(ctx):
    """model.viterbi(seq) -> (logp, [(idx, State)...]) through the pomegranate-compatible class,
    with the reference's error behaviour (hmm.pyx:72-79, 1944-1945, 1967)."""
    from advntr_b200 import pomegranate as pom, read_matcher, synth
    from advntr_b200 import path_utils
    g = Golden("config1")
    loc = synth.config1_locus()
    model = loc.build_model()
    assert [s.name for s in model.states] == g.names
    logp, vpath = model.viterbi(g.reads[0])
    assert logp == g.logp[0]
    assert [i for i, _ in vpath] == list(g.path(0))
    assert all(s is model.states[i] for i, s in vpath)
    assert path_utils.get_number_of_repeats_in_vpath(vpath) == g.ru_count[0]
    assert model.viterbi("")[0] == g.logp[1]
    assert abs(model.log_probability(g.reads[0]) - g.forward[0]) <= 1e-9 * abs(g.forward[0])
    with pytest.raises(ValueError, match="Symbol 'N' is not defined in a distribution"):
        model.viterbi("ACGTN")
    with pytest.raises(ValueError, match="must bake model"):
        pom.HiddenMarkovModel("x").viterbi("A")
    res = model.viterbi_batch(g.reads[:30])
    assert same_bits(res.logp, g.logp[:30])


def test_fp32_mode(ctx, golden):
    """Optional fp32 mode (ADVHMM_FP32): the device result equals a float restatement of the
    reference recurrence bit for bit; against the fp64 reference the stated tolerance is
    |dlogp| <= 2e-5 * |logp| + 2e-5, and repeat counts agree on >= 97 % of the reads (ties
    that fp64 separates can fall differently in fp32)."""
    from advntr_b200 import engine, path_utils
    codes = golden.codes()
    dm = engine.DeviceModel(ctx, golden.baked)
    res = dm.viterbi(codes, precision="fp32")
    dm.close()
    lp32, paths32 = oracle.OracleModel(golden.baked).viterbi(codes, fp32=True)
    assert same_bits(res.logp, lp32)
    assert_paths_equal([res.path(i) for i in range(len(res))], paths32, golden.name + " fp32")
    finite = np.isfinite(golden.logp)
    assert np.all(np.abs(res.logp[finite] - golden.logp[finite]) <= 2e-5 * np.abs(golden.logp[finite]) + 2e-5)

    class _S(object):
        def __init__(self, name):
            self.name = name
    same = total = 0
    for i in range(len(codes)):
        if golden.ru_count[i] < 0:
            continue
        vp = [(int(k), _S(golden.names[k])) for k in res.path(i)]
        same += int(path_utils.get_number_of_repeats_in_vpath(vp) == golden.ru_count[i])
        total += 1
    # north_star: "RU-count concordance reported" -- shown with `pytest -s` / in the -rA summary,
    # and bench.py repeats it in extras.fp32_mode.ru_concordance
    print("fp32 RU-count concordance [%s]: %d/%d = %.2f %%; max |dlogp|/|logp| = %.3g"
          % (golden.name, same, total, 100.0 * same / max(total, 1),
             float(np.max(np.abs(res.logp[finite] - golden.logp[finite]) / np.abs(golden.logp[finite])))))
    assert same >= 0.97 * total, "RU-count concordance %d/%d" % (same, total)


@pytest.mark.parametrize("force_generic", [False, True], ids=["banded", "generic"])
def test_on_device_path_reducers(ctx, golden, force_generic):
    """Summaries computed by the backtrack kernel equal the reference's path consumers
    (golden `consumers` = values of the reference's own hmm_utils functions)."""
    from advntr_b200 import engine, path_utils
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", golden.name + ".npz"))
    want = z["consumers"]
    dm = engine.DeviceModel(ctx, golden.baked)
    dm.set_state_classes(path_utils.state_classes(golden.names, golden.baked["emis"]))
    res = dm.viterbi(golden.codes(), want_summary=True, force_generic=force_generic)
    only = dm.viterbi(golden.codes(), want_summary=True, want_path=False, force_generic=force_generic)
    dm.close()
    assert np.array_equal(res.summaries, only.summaries) and np.array_equal(res.path_len, only.path_len)
    for i, read in enumerate(golden.reads):
        s = res.summaries[i]
        if golden.ru_count[i] < 0:
            assert s["repeats"] == -1
            continue
        assert s["repeats"] == golden.ru_count[i]
        assert [s["n_match"], s["repeat_bp"], s["left_bp"], s["right_bp"]] == list(want[i, :4])
        if read:
            assert res.flank_match_rate(i) == want[i, 4]


@pytest.mark.parametrize("max_len", [1, 20, 33, 70, 100, 129, 150, 161, 200, 225, 257, 290, 320])
def test_every_rows_per_lane_variant(ctx, golden_config1, max_len):
    """The banded kernel is instantiated per rows-per-lane (ceil(longest read / 32) = 1..10);
    each variant, with lengths that are and are not multiples of it, against the oracle."""
    from advntr_b200 import synth
    g = golden_config1
    loc = synth.config1_locus()
    rng = random.Random(max_len)
    seq = loc.sequence
    lens = sorted(set([max_len, max(1, max_len - 1), max(1, max_len - 7), max(1, max_len // 2), 1]))
    reads = []
    for L in lens:
        for _ in range(3):
            s = rng.randrange(0, len(seq) - L + 1)
            reads.append(synth.sequencing_errors(rng, seq[s:s + L + 4], 0.02, 0.01, 0.01)[:L])
    reads = [r for r in reads if r]
    codes = [oracle.encode(r) for r in reads]
    lp, paths = oracle.OracleModel(g.baked).viterbi(codes)
    _, res = _decode(ctx, g.baked, codes)
    assert same_bits(res.logp, lp)
    assert_paths_equal([res.path(i) for i in range(len(res))], paths)


def test_empty_batch_and_all_empty_reads(ctx, golden_config1):
    from advntr_b200 import engine
    g = golden_config1
    dm = engine.DeviceModel(ctx, g.baked)
    res = dm.viterbi([])
    assert len(res) == 0
    res = dm.viterbi([np.zeros(0, dtype=np.uint8)] * 3, both_strands=True)
    assert len(res) == 6 and np.all(res.logp == g.logp[1])
    assert all(np.array_equal(res.path(i), g.path(1)) for i in range(6))
    dm.close()


def test_two_contexts_with_different_model_sizes(golden):
    """Regression: the dynamic shared-memory opt-in is a property of the kernel, not of a context.
    A second context decoding a SMALLER model must not lower the limit the first one relies on."""
    from advntr_b200 import engine
    big = Golden("divergent")            # 150 bp flanks: the largest image of the golden set
    small = Golden("small_b")
    ctx_a, ctx_b = engine.Context(device=0), engine.Context(device=0)
    dm_a = engine.DeviceModel(ctx_a, big.baked)
    dm_b = engine.DeviceModel(ctx_b, small.baked)
    for dm, g in ((dm_a, big), (dm_b, small), (dm_a, big), (dm_b, small)):
        res = dm.viterbi(g.codes())
        assert same_bits(res.logp, g.logp)
        fwd = dm.log_probability(g.codes()[:8])
        assert np.allclose(fwd, g.forward[:8], rtol=1e-9, atol=0)
    dm_a.close(); dm_b.close(); ctx_a.close(); ctx_b.close()


def test_large_host_batch_is_split_and_streamed(ctx):
    """A host-buffer call with >= 262,144 reads of >= 8 models is fed to the device as sub-batches
    (planning overlaps decoding) whose state paths travel home chunk by chunk on a second stream.
    Same answers as small per-model calls, whatever the split."""
    from advntr_b200 import engine, synth
    rng = random.Random(3)
    models, groups, small = [], [], []
    for lid in range(1, 13):
        loc = synth.config2_locus(lid)
        dm = engine.DeviceModel(ctx, loc.build_model().baked)
        mapped, unmapped = synth.config2_reads(loc, coverage=12, decoys=6)
        distinct = [oracle.encode(r) for r in (mapped + unmapped)[:40]]
        reps = 23000 // len(distinct) + rng.randint(0, 3)
        models.append(dm)
        small.append(dm.viterbi(distinct))
        groups.append(distinct * reps)
    assert sum(len(g) for g in groups) >= 262144
    res = ctx.viterbi_multi(models, groups)
    k = 0
    for g, ref in zip(groups, small):
        n = len(ref)
        want_lp = np.tile(ref.logp, len(g) // n)
        assert same_bits(res.logp[k:k + len(g)], want_lp)
        for i in range(0, len(g), 997):                       # a sample of the paths, every replica block
            assert np.array_equal(res.path(k + i), ref.path(i % n))
        assert np.array_equal(res.path_len[k:k + len(g)], np.tile(ref.path_len, len(g) // n))
        k += len(g)
    for dm in models:
        dm.close()


def test_250bp_reads_large_image_takes_the_16_warp_kernel(ctx):
    """250 bp reads: 250-base flanks in the model (vntr_finder.py:131), ~770 columns, a 160 KB image:
    one 16-warp CTA per SM instead of two 8-warp CTAs.  Bit-exact against the oracle."""
    from advntr_b200 import engine, synth
    rng = random.Random(250)
    ru = synth.rand_dna(rng, 24)
    segs = [synth.substitute(rng, ru, 0.03) for _ in range(5)]
    loc = synth.Locus(77, synth.rand_dna(rng, 400), synth.rand_dna(rng, 400), segs, read_length=250, flank=250)
    model = loc.build_model()
    dm = engine.DeviceModel(ctx, model.baked)
    assert dm.kind == "banded" and dm.info.smem_bytes > 113 * 1024
    reads = loc.reads(rng, 70) + [synth.revcomp(r) for r in loc.reads(rng, 10)]
    reads += [loc.reads(rng, 1, length=L)[0] for L in (161, 200, 249, 251, 300, 320)]
    codes = [oracle.encode(r) for r in reads]
    want_lp, want_paths = oracle.OracleModel(model.baked).viterbi(codes)
    res = dm.viterbi(codes)
    assert same_bits(res.logp, want_lp)
    assert_paths_equal([res.path(i) for i in range(len(res))], want_paths)
    dm.close()


def test_path_scores_reproduce_logp_at_scale(ctx):
    """Size-independent property at bench scale (no oracle): 300 config-2 loci, ~46,000 reads in one
    device call; every returned state path, re-scored with the model's tables, gives the returned
    log-probability bit for bit, starts in the start state, ends in the end state and emits the read."""
    from advntr_b200 import engine, path_utils, synth
    models, groups, bakeds = [], [], []
    for lid in range(1, 301):
        loc = synth.config2_locus(lid)
        b = loc.build_model().baked if lid % 50 == 0 else None
        if b is None:
            from advntr_b200 import fast_compile
            b = fast_compile.build_vntr_matcher_hmm(loc.left, loc.right, loc.segments, loc.copies, flank_size=150).baked
        flat, lengths = synth.config2_read_codes(loc)
        cuts = np.concatenate([[0], np.cumsum(lengths)])
        groups.append([flat[cuts[i]:cuts[i + 1]] for i in range(len(lengths))])
        models.append(engine.DeviceModel(ctx, b))
        bakeds.append(b)
    res = ctx.viterbi_multi(models, groups)
    assert (res.path_len > 0).all()
    first = 0
    for b, g in zip(bakeds, groups):
        v = path_utils.rescore_paths(b, g, [res.path(first + i) for i in range(len(g))])
        assert same_bits(v, res.logp[first:first + len(g)])
        first += len(g)
    assert first > 40000
    for dm in models:
        dm.close()


@pytest.mark.parametrize("seed", [101, 202, 303])
def test_random_loci_random_reads_vs_oracle(ctx, seed):
    """Differential test over model SHAPES: random flank lengths (incl. very short), repeat-unit lengths
    1..80, 1..12 unrolled copies, both error rates, divergent repeat segments; reads of length 0..330:
    windows of the locus, pure repeats, homopolymers, random sequence, either strand.  Every model must
    decode bit-identically to the C restatement of the reference, whichever kernel it is routed to."""
    from advntr_b200 import engine, read_matcher, synth
    rng = random.Random(seed)
    kinds = set()
    for _ in range(8):
        R = rng.choice((1, 2, 3, 5, 8, 13, 21, 34, 55, 80))
        copies = rng.randint(1, 12 if R < 30 else 4)
        Ll, Lr = rng.choice((1, 2, 7, 30, 100, 150)), rng.choice((1, 3, 12, 60, 150))
        eps = rng.choice((0.05, 0.3))
        ru = synth.rand_dna(rng, R)
        nseg = rng.randint(1, 6)
        segs = [synth.substitute(rng, ru, rng.choice((0.0, 0.05, 0.3))) for _ in range(nseg)]
        left, right = synth.rand_dna(rng, Ll), synth.rand_dna(rng, Lr)
        if rng.random() < 0.2:
            left = "A" * Ll                                  # homopolymer flank: ties everywhere
        model = read_matcher.get_read_matcher_model(left, right, segs, copies, error_rate=eps)
        locus = left + "".join(segs) * 3 + right
        reads = ["", rng.choice("ACGT"), ru * rng.randint(1, 40), "A" * rng.randint(1, 200), "ACGT" * rng.randint(1, 70)]
        for _ in range(35):
            n = rng.choice((1, 2, 3, 10, 31, 32, 33, 64, 100, 150, 151, 159, 160, 161, 200, 250, 256, 300, 320, 321, 330))
            if rng.random() < 0.6 and len(locus) > 2:
                s = rng.randrange(0, len(locus))
                r = synth.sequencing_errors(rng, (locus * (n // len(locus) + 2))[s:s + n + 8], 0.02, 0.01, 0.01)[:n]
            else:
                r = synth.rand_dna(rng, n)
            reads.append(synth.revcomp(r) if rng.random() < 0.3 else r)
        reads = [r[:330] for r in reads]
        codes = [oracle.encode(r) for r in reads]
        want_lp, want_paths = oracle.OracleModel(model.baked).viterbi(codes)
        dm = engine.DeviceModel(ctx, model.baked)
        kinds.add(dm.kind)
        res = dm.viterbi(codes)
        assert same_bits(res.logp, want_lp), (seed, R, copies, Ll, Lr, eps)
        assert_paths_equal([res.path(i) for i in range(len(res))], want_paths, "R=%d C=%d L=%d/%d" % (R, copies, Ll, Lr))
        gen = dm.viterbi(codes[:12], force_generic=True)
        assert same_bits(gen.logp, want_lp[:12])
        fwd = dm.log_probability(codes[:10])
        assert np.allclose(fwd, oracle.OracleModel(model.baked).log_probability(codes[:10]), rtol=1e-9, atol=0)
        dm.close()
    assert "banded" in kinds


def test_host_call_with_page_locked_result_arrays(ctx):
    """Scores, path offsets and summaries go straight into the caller's arrays when those are page-locked
    (no staging copy); same bits as the pageable route, with and without state paths."""
    import ctypes as C
    import torch
    from advntr_b200 import engine, synth
    lib = engine.load_library()
    models, groups = [], []
    for lid in (3, 4, 5):
        loc = synth.config2_locus(lid)
        dm = engine.DeviceModel(ctx, loc.build_model().baked)
        mapped, unmapped = synth.config2_reads(loc, coverage=10, decoys=5)
        models.append(dm)
        groups.append([oracle.encode(r) for r in mapped + unmapped])
    want = ctx.viterbi_multi(models, groups)                        # numpy arrays: pageable route
    flat = [c for g in groups for c in g]
    seqs, off = engine.pack_reads(flat)
    goff = np.zeros(len(groups) + 1, dtype=np.int64)
    np.cumsum([len(g) for g in groups], out=goff[1:])
    R = len(flat)
    cap = int(want.paths.size) + 64
    h_seqs = torch.from_numpy(seqs).pin_memory()
    h_logp = torch.empty(R, dtype=torch.float64).pin_memory()
    h_plen = torch.empty(R, dtype=torch.int32).pin_memory()
    h_poff = torch.empty(R, dtype=torch.int64).pin_memory()
    h_path = torch.empty(cap, dtype=torch.int32).pin_memory()
    total = C.c_int64(0)
    handles = (C.c_void_p * len(models))(*[m._h for m in models])
    engine._check(lib.advhmm_viterbi_multi(ctx._h, handles, len(models), goff.ctypes.data, h_seqs.data_ptr(),
                                           off.ctypes.data, R, engine.WANT_PATH, h_logp.data_ptr(), h_plen.data_ptr(),
                                           h_poff.data_ptr(), h_path.data_ptr(), cap, C.byref(total)))
    assert same_bits(h_logp.numpy(), want.logp)
    assert np.array_equal(h_plen.numpy(), want.path_len)
    assert total.value == want.paths.size
    for i in range(0, R, 7):
        a = int(h_poff[i])
        assert np.array_equal(h_path.numpy()[a:a + int(h_plen[i])], want.path(i))
    # scores only (no paths): the same route without the path arrays
    h_logp.zero_()
    engine._check(lib.advhmm_viterbi_multi(ctx._h, handles, len(models), goff.ctypes.data, h_seqs.data_ptr(),
                                           off.ctypes.data, R, 0, h_logp.data_ptr(), None, None, None, 0, None))
    assert same_bits(h_logp.numpy(), want.logp)
    for dm in models:
        dm.close()


def test_group_offsets_must_cover_every_read(ctx, golden_config1):
    """Reads outside every group would never be decoded, yet their result slots would come back as if
    they had been: the call is refused instead (ADVHMM_EINVAL)."""
    import ctypes as C
    from advntr_b200 import engine
    lib = engine.load_library()
    dm = engine.DeviceModel(ctx, golden_config1.baked)
    codes = golden_config1.codes()[:6]
    seqs, off = engine.pack_reads(codes)
    handles = (C.c_void_p * 1)(dm._h)
    logp = np.empty(6)
    for goff in ([0, 4], [1, 6], [2, 4]):
        g = np.asarray(goff, dtype=np.int64)
        rc = lib.advhmm_viterbi_multi(ctx._h, handles, 1, g.ctypes.data, seqs.ctypes.data, off.ctypes.data, 6, 0,
                                      logp.ctypes.data, None, None, None, 0, None)
        assert rc == engine.EINVAL, goff
        assert b"group_off" in lib.advhmm_last_error()
    g = np.asarray([0, 6], dtype=np.int64)
    assert lib.advhmm_viterbi_multi(ctx._h, handles, 1, g.ctypes.data, seqs.ctypes.data, off.ctypes.data, 6, 0,
                                    logp.ctypes.data, None, None, None, 0, None) == 0
    assert same_bits(logp, golden_config1.logp[:6])
    dm.close()


def test_one_alphabet_per_call_and_at_most_four_symbols(ctx, golden_config1):
    """Reads are packed 2 bits per symbol and validated against one alphabet size per call."""
    from advntr_b200 import engine
    two = {"n_states": 3, "silent_start": 1, "start_index": 1, "end_index": 2, "finite": 1,
           "in_off": np.array([0, 2, 2, 3], dtype=np.int32), "in_src": np.array([0, 1, 0], dtype=np.int32),
           "in_logp": np.log(np.array([0.5, 1.0, 0.5])), "emis": np.log(np.array([[0.5, 0.5]]))}
    small = engine.DeviceModel(ctx, two)
    assert small.n_symbols == 2
    big = engine.DeviceModel(ctx, golden_config1.baked)
    c = [np.array([0, 1, 0], dtype=np.uint8)]
    lp, paths = oracle.OracleModel(two).viterbi(c)
    assert same_bits(small.viterbi(c).logp, lp)
    with pytest.raises(engine.EngineError) as ei:
        ctx.viterbi_multi([big, small], [c, c])
    assert ei.value.code == engine.EINVAL
    five = dict(two, emis=np.log(np.full((1, 5), 0.2)))
    with pytest.raises(engine.EngineError) as ei:
        engine.DeviceModel(ctx, five)
    assert ei.value.code == engine.EINVAL
    small.close(); big.close()


def test_bad_symbol_in_device_buffer_mode_is_reported(ctx, golden_config1):
    """ADVHMM_DEVICE_BUFFERS calls are asynchronous: a code outside the alphabet cannot fail the call, it
    is reported by advhmm_context_bad_symbol() (the host path returns ADVHMM_ESYMBOL)."""
    import ctypes as C
    import torch
    from advntr_b200 import engine
    lib = engine.load_library()
    dm = engine.DeviceModel(ctx, golden_config1.baked)
    codes = [c.copy() for c in golden_config1.codes()[:5]]
    handles = (C.c_void_p * 1)(dm._h)
    goff = np.asarray([0, 5], dtype=np.int64)
    d_logp = torch.empty(5, dtype=torch.float64, device="cuda")
    for bad_read in (None, 3):
        if bad_read is not None:
            codes[bad_read][7] = 9
        seqs, off = engine.pack_reads(codes)
        d_seqs = torch.from_numpy(seqs).cuda()
        torch.cuda.synchronize()
        engine._check(lib.advhmm_viterbi_multi(ctx._h, handles, 1, goff.ctypes.data, d_seqs.data_ptr(),
                                               off.ctypes.data, 5, engine.DEVICE_BUFFERS, d_logp.data_ptr(),
                                               None, None, None, 0, None))
        assert ctx.bad_symbol() == (-1 if bad_read is None else bad_read)
    dm.close()


@pytest.mark.parametrize("read_length", [100, 250])
def test_locus_decoder_uses_read_length_flanks(ctx, read_length):
    """get_vntr_matcher_hmm builds the matcher with flanking_region_size = read_length
    (vntr_finder.py:131-132): LocusDecoder for a 100 / 250 bp library must decode against the model
    with 100 / 250-base flanks.  Reference tables: read_matcher's literal builder (bit-equal to the
    reference's hmm_utils, tests/test_builder_parity.py); decoding checked against the oracle."""
    from advntr_b200 import locus_batch, read_matcher, synth
    rng = random.Random(read_length)
    ru = synth.rand_dna(rng, 21)
    left, right = synth.rand_dna(rng, 400), synth.rand_dna(rng, 400)
    segs = [synth.substitute(rng, ru, 0.03) for _ in range(4)]
    dec = locus_batch.LocusDecoder(left, right, segs, read_length=read_length)
    copies = read_matcher.copies_for_read_length(read_length, len(ru))
    want = read_matcher.get_read_matcher_model(left[-read_length:], right[:read_length], segs, copies)
    b, w = dec.model.baked, want.baked
    assert b["n_states"] == w["n_states"] and np.array_equal(b["in_src"], w["in_src"])
    assert same_bits(b["in_logp"], w["in_logp"]) and same_bits(b["emis"], w["emis"])
    allele = left + "".join(segs) + right
    reads = []
    for _ in range(40):
        s = rng.randrange(400 - read_length + 5, 400 + 84 - 5)
        reads.append(synth.sequencing_errors(rng, allele[s:s + read_length + 8], 0.01, 0.001, 0.001)[:read_length])
    reads += [synth.rand_dna(rng, read_length) for _ in range(6)]
    lp, paths = oracle.OracleModel(w).viterbi([oracle.encode(r) for r in reads])
    res = dec.model.viterbi_batch(reads)
    assert same_bits(res.logp, lp)
    assert_paths_equal([res.path(i) for i in range(len(res))], paths, "read length %d" % read_length)
    # and the call site: the flank match-rate / recruitment decisions follow from those paths
    selected = dec.select_reads(reads[:40])
    assert len(selected) >= 30


def test_config3_at_its_stated_size(ctx):
    """BASELINE config 3 as stated: 60 bp repeat unit x 100 unrolled copies, 100 bp flanks, error rate 0.3
    (18,918 states, 6,412 columns, 1.3 MB of tables: the striped long-read kernel with the per-warp TMA
    ring), PacBio-like reads of 10-14 kb -- more repeat copies than the model unrolls, CLR-like errors --
    against the CPU oracle: scores as bit patterns, whole state paths, repeat counts.  (The oracle needs
    16 bytes per DP cell, 3-4 GB and a few seconds per read: two long reads, plus short and empty ones.)"""
    from advntr_b200 import fast_compile, path_utils, synth
    loc = synth.config3_locus()
    model = fast_compile.get_read_matcher_model(loc.left, loc.right, loc.segments, loc.copies, error_rate=0.3)
    dm = model._device_model()
    assert dm.kind == "banded" and dm.info.n_states == 18918 and dm.info.smem_bytes > 227 * 1024
    rng = random.Random(303)
    reads = []
    for copies in (170, 225):                      # ~10.5 kb and ~14 kb
        reads.append(synth.sequencing_errors(rng, loc.left + loc.pattern * copies + loc.right, 0.02, 0.05, 0.05))
    reads.append(synth.sequencing_errors(rng, loc.left + loc.pattern * 40 + loc.right, 0.02, 0.05, 0.05))   # spans, 40 copies
    reads += [reads[0][:161], reads[1][:160], reads[1][5000:5321], ""]
    assert 10000 <= len(reads[0]) <= 20000 and 10000 <= len(reads[1]) <= 20000
    codes = [oracle.encode(r) for r in reads]
    want_lp, want_paths = oracle.OracleModel(model.baked).viterbi(codes)
    res = model.viterbi_batch(reads)
    assert same_bits(res.logp, want_lp)
    assert_paths_equal([res.path(i) for i in range(len(res))], want_paths, "config 3")
    st = model.states
    assert path_utils.get_number_of_repeats_in_vpath([(int(k), st[k]) for k in res.path(2)]) == 40
    summ = model.viterbi_batch(reads, want_path=False, want_summary=True)
    assert same_bits(summ.logp, want_lp) and summ.summaries["repeats"][2] == 40


@pytest.mark.parametrize("wpr", [1, 2, 4, 8])
def test_long_read_kernel_with_every_warps_per_read_setting(ctx, wpr, monkeypatch):
    """The striped long-read kernel deals the stripes of a read to 1, 2, 4 or 8 warps that trail each other
    by three column blocks (carry hand-over through memory + progress flags).  Whatever the split, the
    result is the oracle's: reads of 1 .. 23 stripes, fewer stripes than warps included."""
    from advntr_b200 import engine, read_matcher, synth
    monkeypatch.setenv("ADVHMM_LONG_WPR", str(wpr))
    rng = random.Random(1000 + wpr)
    ru = synth.rand_dna(rng, 41)
    left, right = synth.rand_dna(rng, 100), synth.rand_dna(rng, 100)
    model = read_matcher.get_read_matcher_model(left, right, [ru], 30, error_rate=0.3)
    dm = engine.DeviceModel(ctx, model.baked)
    assert dm.kind == "banded" and dm.info.smem_bytes > 227 * 1024
    reads = []
    for copies in (1, 3, 7, 12, 19, 30, 44, 80):
        reads.append(synth.sequencing_errors(rng, left + ru * copies + right, 0.02, 0.05, 0.05))
    reads += [reads[3][:159], reads[3][:160], reads[3][:161], reads[5][:321], "", "ACGT"]
    codes = [oracle.encode(r) for r in reads]
    lp, paths = oracle.OracleModel(model.baked).viterbi(codes)
    for _ in range(2):                                   # twice: flags and rings start clean every launch
        res = dm.viterbi(codes)
        assert same_bits(res.logp, lp)
        assert_paths_equal([res.path(i) for i in range(len(res))], paths, "wpr %d" % wpr)
    dm.close()


def test_long_reads_in_several_chunks_backtrack_overlaps_the_next_fill(monkeypatch):
    """A workspace too small for all long reads of a batch: the long-read family runs in chunks, two
    workspace halves, the backtrack of chunk k on a second stream next to the fill of chunk k+1.  Same
    scores and paths as the oracle, through both front ends (host buffers with chunk-wise path copies)."""
    from advntr_b200 import engine, read_matcher, synth
    monkeypatch.setenv("ADVHMM_WORKSPACE_MB", "24")
    c2 = engine.Context(device=0)
    rng = random.Random(77)
    ru = synth.rand_dna(rng, 41)
    left, right = synth.rand_dna(rng, 100), synth.rand_dna(rng, 100)
    model = read_matcher.get_read_matcher_model(left, right, [ru], 30, error_rate=0.3)
    dm = engine.DeviceModel(c2, model.baked)
    reads = [synth.sequencing_errors(rng, left + ru * rng.randint(5, 60) + right, 0.02, 0.05, 0.05) for _ in range(21)]
    reads += [reads[0][:100], ""]                       # a short-kernel read and an empty one in the same call
    codes = [oracle.encode(r) for r in reads]
    lp, paths = oracle.OracleModel(model.baked).viterbi(codes)
    launches0 = c2.launch_count
    for _ in range(2):
        res = dm.viterbi(codes, want_summary=True)
        assert same_bits(res.logp, lp)
        assert_paths_equal([res.path(i) for i in range(len(res))], paths, "chunked long reads")
    assert c2.launch_count - launches0 >= 2 * (1 + 2 * 3), "expected the long reads to be split into several chunks"
    only = dm.viterbi(codes, want_path=False)
    assert same_bits(only.logp, lp)
    dm.close()
    c2.close()


@pytest.mark.parametrize("seed", [11, 22])
def test_random_shapes_native_models_and_long_reads_vs_oracle(ctx, seed, monkeypatch):
    """Differential test of this round's two new pieces together: models made by the NATIVE compiler
    (advhmm_models_create_for_loci; random shapes incl. gapped alignments) and reads of 321..2,600 bases
    (the striped long-read kernel; a random warps-per-read setting per model) against the C restatement of
    the reference running on the LITERAL builder's tables."""
    from advntr_b200 import engine, read_matcher, synth
    rng = random.Random(seed)
    specs = []
    for _ in range(6):
        R = rng.choice((2, 3, 5, 8, 13, 21, 34, 55))
        copies = rng.randint(1, 10 if R < 30 else 4)
        Ll, Lr = rng.choice((1, 2, 7, 30, 100, 150)), rng.choice((1, 3, 12, 60, 150))
        eps = rng.choice((0.05, 0.3))
        ru = synth.rand_dna(rng, R)
        rows = [synth.substitute(rng, ru, rng.choice((0.0, 0.05, 0.3))) for _ in range(rng.randint(1, 6))]
        if rng.random() < 0.5:                               # gapped alignment: one extra column, gaps sprinkled
            rows = [r[:R // 2] + (rng.choice("ACGT") if rng.random() < 0.3 else "-") + r[R // 2:] for r in rows]
            rows = ["".join("-" if (c != "-" and rng.random() < 0.08) else c for c in r) for r in rows]
            rows = [r if r.strip("-") else "A" + r[1:] for r in rows]
        left, right = synth.rand_dna(rng, Ll), synth.rand_dna(rng, Lr)
        try:
            read_matcher.repeat_profile(rows, eps)
        except Exception:
            rows = [r.replace("-", "A") for r in rows]
        specs.append((left, right, rows, copies, eps))
    cols = engine.LociColumns.from_lists(*[list(x) for x in zip(*specs)])
    models = ctx.compile_loci(cols)
    for (left, right, rows, copies, eps), dm in zip(specs, models):
        lit = read_matcher.get_read_matcher_model(left, right, rows, copies, error_rate=eps)
        t = dm.tables()
        assert same_bits(t["in_logp"], lit.baked["in_logp"]) and same_bits(t["emis"], lit.baked["emis"])
        assert np.array_equal(t["in_src"], lit.baked["in_src"])
        locus = left + "".join(r.replace("-", "") for r in rows) * 4 + right
        reads = ["", "G"]
        for n in (150, 321, 400, 640, 801, 1281, rng.randint(1500, 2600)):
            s0 = rng.randrange(0, len(locus))
            r = synth.sequencing_errors(rng, (locus * (n // len(locus) + 2))[s0:s0 + n + 40], 0.02, 0.03, 0.03)[:n]
            reads.append(synth.revcomp(r) if rng.random() < 0.3 else r)
        reads.append(synth.rand_dna(rng, 700))
        codes = [oracle.encode(r) for r in reads]
        want_lp, want_paths = oracle.OracleModel(lit.baked).viterbi(codes)
        monkeypatch.setenv("ADVHMM_LONG_WPR", str(rng.choice((1, 2, 4, 8))))
        res = dm.viterbi(codes, want_summary=True)
        assert same_bits(res.logp, want_lp), (seed, len(left), len(right), copies, eps)
        assert_paths_equal([res.path(i) for i in range(len(res))], want_paths, "native + long reads")
        dm.close()


@pytest.mark.parametrize("wpr", [1, 8])
def test_fp32_mode_on_long_reads(ctx, wpr, monkeypatch):
    """The optional fp32 mode covers the striped long-read kernel too (the same kernel instantiated for
    float: float image, float carry rows): bit-equal to the float restatement of the reference recurrence
    (oracle_viterbi_f32); against the fp64 result the stated tolerance for reads of thousands of bases is
    |dlogp| <= 1e-4 |logp| + 1e-4 (short reads: 2e-5, test_fp32_mode)."""
    from advntr_b200 import engine, read_matcher, synth
    monkeypatch.setenv("ADVHMM_LONG_WPR", str(wpr))
    rng = random.Random(3200 + wpr)
    ru = synth.rand_dna(rng, 37)
    left, right = synth.rand_dna(rng, 100), synth.rand_dna(rng, 100)
    model = read_matcher.get_read_matcher_model(left, right, [ru, synth.substitute(rng, ru, 0.1)], 40, error_rate=0.3)
    dm = engine.DeviceModel(ctx, model.baked)
    assert dm.info.smem_bytes > 227 * 1024
    reads = [synth.sequencing_errors(rng, left + ru * k + right, 0.02, 0.05, 0.05) for k in (2, 9, 17, 33, 60)]
    reads += [reads[2][:161], reads[3][:320], reads[3][:321], ""]
    codes = [oracle.encode(r) for r in reads]
    om = oracle.OracleModel(model.baked)
    lp32, paths32 = om.viterbi(codes, fp32=True)
    res = dm.viterbi(codes, precision="fp32")
    assert same_bits(res.logp, lp32)
    assert_paths_equal([res.path(i) for i in range(len(res))], paths32, "fp32 long reads")
    # stated tolerance for long reads: float rounding accumulates along thousands of additions
    # (observed: 2.4e-5 relative on a 2.5 kb read)
    lp64, _ = om.viterbi(codes)
    assert np.all(np.abs(res.logp - lp64) <= 1e-4 * np.abs(lp64) + 1e-4)
    # and a locus model (native compiler) builds its float tables on first use
    from advntr_b200 import fast_compile
    nat = fast_compile.compile_many([(left, right, [ru, synth.substitute(random.Random(3200 + wpr + 1), ru, 0.1)], 40, 0.3)], ctx)[0]
    r2 = nat._device_model().viterbi(codes[:4], precision="fp32")
    lp_n, _ = oracle.OracleModel(nat.baked).viterbi(codes[:4], fp32=True)
    assert same_bits(r2.logp, lp_n)
    nat._release_engine()
    dm.close()


@pytest.mark.parametrize("wpr", [1, 4])
def test_forward_on_long_reads_banded(ctx, wpr, monkeypatch):
    """log_probability of reads beyond the register wavefront's 320 bases, and of any read of a model whose
    tables exceed shared memory: the striped long-read kernel instantiated with log-sum-exp (round 1 sent
    these to the generic kernel).  1e-9 relative against the C restatement of hmm.pyx:1371-1484; the generic
    forward kernel must agree too."""
    from advntr_b200 import engine, read_matcher, synth
    monkeypatch.setenv("ADVHMM_LONG_WPR", str(wpr))
    rng = random.Random(640 + wpr)
    ru = synth.rand_dna(rng, 29)
    left, right = synth.rand_dna(rng, 100), synth.rand_dna(rng, 100)
    small = read_matcher.get_read_matcher_model(left, right, [ru], 6, error_rate=0.05)       # image fits in shared memory
    big = read_matcher.get_read_matcher_model(left, right, [ru], 45, error_rate=0.3)          # it does not
    for model in (small, big):
        dm = engine.DeviceModel(ctx, model.baked)
        reads = [synth.sequencing_errors(rng, left + ru * k + right, 0.02, 0.03, 0.03) for k in (1, 5, 12, 30)]
        reads += [reads[3][:321], reads[3][:160], reads[2][:100], "", "T"]
        codes = [oracle.encode(r) for r in reads]
        want = oracle.OracleModel(model.baked).log_probability(codes)
        got = dm.log_probability(codes)
        assert np.allclose(got, want, rtol=1e-9, atol=0), (got, want)
        gen = dm.log_probability(codes, force_generic=True)
        assert np.allclose(gen, want, rtol=1e-9, atol=0)
        dm.close()
