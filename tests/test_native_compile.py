"""Native model compilation (csrc/locus_compile.hpp behind advhmm_models_create_for_loci) against the
literal builder, which tests/test_builder_parity.py pins bit for bit on the compiled reference
(hmm_utils.get_read_matcher_model on the reference's own pomegranate).

CPU tests use a host-only context (no device needed: structure, parameter chains, and the tables
the banded kernels read are all made on the host).  What is compared: state names and order, CSR
in-edge arrays, the 64-bit patterns of every transition / emission log-probability, and -- against a
descriptor-made model of the same arrays -- the shared-memory image, first-row tables, row-0
closure and state classes byte for byte."""
import ctypes as C
import random

import numpy as np
import pytest

from advntr_b200 import engine, fast_compile, path_utils, read_matcher, synth
from conftest import same_bits


@pytest.fixture(scope="module")
def hctx():
    c = engine.Context(device=-1)
    yield c
    c.close()


def _check_locus(hctx, left, right, segs, copies, eps=0.05, kernel_tables=True):
    cols = engine.LociColumns.from_lists([left], [right], [segs], [copies], eps)
    dm = hctx.compile_loci(cols)[0]
    t = dm.tables()
    lit = read_matcher.get_read_matcher_model(left, right, segs, copies, error_rate=eps)
    b = lit.baked
    names = [s.name for s in lit.states]
    assert t["names"] == names
    for k in ("n_states", "silent_start", "start_index", "end_index", "finite"):
        assert t[k] == b[k], k
    assert np.array_equal(t["in_off"], b["in_off"]) and np.array_equal(t["in_src"], b["in_src"])
    assert same_bits(t["in_logp"], b["in_logp"])
    assert same_bits(t["emis"], b["emis"])
    if kernel_tables:
        legacy = engine.DeviceModel(hctx, b)
        mine = dm.banded_tables(b["n_states"], b["silent_start"], len(b["in_src"]))
        want = legacy.banded_tables(b["n_states"], b["silent_start"], len(b["in_src"]))
        for k in ("image", "tb1", "tb0"):
            assert np.array_equal(mine[k], want[k]), k
        assert same_bits(mine["fin_w"], want["fin_w"])
        assert mine["logp_empty"] == want["logp_empty"]
        assert np.array_equal(mine["classes"], path_utils.state_classes(names, b["emis"]))
        assert dm.info.kind == legacy.info.kind == engine.KIND_BANDED
        for f in ("n_states", "n_edges", "n_columns", "n_final_states", "smem_bytes", "max_in_degree"):
            assert getattr(dm.info, f) == getattr(legacy.info, f), f
        legacy.close()
    dm.close()
    return t


def test_config1_and_config2_loci(hctx):
    loc = synth.config1_locus()
    t = _check_locus(hctx, loc.left[-150:], loc.right[:150], loc.segments, loc.copies)
    assert (t["n_states"], t["silent_start"], len(t["in_src"])) == (1176, 768, 3877)     # config-1 sizes
    for lid in (1, 2, 5, 7, 11):
        l = synth.config2_locus(lid)
        _check_locus(hctx, l.left[-150:], l.right[:150], l.segments, l.copies)


@pytest.mark.parametrize("shape", [(1, 1, 1, 1), (1, 2, 1, 2), (2, 1, 2, 1), (3, 3, 1, 3), (5, 4, 2, 2),
                                   (10, 7, 3, 12), (20, 20, 11, 3), (7, 30, 6, 11), (100, 100, 13, 2)])
def test_degenerate_and_odd_shapes(hctx, shape):
    """One-column flanks, one-column repeat units, one copy, two-digit copy numbers: the bake orders
    (name sort, DFS topological order, in-edge order) come out of the same graph operations."""
    Ll, Lr, R, C = shape
    rng = random.Random(hash(shape) & 0xffff)
    for eps in (0.05, 0.3):
        ru = synth.rand_dna(rng, R)
        segs = [synth.substitute(rng, ru, 0.1) for _ in range(rng.randint(1, 4))]
        _check_locus(hctx, synth.rand_dna(rng, Ll), synth.rand_dna(rng, Lr), segs, C, eps)


def test_gapped_alignments(hctx):
    """Aligned repeat segments with gaps: insert columns (>= 50 % gaps), delete states in the walks,
    pseudocounts (profile_hmm.py:13-161) -- what MUSCLE's output goes through in the reference."""
    rng = random.Random(9)
    done = 0
    for trial in range(60):
        R, n = rng.randint(2, 25), rng.randint(1, 7)
        ru = synth.rand_dna(rng, R)
        width = R + rng.randint(0, 4)
        rows = []
        for _ in range(n):
            row = list(synth.substitute(rng, (ru + synth.rand_dna(rng, 8))[:width], 0.1))
            for j in range(width):
                if rng.random() < 0.25:
                    row[j] = "-"
            if all(c == "-" for c in row):
                row[0] = "A"
            rows.append("".join(row))
        left, right = synth.rand_dna(rng, rng.randint(1, 40)), synth.rand_dna(rng, rng.randint(1, 40))
        try:
            read_matcher.repeat_profile(rows, 0.05)
        except Exception:
            continue                       # the reference's profile code itself rejects this alignment
        _check_locus(hctx, left, right, rows, rng.randint(1, 6), rng.choice((0.05, 0.3)), kernel_tables=(trial % 4 == 0))
        done += 1
    assert done >= 40


def test_batch_equals_one_by_one_and_thread_count_does_not_matter(hctx):
    loci = [synth.config2_locus(i) for i in range(20, 60)]
    specs = [(l.left[-150:], l.right[:150], l.segments, l.copies, 0.05) for l in loci]
    cols = engine.LociColumns.from_lists(*[list(x) for x in zip(*specs)])
    batches = [hctx.compile_loci(cols, n_threads=nt) for nt in (1, 3, 0)]
    for i, l in enumerate(loci):
        one = fast_compile.get_read_matcher_model(*specs[i][:4]).baked
        for models in batches:
            t = models[i].tables()
            assert same_bits(t["in_logp"], one["in_logp"]) and same_bits(t["emis"], one["emis"])
            assert np.array_equal(t["in_src"], one["in_src"])
    # a sub-range of the columns
    part = hctx.compile_loci(cols, lo=7, hi=12)
    assert len(part) == 5
    assert same_bits(part[0].tables()["in_logp"], batches[0][7].tables()["in_logp"])
    for models in batches + [part]:
        for m in models:
            m.close()


def test_shape_cache_can_be_cleared_and_rebuilt(hctx):
    l = synth.config2_locus(77)
    spec = (l.left[-150:], l.right[:150], l.segments, l.copies)
    a = fast_compile.get_read_matcher_model(*spec).baked
    engine.load_library().advhmm_shape_cache_clear()
    b = fast_compile.get_read_matcher_model(*spec).baked
    assert same_bits(a["in_logp"], b["in_logp"]) and a["names"] == b["names"]


def test_libm_exp_is_not_numpy_exp_and_the_binding_installs_numpy(hctx):
    """The reference's dense round trips apply numpy.exp (hmm.pyx:514).  numpy's SIMD exp and libm's
    differ in the last bit for a few percent of the arguments, so the library takes the vector exp from
    its caller: with libm the tables drift by an ulp here and there, with numpy's they are the
    reference's."""
    lib = engine.load_library()
    l = synth.config2_locus(3)
    cols = engine.LociColumns.from_lists([l.left[-150:]], [l.right[:150]], [l.segments], [l.copies], 0.05)
    want = read_matcher.build_vntr_matcher_hmm(l.left, l.right, l.segments, l.copies, flank_size=150).baked
    try:
        lib.advhmm_set_vexp(engine.VEXP_FN(), None)                    # NULL -> libm
        m = hctx.compile_loci(cols)[0]
        libm = m.tables()["in_logp"]
        m.close()
    finally:
        lib.advhmm_set_vexp(engine._vexp_keepalive, None)
    m = hctx.compile_loci(cols)[0]
    assert same_bits(m.tables()["in_logp"], want["in_logp"])
    m.close()
    assert np.allclose(libm, want["in_logp"], rtol=1e-15, atol=0)      # at most an ulp away


def test_bad_loci_are_refused(hctx):
    ok = ("ACGTACGTAC", "TTGACCATGA", ["ACGTT", "ACGAT"], 3, 0.05)
    for bad, what in (
        (("ACGTACGTAC", "TTGACCATGA", ["ACGTT", "ACNAT"], 3, 0.05), "ACGT-"),
        (("ACGTACGTAC", "TTGACCATGA", ["-----"], 3, 0.05), "match column"),
        (("", "TTGACCATGA", ["ACGTT"], 3, 0.05), "flanks"),
        (("ACGTACGTAC", "TTGACCATGA", ["ACGTT"], 0, 0.05), "copies"),
    ):
        cols = engine.LociColumns.from_lists(*[[ok[k], bad[k]] for k in range(4)], [0.05, bad[4]])
        with pytest.raises(engine.EngineError) as ei:
            hctx.compile_loci(cols)
        assert ei.value.code == engine.EINVAL and "locus 1" in str(ei.value) and what in str(ei.value), (what, str(ei.value))
    with pytest.raises(ValueError, match="one width"):
        engine.LociColumns.from_lists(["ACGT"], ["ACGT"], [["ACGTT", "ACG"]], [3], 0.05)


def test_locus_model_can_become_a_full_model_on_the_host(hctx):
    """Generic kernel / forward / fp32 need the full set of tables: a locus model builds them on
    demand from its own arrays; here (no device) only the analysis runs."""
    l = synth.config2_locus(8)
    m = fast_compile.get_read_matcher_model(l.left[-150:], l.right[:150], l.segments, l.copies)
    b = m.baked
    legacy = engine.DeviceModel(hctx, b)
    assert legacy.info.n_columns == m._any_model().info.n_columns
    legacy.close()


def test_caller_supplied_alignment_equals_the_reference_tables():
    """tests/golden/aligned.npz: the reference's own get_read_matcher_model on repeat segments of UNEQUAL
    length, its MUSCLE wrapper answering with a fixed gapped alignment (make_golden_aligned.py).  The
    native compiler and the literal builder, given that alignment, must reproduce the reference's tables."""
    from conftest import Golden
    g = Golden("aligned")
    i = g.inputs
    assert len(set(map(len, i["segments"]))) > 1 and any("-" in a for a in i["alignment"])
    for build in (fast_compile.get_read_matcher_model, read_matcher.get_read_matcher_model):
        m = build(i["left"], i["right"], i["alignment"], i["copies"], error_rate=i["error_rate"])
        b = m.baked
        assert [s.name for s in m.states] == g.names
        assert np.array_equal(b["in_off"], g.baked["in_off"]) and np.array_equal(b["in_src"], g.baked["in_src"])
        assert same_bits(b["in_logp"], g.baked["in_logp"]) and same_bits(b["emis"], g.baked["emis"])
    with pytest.raises(ValueError, match="unequal length"):
        fast_compile.get_read_matcher_model(i["left"], i["right"], i["segments"], i["copies"])


@pytest.mark.gpu
def test_locus_with_aligned_segments_end_to_end():
    """The aligned-segment path through LocusDecoder and GenotypingRun on the device against the
    reference's decode of the golden reads (scores as bit patterns, paths, repeat counts)."""
    from conftest import Golden
    from advntr_b200 import locus_batch, pipeline
    g = Golden("aligned")
    i = g.inputs
    dec = locus_batch.LocusDecoder(i["left"], i["right"], i["segments"], read_length=150, aligned_segments=i["alignment"])
    res = dec.model.viterbi_batch(g.reads, want_summary=True)
    assert same_bits(res.logp, g.logp)
    for k in range(len(g.reads)):
        assert np.array_equal(res.path(k), g.path(k)), k
    assert np.array_equal(res.summaries["repeats"], g.ru_count)
    with pytest.raises(ValueError, match="not an alignment"):
        locus_batch.LocusDecoder(i["left"], i["right"], i["segments"], aligned_segments=i["alignment"][:-1] + ["ACGT"])
    reads = [r for r in g.reads if len(r) == 150]
    run = pipeline.GenotypingRun([pipeline.LocusSpec(7, i["left"], i["right"], i["segments"],
                                                     aligned_segments=i["alignment"])])
    got = run.genotype({7: reads})[7]
    want = dec.genotype(dec.select_reads(reads))
    for key in ("copy_numbers", "recruited_reads_count", "spanning_reads_count", "flanking_reads_count"):
        assert got[key] == want[key], key
    run.close()


@pytest.mark.gpu
def test_locus_models_decode_on_device_like_descriptor_models_and_the_oracle():
    """Device side of the native route: tables written by all host threads into pinned staging, one
    arena per batch, structural tables shared per shape.  Decoding must equal the oracle on the
    model's own arrays (scores as bit patterns, paths as arrays), the summaries those of a
    descriptor-made model; fp32 / generic / forward build the full tables on demand."""
    import oracle
    from conftest import assert_paths_equal
    ctx = engine.Context(device=0)
    ids = list(range(3, 19)) + [3, 4]                     # two loci twice: same shape, same locus
    loci = [synth.config2_locus(i) for i in ids]
    cols = engine.LociColumns.from_lists([l.left[-150:] for l in loci], [l.right[:150] for l in loci],
                                         [l.segments for l in loci], [l.copies for l in loci], 0.05)
    models = ctx.compile_loci(cols)
    groups, want = [], []
    for l, dm in zip(loci, models):
        mapped, unmapped = synth.config2_reads(l, coverage=6, decoys=4)
        codes = [oracle.encode(r) for r in mapped + unmapped + [""]]
        groups.append(codes)
        want.append(oracle.OracleModel(dm.tables()).viterbi(codes))
    res = ctx.viterbi_multi(models, groups, want_summary=True)
    k = 0
    for (lp, paths) in want:
        n = len(lp)
        assert same_bits(res.logp[k:k + n], lp)
        assert_paths_equal([res.path(i) for i in range(k, k + n)], paths)
        k += n
    legacy = [engine.DeviceModel(ctx, dm.tables()) for dm in models[:4]]
    for dm, lg, codes in zip(models, legacy, groups):
        lg.set_state_classes(path_utils.state_classes(dm.tables()["names"], dm.tables()["emis"]))
        a = dm.viterbi(codes, want_summary=True)
        b = lg.viterbi(codes, want_summary=True)
        assert same_bits(a.logp, b.logp) and np.array_equal(a.summaries, b.summaries)
        assert np.array_equal(a.paths, b.paths)
        # the full tables, built on first use
        assert same_bits(dm.viterbi(codes, precision="fp32").logp, lg.viterbi(codes, precision="fp32").logp)
        assert same_bits(dm.viterbi(codes, force_generic=True).logp, a.logp)
        assert np.array_equal(dm.log_probability(codes[:6]).view(np.int64), lg.log_probability(codes[:6]).view(np.int64))
        # ... and the model still decodes on the banded path afterwards, summaries included
        c = dm.viterbi(codes, want_summary=True)
        assert same_bits(c.logp, a.logp) and np.array_equal(c.summaries, a.summaries)
    for m in models + legacy:
        m.close()
    ctx.close()


@pytest.mark.gpu
def test_many_batches_and_model_lifetimes():
    """Arenas are freed when the last model of their batch goes; models of different batches and of
    descriptor origin mix in one call."""
    import oracle
    ctx = engine.Context(device=0)
    rng = random.Random(4)
    keep = []
    for rep in range(6):
        loci = [synth.config2_locus(100 + 7 * rep + j) for j in range(5)]
        cols = engine.LociColumns.from_lists([l.left[-150:] for l in loci], [l.right[:150] for l in loci],
                                             [l.segments for l in loci], [l.copies for l in loci], 0.05)
        models = ctx.compile_loci(cols)
        keep.append((loci[rep % 5], models[rep % 5]))
        for j, m in enumerate(models):
            if j != rep % 5:
                m.close()
    groups, want = [], []
    for l, dm in keep:
        codes = [oracle.encode(r) for r in synth.config2_reads(l, coverage=3, decoys=2)[0]]
        groups.append(codes)
        want.append(oracle.OracleModel(dm.tables()).viterbi(codes)[0])
    res = ctx.viterbi_multi([dm for _, dm in keep], groups, want_path=False)
    assert same_bits(res.logp, np.concatenate(want))
    for _, dm in keep:
        dm.close()
    ctx.close()


@pytest.mark.gpu
def test_shape_cache_clear_does_not_leave_stale_device_tables():
    """Regression: the per-context structural tables were keyed by the address of the shape structure;
    after advhmm_shape_cache_clear() a NEW structure could get the address of a dropped one and inherit
    its (wrong) device tables."""
    import oracle
    ctx = engine.Context(device=0)
    lib = engine.load_library()
    rng = random.Random(8)
    for rep in range(12):
        ids = [rng.randrange(1, 5000) for _ in range(6)]
        loci = [synth.config2_locus(i) for i in ids]
        cols = engine.LociColumns.from_lists([l.left[-150:] for l in loci], [l.right[:150] for l in loci],
                                             [l.segments for l in loci], [l.copies for l in loci], 0.05)
        models = ctx.compile_loci(cols)
        groups = [[oracle.encode(r) for r in synth.config2_reads(l, coverage=2, decoys=1)[0]] for l in loci]
        res = ctx.viterbi_multi(models, groups, want_summary=True)
        k = 0
        for dm, codes in zip(models, groups):
            lp, paths = oracle.OracleModel(dm.tables()).viterbi(codes)
            assert same_bits(res.logp[k:k + len(codes)], lp), rep
            for i, p in enumerate(paths):
                assert np.array_equal(res.path(k + i), p), rep
            k += len(codes)
        for m in models:
            m.close()
        lib.advhmm_shape_cache_clear()
    ctx.close()
