"""Keyword pre-filter (the step before the hot path).

CPU: the restatement ``oracle/kfilter_oracle.py`` against the stdout of the reference's own
``adVNTR-Filtering`` binary -- the committed golden capture, and (where the compiled binary is
present) fresh random cases.  GPU: ``advntr_b200.keyword_filter`` through the C-ABI against both.
"""
import gzip
import json
import os
import random
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

import kfilter_oracle
from advntr_b200 import synth

FILTER_BIN = os.path.join(ROOT, "oracle", "_ref", "adVNTR-Filtering")


def _golden_stdout():
    with gzip.open(os.path.join(GOLDEN, "kfilter_reference_stdout.json.gz"), "rt") as fh:
        return json.load(fh)


def _run_reference_binary(kw, names, seqs, min_matches):
    with tempfile.TemporaryDirectory() as d:
        fa, kf = os.path.join(d, "reads.fa"), os.path.join(d, "kw.txt")
        with open(fa, "w") as fh:
            for n, s in zip(names, seqs):
                fh.write(">%s\n%s\n" % (n, s))
        with open(kf, "w") as fh:
            for vid, words in kw:
                fh.write("%s %s\n" % (vid, " ".join(words)))
        with open(kf) as stdin:
            return subprocess.run([FILTER_BIN, fa, "--min_matches", str(min_matches)], stdin=stdin,
                                  capture_output=True, text=True, check=True).stdout


def _random_case(seed, lengths=(15,), n_loci=12, n_reads=150, read_len=(20, 200), share=True):
    """Small adversarial inputs: short alphabets (many repeated k-mers), keywords shared by several
    loci, overlapping occurrences, reads shorter than k, non-ACGT symbols, mixed keyword lengths."""
    rng = random.Random(seed)
    alphabet = "ACGT" if seed % 2 else "AC"
    genome = "".join(rng.choice(alphabet) for _ in range(600))
    kw = []
    for v in range(n_loci):
        words = []
        for _ in range(rng.randint(1, 9)):
            k = rng.choice(lengths)
            p = rng.randrange(0, len(genome) - k)
            words.append(genome[p:p + k])
        if share and kw and rng.random() < 0.4:
            words.append(rng.choice(kw[rng.randrange(len(kw))][1]))
        if rng.random() < 0.2:
            words.append(words[0])                       # duplicate within a line collapses
        if rng.random() < 0.15:
            w = words[0]
            words.append(w[:3] + "N" + w[4:])            # the fifth symbol inside a keyword
        kw.append((100 + 7 * v, words))
    names, seqs = [], []
    for r in range(n_reads):
        n = rng.randint(*read_len)
        p = rng.randrange(0, len(genome) - n) if n < len(genome) else 0
        s = genome[p:p + n]
        if rng.random() < 0.3:
            q = rng.randrange(len(s))
            s = s[:q] + rng.choice("NnRx") + s[q + 1:]
        if rng.random() < 0.1:
            s = s.lower()
        names.append("read_%03d" % r)
        seqs.append(s)
    return kw, names, seqs


# ------------------------------------------------------------------------------------ CPU
def test_oracle_reproduces_reference_binary_stdout():
    kw, names, seqs = synth.kfilter_case()
    want = _golden_stdout()
    for mm in (5, 2):
        assert kfilter_oracle.filter_output(kw, names, seqs, min_matches=mm) == want[str(mm)]


@pytest.mark.skipif(not os.path.exists(FILTER_BIN), reason="oracle/_ref/adVNTR-Filtering not built here")
@pytest.mark.parametrize("seed,lengths", [(1, (15,)), (2, (15,)), (3, (7, 15, 21)), (4, (5,)), (5, (15, 30, 80)),
                                          (6, (3, 4))])
def test_oracle_matches_compiled_reference_on_random_cases(seed, lengths):
    kw, names, seqs = _random_case(seed, lengths)
    for mm in (1, 3, 5):
        assert kfilter_oracle.filter_output(kw, names, seqs, min_matches=mm) == \
            _run_reference_binary(kw, names, seqs, mm)


def test_keywords_for_filtering_follow_the_reference_rule():
    from advntr_b200 import keyword_filter
    loc = synth.config2_locus(5)
    words = keyword_filter.get_keywords_for_filtering(loc.left, loc.right, loc.segments, loc.pattern, keyword_size=15)
    locus = loc.left[-15:] + "".join(loc.segments) + loc.right[:15]
    assert words == {locus[i:i + 15] for i in range(0, len(locus) - 14, 5)}
    long_words = keyword_filter.get_keywords_for_filtering(loc.left, loc.right, loc.segments, loc.pattern,
                                                           short_reads=False)
    assert long_words == {loc.left[-80:], loc.right[:80]}


# ------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def ctx():
    from advntr_b200 import engine
    c = engine.Context(device=0)
    yield c
    c.close()


@pytest.mark.gpu
def test_device_filter_reproduces_reference_binary_stdout(ctx):
    from advntr_b200 import keyword_filter
    kw, names, seqs = synth.kfilter_case()
    want = _golden_stdout()
    kf = keyword_filter.KeywordFilter(kw, ctx=ctx)
    for mm in (5, 2):
        assert kf.format_output(names, seqs, min_matches=mm) == want[str(mm)]
    kf.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,lengths", [(1, (15,)), (2, (15,)), (3, (7, 15, 21)), (4, (5,)), (5, (15, 30, 80)),
                                          (6, (3, 4)), (7, (1, 2)), (8, (21, 22, 43, 64, 65))])
def test_device_filter_matches_oracle_on_random_cases(ctx, seed, lengths):
    from advntr_b200 import keyword_filter
    kw, names, seqs = _random_case(seed, lengths)
    kf = keyword_filter.KeywordFilter(kw, ctx=ctx)
    for mm in (1, 3, 5):
        assert kf.format_output(names, seqs, min_matches=mm) == \
            kfilter_oracle.filter_output(kw, names, seqs, min_matches=mm)
    kf.close()


@pytest.mark.gpu
def test_device_filter_edge_cases(ctx):
    from advntr_b200 import keyword_filter
    kw = [(1, ["ACGTACGTACGTACG"]), (2, ["ACGTACGTACGTACG", "TTTTTTTTTTTTTTT"]), (3, [])]
    kf = keyword_filter.KeywordFilter(kw, ctx=ctx)
    # no reads at all; reads shorter than the keywords; empty read; homopolymer with overlapping hits
    assert kf.format_output([], [], min_matches=1) == kfilter_oracle.filter_output(kw, [], [], min_matches=1)
    names = ["a", "b", "c", "d"]
    seqs = ["ACGT", "", "T" * 40, "ACGTACGTACGTACGTACGTACG"]
    for mm in (1, 2, 5, 26, 27):
        assert kf.format_output(names, seqs, min_matches=mm) == \
            kfilter_oracle.filter_output(kw, names, seqs, min_matches=mm)
    kf.close()
    # a filter without any keyword
    kf = keyword_filter.KeywordFilter([(9, [])], ctx=ctx)
    assert kf.format_output(names, seqs, min_matches=1) == "9 0\n"
    kf.close()


@pytest.mark.gpu
def test_device_filter_long_reads_and_many_hits_per_read(ctx):
    """PacBio-like reads against many loci: every read hits dozens of loci (exercises the
    occurrence-counter growth path) -- counts compared pair by pair with the oracle's."""
    from advntr_b200 import keyword_filter
    rng = random.Random(5)
    genome = synth.rand_dna(rng, 20000)
    kw = [(v, [genome[p:p + 15] for p in range(v * 60, v * 60 + 60, 5)]) for v in range(300)]
    names = ["long%d" % i for i in range(24)]
    seqs = []
    for i in range(24):
        a = rng.randrange(0, 8000)
        seqs.append(genome[a:a + rng.randrange(3000, 12000)])
    kf = keyword_filter.KeywordFilter(kw, ctx=ctx)
    assert kf.format_output(names, seqs, min_matches=5) == kfilter_oracle.filter_output(kw, names, seqs, 5)
    kf.close()


@pytest.mark.gpu
def test_device_filter_per_locus_cap(ctx):
    """More than 3 x max_reads accepted reads for one locus: the reference stops accepting in file
    order (main.cc:283) and lists max_reads + 1 names (main.cc:318-322)."""
    from advntr_b200 import keyword_filter
    kw = [(4, ["ACGTTGCAAGGCTTA", "GGGGGCCCCCAAAAA"]), (2, ["ACGTTGCAAGGCTTA"])]
    rng = random.Random(9)
    names, seqs = [], []
    for i in range(80):
        names.append("q%02d" % i)
        seqs.append(synth.rand_dna(rng, 10) + "ACGTTGCAAGGCTTA" * rng.randint(1, 3) + synth.rand_dna(rng, 5))
    kf = keyword_filter.KeywordFilter(kw, ctx=ctx)
    per_locus, reads = kf.filter_reads(names, seqs, min_matches=1, max_reads=7)
    assert kf.format_output(names, seqs, min_matches=1, max_reads=7) == \
        kfilter_oracle.filter_output(kw, names, seqs, min_matches=1, max_reads=7)
    assert [len(shown) for _, _, shown in per_locus] == [8, 8] and [c for _, c, _ in per_locus] == [7, 7]
    kf.close()


@pytest.mark.gpu
def test_device_resident_scan_equals_host_scan(ctx):
    """ADVHMM_DEVICE_BUFFERS (+ ADVHMM_DEVICE_OFFSETS): reads, offsets and the triples stay on the device."""
    import torch
    from advntr_b200 import keyword_filter
    kw, names, seqs = synth.kfilter_case(n_loci=40, reads_per_locus=20, decoys=3000, seed=5)
    kf = keyword_filter.KeywordFilter(kw, ctx=ctx)
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    flat = np.frombuffer("".join(seqs).encode("ascii"), dtype=np.uint8)
    hr, hl, hc = kf.filter.scan(flat, off, 3)
    want = sorted(zip(hr.tolist(), hl.tolist(), hc.tolist()))
    assert len(want) > 100
    d_seqs = torch.zeros(len(flat) + 16, dtype=torch.uint8, device="cuda")
    d_seqs[:len(flat)] = torch.from_numpy(flat.copy())
    d_off = torch.from_numpy(off).cuda()
    cap = 4 * len(want)
    d_r, d_l, d_c = (torch.empty(cap, dtype=torch.int32, device="cuda") for _ in range(3))
    d_n = torch.zeros(1, dtype=torch.int64, device="cuda")
    for offsets in (off, d_off.data_ptr()):
        d_n.zero_()
        kf.filter.scan_device(d_seqs.data_ptr(), offsets, len(seqs), 3, d_r.data_ptr(), d_l.data_ptr(),
                              d_c.data_ptr(), cap, d_n.data_ptr())
        ctx.synchronize()
        n = int(d_n.item())
        got = sorted(zip(d_r[:n].tolist(), d_l[:n].tolist(), d_c[:n].tolist()))
        assert got == want
    kf.close()
