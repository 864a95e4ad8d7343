"""Synthetic loci and reads of the shapes BASELINE.json names.

There is no network and no real data set; every benchmark and parity test runs on these
generators.  Everything is driven by ``random.Random(seed)`` so the same inputs are
produced in this container and on the GPU box.
"""
from __future__ import annotations

import random

from . import read_matcher

_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def rand_dna(rng, n):
    return "".join(rng.choice("ACGT") for _ in range(n))


def revcomp(s):
    return "".join(_COMP[c] for c in reversed(s))


def substitute(rng, s, rate):
    return "".join(rng.choice("ACGT") if rng.random() < rate else c for c in s)


def sequencing_errors(rng, s, sub, ins, dele):
    out = []
    for c in s:
        x = rng.random()
        if x < dele:
            continue
        if x < dele + ins:
            out.append(c)
            out.append(rng.choice("ACGT"))
        elif x < dele + ins + sub:
            out.append(rng.choice("ACGT"))
        else:
            out.append(c)
    return "".join(out)


class Locus(object):
    """A synthetic VNTR locus: flanks, repeat segments of the reference allele, read model."""

    def __init__(self, locus_id, left, right, segments, read_length=150, flank=150, error_rate=0.05):
        self.id = locus_id
        self.left, self.right, self.segments = left, right, list(segments)
        self.pattern = segments[0]
        self.read_length = read_length
        self.flank = flank
        self.error_rate = error_rate
        self.copies = read_matcher.copies_for_read_length(read_length, len(self.pattern))

    @property
    def sequence(self):
        return self.left + "".join(self.segments) + self.right

    def build_model(self):
        """vntr_finder.py:117-138 -> get_read_matcher_model on the trimmed flanks."""
        return read_matcher.build_vntr_matcher_hmm(self.left, self.right, self.segments, self.copies,
                                                   flank_size=self.flank, error_rate=self.error_rate)

    def reads(self, rng, n, sub=0.01, ins=0.001, dele=0.001, length=None):
        """Uniform windows of the locus with Illumina-like errors, trimmed to the read length."""
        length = length or self.read_length
        seq = self.sequence
        out = []
        for _ in range(n):
            s = rng.randrange(0, max(1, len(seq) - length))
            out.append(sequencing_errors(rng, seq[s:s + length + 8], sub, ins, dele)[:length])
        return out


def config1_locus():
    """BASELINE config 1: 30 bp RU x 10 identical copies, 100 bp flanks."""
    rng = random.Random(1)
    ru = rand_dna(rng, 30)
    left = rand_dna(rng, 100)
    right = rand_dna(rng, 100)
    return Locus(1, left, right, [ru] * 10, read_length=150, flank=150)


def config1_reads(n=1000, seed=11):
    loc = config1_locus()
    return loc.reads(random.Random(seed), n)


def config2_locus(locus_id, read_length=150):
    """BASELINE config 2: loci shaped like the recommended hg19 Illumina set (section 8d):
    RU length 6..70, total VNTR length < 140 bp, 2 % divergence between copies, 500 bp flanks
    of which the 150 nearest bases enter the model."""
    rng = random.Random(1000003 * locus_id + 17)
    R = rng.randint(6, 70)
    ncopies = max(2, 139 // R)
    ru = rand_dna(rng, R)
    segments = [substitute(rng, ru, 0.02) for _ in range(ncopies)]
    left = rand_dna(rng, 500)
    right = rand_dna(rng, 500)
    return Locus(locus_id, left, right, segments, read_length=read_length, flank=150)


def config5_locus(locus_id, read_length=150):
    """BASELINE config 5 (the 158,522-locus genic set, README.md:31-35): the config-2 generator with
    repeat units up to 100 bp and VNTRs up to 1 kb in the reference."""
    rng = random.Random(1000003 * locus_id + 29)
    R = rng.randint(6, 100)
    total = rng.randint(2 * R, max(2 * R, 1000))
    ncopies = max(2, total // R)
    ru = rand_dna(rng, R)
    segments = [substitute(rng, ru, 0.02) for _ in range(ncopies)]
    left = rand_dna(rng, 500)
    right = rand_dna(rng, 500)
    return Locus(locus_id, left, right, segments, read_length=read_length, flank=150)


def locus_cost_estimate(locus_id, generator="config2", read_length=150, coverage=30, decoys=50, flank=150):
    """(reads, DP cells) of a synthetic locus from the FIRST draws of its generator only -- repeat-unit
    length and copy number -- without making its sequences, reads or model: the reads
    config2_read_codes will make times the state count m = 3 L_l + 3 L_r + C (3 R + 3) + 18 (SURVEY.md
    section 8).  Cheap enough for every rank to balance all loci of a run by itself."""
    if generator == "config5":
        rng = random.Random(1000003 * locus_id + 29)
        R = rng.randint(6, 100)
        ncopies = max(2, rng.randint(2 * R, max(2 * R, 1000)) // R)
    else:
        rng = random.Random(1000003 * locus_id + 17)
        R = rng.randint(6, 70)
        ncopies = max(2, 139 // R)
    L = read_length
    n_reads = max(1, int(round((R * ncopies + L) * coverage / float(L)))) + 2 * decoys
    C = read_matcher.copies_for_read_length(L, R)
    m = 6 * flank + C * (3 * R + 3) + 18
    return n_reads, n_reads * L * m


def config2_reads(locus, coverage=30, decoys=50, seed=None):
    """Mapped reads overlapping the VNTR at `coverage`x (decoded on one strand) and decoy
    unmapped reads (decoded on both strands, vntr_finder.py:235-254)."""
    rng = random.Random(7919 * locus.id + 3 if seed is None else seed)
    vntr_len = sum(len(s) for s in locus.segments)
    L = locus.read_length
    n_mapped = max(1, int(round((vntr_len + L) * coverage / float(L))))
    lo = max(0, len(locus.left) - L + 1)
    hi = len(locus.left) + vntr_len - 1
    seq = locus.sequence
    mapped = []
    for _ in range(n_mapped):
        s = rng.randint(lo, max(lo, min(hi, len(seq) - L)))
        mapped.append(sequencing_errors(rng, seq[s:s + L + 8], 0.01, 0.001, 0.001)[:L])
    unmapped = [rand_dna(rng, L) for _ in range(decoys)]
    return mapped, unmapped


def config3_locus(copies=100, R=60, flank=100):
    """BASELINE config 3: long-RU VNTR for PacBio-like reads (error rate 0.3)."""
    rng = random.Random(3)
    ru = rand_dna(rng, R)
    loc = Locus(3, rand_dna(rng, flank), rand_dna(rng, flank), [ru] * copies, read_length=R * copies + 2 * flank,
                flank=flank, error_rate=0.3)
    loc.copies = copies
    return loc


# ------------------------------------------------------------------ bulk (numpy) read generators
import numpy as np

_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    _CODE[ord(_c)] = _i


def encode(seq):
    """ASCII DNA -> uint8 codes A,C,G,T = 0..3 (host-side helper for synthetic inputs)."""
    return _CODE[np.frombuffer(seq.encode("ascii"), dtype=np.uint8)]


def revcomp_codes(codes):
    return (3 - codes[::-1]).astype(np.uint8)


def config2_read_codes(locus, coverage=30, decoys=50, seed=None):
    """Bench-scale version of :func:`config2_reads`: returns ``(flat uint8 codes, lengths)`` for
    one locus -- mapped reads (one strand) followed by decoy reads, each decoy as forward and
    reverse complement (the both-strands call site, vntr_finder.py:239-246)."""
    rng = np.random.Generator(np.random.PCG64(7919 * locus.id + 3 if seed is None else seed))
    L = locus.read_length
    seq = encode(locus.sequence)
    vntr_len = sum(len(s) for s in locus.segments)
    n_mapped = max(1, int(round((vntr_len + L) * coverage / float(L))))
    lo = max(0, len(locus.left) - L + 1)
    hi = max(lo, min(len(locus.left) + vntr_len - 1, len(seq) - L - 8))
    starts = rng.integers(lo, hi + 1, size=n_mapped)
    win = seq[starts[:, None] + np.arange(L + 8)[None, :]]
    sub = rng.random(win.shape) < 0.01
    win = np.where(sub, rng.integers(0, 4, size=win.shape, dtype=np.uint8), win).astype(np.uint8)
    reads = []
    indel = rng.random(win.shape) < 0.002          # 0.1 % insertions + 0.1 % deletions
    for r in range(n_mapped):
        row = win[r]
        pos = np.nonzero(indel[r])[0]
        if len(pos):
            for p in pos[::-1]:
                if rng.random() < 0.5:
                    row = np.delete(row, p)
                else:
                    row = np.insert(row, p, rng.integers(0, 4))
        reads.append(row[:L])
    dec = rng.integers(0, 4, size=(decoys, L), dtype=np.uint8)
    for d in dec:
        reads.append(d)
        reads.append(revcomp_codes(d))
    lengths = np.fromiter((len(r) for r in reads), dtype=np.int64, count=len(reads))
    return np.concatenate(reads), lengths


def kfilter_case(n_loci=30, reads_per_locus=8, decoys=400, seed=77, keyword_size=15):
    """Input of the keyword pre-filter parity case: keywords of config-2 loci as
    ``genome_analyzer.py:181`` writes them (keyword_size=15) + unmapped reads: locus reads on either
    strand and random decoys, a few with N and with lower-case stretches.
    -> (keywords_by_locus, names, seqs)."""
    from . import keyword_filter
    rng = random.Random(seed)
    loci = [config2_locus(i) for i in range(1, n_loci + 1)]
    kw = [(l.id, sorted(keyword_filter.get_keywords_for_filtering(l.left, l.right, l.segments, l.pattern,
                                                                   keyword_size=keyword_size))) for l in loci]
    names, seqs = [], []
    for l in loci:
        for _ in range(reads_per_locus):
            s = rng.randrange(300, 560)
            r = sequencing_errors(rng, l.sequence[s:s + 158], 0.01, 0.001, 0.001)[:150]
            if rng.random() < 0.5:
                r = revcomp(r)
            if rng.random() < 0.1:
                p = rng.randrange(len(r))
                r = r[:p] + "N" + r[p + 1:]
            if rng.random() < 0.05:
                r = r[:40] + r[40:60].lower() + r[60:]
            names.append("r%d" % len(names)); seqs.append(r)
    for _ in range(decoys):
        names.append("r%d" % len(names)); seqs.append(rand_dna(rng, 150))
    return kw, names, seqs
