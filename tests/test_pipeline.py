"""Many loci in one go (advntr_b200/pipeline.py): keyword filter -> one batched device call with
on-device path reducers -> genotypes.  Every decision must equal what LocusDecoder (validated against
the reference's VNTRFinder in test_callsites.py / test_downstream.py) makes locus by locus from full
paths, and the synthetic diploid alleles must come out."""
import random

import pytest

import kfilter_oracle
from advntr_b200 import synth


def _sample(n_loci=14, seed=5):
    rng = random.Random(seed)
    loci, mapped, names, seqs, truth = [], {}, [], [], {}
    for lid in range(1, n_loci + 1):
        R = rng.choice((8, 12, 17, 24, 31, 40))
        ru = synth.rand_dna(rng, R)
        left, right = synth.rand_dna(rng, 300), synth.rand_dna(rng, 300)
        nref = max(2, 100 // R)
        a = rng.randint(2, max(2, 110 // R))
        b = a if rng.random() < 0.4 else rng.randint(2, max(2, 110 // R))
        truth[lid] = sorted((a, b))
        loci.append((lid, left, right, [ru] * nref))
        mapped[lid] = []
        for copies in (a, b):
            allele = left + ru * copies + right
            for _ in range(int(round((R * copies + 150) * 15 / 150.0))):
                s = rng.randrange(300 - 149, 300 + R * copies - 1)
                read = synth.sequencing_errors(rng, allele[s:s + 158], 0.004, 0.0003, 0.0003)[:150]
                if len(read) < 150:
                    continue
                if rng.random() < 0.2:                                   # unmapped, either strand
                    names.append("u%05d" % len(names))
                    seqs.append(synth.revcomp(read) if rng.random() < 0.5 else read)
                else:
                    mapped[lid].append(read)
    for _ in range(300):                                                 # decoys
        names.append("u%05d" % len(names))
        seqs.append(synth.rand_dna(rng, 150))
    order = list(range(len(names)))
    rng.shuffle(order)
    return loci, mapped, [names[i] for i in order], [seqs[i] for i in order], truth


@pytest.mark.gpu
def test_pipeline_equals_per_locus_decoders_and_recovers_alleles():
    from advntr_b200 import keyword_filter, locus_batch, pipeline
    loci, mapped, names, seqs, truth = _sample()
    run = pipeline.GenotypingRun([pipeline.LocusSpec(*l) for l in loci])
    # step 1: the device filter hands every locus the reads the reference binary would
    filtered = run.filter_unmapped(names, seqs)
    kw = [(lid, sorted(keyword_filter.get_keywords_for_filtering(left, right, segs, segs[0], keyword_size=15)))
          for lid, left, right, segs in loci]
    text = kfilter_oracle.filter_output(kw, names, seqs)
    listed = {}
    for line in text.split("\n"):
        tok = line.split()
        if len(tok) >= 2 and tok[0].isdigit() and tok[1].isdigit():
            listed[int(tok[0])] = set(tok[2:])
    assert {lid: {n for n, _ in rs} for lid, rs in filtered.items()} == listed
    assert sum(len(v) for v in listed.values()) > 20
    # steps 2 + 3 against the read-by-read decoders
    for acc in (False, True):
        got = run.genotype(mapped, names, seqs, accuracy_filter=acc)
        for lid, left, right, segs in loci:
            dec = locus_batch.LocusDecoder(left, right, segs, read_length=150, locus_id=lid)
            selected = dec.select_reads(mapped[lid], [s for _, s in filtered[lid]])
            want = dec.genotype(selected, accuracy_filter=acc)
            covered, flanking = dec.observed_repeats(selected, acc)
            g = got[lid]
            assert (g["covered_repeats"], g["flanking_repeats"]) == (covered, flanking), lid
            for key in ("copy_numbers", "recruited_reads_count", "spanning_reads_count", "flanking_reads_count",
                        "maximum_likelihood"):
                assert g[key] == want[key], (lid, key)
    plain = run.genotype(mapped, names, seqs)
    right_calls = sum(1 for lid in truth if plain[lid]["copy_numbers"] is not None and
                      sorted(plain[lid]["copy_numbers"]) == truth[lid])
    assert right_calls >= len(truth) - 2, (right_calls, {l: (plain[l]["copy_numbers"], truth[l]) for l in truth})
    run.close()


@pytest.mark.gpu
def test_frameshift_mode_equals_per_locus_python_consumers():
    """GenotypingRun.find_frameshifts (one device call with full paths + the native consumer) against the per-locus
    route a reference-shaped caller takes (LocusDecoder.select_reads -> frameshift_candidate -> identify_frameshift)."""
    import bench_workloads
    from advntr_b200 import genotype, locus_batch, path_utils, pipeline
    specs, reads = [], {}
    for lid in (2, 5, 9, 14, 23, 31, 40):
        loc, rs = bench_workloads.frameshift_locus_reads(lid, coverage=30)
        specs.append(pipeline.LocusSpec(lid, loc.left, loc.right, loc.segments))
        reads[lid] = rs
    run = pipeline.GenotypingRun(specs)
    got = run.find_frameshifts(reads)
    n_calls = 0
    for spec, rec in zip(specs, run.frameshift_records):
        dec = locus_batch.LocusDecoder(spec.left_flank, spec.right_flank, spec.repeat_segments, 150, locus_id=spec.id)
        selected = dec.select_reads(reads[spec.id])
        (label, count), repeat_bp = dec.frameshift_candidate(selected)
        assert (path_utils.frameshift_label(rec), int(rec["count"]), int(rec["repeat_bp"]), int(rec["selected"])) == \
            (label, count, repeat_bp, len(selected))
        coverage = float(repeat_bp) / sum(len(s) for s in spec.repeat_segments) / 2
        want = label if genotype.identify_frameshift(coverage, count, 1 / coverage) else None
        assert got[spec.id] == want
        n_calls += label is not None and count >= 3
    # (whether the binomial test then fires is the reference's business: scipy's binom.pmf is nan for the
    # non-integer coverage it is handed, vntr_finder.py:256-263, so most of these stay None there as well)
    assert n_calls >= 3
    run.close()
