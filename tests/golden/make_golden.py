#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference and oracle/_ref):

    python oracle/build_ref.py && python tests/golden/make_golden.py

For every case the model is built by the reference's own ``advntr/hmm_utils.py``
(``get_read_matcher_model``) on the compiled, unmodified vendored pomegranate, and decoded
by that engine's ``viterbi`` / ``log_probability``.  Stored per case (``<case>.npz``):

  baked arrays   in_off, in_src, in_logp, emis, names, scalars   (what bake() produced)
  inputs         left, right, segments, copies, error_rate, reads
  outputs        logp[R] (bit patterns), paths (concatenated state indices) + path_off,
                 forward[R], ru_count[R] (hmm_utils.get_number_of_repeats_in_vpath)

plus ``fingerprint.npz`` (numpy.exp / libm log of a fixed vector: the builder round-trips
through both, so bit-exact builder parity is only expected where they reproduce).
"""
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)

import refenv   # noqa: E402
import oracle   # noqa: E402
from advntr_b200 import synth   # noqa: E402  (input generators only)


def case_inputs(name):
    if name == "config1":
        loc = synth.config1_locus()
        rng = random.Random(11)
        reads = [loc.left[-30:] + loc.pattern * 3 + loc.right[:30], "", "A", synth.rand_dna(rng, 150),
                 loc.pattern * 5, loc.left[-100:] + loc.pattern[:20], loc.pattern[10:] + loc.right[:120],
                 "ACGT" * 37, "A" * 150, loc.left[-100:] + loc.right[:50], synth.rand_dna(rng, 7)]
        reads += loc.reads(rng, 85)
        reads += [synth.revcomp(r) for r in reads[11:27]]
        reads += [loc.reads(rng, 1, length=L)[0] for L in (1, 2, 31, 32, 33, 64, 100, 149, 151, 160, 200, 250)]
        return loc.left[-150:], loc.right[:150], loc.segments, loc.copies, 0.05, reads
    rng = random.Random({"small_a": 2, "small_b": 3, "divergent": 4}[name])
    if name == "small_a":
        R, nseg, Ll, Lr, copies, eps = 12, 5, 60, 70, 4, 0.05
    elif name == "small_b":
        R, nseg, Ll, Lr, copies, eps = 7, 8, 40, 33, 9, 0.3
    else:
        R, nseg, Ll, Lr, copies, eps = 23, 6, 150, 150, 7, 0.05
    ru = synth.rand_dna(rng, R)
    left, right = synth.rand_dna(rng, Ll), synth.rand_dna(rng, Lr)
    segs = [synth.substitute(rng, ru, 0.1) for _ in range(nseg)]
    locus = left + "".join(segs) + right
    reads = ["", "C", synth.rand_dna(rng, 50)]
    for _ in range(45):
        L = rng.choice((20, 47, 80, 100, 150))
        s = rng.randrange(0, max(1, len(locus) - L))
        reads.append(synth.sequencing_errors(rng, locus[s:s + L + 8], 0.02, 0.01, 0.01)[:L])
    return left, right, segs, copies, eps, reads


def callsite_inputs():
    """Config-4 style inputs: a coding-VNTR-like locus, mapped reads of which half carry a 1 bp
    deletion inside one repeat unit (a frameshift), and unmapped reads (reverse-strand reads of
    the locus + random decoys)."""
    rng = random.Random(404)
    ru = synth.rand_dna(rng, 30)
    left, right = synth.rand_dna(rng, 200), synth.rand_dna(rng, 200)
    segs = [ru] * 6
    locus = left + "".join(segs) + right
    shifted = left + ru * 2 + ru[:13] + ru[14:] + ru * 3 + right      # 1 bp deleted in unit 3
    mapped = []
    for i in range(60):
        src = shifted if i % 2 else locus
        s = rng.randrange(120, 170)
        mapped.append(synth.sequencing_errors(rng, src[s:s + 158], 0.005, 0.0, 0.0)[:150])
    unmapped = [synth.revcomp(locus[s:s + 150]) for s in range(125, 165, 4)]
    unmapped += [synth.rand_dna(rng, 150) for _ in range(10)]
    return left, right, segs, mapped, unmapped


def make_callsite_case(pom):
    """Decisions of the reference's own VNTRFinder (vntr_finder.py) on those reads: which mapped
    reads recruit_read() accepts, what process_unmapped_read() selects, repeat counts of
    spanning / flanking reads, and the frameshift candidate + its count."""
    vf = refenv.reference_vntr_finder(pom, "ref")
    left, right, segs, mapped, unmapped = callsite_inputs()

    class FakeVNTR(object):
        id, pattern, chromosome, start_point, scaled_score = 4, segs[0], "chr1", 1000, None
        left_flanking_region, right_flanking_region = left, right

        def get_repeat_segments(self):
            return segs

        def get_length(self):
            return sum(len(x) for x in segs)

    finder = vf.VNTRFinder(FakeVNTR())
    hmm = finder.get_vntr_matcher_hmm(read_length=150)
    score = finder.get_min_score_to_select_a_read(150)
    selected, mapped_ok = [], []
    for r in mapped:                                       # vntr_finder.py:736-748
        logp, vpath = hmm.viterbi(r)
        ok = bool(finder.recruit_read(logp, vpath, score, r))
        mapped_ok.append(ok)
        if ok:
            selected.append(vf.SelectedRead(sequence=r, logp=logp, vpath=vpath, reference_start=1))
    n_mapped_selected = len(selected)

    class Acc(object):
        value = 0.0
    for r in unmapped:                                     # vntr_finder.py:757-764
        finder.process_unmapped_read(None, r, hmm, score, Acc(), selected)
    covered, flanking = [], []
    for sr in selected:                                    # vntr_finder.py:810-847
        n = vf.get_number_of_repeats_in_vpath(sr.vpath)
        if finder.read_flanks_repeats_with_confidence(sr.vpath, sr.sequence):
            covered.append(n)
        else:
            flanking.append(n)
    seen = {}
    finder.identify_frameshift = lambda cov, cnt, exp, error_rate=0.01: seen.update(cov=cov, cnt=cnt) or False
    import logging

    class Grab(logging.Handler):
        def emit(self, record):
            msg = record.getMessage()
            if msg.startswith("Frameshift Candidate and Occurrence"):
                seen["candidate"] = msg.split("Occurrence ")[1].split(":")[0]
    grab = Grab()
    logging.getLogger().addHandler(grab)
    logging.getLogger().setLevel(logging.INFO)
    finder.find_frameshift_from_selected_reads(selected)
    logging.getLogger().removeHandler(grab)
    real = vf.VNTRFinder(FakeVNTR())
    with open(os.path.join(HERE, "callsite_frameshift.json"), "w") as fh:
        json.dump({"left": left, "right": right, "segments": segs, "mapped": mapped, "unmapped": unmapped,
                   "mapped_recruited": mapped_ok,
                   "selected_sequences": [sr.sequence for sr in selected],
                   "selected_logp": [sr.logp for sr in selected],
                   "n_mapped_selected": n_mapped_selected,
                   "covered_repeats": covered, "flanking_repeats": sorted(flanking),
                   "frameshift_candidate": seen.get("candidate"),
                   "frameshift_count": seen["cnt"], "avg_bp_coverage": seen["cov"],
                   "frameshift_result": real.find_frameshift_from_selected_reads(selected)}, fh)
    print("callsite: %d/%d mapped recruited, %d selected in total, covered %s, frameshift %r x%d" % (
        sum(mapped_ok), len(mapped), len(selected), covered[:8], real.find_frameshift_from_selected_reads(selected),
        seen["cnt"]))


def make_kfilter_case():
    """stdout of the reference's adVNTR-Filtering binary (filtering/main.cc, compiled unmodified)."""
    import subprocess, tempfile, gzip
    import build_ref
    assert build_ref.build_filter()
    kw, names, seqs = synth.kfilter_case()
    with tempfile.TemporaryDirectory() as d:
        fa, kf = os.path.join(d, "reads.fa"), os.path.join(d, "kw.txt")
        with open(fa, "w") as fh:
            for n, s in zip(names, seqs):
                fh.write(">%s\n%s\n" % (n, s))
        with open(kf, "w") as fh:
            for vid, words in kw:
                fh.write("%d %s\n" % (vid, " ".join(words)))
        out = {}
        for mm in (5, 2):
            res = subprocess.run([build_ref.FILTER_BIN, fa, "--min_matches", str(mm)], stdin=open(kf),
                                 capture_output=True, text=True, check=True)
            out[str(mm)] = res.stdout
    with gzip.open(os.path.join(HERE, "kfilter_reference_stdout.json.gz"), "wt") as fh:
        json.dump(out, fh)
    print("kfilter: %d loci, %d reads, %d output lines (min 5), %d (min 2)" % (
        len(kw), len(names), out["5"].count("\n"), out["2"].count("\n")))


def main():
    make_kfilter_case()
    pom = refenv.reference_pomegranate()
    hu = refenv.reference_hmm_utils(pom, "ref")
    settings = refenv.reference_settings()
    for name in ("config1", "small_a", "small_b", "divergent"):
        left, right, segs, copies, eps, reads = case_inputs(name)
        settings.MAX_ERROR_RATE = eps
        model = hu.get_read_matcher_model(left, right, segs, copies=copies)
        settings.MAX_ERROR_RATE = 0.05
        b = oracle.baked_from_reference_model(model)
        logp, fwd, ru, paths, off, consumers = [], [], [], [], [0], []
        for r in reads:
            lp, vp = model.viterbi(r)
            logp.append(lp)
            fwd.append(model.log_probability(r))
            if vp is None:
                ru.append(-1)
                consumers.append([0, 0, 0, 0, 0.0])
            else:
                paths.extend(i for i, _ in vp)
                ru.append(hu.get_number_of_repeats_in_vpath(vp))
                # the other path consumers of hmm_utils.py:191-286 (rate needs a non-empty read)
                consumers.append([hu.get_number_of_matches_in_vpath(vp),
                                  hu.get_number_of_repeat_bp_matches_in_vpath(vp),
                                  hu.get_left_flanking_region_size_in_vpath(vp),
                                  hu.get_right_flanking_region_size_in_vpath(vp),
                                  hu.get_flanking_regions_matching_rate(vp, r, left, right) if r else -1.0])
            off.append(len(paths))
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            in_off=b["in_off"], in_src=b["in_src"], in_logp=b["in_logp"], emis=b["emis"],
            scalars=np.array([b["n_states"], b["silent_start"], b["start_index"], b["end_index"], b["finite"]]),
            names=np.array("\n".join(b["names"])),
            inputs=np.array(json.dumps({"left": left, "right": right, "segments": segs, "copies": copies,
                                        "error_rate": eps, "reads": reads})),
            logp=np.array(logp, dtype=np.float64), forward=np.array(fwd, dtype=np.float64),
            ru_count=np.array(ru, dtype=np.int32), consumers=np.array(consumers, dtype=np.float64),
            paths=np.array(paths, dtype=np.int32),
            path_off=np.array(off, dtype=np.int64))
        print(name, "states", b["n_states"], "edges", len(b["in_src"]), "reads", len(reads))
    make_callsite_case(pom)
    # the one fixture the reference's own tests hold for this path (tests/test_hmm_utils.py:15-17):
    # a recorded Viterbi path (state names) + read + the repeat segments it must yield
    with open(os.path.join(refenv.REF_ROOT, "tests", "data", "hmm_utils.json")) as fh:
        fx = json.load(fh)
    with open(os.path.join(HERE, "ref_tests_hmm_utils.json"), "w") as fh:
        json.dump({"provenance": "reference tests/data/hmm_utils.json (test fixture, copied verbatim)",
                   "visited_states": fx["visited_states"], "sequence": fx["sequence"],
                   "correct_repeats": fx["correct_repeats"]}, fh)
    x = np.linspace(-30.0, 0.0, 4097)
    import math
    np.savez_compressed(os.path.join(HERE, "fingerprint.npz"), x=x, exp=np.exp(x),
                        log=np.array([math.log(v) for v in np.exp(x)]))


if __name__ == "__main__":
    main()
