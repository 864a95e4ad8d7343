#!/usr/bin/env python
"""Golden values for what happens DOWNSTREAM of the Viterbi path, FROM THE REFERENCE ITSELF
(run in the build container only: needs /root/reference and oracle/_ref):

    python oracle/build_ref.py && python tests/golden/make_golden_downstream.py

Writes tests/golden/downstream.json and tests/golden/update_model.npz:

  genotype      find_genotype_based_on_observed_repeats (vntr_finder.py:485-532) on the four count
                lists of the reference's tests/test_genotyping.py and on random count lists,
                diploid and haploid
  frameshift    identify_frameshift (vntr_finder.py:256-263) on the seven cases of the reference's
                tests/test_frameshift_identification.py and on random ones
  msa           get_multiple_alignment_of_viterbi_paths: the two fixtures of tests/test_hmm_utils.py
  loci          Illumina genotyping of three diploid synthetic loci through the reference's own
                find_repeat_count_from_alignment_file (read IO replaced by the read lists): selected
                reads, spanning / flanking repeat counts, genotype, max_prob
  pacbio        get_dominant_copy_numbers_from_spanning_reads (vntr_finder.py:534-585) on spanning
                reads of a diploid long locus: observed repeat counts, genotype
  segmentation  build_reference_repeat_finder_hmm (merge='All' bake) + find_repeat_segments on a
                reference region (segmentation.npz)
  update_model  get_read_matcher_model(..., vpaths) (--update, hmm_utils.py:427-429, 553-595): the
                baked tables of the re-estimated model (update_model.npz) and reads decoded on it
"""
import json
import logging
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


def fake_vntr(left, right, segs, vid=7):
    class FakeVNTR(object):
        id, pattern, chromosome, start_point, scaled_score = vid, segs[0], "chr1", 1000, None
        left_flanking_region, right_flanking_region = left, right

        def get_repeat_segments(self):
            return segs

        def get_length(self):
            return sum(len(x) for x in segs)
    return FakeVNTR()


def diploid_locus_inputs(seed, R, copies_a, copies_b, n_ref_copies, coverage=30, decoys=6):
    """A locus whose sample carries two alleles; 150 bp reads from both, plus reverse-strand and
    decoy unmapped reads."""
    rng = random.Random(seed)
    ru = synth.rand_dna(rng, R)
    left, right = synth.rand_dna(rng, 300), synth.rand_dna(rng, 300)
    segs = [ru] * n_ref_copies
    mapped, unmapped = [], []
    for copies in (copies_a, copies_b):
        allele = left + ru * copies + right
        n = int(round((R * copies + 150) * coverage / 2 / 150.0))
        for _ in range(n):
            s = rng.randrange(300 - 149, 300 + R * copies - 1)
            read = synth.sequencing_errors(rng, allele[s:s + 158], 0.005, 0.0005, 0.0005)[:150]
            if len(read) < 150:
                continue
            (unmapped if rng.random() < 0.15 else mapped).append(read)
    unmapped = [synth.revcomp(r) if i % 2 else r for i, r in enumerate(unmapped)]
    unmapped += [synth.rand_dna(rng, 150) for _ in range(decoys)]
    return left, right, segs, mapped, unmapped


def illumina_locus_case(vf, seed, R, a, b, nref):
    left, right, segs, mapped, unmapped = diploid_locus_inputs(seed, R, a, b, nref)
    finder = vf.VNTRFinder(fake_vntr(left, right, segs))
    hmm = finder.get_vntr_matcher_hmm(read_length=150)
    score = finder.get_min_score_to_select_a_read(150)
    selected = []
    for r in mapped:                                       # vntr_finder.py:736-748
        logp, vpath = hmm.viterbi(r)
        if finder.recruit_read(logp, vpath, score, r):
            selected.append(vf.SelectedRead(sequence=r, logp=logp, vpath=vpath, reference_start=1))

    class Acc(object):
        value = 0.0
    for r in unmapped:                                     # vntr_finder.py:757-764
        finder.process_unmapped_read(None, r, hmm, score, Acc(), selected)
    finder.select_illumina_reads = lambda *a, **k: selected    # the read IO of :701-773, done above
    seen = {}

    class Grab(logging.Handler):
        def emit(self, record):
            msg = record.getMessage()
            for key in ("covered repeats: ", "flanking repeats: "):
                if msg.startswith(key):
                    seen[key.split()[0]] = json.loads(msg[len(key):])
    grab = Grab()
    logging.getLogger().addHandler(grab)
    logging.getLogger().setLevel(logging.INFO)
    out = {}
    for acc_filter in (False, True):
        res = finder.find_repeat_count_from_alignment_file(None, None, acc_filter)
        out["accuracy_filter" if acc_filter else "plain"] = {
            "copy_numbers": list(res.copy_numbers) if res.copy_numbers is not None else None,
            "max_prob": float(res.maximum_likelihood), "recruited": res.recruited_reads_count,
            "spanning": res.spanning_reads_count, "flanking": res.flanking_reads_count,
            "covered_repeats": seen["covered"], "flanking_repeats": seen["flanking"]}
    logging.getLogger().removeHandler(grab)
    return {"left": left, "right": right, "segments": segs, "mapped": mapped, "unmapped": unmapped,
            "alleles": [a, b], "selected_sequences": [s.sequence for s in selected],
            "covered_repeats": out["plain"]["covered_repeats"], "flanking_repeats": out["plain"]["flanking_repeats"],
            "result": out}, selected, finder


def pacbio_case(vf, settings):
    """Spanning reads of a diploid long locus, CLR-like errors; MAX_ERROR_RATE 0.3 as the pacbio
    command sets it (advntr_commands.py:66-69)."""
    rng = random.Random(77)
    R, a, b = 14, 17, 21
    ru = synth.rand_dna(rng, R)
    left, right = synth.rand_dna(rng, 200), synth.rand_dna(rng, 200)
    segs = [synth.substitute(rng, ru, 0.03) for _ in range(12)]
    segs[0] = ru

    class Read(object):
        def __init__(self, seq, rid):
            self.sequence, self.read_id = seq, rid
            self.source = type("Src", (), {"name": "synthetic"})()
    reads = []
    for copies in (a, b):
        allele = left[-100:] + ru * copies + right[:100]
        for _ in range(9):
            reads.append(Read(synth.sequencing_errors(rng, allele, 0.02, 0.04, 0.04), "r%d" % len(reads)))
    saved = settings.MAX_ERROR_RATE
    settings.MAX_ERROR_RATE = 0.3
    try:
        finder = vf.VNTRFinder(fake_vntr(left, right, segs))
        seen = {}

        class Grab(logging.Handler):
            def emit(self, record):
                msg = record.getMessage()
                if msg.startswith("observed repeats: "):
                    seen["observed"] = json.loads(msg[len("observed repeats: "):])
        grab = Grab()
        logging.getLogger().addHandler(grab)
        logging.getLogger().setLevel(logging.INFO)
        out = {}
        for acc_filter in (False, True):
            finder = vf.VNTRFinder(fake_vntr(left, right, segs))
            cn, max_prob = finder.get_dominant_copy_numbers_from_spanning_reads(reads, False, acc_filter)
            out["accuracy_filter" if acc_filter else "plain"] = {
                "copy_numbers": list(cn) if cn is not None else None, "max_prob": float(max_prob)}
        logging.getLogger().removeHandler(grab)
    finally:
        settings.MAX_ERROR_RATE = saved
    return {"left": left, "right": right, "segments": segs, "reads": [r.sequence for r in reads],
            "alleles": [a, b], "error_rate": 0.3, "observed_repeats": seen["observed"], "result": out}


def update_model_case(pom, hu, locus_case, selected, finder):
    """One --update iteration (vntr_finder.py:667-698): profile re-estimated from the selected reads
    and the reference repeats, then the mapped reads decoded against the new model."""
    left, right, segs = locus_case["left"], locus_case["right"], locus_case["segments"]
    hmm = finder.get_vntr_matcher_hmm(read_length=150)
    ref_repeats = []
    for seg in segs:
        logp, vpath = hmm.viterbi(seg.upper())
        ref_repeats.append((seg.upper(), vpath))
    vpaths = [(s.sequence, s.vpath) for s in selected] + ref_repeats
    copies = finder.get_copies_for_hmm(150)
    alignment = hu.get_multiple_alignment_of_repeats_from_reads(vpaths)
    new = hu.get_read_matcher_model(left[-150:], right[:150], None, copies, vpaths)
    b = oracle.baked_from_reference_model(new)
    reads = locus_case["mapped"][:25]
    logp, paths, off = [], [], [0]
    for r in reads:
        lp, vp = new.viterbi(r)
        logp.append(lp)
        paths.extend(i for i, _ in vp)
        off.append(len(paths))
    np.savez_compressed(
        os.path.join(HERE, "update_model.npz"),
        in_off=b["in_off"], in_src=b["in_src"], in_logp=b["in_logp"], emis=b["emis"],
        scalars=np.array([b["n_states"], b["silent_start"], b["start_index"], b["end_index"], b["finite"]]),
        names=np.array("\n".join(b["names"])),
        inputs=np.array(json.dumps({"left": left, "right": right, "segments": segs, "copies": copies,
                                    "selected_sequences": [s.sequence for s in selected],
                                    "alignment": alignment, "reads": reads})),
        logp=np.array(logp, dtype=np.float64), paths=np.array(paths, dtype=np.int32),
        path_off=np.array(off, dtype=np.int64))
    print("update_model: %d vpaths, alignment %d rows x %d cols, %d states" % (
        len(vpaths), len(alignment), len(alignment[0]), b["n_states"]))


def main():
    pom = refenv.reference_pomegranate()
    hu = refenv.reference_hmm_utils(pom, "ref")
    settings = refenv.reference_settings()
    vf = refenv.reference_vntr_finder(pom, "ref")
    settings.MAX_ERROR_RATE = 0.05
    out = {}

    # ---- genotype statistics -----------------------------------------------------------------
    rng = random.Random(2024)
    lists = [[3, 3, 3, 3, 3], [2, 3, 3, 3, 3], [2, 2, 3, 3, 3], [4, 5, 5, 5, 7, 8, 8, 8, 9], [], [6], [0, 0, 4]]
    for _ in range(300):
        n = rng.randint(1, 40)
        a, b = rng.randint(1, 30), rng.randint(1, 30)
        lists.append([max(0, rng.choice((a, b)) + (rng.choice((-2, -1, 1, 2)) if rng.random() < 0.15 else 0))
                      for _ in range(n)])
    geno = []
    for haploid in (False, True):
        finder = vf.VNTRFinder(fake_vntr("A" * 20, "C" * 20, ["CACA"]), is_haploid=haploid)
        for obs in lists:
            res, p = finder.find_genotype_based_on_observed_repeats(list(obs))
            geno.append({"observed": obs, "haploid": haploid,
                         "genotype": list(res) if res is not None else None, "max_prob": float(p)})
    out["genotype"] = geno

    # ---- frameshift binomial test ------------------------------------------------------------
    finder = vf.VNTRFinder(fake_vntr("A" * 20, "C" * 20, ["CACA"]))
    fs = [(20, 10, 0.5), (20, 1, 0.5), (40, 17, 0.5), (40, 3, 0.5), (100, 42, 0.5), (100, 9, 0.5), (10, 10, 0.5)]
    for _ in range(200):
        cov = rng.randint(1, 200)
        fs.append((cov, rng.randint(0, cov), rng.choice((0.5, 0.25, 1.0 / 3))))
    out["frameshift"] = [{"coverage": c, "observed": o, "expected": e,
                          "result": bool(finder.identify_frameshift(c, o, e))} for c, o, e in fs]

    # ---- the reference's own MSA fixtures ----------------------------------------------------
    with open(os.path.join(refenv.REF_ROOT, "tests", "data", "hmm_utils.json")) as fh:
        fx = json.load(fh)
    out["msa"] = {"real_data_alignment": fx["alignment"],
                  "two_sequences": {"repeats": ["ACTTA", "ATTGA"],
                                    "states": [["M1", "M2", "M3", "M4", "M5"], ["M1", "D2", "M3", "M4", "I4", "M5"]],
                                    "alignment": ["ACTT-A", "A-TTGA"]}}

    # ---- Illumina loci, end to end -----------------------------------------------------------
    loci = []
    keep = None
    for seed, R, a, b, nref in ((11, 20, 3, 5, 4), (12, 12, 6, 6, 6), (13, 35, 2, 3, 2)):
        case, selected, finder = illumina_locus_case(vf, seed, R, a, b, nref)
        loci.append(case)
        print("locus R=%d alleles %d/%d: %d selected, covered %s, flanking %s -> %s" % (
            R, a, b, len(selected), case["covered_repeats"], case["flanking_repeats"], case["result"]))
        if keep is None:
            keep = (case, selected, finder)
    out["loci"] = loci

    # ---- --update ----------------------------------------------------------------------------
    update_model_case(pom, hu, *keep)

    # ---- PacBio ------------------------------------------------------------------------------
    out["pacbio"] = pacbio_case(vf, settings)
    print("pacbio: observed", out["pacbio"]["observed_repeats"], "->", out["pacbio"]["result"])

    # ---- reference segmentation (addmodel: reference_vntr.py:80-87, hmm_utils.py:598-680) -----
    rng = random.Random(31)
    pattern = synth.rand_dna(rng, 17)
    units = [synth.substitute(rng, pattern, 0.06) for _ in range(6)]
    units[2] = units[2][:5] + units[2][6:]                  # one unit with a deletion
    units[4] = units[4][:9] + "G" + units[4][9:]            # ... and one with an insertion
    region = "".join(units)
    model = hu.build_reference_repeat_finder_hmm([pattern], copies=6)
    b = oracle.baked_from_reference_model(model)
    lp, vp = model.viterbi(region)
    visited = [st.name for _, st in vp[1:-1]]
    segments = hu.get_repeat_segments_from_visited_states_and_region(visited, region)
    np.savez_compressed(
        os.path.join(HERE, "segmentation.npz"),
        in_off=b["in_off"], in_src=b["in_src"], in_logp=b["in_logp"], emis=b["emis"],
        scalars=np.array([b["n_states"], b["silent_start"], b["start_index"], b["end_index"], b["finite"]]),
        names=np.array("\n".join(b["names"])),
        inputs=np.array(json.dumps({"pattern": pattern, "copies": 6, "region": region, "units": units,
                                    "segments": segments})),
        logp=np.array([lp], dtype=np.float64), paths=np.array([i for i, _ in vp], dtype=np.int32))
    print("segmentation:", segments == units, segments)

    with open(os.path.join(HERE, "downstream.json"), "w") as fh:
        json.dump(out, fh)


if __name__ == "__main__":
    main()
