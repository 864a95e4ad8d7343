#!/usr/bin/env python
"""Golden vector for a locus whose repeat segments have UNEQUAL lengths (tests/golden/aligned.npz).

The reference hands such segments to MUSCLE (profile_hmm.py:165-171) and builds the repeat-unit
profile from MUSCLE's alignment with ``build_profile_hmm_pseudocounts_for_alignment``
(profile_hmm.py:13-161).  MUSCLE is an external binary that is absent here, so its output cannot
be pinned; everything AFTER it can: this script lets the reference's own code path run --
``hmm_utils.get_read_matcher_model`` -> ``build_profile_hmm_for_repeats`` -> the Biopython MUSCLE
wrapper -- with the wrapper returning a FIXED gapped alignment of the segments (hand-made, stored in
the fixture), then decodes reads on the compiled reference engine.  What the fixture pins: insert
columns (>= 50 % gaps), delete walks, pseudocounts, the resulting tables (bit patterns), Viterbi
scores, paths, repeat counts for a caller-supplied alignment.

Run in the build container only (needs /root/reference and oracle/_ref):
    python tests/golden/make_golden_aligned.py
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


def inputs():
    rng = random.Random(808)
    ru = synth.rand_dna(rng, 17)
    left, right = synth.rand_dna(rng, 150), synth.rand_dna(rng, 150)
    # six copies of a 17 bp unit: substitutions, one copy with a 2 bp insertion, two with a deletion,
    # one with both; the alignment has 19 columns, two of them insert columns (gaps in >= half the rows)
    alignment = []
    for k in range(6):
        row = list(synth.substitute(rng, ru, 0.08))
        ins1 = rng.choice("ACGT") if k in (1, 4) else "-"
        ins2 = rng.choice("ACGT") if k == 1 else "-"
        row = row[:6] + [ins1, ins2] + row[6:]
        if k in (2, 4):
            row[11] = "-"
        if k == 3:
            row[2] = "-"
        alignment.append("".join(row))
    segments = [a.replace("-", "") for a in alignment]
    assert len(set(map(len, segments))) > 1
    locus = left + "".join(segments) + right
    reads = ["", "G", synth.rand_dna(rng, 150)]
    for _ in range(60):
        s = rng.randrange(0, len(locus) - 150)
        reads.append(synth.sequencing_errors(rng, locus[s:s + 158], 0.01, 0.002, 0.002)[:150])
    reads += [synth.revcomp(r) for r in reads[3:11]]
    copies = int(round(150.0 / len(segments[0]) + 0.5))          # vntr_finder.py:98-99 for 150 bp reads
    return left, right, segments, alignment, copies, 0.05, reads


def main():
    pom = refenv.reference_pomegranate()
    hu = refenv.reference_hmm_utils(pom, "ref")
    settings = refenv.reference_settings()
    left, right, segments, alignment, copies, eps, reads = inputs()

    class _Rec(object):
        def __init__(self, seq):
            self.seq = seq
    calls = []

    def fixed_alignment(handle, fmt):
        """AlignIO.read of MUSCLE's stdout: the fixed alignment, after checking that the reference asked for
        exactly these segments."""
        asked = [ln.strip() for ln in handle.read().splitlines() if ln.strip() and not ln.startswith(">")]
        assert asked == segments, "the reference asked MUSCLE for other segments"
        calls.append(1)
        return [_Rec(a) for a in alignment]
    alignio = sys.modules["Bio.AlignIO"]
    saved = alignio.read
    alignio.read = fixed_alignment
    try:
        settings.MAX_ERROR_RATE = eps
        model = hu.get_read_matcher_model(left, right, segments, copies=copies)
    finally:
        alignio.read = saved
        settings.MAX_ERROR_RATE = 0.05
    assert calls, "the MUSCLE path was not taken"
    b = oracle.baked_from_reference_model(model)
    logp, ru, paths, off = [], [], [], [0]
    for r in reads:
        lp, vp = model.viterbi(r)
        logp.append(lp)
        paths.extend(i for i, _ in vp)
        ru.append(hu.get_number_of_repeats_in_vpath(vp))
        off.append(len(paths))
    np.savez_compressed(
        os.path.join(HERE, "aligned.npz"),
        in_off=b["in_off"], in_src=b["in_src"], in_logp=b["in_logp"], emis=b["emis"],
        scalars=np.array([b["n_states"], b["silent_start"], b["start_index"], b["end_index"], b["finite"]]),
        names=np.array("\n".join(b["names"])),
        inputs=np.array(json.dumps({"left": left, "right": right, "segments": segments, "alignment": alignment,
                                    "copies": copies, "error_rate": eps, "reads": reads})),
        logp=np.array(logp, dtype=np.float64), forward=np.zeros(len(reads)), ru_count=np.array(ru, dtype=np.int32),
        paths=np.array(paths, dtype=np.int32), path_off=np.array(off, dtype=np.int64))
    print("aligned: states", b["n_states"], "edges", len(b["in_src"]), "reads", len(reads),
          "segment lengths", sorted(set(map(len, segments))), "alignment width", len(alignment[0]))


if __name__ == "__main__":
    main()
