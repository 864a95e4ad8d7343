#!/usr/bin/env python
"""Golden values for the stored-model round trip (``vntr_finder.py:116-137``: ``to_json`` when a model
is first built, ``from_json`` on later runs), FROM THE REFERENCE ITSELF (build container only):

    python oracle/build_ref.py && python tests/golden/make_golden_json.py

Writes tests/golden/stored_model.npz:

  json     the reference's ``to_json()`` of a read matcher built by its ``hmm_utils`` (gzip'ed text)
  tables   the baked tables of the reference's ``from_json(json)`` -- baked with the DEFAULT
           merge='All' (``hmm.pyx:3143``), i.e. NOT the model that was stored: orphans removed, silent
           states with a probability-1 out-edge merged away
  decode   reads decoded by the reference on the reloaded model (log-probabilities, state paths)
"""
import gzip
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


def main():
    ref = refenv.reference_pomegranate()
    hu = refenv.reference_hmm_utils(ref, "ref")
    rng = random.Random(41)
    left, right = synth.rand_dna(rng, 60), synth.rand_dna(rng, 60)
    ru = synth.rand_dna(rng, 14)
    segs = [ru, synth.sequencing_errors(rng, ru, 0.1, 0.0, 0.0)[:14].ljust(14, "A"), ru]
    copies = 4
    model = hu.get_read_matcher_model(left, right, segs, copies=copies)
    text = model.to_json()
    again = type(model).from_json(text)
    b = oracle.baked_from_reference_model(again)
    reads = []
    for k in range(40):
        n = rng.randint(1, 4)
        allele = left + ru * n + right
        s = rng.randrange(0, len(allele) - 30)
        reads.append(synth.sequencing_errors(rng, allele[s:s + rng.randint(25, 90)], 0.02, 0.005, 0.005) or "A")
    logp, paths, off = [], [], [0]
    for r in reads:
        lp, vp = again.viterbi(r)
        logp.append(lp)
        paths.extend(i for i, _ in (vp or []))
        off.append(len(paths))
    np.savez_compressed(
        os.path.join(HERE, "stored_model.npz"),
        json_gz=np.frombuffer(gzip.compress(text.encode(), 9, mtime=0), dtype=np.uint8),
        in_off=b["in_off"], in_src=b["in_src"], in_logp=b["in_logp"], emis=b["emis"],
        scalars=np.array([b["n_states"], b["silent_start"], b["start_index"], b["end_index"], b["finite"]]),
        names=np.array("\n".join(b["names"])),
        inputs=np.array(json.dumps({"left": left, "right": right, "segments": segs, "copies": copies, "reads": reads})),
        n_states_stored=np.array(len(model.states)),
        json_again_sha=np.array(__import__("hashlib").sha256(again.to_json().encode()).hexdigest()),
        logp=np.array(logp, dtype=np.float64), paths=np.array(paths, dtype=np.int32),
        path_off=np.array(off, dtype=np.int64))
    print("stored model: %d states -> %d after from_json; json %d bytes; %d reads decoded" % (
        len(model.states), b["n_states"], len(text), len(reads)))


if __name__ == "__main__":
    main()
