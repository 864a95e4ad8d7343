"""Path consumers (advntr_b200/path_utils.py) against the reference's own fixture and against
values computed by the reference's hmm_utils on the golden paths."""
import json
import os

import numpy as np

from advntr_b200 import path_utils
from conftest import GOLDEN


class _S(object):
    def __init__(self, name):
        self.name = name


def test_reference_fixture_repeat_segments():
    fx = json.load(open(os.path.join(GOLDEN, "ref_tests_hmm_utils.json")))
    states = fx["visited_states"].split(",")
    repeats, runs = path_utils.extract_repeating_segments_from_read(fx["sequence"], states)
    assert repeats == fx["correct_repeats"]
    assert len(runs) == len(repeats)
    assert path_utils.get_repeating_pattern_lengths(states) == [len(r) for r in repeats]


def test_consumers_match_reference_values(golden):
    left, right = golden.inputs["left"], golden.inputs["right"]
    z = np.load(os.path.join(GOLDEN, golden.name + ".npz"))
    want = z["consumers"]
    for i, read in enumerate(golden.reads):
        p = golden.path(i)
        if p is None:
            continue
        vpath = [(int(k), _S(golden.names[k])) for k in p]
        assert path_utils.get_number_of_repeats_in_vpath(vpath) == golden.ru_count[i]
        s = path_utils.summarize(vpath)
        assert [s.n_match, s.repeat_bp, s.left_bp, s.right_bp] == list(want[i, :4])
        assert s.n_emitted == len(read)
        if read:
            assert path_utils.get_flanking_regions_matching_rate(vpath, read, left, right) == want[i, 4]


def test_rescore_paths_reproduces_reference_logp(golden):
    """The reference's own paths, re-scored with the baked tables in its operation order, give its own
    log-probabilities bit for bit (the size-independent check bench.py and the GPU tests rely on)."""
    import oracle
    keep = [i for i in range(len(golden.reads)) if golden.path(i) is not None]
    codes = [oracle.encode(golden.reads[i]) for i in keep]
    scores = path_utils.rescore_paths(golden.baked, codes, [golden.path(i) for i in keep])
    assert np.array_equal(scores.view(np.int64), golden.logp[keep].view(np.int64))
    # and it notices a corrupted path
    bad = [np.array(golden.path(i)) for i in keep]
    victim = max(range(len(bad)), key=lambda i: len(bad[i]))
    k = next(k for k in range(len(bad[victim]) // 2, len(bad[victim])) if bad[victim][k] != bad[victim][k - 1])
    bad[victim][k] = bad[victim][k - 1]
    try:
        noticed = path_utils.rescore_paths(golden.baked, codes, bad)[victim] != golden.logp[keep][victim]
    except AssertionError:
        noticed = True
    assert noticed
