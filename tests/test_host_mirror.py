"""Host-side model compiler: the generic level schedule and the banded profile tables, walked
sequentially by the test-only mirror, reproduce the golden vectors bit for bit."""
import ctypes as C

import numpy as np
import pytest

import oracle
from conftest import assert_paths_equal, mirror_viterbi, same_bits


@pytest.mark.parametrize("kind", [0, 1], ids=["generic-tables", "banded-tables"])
def test_mirror_matches_golden(golden, host_mirror, kind):
    logp, paths = mirror_viterbi(host_mirror, golden.baked, golden.codes(), kind)
    assert same_bits(logp, golden.logp)
    assert_paths_equal(paths, [golden.path(i) for i in range(len(golden.reads))], golden.name)


def test_read_matcher_family_is_banded(golden, host_mirror):
    om = oracle.OracleModel(golden.baked)
    info = np.zeros(6, dtype=np.int32)
    why = C.create_string_buffer(256)
    assert host_mirror.mirror_info(C.byref(om.c), C.c_void_p(info.ctypes.data), why, 256) == 0
    valid, NC, n_final, acc_col, n_levels, n_edges = info
    assert valid == 1, why.value
    assert n_final == 4 and acc_col > 0
    assert n_edges == len(golden.baked["in_src"])
    # one column per live silent state; the silent chain is as deep as the level schedule
    assert NC == golden.baked["n_states"] - golden.baked["silent_start"] - n_final
    assert n_levels == NC + 3


def test_non_profile_model_falls_back_to_generic(host_mirror):
    """A two-state loop with no silent structure is not banded; the generic tables still decode it
    exactly like the oracle (non-finite termination included)."""
    in_off = np.array([0, 2, 4, 4, 4], dtype=np.int32)        # states: e0, e1, start, end
    in_src = np.array([0, 2, 0, 1], dtype=np.int32)
    in_logp = np.log(np.array([0.6, 1.0, 0.4, 1.0]))
    emis = np.log(np.array([[0.7, 0.1, 0.1, 0.1], [0.1, 0.1, 0.1, 0.7]]))
    baked = {"n_states": 4, "silent_start": 2, "start_index": 2, "end_index": 3, "finite": 0,
             "in_off": in_off, "in_src": in_src, "in_logp": in_logp, "emis": emis}
    om = oracle.OracleModel(baked)
    codes = [np.array(x, dtype=np.uint8) for x in ([0, 0, 3, 3], [3], [0, 1, 2, 3, 0], [])]
    want_lp, want_paths = om.viterbi(codes)
    lp, paths = mirror_viterbi(host_mirror, baked, codes, 0)
    assert same_bits(lp, want_lp)
    assert_paths_equal(paths, want_paths)
    info = np.zeros(6, dtype=np.int32)
    why = C.create_string_buffer(256)
    host_mirror.mirror_info(C.byref(om.c), C.c_void_p(info.ctypes.data), why, 256)
    assert info[0] == 0 and why.value
