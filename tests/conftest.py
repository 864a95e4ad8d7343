"""Shared fixtures.  ``-m "not gpu"`` covers the oracle against the golden vectors, the host
logic (builders, model compiler via the test-only host mirror, sharding) and that the C-ABI
library loads and exports every symbol of include/advhmm.h.  ``-m gpu`` is the parity suite
proper: every call goes through the C-ABI onto the device."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ("config1", "small_a", "small_b", "divergent")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


class Golden(object):
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        sc = z["scalars"]
        self.name = name
        self.baked = {"n_states": int(sc[0]), "silent_start": int(sc[1]), "start_index": int(sc[2]),
                      "end_index": int(sc[3]), "finite": int(sc[4]), "in_off": z["in_off"],
                      "in_src": z["in_src"], "in_logp": z["in_logp"], "emis": z["emis"]}
        self.names = str(z["names"]).split("\n")
        self.inputs = json.loads(str(z["inputs"]))
        self.reads = self.inputs["reads"]
        self.logp, self.forward, self.ru_count = z["logp"], z["forward"], z["ru_count"]
        self._paths, self._off = z["paths"], z["path_off"]

    def path(self, i):
        if self.ru_count[i] < 0 and self._off[i] == self._off[i + 1]:
            return None
        return self._paths[self._off[i]:self._off[i + 1]]

    def codes(self):
        import oracle
        return [oracle.encode(r) for r in self.reads]


@pytest.fixture(scope="session", params=GOLDEN_CASES)
def golden(request):
    return Golden(request.param)


@pytest.fixture(scope="session")
def golden_config1():
    return Golden("config1")


@pytest.fixture(scope="session")
def host_mirror():
    """tests/_host_mirror.so: sequential walk of the tables the CUDA kernels read (test-only)."""
    src = os.path.join(ROOT, "tests", "host_mirror.cpp")
    lib = os.path.join(ROOT, "tests", "_host_mirror.so")
    deps = [src, os.path.join(ROOT, "advntr_b200", "csrc", "model_compile.hpp")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", lib, src])
    return C.CDLL(lib)


def mirror_viterbi(lib, baked, codes, kind):
    import oracle
    om = oracle.OracleModel(baked)
    flat, off = om._pack(codes)
    R = len(codes)
    stride = int((off[1:] - off[:-1]).max() if R else 0) + om.n_states
    lp = np.empty(R)
    pl = np.empty(R, dtype=np.int32)
    paths = np.empty((R, stride), dtype=np.int32)
    why = C.create_string_buffer(256)
    rc = lib.mirror_viterbi(C.byref(om.c), kind, C.c_void_p(flat.ctypes.data), C.c_void_p(off.ctypes.data), R,
                            C.c_void_p(lp.ctypes.data), C.c_void_p(pl.ctypes.data),
                            C.c_void_p(paths.ctypes.data), C.c_int64(stride), why, 256)
    assert rc == 0, why.value
    return lp, [paths[r, :pl[r]].copy() if pl[r] >= 0 else None for r in range(R)]


def same_bits(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.array_equal(a.view(np.int64), b.view(np.int64)))


def assert_paths_equal(got, want, what=""):
    assert len(got) == len(want)
    for i, (g, w) in enumerate(zip(got, want)):
        if w is None:
            assert g is None, "%s read %d: expected impossible" % (what, i)
        else:
            assert g is not None and np.array_equal(np.asarray(g), np.asarray(w)), \
                "%s read %d: state path differs" % (what, i)


def gpu_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
