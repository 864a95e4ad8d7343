"""The C-ABI shared library loads without a GPU, exports every symbol include/advhmm.h declares,
and its host-side model analysis works in a host-only context (no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "advhmm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(advhmm_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from advntr_b200 import build, engine
    build.build_library()
    lib = engine.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), "libadvhmm.so does not export " + s
    assert set(syms) == set(engine.EXPORTS)
    assert lib.advhmm_abi_version() == 1


def test_host_only_context_analyses_models(golden):
    from advntr_b200 import engine
    ctx = engine.Context(device=-1)
    dm = engine.DeviceModel(ctx, golden.baked)
    assert dm.kind == "banded"
    assert dm.info.n_states == golden.baked["n_states"]
    assert dm.info.n_edges == len(golden.baked["in_src"])
    assert dm.info.n_final_states == 4
    assert 0 < dm.info.smem_bytes < 227 * 1024
    # no CPU fallback: decoding on a context without a device fails loudly
    with pytest.raises(engine.EngineError) as ei:
        dm.viterbi(golden.codes()[:2])
    assert ei.value.code == engine.ECUDA
    dm.close()
    ctx.close()


def test_malformed_model_is_rejected():
    from advntr_b200 import engine
    ctx = engine.Context(device=-1)
    bad = {"n_states": 3, "silent_start": 1, "start_index": 1, "end_index": 2, "finite": 1,
           "in_off": np.array([0, 1, 1, 2], dtype=np.int32), "in_src": np.array([1, 9], dtype=np.int32),
           "in_logp": np.zeros(2), "emis": np.zeros((1, 4))}
    with pytest.raises(engine.EngineError) as ei:
        engine.DeviceModel(ctx, bad)
    assert ei.value.code == engine.EINVAL
    ctx.close()


def test_encode_acgt_reports_first_bad_symbol():
    from advntr_b200 import engine
    codes, bad = engine.encode_acgt("ACGTacgt")
    assert bad == -1 and list(codes) == [0, 1, 2, 3, 0, 1, 2, 3]
    assert engine.encode_acgt("ACNGT")[1] == 2
