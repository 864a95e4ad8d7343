"""ctypes front-end of the C restatement (``oracle/hmm_oracle.c``) + helpers to read the
baked arrays out of the compiled reference engine.

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs.  Never by ``advntr_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libhmm_oracle.so")
SRC = os.path.join(HERE, "hmm_oracle.c")


class _Model(C.Structure):
    _fields_ = [("n_states", C.c_int32), ("silent_start", C.c_int32),
                ("start_index", C.c_int32), ("end_index", C.c_int32),
                ("finite", C.c_int32), ("n_symbols", C.c_int32),
                ("in_off", C.c_void_p), ("in_src", C.c_void_p),
                ("in_logp", C.c_void_p), ("emis", C.c_void_p)]


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_viterbi_batch.restype = None
        _lib.oracle_viterbi_f32_batch.restype = None
        _lib.oracle_log_probability_batch.restype = None
    return _lib


class OracleModel(object):
    """Holds the baked arrays (C-contiguous copies) and the C struct that points at them."""

    def __init__(self, baked: dict):
        self.n_states = int(baked["n_states"])
        self.in_off = np.ascontiguousarray(baked["in_off"], dtype=np.int32)
        self.in_src = np.ascontiguousarray(baked["in_src"], dtype=np.int32)
        self.in_logp = np.ascontiguousarray(baked["in_logp"], dtype=np.float64)
        self.emis = np.ascontiguousarray(baked["emis"], dtype=np.float64)
        self.c = _Model(self.n_states, int(baked["silent_start"]), int(baked["start_index"]),
                        int(baked["end_index"]), int(baked["finite"]), int(self.emis.shape[1]),
                        self.in_off.ctypes.data, self.in_src.ctypes.data,
                        self.in_logp.ctypes.data, self.emis.ctypes.data)

    @staticmethod
    def _pack(codes):
        lens = np.fromiter((len(c) for c in codes), dtype=np.int64, count=len(codes))
        off = np.zeros(len(codes) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        flat = (np.concatenate([np.asarray(c, dtype=np.uint8) for c in codes])
                if len(codes) and off[-1] else np.zeros(1, dtype=np.uint8))
        return np.ascontiguousarray(flat), off

    def viterbi(self, codes, fp32=False):
        """codes: list of uint8 arrays.  Returns (logp[R], [path arrays or None]).  ``fp32``: the
        float restatement (checker of the engine's optional fp32 mode)."""
        flat, off = self._pack(codes)
        R = len(codes)
        stride = int((off[1:] - off[:-1]).max() if R else 0) + self.n_states
        logp = np.empty(R, dtype=np.float64)
        plen = np.empty(R, dtype=np.int32)
        paths = np.empty((R, stride), dtype=np.int32)
        fn = lib().oracle_viterbi_f32_batch if fp32 else lib().oracle_viterbi_batch
        fn(C.byref(self.c), C.c_void_p(flat.ctypes.data),
                                   C.c_void_p(off.ctypes.data), C.c_int32(R),
                                   C.c_void_p(logp.ctypes.data), C.c_void_p(plen.ctypes.data),
                                   C.c_void_p(paths.ctypes.data), C.c_int64(stride))
        return logp, [paths[r, :plen[r]].copy() if plen[r] >= 0 else None for r in range(R)]

    def log_probability(self, codes):
        flat, off = self._pack(codes)
        R = len(codes)
        logp = np.empty(R, dtype=np.float64)
        lib().oracle_log_probability_batch(C.byref(self.c), C.c_void_p(flat.ctypes.data),
                                           C.c_void_p(off.ctypes.data), C.c_int32(R),
                                           C.c_void_p(logp.ctypes.data))
        return logp


def baked_from_reference_model(model, alphabet="ACGT") -> dict:
    """Read the baked arrays out of a compiled-reference ``HiddenMarkovModel``.

    The C arrays are private to the Cython class, but ``states`` (baked order) and ``graph``
    (whose ``edges_iter`` order bake() walked, hmm.pyx:994-1011) are public; in-edge CSR is
    the stable sort of that walk by target.  Weights are the stored logs, bit-exact.
    """
    states = model.states
    m = len(states)
    idx = {s: i for i, s in enumerate(states)}
    src, dst, wts = [], [], []
    for a, b, data in model.graph.edges_iter(data=True):
        src.append(idx[a]); dst.append(idx[b]); wts.append(data["probability"])
    src = np.asarray(src, dtype=np.int32); dst = np.asarray(dst, dtype=np.int32)
    wts = np.asarray(wts, dtype=np.float64)
    order = np.argsort(dst, kind="stable")
    in_off = np.zeros(m + 1, dtype=np.int32)
    np.cumsum(np.bincount(dst, minlength=m), out=in_off[1:])
    S = model.silent_start
    emis = np.array([[states[l].distribution.log_probability(ch) for ch in alphabet]
                     for l in range(S)], dtype=np.float64).reshape(S, len(alphabet))
    end = model.end_index
    return {"n_states": m, "silent_start": S, "start_index": model.start_index,
            "end_index": end, "finite": int(in_off[end + 1] - in_off[end] > 0),
            "in_off": in_off, "in_src": src[order], "in_logp": wts[order], "emis": emis,
            "names": [s.name for s in states]}


def encode(seq: str) -> np.ndarray:
    lut = np.full(256, 255, dtype=np.uint8)
    for i, ch in enumerate("ACGT"):
        lut[ord(ch)] = i
        lut[ord(ch.lower())] = i
    out = lut[np.frombuffer(seq.encode("ascii"), dtype=np.uint8)]
    if (out == 255).any():
        raise ValueError("non-ACGT symbol")
    return out
