"""ctypes binding of the C-ABI (``include/advhmm.h``) -- the only way this package decodes.

There is no CPU fallback: if ``libadvhmm.so`` is missing or no sm_100 GPU is visible, the
decoding calls raise.  Build the library with ``python -m advntr_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ADVHMM_LIB") or os.path.join(_PKG, "libadvhmm.so")

OK, EINVAL, ECUDA, ENOMEM, ESYMBOL, ECAPACITY, EUNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
WANT_PATH, BOTH_STRANDS, FP32, FORCE_GENERIC, WANT_SUMMARY, DEVICE_BUFFERS, DEVICE_OFFSETS = 0x1, 0x2, 0x4, 0x8, 0x10, 0x100, 0x200
SUMMARY_DTYPE = np.dtype([(n, np.int32) for n in ("repeats", "n_match", "repeat_bp", "left_bp", "right_bp",
                                                   "left_hits", "right_hits", "unit_starts_ends")])
KIND_GENERIC, KIND_BANDED = 0, 1
CALL_ACCURACY_FILTER, CALL_HAPLOID = 0x1, 0x2
# ``advhmm_locus_call``: what the native stage after the decode returns per locus
CALL_DTYPE = np.dtype([("has_call", np.int32), ("c1", np.int32), ("c2", np.int32), ("recruited", np.int32),
                       ("spanning", np.int32), ("flanking", np.int32), ("max_prob", np.float64)], align=True)
# ``advhmm_frameshift_call``
FRAMESHIFT_DTYPE = np.dtype([("kind", np.int32), ("column", np.int32), ("base", np.int32), ("count", np.int32),
                             ("selected", np.int32), ("reserved", np.int32), ("repeat_bp", np.int64)], align=True)

EXPORTS = (
    "advhmm_context_create", "advhmm_context_destroy", "advhmm_context_synchronize",
    "advhmm_context_stream", "advhmm_context_launch_count", "advhmm_context_profile",
    "advhmm_context_profile_read", "advhmm_fp64_add_peak", "advhmm_context_bad_symbol",
    "advhmm_model_create", "advhmm_model_destroy", "advhmm_model_info_get",
    "advhmm_viterbi_batch", "advhmm_log_probability_batch", "advhmm_viterbi_multi",
    "advhmm_viterbi_multi_summary", "advhmm_model_set_state_classes",
    "advhmm_kfilter_create", "advhmm_kfilter_destroy", "advhmm_kfilter_scan",
    "advhmm_last_error", "advhmm_abi_version", "advhmm_encode_acgt",
    "advhmm_set_vexp", "advhmm_models_create_for_loci", "advhmm_shape_cache_clear",
    "advhmm_model_dims_get", "advhmm_model_tables_get", "advhmm_model_banded_tables_get",
    "advhmm_genotypes_from_summaries", "advhmm_genotypes_from_counts", "advhmm_frameshift_candidates",
)


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "advhmm error %d: %s" % (code, msg))
        self.code = code


class ModelDesc(C.Structure):
    _fields_ = [("n_states", C.c_int32), ("silent_start", C.c_int32), ("start_index", C.c_int32),
                ("end_index", C.c_int32), ("finite", C.c_int32), ("n_symbols", C.c_int32),
                ("in_off", C.c_void_p), ("in_src", C.c_void_p), ("in_logp", C.c_void_p),
                ("emis", C.c_void_p)]


class ModelInfo(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_states", C.c_int32), ("n_edges", C.c_int32),
                ("n_columns", C.c_int32), ("n_final_states", C.c_int32), ("smem_bytes", C.c_int32),
                ("max_in_degree", C.c_int32), ("reserved", C.c_int32)]


class LociDesc(C.Structure):
    """``advhmm_loci``: the columns of a batch of loci (include/advhmm.h)."""
    _fields_ = [("n_loci", C.c_int32),
                ("left", C.c_void_p), ("left_off", C.c_void_p),
                ("right", C.c_void_p), ("right_off", C.c_void_p),
                ("segments", C.c_void_p), ("seg_off", C.c_void_p),
                ("n_segments", C.c_void_p), ("copies", C.c_void_p), ("error_rate", C.c_void_p)]


class ModelDims(C.Structure):
    _fields_ = [("n_states", C.c_int32), ("silent_start", C.c_int32), ("start_index", C.c_int32),
                ("end_index", C.c_int32), ("finite", C.c_int32), ("n_symbols", C.c_int32),
                ("n_edges", C.c_int64), ("names_bytes", C.c_int64), ("shape", C.c_int32 * 4)]


VEXP_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)


def _numpy_vexp(inp, out, n, user):
    """The vector exp of the reference: ``numpy.exp`` (``hmm.pyx:514``), on the library's buffers."""
    a = np.frombuffer((C.c_double * n).from_address(inp), dtype=np.float64)
    b = np.frombuffer((C.c_double * n).from_address(out), dtype=np.float64)
    np.exp(a, out=b)


_vexp_keepalive = VEXP_FN(_numpy_vexp)

_lib = None
_lib_lock = threading.Lock()


def load_library():
    """Load ``libadvhmm.so`` (no compute is started).  Raises ImportError when it is not built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError("advntr_b200: %s is not built -- run `python -m advntr_b200.build` "
                              "(there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        vp, i32, i64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
        lib.advhmm_context_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
        lib.advhmm_context_destroy.argtypes = [vp]
        lib.advhmm_context_destroy.restype = None
        lib.advhmm_context_synchronize.argtypes = [vp]
        lib.advhmm_context_stream.argtypes = [vp]
        lib.advhmm_context_stream.restype = vp
        lib.advhmm_context_launch_count.argtypes = [vp]
        lib.advhmm_context_launch_count.restype = i64
        lib.advhmm_context_profile.argtypes = [vp, C.c_int]
        lib.advhmm_context_profile_read.argtypes = [vp, vp, vp, vp, vp]
        lib.advhmm_fp64_add_peak.argtypes = [vp, vp]
        lib.advhmm_context_bad_symbol.argtypes = [vp, vp]
        lib.advhmm_model_create.argtypes = [vp, C.POINTER(ModelDesc), C.POINTER(vp)]
        lib.advhmm_model_destroy.argtypes = [vp]
        lib.advhmm_model_destroy.restype = None
        lib.advhmm_model_info_get.argtypes = [vp, C.POINTER(ModelInfo)]
        lib.advhmm_viterbi_batch.argtypes = [vp, vp, vp, i32, u32, vp, vp, vp, vp, i64, vp]
        lib.advhmm_log_probability_batch.argtypes = [vp, vp, vp, i32, u32, vp]
        lib.advhmm_viterbi_multi.argtypes = [vp, vp, i32, vp, vp, vp, i32, u32, vp, vp, vp, vp, i64, vp]
        lib.advhmm_viterbi_multi_summary.argtypes = [vp, vp, i32, vp, vp, vp, i32, u32, vp, vp, vp, vp, i64, vp, vp]
        lib.advhmm_model_set_state_classes.argtypes = [vp, vp]
        lib.advhmm_kfilter_create.argtypes = [vp, i64, C.c_char_p, vp, vp, C.POINTER(vp)]
        lib.advhmm_kfilter_destroy.argtypes = [vp]
        lib.advhmm_kfilter_destroy.restype = None
        lib.advhmm_kfilter_scan.argtypes = [vp, vp, vp, i32, i32, u32, vp, vp, vp, i64, vp]
        lib.advhmm_last_error.restype = C.c_char_p
        lib.advhmm_abi_version.restype = C.c_int
        lib.advhmm_encode_acgt.argtypes = [C.c_char_p, i64, vp]
        lib.advhmm_encode_acgt.restype = i64
        lib.advhmm_set_vexp.argtypes = [VEXP_FN, vp]
        lib.advhmm_models_create_for_loci.argtypes = [vp, C.POINTER(LociDesc), i32, vp]
        lib.advhmm_shape_cache_clear.restype = None
        lib.advhmm_model_dims_get.argtypes = [vp, C.POINTER(ModelDims)]
        lib.advhmm_model_tables_get.argtypes = [vp, vp, vp, vp, vp, vp]
        lib.advhmm_model_banded_tables_get.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, vp]
        lib.advhmm_genotypes_from_summaries.argtypes = [i64, vp, vp, vp, vp, vp, vp, vp, vp, u32, i32, i32, vp, vp]
        lib.advhmm_genotypes_from_counts.argtypes = [i64, vp, vp, u32, vp]
        lib.advhmm_frameshift_candidates.argtypes = [i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp]
        # parameter chains go through numpy.exp, as the reference's do (the library's default is libm)
        lib.advhmm_set_vexp(_vexp_keepalive, None)
        _lib = lib
        return lib


def _check(rc):
    if rc != OK:
        raise EngineError(rc, load_library().advhmm_last_error().decode("utf-8", "replace"))


def encode_acgt(seq):
    """ASCII DNA -> uint8 codes (A,C,G,T = 0..3).  Returns (codes, index of first bad byte or -1)."""
    raw = seq.encode("ascii", "replace") if isinstance(seq, str) else bytes(seq)
    out = np.empty(len(raw), dtype=np.uint8)
    bad = load_library().advhmm_encode_acgt(raw, len(raw), out.ctypes.data)
    return out, int(bad)


def encode_batch(sequences):
    """list of ASCII DNA strings -> (flat uint8 codes, int64 offsets[R+1]) with ONE encode call.
    Raises the reference's ValueError for the first symbol that is not ACGT (hmm.pyx:72-79)."""
    R = len(sequences)
    off = np.zeros(R + 1, dtype=np.int64)
    if R:
        np.cumsum(np.fromiter(map(len, sequences), dtype=np.int64, count=R), out=off[1:])
    raw = "".join(sequences).encode("ascii", "replace")
    flat = np.empty(max(len(raw), 1), dtype=np.uint8)
    bad = load_library().advhmm_encode_acgt(raw, len(raw), flat.ctypes.data)
    if bad >= 0:
        raise ValueError("Symbol '{}' is not defined in a distribution".format(chr(raw[bad])))
    return flat, off


def pack_reads(codes):
    """list of uint8 arrays -> (flat uint8, int64 offsets[R+1])."""
    R = len(codes)
    off = np.zeros(R + 1, dtype=np.int64)
    if R:
        np.cumsum(np.fromiter((len(c) for c in codes), dtype=np.int64, count=R), out=off[1:])
    flat = np.empty(max(int(off[-1]), 1), dtype=np.uint8)
    for c, a, b in zip(codes, off[:-1], off[1:]):
        flat[a:b] = c
    return flat, off


def genotypes_from_summaries(group_off, n_mapped, n_unmapped, min_score, logp, summaries, path_len, seq_off,
                             accuracy_filter=False, is_haploid=False, min_repeat_bp=2, threads=0, want_read_class=False):
    """``advhmm_genotypes_from_summaries``: the per-read results of many loci (what
    ``advhmm_viterbi_multi_summary`` returned) -> one ``CALL_DTYPE`` record per locus, on all host
    threads.  ``min_score``: one float per locus, NaN (or None for all) = no minimum score known.
    -> (calls, read_class or None)."""
    group_off = np.ascontiguousarray(group_off, dtype=np.int64)
    n_loci = len(group_off) - 1
    n_mapped = np.ascontiguousarray(n_mapped, dtype=np.int32)
    n_unmapped = np.ascontiguousarray(n_unmapped, dtype=np.int32)
    if len(n_mapped) != n_loci or len(n_unmapped) != n_loci:
        raise ValueError("one (mapped, unmapped) pair per locus")
    score = None if min_score is None else np.ascontiguousarray(min_score, dtype=np.float64)
    if score is not None and len(score) != n_loci:
        raise ValueError("one minimum score per locus")
    logp = np.ascontiguousarray(logp, dtype=np.float64)
    summaries = np.ascontiguousarray(summaries, dtype=SUMMARY_DTYPE)
    path_len = np.ascontiguousarray(path_len, dtype=np.int32)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
    n_reads = int(group_off[-1]) if n_loci else 0
    if min(len(logp), len(summaries), len(path_len), len(seq_off) - 1) < n_reads:
        raise ValueError("per-read arrays are shorter than group_off says")
    calls = np.zeros(n_loci, dtype=CALL_DTYPE)
    read_class = np.zeros(max(n_reads, 1), dtype=np.uint8) if want_read_class else None
    flags = (CALL_ACCURACY_FILTER if accuracy_filter else 0) | (CALL_HAPLOID if is_haploid else 0)
    _check(load_library().advhmm_genotypes_from_summaries(
        n_loci, group_off.ctypes.data, n_mapped.ctypes.data, n_unmapped.ctypes.data,
        score.ctypes.data if score is not None else None, logp.ctypes.data, summaries.ctypes.data, path_len.ctypes.data,
        seq_off.ctypes.data, flags, int(min_repeat_bp), int(threads), calls.ctypes.data,
        read_class.ctypes.data if want_read_class else None))
    return calls, (read_class[:n_reads] if want_read_class else None)


def genotypes_from_counts(count_lists, accuracy_filter=False, is_haploid=False):
    """``advhmm_genotypes_from_counts``: ``find_genotype_based_on_observed_repeats`` for many lists of
    observed repeat counts -> ``CALL_DTYPE`` records (``has_call`` 0 = the reference's ``None``)."""
    n = len(count_lists)
    off = np.zeros(n + 1, dtype=np.int64)
    if n:
        np.cumsum(np.fromiter(map(len, count_lists), dtype=np.int64, count=n), out=off[1:])
    flat = np.zeros(max(int(off[-1]), 1), dtype=np.int32)
    for c, a, b in zip(count_lists, off[:-1], off[1:]):
        flat[a:b] = c
    calls = np.zeros(n, dtype=CALL_DTYPE)
    flags = (CALL_ACCURACY_FILTER if accuracy_filter else 0) | (CALL_HAPLOID if is_haploid else 0)
    _check(load_library().advhmm_genotypes_from_counts(n, flat.ctypes.data, off.ctypes.data, flags, calls.ctypes.data))
    return calls


def frameshift_candidates(group_off, pattern_len, min_score, state_tables, logp, summaries, path_len, path_off, paths,
                          seqs, seq_off, threads=0):
    """``advhmm_frameshift_candidates``: the path-walking part of ``find_frameshift_from_selected_reads``
    (``vntr_finder.py:265-300``) for many loci on all host threads.  ``state_tables[g]`` = (class bytes,
    name numbers) of locus g's model (``path_utils.frameshift_state_tables``); the per-read arrays are what a
    ``WANT_PATH | WANT_SUMMARY`` call returned.  -> ``FRAMESHIFT_DTYPE`` records."""
    group_off = np.ascontiguousarray(group_off, dtype=np.int64)
    n_loci = len(group_off) - 1
    if len(state_tables) != n_loci or len(pattern_len) != n_loci:
        raise ValueError("one pattern length and one pair of state tables per locus")
    state_off = np.zeros(n_loci + 1, dtype=np.int64)
    if n_loci:
        np.cumsum([len(c) for c, _ in state_tables], out=state_off[1:])
    cls = np.ascontiguousarray(np.concatenate([c for c, _ in state_tables]) if n_loci else np.zeros(1), dtype=np.uint8)
    lab = np.ascontiguousarray(np.concatenate([l for _, l in state_tables]) if n_loci else np.zeros(1), dtype=np.int32)
    plen_ = np.ascontiguousarray(pattern_len, dtype=np.int32)
    score = None if min_score is None else np.ascontiguousarray(min_score, dtype=np.float64)
    logp = np.ascontiguousarray(logp, dtype=np.float64)
    summaries = np.ascontiguousarray(summaries, dtype=SUMMARY_DTYPE)
    path_len = np.ascontiguousarray(path_len, dtype=np.int32)
    path_off = np.ascontiguousarray(path_off, dtype=np.int64)
    paths = np.ascontiguousarray(paths, dtype=np.int32)
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
    out = np.zeros(n_loci, dtype=FRAMESHIFT_DTYPE)
    _check(load_library().advhmm_frameshift_candidates(
        n_loci, group_off.ctypes.data, plen_.ctypes.data, score.ctypes.data if score is not None else None,
        state_off.ctypes.data, cls.ctypes.data, lab.ctypes.data, logp.ctypes.data, summaries.ctypes.data,
        path_len.ctypes.data, path_off.ctypes.data, paths.ctypes.data, seqs.ctypes.data, seq_off.ctypes.data, int(threads),
        out.ctypes.data))
    return out


class LociColumns(object):
    """Host arrays behind one ``advhmm_loci``: flank codes, aligned repeat segments, copies, error rates."""

    def __init__(self, left, left_off, right, right_off, segments, seg_off, n_segments, copies, error_rate):
        self.left = np.ascontiguousarray(left, dtype=np.uint8)
        self.right = np.ascontiguousarray(right, dtype=np.uint8)
        self.left_off = np.ascontiguousarray(left_off, dtype=np.int64)
        self.right_off = np.ascontiguousarray(right_off, dtype=np.int64)
        self.segments = np.ascontiguousarray(segments, dtype=np.uint8)
        self.seg_off = np.ascontiguousarray(seg_off, dtype=np.int64)
        self.n_segments = np.ascontiguousarray(n_segments, dtype=np.int32)
        self.copies = np.ascontiguousarray(copies, dtype=np.int32)
        self.error_rate = np.ascontiguousarray(error_rate, dtype=np.float64)
        self.n = len(self.copies)

    @classmethod
    def from_lists(cls, lefts, rights, segment_lists, copies, error_rates):
        """``lefts`` / ``rights``: the flank strings that enter the model (already trimmed to the flank size);
        ``segment_lists``: per locus the aligned repeat segments (equal-length strings over ACGT-)."""
        n = len(copies)

        def flat_codes(strings):
            off = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(np.fromiter(map(len, strings), dtype=np.int64, count=n), out=off[1:])
            raw = "".join(strings).encode("ascii", "replace")
            out = np.empty(max(len(raw), 1), dtype=np.uint8)
            bad = load_library().advhmm_encode_acgt(raw, len(raw), out.ctypes.data)
            if bad >= 0:
                raise ValueError("flank contains a non-ACGT symbol")
            return out, off
        for segs in segment_lists:
            if len(segs) < 1 or len(set(map(len, segs))) != 1:
                raise ValueError("the aligned repeat segments of a locus must have one width")
        left, left_off = flat_codes(lefts)
        right, right_off = flat_codes(rights)
        seg_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.fromiter((sum(map(len, segs)) for segs in segment_lists), dtype=np.int64, count=n), out=seg_off[1:])
        raw = "".join("".join(segs) for segs in segment_lists).upper().encode("ascii", "replace")
        segments = np.frombuffer(raw, dtype=np.uint8) if raw else np.zeros(1, dtype=np.uint8)
        n_seg = np.fromiter(map(len, segment_lists), dtype=np.int32, count=n)
        rates = np.full(n, error_rates, dtype=np.float64) if np.isscalar(error_rates) else error_rates
        return cls(left, left_off, right, right_off, segments, seg_off, n_seg, copies, rates)

    def desc(self, lo=0, hi=None):
        """The C struct for loci [lo, hi) (a view: the arrays of ``self`` must stay alive)."""
        hi = self.n if hi is None else hi
        return LociDesc(hi - lo, self.left.ctypes.data, self.left_off[lo:].ctypes.data,
                        self.right.ctypes.data, self.right_off[lo:].ctypes.data,
                        self.segments.ctypes.data, self.seg_off[lo:].ctypes.data,
                        self.n_segments[lo:].ctypes.data, self.copies[lo:].ctypes.data, self.error_rate[lo:].ctypes.data)


class Context(object):
    """One engine context (device + stream + workspaces).  ``device=-1``: host-only analysis."""

    _default = {}

    def __init__(self, device=0, stream=None):
        self._lib = load_library()
        h = C.c_void_p()
        _check(self._lib.advhmm_context_create(int(device), C.c_void_p(stream or 0), C.byref(h)))
        self._h = h
        self.device = int(device)

    @classmethod
    def default(cls, device=None):
        if device is None:
            device = int(os.environ.get("ADVHMM_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        ctx = cls._default.get(device)
        if ctx is None:
            ctx = cls._default[device] = cls(device)
        return ctx

    @property
    def stream(self):
        return self._lib.advhmm_context_stream(self._h)

    @property
    def launch_count(self):
        return int(self._lib.advhmm_context_launch_count(self._h))

    def synchronize(self):
        _check(self._lib.advhmm_context_synchronize(self._h))

    def profile(self, enable=True):
        _check(self._lib.advhmm_context_profile(self._h, 1 if enable else 0))

    def profile_read(self):
        """(fill_ms, fill_launches, backtrack_ms, backtrack_launches) since the last read."""
        fm, bm = C.c_double(0), C.c_double(0)
        fl, bl = C.c_int64(0), C.c_int64(0)
        _check(self._lib.advhmm_context_profile_read(self._h, C.addressof(fm), C.addressof(fl),
                                                     C.addressof(bm), C.addressof(bl)))
        return fm.value, fl.value, bm.value, bl.value

    def bad_symbol(self):
        """After a device-buffer call: index of its first read with a code outside the alphabet, or -1."""
        bad = C.c_int32(-1)
        _check(self._lib.advhmm_context_bad_symbol(self._h, C.addressof(bad)))
        return bad.value

    def fp64_add_peak(self):
        g = C.c_double(0)
        _check(self._lib.advhmm_fp64_add_peak(self._h, C.addressof(g)))
        return g.value

    def close(self):
        if getattr(self, "_h", None):
            self._lib.advhmm_context_destroy(self._h)
            self._h = None

    def compile_loci(self, columns, n_threads=0, lo=0, hi=None):
        """``advhmm_models_create_for_loci``: the read-matcher models of a batch of loci, compiled and
        uploaded natively.  -> list of :class:`DeviceModel`."""
        hi = columns.n if hi is None else hi
        n = hi - lo
        handles = (C.c_void_p * max(n, 1))()
        d = columns.desc(lo, hi)
        _check(self._lib.advhmm_models_create_for_loci(self._h, C.byref(d), int(n_threads), handles))
        return [DeviceModel.from_handle(self, handles[i]) for i in range(n)]

    # -- many loci in one call -----------------------------------------------------------
    def viterbi_multi(self, models, groups, both_strands=False, want_path=True,
                      force_generic=False, path_cap=None, precision="fp64", want_summary=False):
        """``groups[g]`` = list of uint8 code arrays decoded against ``models[g]``."""
        flat_codes = [c for grp in groups for c in grp]
        goff = np.zeros(len(groups) + 1, dtype=np.int64)
        np.cumsum([len(g) for g in groups], out=goff[1:])
        seqs, off = pack_reads(flat_codes)
        return self._run(models, goff, seqs, off, both_strands, want_path, force_generic, path_cap,
                         fp32=(precision == "fp32"), want_summary=want_summary)

    def _run(self, models, goff, seqs, off, both_strands, want_path, force_generic, path_cap, fp32=False,
             want_summary=False):
        R = len(off) - 1
        strands = 2 if both_strands else 1
        n_out = R * strands
        flags = (WANT_PATH if want_path else 0) | (BOTH_STRANDS if both_strands else 0) | \
                (FORCE_GENERIC if force_generic else 0) | (FP32 if fp32 else 0) | \
                (WANT_SUMMARY if want_summary else 0)
        summ = np.zeros(n_out if want_summary else 0, dtype=SUMMARY_DTYPE)
        handles = (C.c_void_p * len(models))(*[m._h for m in models])
        logp = np.empty(n_out, dtype=np.float64)
        plen = np.full(n_out, -1, dtype=np.int32)
        poff = np.zeros(n_out, dtype=np.int64)
        total = C.c_int64(0)
        if want_path and path_cap is None:
            lens = np.repeat(off[1:] - off[:-1], strands)
            extra = max(m.path_extra for m in models) if models else 0
            path_cap = int(lens.sum() + n_out * extra) + 16
        cap = int(path_cap or 0)
        while True:
            path = np.empty(max(cap, 1), dtype=np.int32)
            rc = self._lib.advhmm_viterbi_multi_summary(
                self._h, handles, len(models), goff.ctypes.data, seqs.ctypes.data, off.ctypes.data, R,
                flags, logp.ctypes.data, plen.ctypes.data, poff.ctypes.data, path.ctypes.data, cap,
                C.byref(total), summ.ctypes.data if want_summary else None)
            if rc == ECAPACITY and total.value > cap:
                cap = int(total.value) + 16
                continue
            _check(rc)
            break
        return ViterbiResult(logp, plen, poff, path[:max(int(total.value), 0)], summ if want_summary else None)


class ViterbiResult(object):
    """Arrays returned by one batched decode; ``path(i)`` is read i's state-index path."""

    __slots__ = ("logp", "path_len", "path_off", "paths", "summaries")

    def __init__(self, logp, path_len, path_off, paths, summaries=None):
        self.logp, self.path_len, self.path_off, self.paths = logp, path_len, path_off, paths
        self.summaries = summaries      # SUMMARY_DTYPE records (on-device path reducers) or None

    def flank_match_rate(self, i, accuracy_filter=False):
        """get_flanking_regions_matching_rate (hmm_utils.py:209-268) from the device summary."""
        s = self.summaries[i]
        empty = 0.00001 if accuracy_filter else 1
        right = float(s["right_hits"]) / s["right_bp"] if s["right_bp"] else empty
        left = float(s["left_hits"]) / s["left_bp"] if s["left_bp"] else empty
        return min(right, left)

    def __len__(self):
        return len(self.logp)

    def path(self, i):
        n = int(self.path_len[i])
        if n < 0:
            return None
        o = int(self.path_off[i])
        return self.paths[o:o + n]


class DeviceModel(object):
    """A baked model uploaded to one context (``advhmm_model``)."""

    def __init__(self, ctx, baked):
        self._lib = load_library()
        self.ctx = ctx
        # keep the arrays alive for the duration of the create call
        in_off = np.ascontiguousarray(baked["in_off"], dtype=np.int32)
        in_src = np.ascontiguousarray(baked["in_src"], dtype=np.int32)
        in_logp = np.ascontiguousarray(baked["in_logp"], dtype=np.float64)
        emis = np.ascontiguousarray(baked["emis"], dtype=np.float64)
        S = int(baked["silent_start"])
        K = int(emis.shape[1]) if emis.ndim == 2 and S else 4
        desc = ModelDesc(int(baked["n_states"]), S, int(baked["start_index"]), int(baked["end_index"]),
                         int(baked["finite"]), K, in_off.ctypes.data, in_src.ctypes.data,
                         in_logp.ctypes.data, emis.ctypes.data)
        h = C.c_void_p()
        _check(self._lib.advhmm_model_create(ctx._h, C.byref(desc), C.byref(h)))
        self._h = h
        info = ModelInfo()
        _check(self._lib.advhmm_model_info_get(h, C.byref(info)))
        self.info = info
        self.n_states = int(baked["n_states"])
        self.n_symbols = K
        # upper bound on (path length - read length): every silent state once, plus ends
        self.path_extra = self.n_states - S + 2

    @classmethod
    def from_baked(cls, baked, ctx=None):
        return cls(ctx or Context.default(), baked)

    @classmethod
    def from_handle(cls, ctx, handle):
        """Wrap an ``advhmm_model*`` made by ``advhmm_models_create_for_loci``."""
        self = cls.__new__(cls)
        self._lib = load_library()
        self.ctx = ctx
        self._h = C.c_void_p(handle)
        info = ModelInfo()
        _check(self._lib.advhmm_model_info_get(self._h, C.byref(info)))
        self.info = info
        dims = self.dims()
        self.n_states = dims.n_states
        self.n_symbols = dims.n_symbols
        self.path_extra = dims.n_states - dims.silent_start + 2
        return self

    def dims(self):
        d = ModelDims()
        _check(self._lib.advhmm_model_dims_get(self._h, C.byref(d)))
        return d

    def tables(self):
        """The baked arrays (+ state names) of a locus model, as ``pomegranate.bake`` would hold them."""
        d = self.dims()
        in_off = np.empty(d.n_states + 1, dtype=np.int32)
        in_src = np.empty(d.n_edges, dtype=np.int32)
        in_logp = np.empty(d.n_edges, dtype=np.float64)
        emis = np.empty((d.silent_start, d.n_symbols), dtype=np.float64)
        names = C.create_string_buffer(int(d.names_bytes) + 1)
        _check(self._lib.advhmm_model_tables_get(self._h, in_off.ctypes.data, in_src.ctypes.data, in_logp.ctypes.data,
                                                 emis.ctypes.data, names))
        return {"n_states": d.n_states, "silent_start": d.silent_start, "start_index": d.start_index,
                "end_index": d.end_index, "finite": d.finite, "in_off": in_off, "in_src": in_src,
                "in_logp": in_logp, "emis": emis, "alphabet": "ACGT",
                "names": names.raw[:int(d.names_bytes)].decode("ascii").split("\n")[:-1], "shape": tuple(d.shape)}

    def banded_tables(self, n_states, silent_start, n_edges):
        """Host copies of what the banded kernels read for this model (tests)."""
        nb = C.c_int64(0)
        _check(self._lib.advhmm_model_banded_tables_get(self._h, None, 0, None, None, None, None, None, C.addressof(nb)))
        image = np.zeros(nb.value, dtype=np.uint8)
        tb1 = np.zeros(4 * silent_start, dtype=np.int32)
        tb0 = np.zeros(n_states, dtype=np.int32)
        fin_w = np.full(n_edges, np.nan)
        classes = np.zeros(n_states, dtype=np.uint8)
        empty = C.c_double(0)
        _check(self._lib.advhmm_model_banded_tables_get(self._h, image.ctypes.data, nb.value, tb1.ctypes.data,
                                                        tb0.ctypes.data, fin_w.ctypes.data, classes.ctypes.data,
                                                        C.addressof(empty), C.addressof(nb)))
        return {"image": image, "tb1": tb1, "tb0": tb0, "fin_w": fin_w[~np.isnan(fin_w)], "classes": classes,
                "logp_empty": empty.value}

    @property
    def kind(self):
        return "banded" if self.info.kind == KIND_BANDED else "generic"

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self._lib.advhmm_model_destroy(self._h)
        self._h = None

    def set_state_classes(self, classes):
        """Per-state class bytes for the on-device path reducers (see include/advhmm.h)."""
        cls = np.ascontiguousarray(classes, dtype=np.uint8)
        if len(cls) != self.n_states:
            raise ValueError("need one class byte per state")
        _check(self._lib.advhmm_model_set_state_classes(self._h, cls.ctypes.data))

    def viterbi(self, codes, both_strands=False, want_path=True, precision="fp64",
                force_generic=False, path_cap=None, want_summary=False):
        if precision not in ("fp64", "fp32"):
            raise EngineError(EUNSUPPORTED, "precision must be 'fp64' (default, bit-exact) or 'fp32'")
        seqs, off = pack_reads(codes)
        goff = np.array([0, len(codes)], dtype=np.int64)
        return self.ctx._run([self], goff, seqs, off, both_strands, want_path, force_generic, path_cap,
                             fp32=(precision == "fp32"), want_summary=want_summary)

    def viterbi_packed(self, seqs, off, both_strands=False, want_path=True, precision="fp64",
                       force_generic=False, path_cap=None, want_summary=False):
        """``viterbi`` for reads already flattened to (uint8 codes, int64 offsets)."""
        goff = np.array([0, len(off) - 1], dtype=np.int64)
        return self.ctx._run([self], goff, seqs, off, both_strands, want_path, force_generic, path_cap,
                             fp32=(precision == "fp32"), want_summary=want_summary)

    def log_probability(self, codes, force_generic=False):
        """Forward log-probabilities (``hmm.pyx:1258-1313``): profile-shaped models run on the banded kernels
        (register wavefront up to 320 bases, the striped long-read kernel beyond and for models larger than
        shared memory), anything else -- or everything with ``force_generic`` -- on the generic kernel."""
        seqs, off = pack_reads(codes)
        R = len(codes)
        logp = np.empty(R, dtype=np.float64)
        _check(self._lib.advhmm_log_probability_batch(self._h, seqs.ctypes.data, off.ctypes.data, R,
                                                      FORCE_GENERIC if force_generic else 0, logp.ctypes.data))
        return logp

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceKeywordFilter(object):
    """Keywords (any mix of lengths) on one context (``advhmm_kfilter``)."""

    def __init__(self, ctx, keywords, locus_ids):
        self._lib = load_library()
        self.ctx = ctx
        raw = "".join(keywords).encode("ascii", "replace")
        off = np.zeros(len(keywords) + 1, dtype=np.int64)
        if len(keywords):
            np.cumsum(np.fromiter(map(len, keywords), dtype=np.int64, count=len(keywords)), out=off[1:])
        ids = np.ascontiguousarray(locus_ids, dtype=np.int32)
        if len(ids) != len(keywords):
            raise ValueError("one locus id per keyword")
        h = C.c_void_p()
        _check(self._lib.advhmm_kfilter_create(ctx._h, len(keywords), raw, off.ctypes.data, ids.ctypes.data, C.byref(h)))
        self._h = h

    def scan(self, flat, off, min_matches):
        """flat: ASCII bytes of all reads back to back (uint8 array); off: int64 offsets.  Returns
        the arrays (read index, locus id, occurrences) of the pairs with occurrences >= min_matches."""
        R = len(off) - 1
        cap = max(1024, 2 * R)
        total = C.c_int64(0)
        while True:
            hr, hl, hc = (np.empty(cap, dtype=np.int32) for _ in range(3))
            rc = self._lib.advhmm_kfilter_scan(self._h, flat.ctypes.data, off.ctypes.data, R, int(min_matches), 0,
                                               hr.ctypes.data, hl.ctypes.data, hc.ctypes.data, cap, C.byref(total))
            if rc == ECAPACITY and total.value > cap:
                cap = int(total.value) + 16
                continue
            _check(rc)
            n = int(total.value)
            return hr[:n], hl[:n], hc[:n]

    def scan_device(self, d_seqs_ptr, off, n_reads, min_matches, d_hit_read, d_hit_locus, d_hit_count, cap, d_n_hits):
        """Device-resident form (``ADVHMM_DEVICE_BUFFERS``): raw device pointers, results stay on the
        device.  ``off``: host int64 array, or an int (device pointer, ``ADVHMM_DEVICE_OFFSETS``)."""
        flags = DEVICE_BUFFERS
        if isinstance(off, int):
            flags |= DEVICE_OFFSETS
            off_ptr = off
        else:
            off_ptr = off.ctypes.data
        _check(self._lib.advhmm_kfilter_scan(self._h, d_seqs_ptr, off_ptr, int(n_reads), int(min_matches),
                                             flags, d_hit_read, d_hit_locus, d_hit_count, int(cap), d_n_hits))

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self._lib.advhmm_kfilter_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
