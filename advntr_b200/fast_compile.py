"""The read-matcher HMM of a locus through the NATIVE compiler.

``read_matcher.get_read_matcher_model`` follows the reference literally: three sub-models, two
matrix round trips and eight ``bake`` calls per locus in Python (0.1 s here, 1 s in the reference).
But for a given *shape* -- flank lengths, repeat-unit match columns R, unrolled copies C -- every
locus has the same states and the same edges in the same order; only the numbers differ.  The
engine's C++ side (``csrc/locus_compile.hpp``, ``advhmm_models_create_for_loci``) therefore builds
the structure of a shape once (symbolically, by the reference's own sequence of graph operations),
evaluates the few hundred parameters of a locus through the same chain of float operations
(libm ``log``, ``numpy.exp`` through a callback) and writes the device tables of whole batches of
loci with all host threads.  This module is the Python face of it:

* :class:`CompiledHMM` -- same decoding surface as ``pomegranate.HiddenMarkovModel``; the baked
  tables and the ``State`` list are fetched from the library on demand;
* :func:`get_read_matcher_model` / :func:`build_vntr_matcher_hmm` -- drop-ins for the literal builders
  (no ``vpaths``); :func:`compile_many` -- ONE native call for the models of many loci.

Exactness is tested, not assumed: ``tests/test_fast_compile.py`` and ``tests/test_native_compile.py``
compare names, CSR arrays and the bit patterns of every table entry with the literal build over
loci, shapes (down to one-column flanks and repeat units) and gapped alignments.
"""
from __future__ import annotations

from . import engine, read_matcher
from .pomegranate import HiddenMarkovModel, State, DiscreteDistribution

_host_ctx = None
_states_by_shape = {}


def _host_context():
    """A context without a device: compiles tables (for ``.baked``) where no GPU is involved."""
    global _host_ctx
    if _host_ctx is None:
        _host_ctx = engine.Context(device=-1)
    return _host_ctx


class LocusModelSpec(object):
    """What defines the read matcher of a locus: the flank bases that enter the model, the aligned
    repeat segments, the unrolled copies and the error rate."""

    __slots__ = ("left", "right", "aligned", "copies", "error_rate")

    def __init__(self, left, right, aligned, copies, error_rate):
        self.left, self.right, self.aligned = left, right, list(aligned)
        self.copies, self.error_rate = int(copies), float(error_rate)


def _columns(specs):
    return engine.LociColumns.from_lists([s.left for s in specs], [s.right for s in specs],
                                         [s.aligned for s in specs], [s.copies for s in specs],
                                         [s.error_rate for s in specs])


class CompiledHMM(HiddenMarkovModel):
    """A baked read-matcher model made by the native compiler: same decoding surface."""

    def __init__(self, spec, device_model=None):
        self.name = "Read Matcher"
        self.spec = spec
        self._engine = device_model          # engine.DeviceModel on a device context (decoding)
        self._host_model = None              # engine.DeviceModel on the host-only context (tables)
        self._tables = None
        self._states = None
        self.d = 1
        self.discrete = 1
        self.multivariate = 0
        self.keymap = [{c: i for i, c in enumerate("ACGT")}]
        self.graph = None
        self._pseudo = {}

    # -- tables / states on demand -----------------------------------------------------------------
    def _any_model(self):
        if self._engine is not None:
            return self._engine
        if self._host_model is None:
            self._host_model = _host_context().compile_loci(_columns([self.spec]))[0]
        return self._host_model

    @property
    def _baked(self):
        if self._tables is None:
            self._tables = self._any_model().tables()
        return self._tables

    @property
    def states(self):
        if self._states is None:
            t = self._baked
            cached = _states_by_shape.get(t["shape"])
            if cached is None:
                dummy = DiscreteDistribution({"A": 0.25, "C": 0.25, "G": 0.25, "T": 0.25})
                S = t["silent_start"]
                cached = _states_by_shape[t["shape"]] = [State(dummy if i < S else None, name=nm)
                                                         for i, nm in enumerate(t["names"])]
            self._states = cached
        return self._states

    n_states = property(lambda self: self._baked["n_states"])
    n_edges = property(lambda self: len(self._baked["in_src"]))
    silent_start = property(lambda self: self._baked["silent_start"])
    start_index = property(lambda self: self._baked["start_index"])
    end_index = property(lambda self: self._baked["end_index"])
    finite = property(lambda self: self._baked["finite"])
    start = property(lambda self: self.states[self.start_index])
    end = property(lambda self: self.states[self.end_index])

    def bake(self, verbose=False, merge="All"):
        raise ValueError("a compiled model is already baked")

    def _device_model(self):
        if self._engine is None:
            self._engine = engine.Context.default().compile_loci(_columns([self.spec]))[0]
        return self._engine

    def _release_engine(self):
        for attr in ("_engine", "_host_model"):
            dm = getattr(self, attr, None)
            if dm is not None:
                dm.close()
                setattr(self, attr, None)


def _spec(left, right, segments, copies, error_rate):
    return LocusModelSpec(left, right, read_matcher.align_repeat_segments([s.upper() for s in segments]),
                          copies, error_rate)


def get_read_matcher_model(left, right, segments, copies, error_rate=read_matcher.DEFAULT_MAX_ERROR_RATE):
    """Drop-in for ``read_matcher.get_read_matcher_model`` (no ``vpaths``): same tables, compiled natively.
    ``segments``: equal-length repeat segments, or an alignment of them (strings over ``ACGT-``)."""
    model = CompiledHMM(_spec(left, right, segments, copies, error_rate))
    if error_rate > 0 and len(left) > 0 and len(right) > 0 and copies >= 1:
        # every transition probability of the builders is positive then (pseudocounts), the structure is the
        # shape's: nothing is compiled until the tables or a decode are asked for -- or until compile_many /
        # attach_device_models compiles the models of many loci in one native call
        return model
    try:
        model._any_model()
    except engine.EngineError as e:
        if e.code != engine.EUNSUPPORTED:          # zero-probability transition: the structure differs
            raise ValueError(str(e))
        return read_matcher.get_read_matcher_model(left, right, segments, copies, error_rate=error_rate)
    return model


def build_vntr_matcher_hmm(left_flank, right_flank, repeat_segments, copies, flank_size=100,
                           error_rate=read_matcher.DEFAULT_MAX_ERROR_RATE):
    return get_read_matcher_model(left_flank[-flank_size:], right_flank[:flank_size],
                                  repeat_segments, copies, error_rate=error_rate)


def compile_many(loci, ctx=None, n_threads=0):
    """The models of many loci with ONE native call (``advhmm_models_create_for_loci``): parsing the
    repeat segments, parameter chains, device tables and the upload run in the library on all host
    threads.  ``loci``: iterable of ``(left, right, segments, copies, error_rate)`` with the flanks
    already trimmed to the flank size.  -> list of :class:`CompiledHMM` with their device models."""
    ctx = ctx or engine.Context.default()
    specs = [_spec(*l) for l in loci]
    models = ctx.compile_loci(_columns(specs), n_threads=n_threads)
    return [CompiledHMM(s, dm) for s, dm in zip(specs, models)]


def attach_device_models(models, ctx=None, n_threads=0):
    """Give every :class:`CompiledHMM` of ``models`` that has none its device model, in one native call."""
    ctx = ctx or engine.Context.default()
    todo = [m for m in models if isinstance(m, CompiledHMM) and m._engine is None]
    if todo:
        for m, dm in zip(todo, ctx.compile_loci(_columns([m.spec for m in todo]), n_threads=n_threads)):
            m._engine = dm
