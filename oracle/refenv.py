"""Loader for the compiled reference engine and the reference's own Python builders.

TEST INFRASTRUCTURE ONLY -- imported by ``tests/``, ``tests/golden/make_golden.py``,
``__graft_entry__.smoke`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs,
never by ``advntr_b200``.

* ``reference_pomegranate()``  -> the compiled, unmodified vendored pomegranate
  (``oracle/_ref/pomegranate/*.so``, built by ``oracle/build_ref.py``) bound to the
  networkx-1.11 shim in ``oracle/nx111``.  Works on the GPU box too (the ``.so``
  files travel; ``/root/reference`` does not).
* ``reference_hmm_utils(backend)`` -> the reference's ``advntr/hmm_utils.py`` loaded
  from ``/root/reference`` (THIS container only) with its ``from pomegranate import``
  bound to ``backend`` (the compiled reference, or ``advntr_b200.pomegranate`` to
  prove that the reference's builders run unmodified on the new engine).
  ``Bio`` (MUSCLE wrapper) is stubbed with the identity alignment, valid for the
  equal-length repeat segments every synthetic locus here uses.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("ADVNTR_REFERENCE", "/root/reference")
_REF_POM = None


def have_reference_engine() -> bool:
    sys.path.insert(0, HERE)
    try:
        import build_ref
        return build_ref.have_ref()
    finally:
        sys.path.remove(HERE)


def have_reference_sources() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "advntr", "hmm_utils.py"))


def _nx_shim():
    name = "_advntr_oracle_nx111"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(HERE, "nx111", "networkx", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def reference_pomegranate():
    """Import ``oracle/_ref/pomegranate`` with ``networkx`` resolved to the 1.11 shim."""
    global _REF_POM
    if _REF_POM is not None:
        return _REF_POM
    if not have_reference_engine():
        raise ImportError("oracle/_ref is not built (run: python oracle/build_ref.py)")
    saved_nx = sys.modules.get("networkx")
    saved_pom = sys.modules.get("pomegranate")
    sys.modules["networkx"] = _nx_shim()
    sys.modules.pop("pomegranate", None)
    sys.path.insert(0, os.path.join(HERE, "_ref"))
    try:
        mod = importlib.import_module("pomegranate")
        for sub in ("hmm", "base", "distributions", "utils"):
            importlib.import_module("pomegranate." + sub)
    finally:
        sys.path.remove(os.path.join(HERE, "_ref"))
        if saved_nx is not None:
            sys.modules["networkx"] = saved_nx
        else:
            sys.modules.pop("networkx", None)
        # keep the compiled package importable under a private name only
        for k in [k for k in sys.modules if k == "pomegranate" or k.startswith("pomegranate.")]:
            sys.modules["_advntr_ref_" + k] = sys.modules[k]
        # (pomegranate.* stay registered: the extension modules cimport each other by name)
        if saved_pom is not None:
            sys.modules["pomegranate"] = saved_pom
    _REF_POM = mod
    return mod


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def _install_bio_stubs():
    """Stand-ins for Biopython's MUSCLE wrapper (``profile_hmm.py:6-7,165-171``).

    MUSCLE (an external binary) is absent, so multi-segment profiles are only built
    for EQUAL-LENGTH repeat segments, whose multiple alignment is the identity; the
    stub hands the sequences back unaligned and refuses ragged input.  The profile
    counts are row-order invariant, so MUSCLE's output ordering does not matter.
    (MUSCLE output itself is therefore "parity unpinned", see DESIGN.md.)
    """
    if "Bio" in sys.modules:
        return

    class _Rec(object):
        def __init__(self, seq):
            self.seq = seq

    def _muscle(*a, **k):
        def run(stdin=None):
            return stdin, ""
        return run

    def _read(handle, fmt):
        seqs = [ln.strip() for ln in handle.read().splitlines()
                if ln.strip() and not ln.startswith(">")]
        if len(set(len(x) for x in seqs)) != 1:
            raise RuntimeError("MUSCLE is not available: only equal-length repeat "
                               "segments can be profiled (oracle stub)")
        return [_Rec(x) for x in seqs]

    bio = _stub("Bio")
    align = _stub("Bio.Align")
    apps = _stub("Bio.Align.Applications", MuscleCommandline=_muscle)
    alignio = _stub("Bio.AlignIO", read=_read)
    bio.Align, bio.AlignIO, align.Applications = align, alignio, apps
    sys.modules.update({"Bio": bio, "Bio.Align": align,
                        "Bio.Align.Applications": apps, "Bio.AlignIO": alignio})


def reference_settings():
    """The reference's ``advntr.settings`` module (MAX_ERROR_RATE lives there)."""
    if not have_reference_sources():
        raise ImportError("reference sources not present")
    if REF_ROOT not in sys.path:
        sys.path.append(REF_ROOT)  # appended: never shadows anything of ours
    return importlib.import_module("advntr.settings")


def reference_hmm_utils(backend, tag: str):
    """Load the reference's hmm_utils.py bound to ``backend`` as its pomegranate."""
    name = "_advntr_ref_hmm_utils_" + tag
    if name in sys.modules:
        return sys.modules[name]
    reference_settings()
    _install_bio_stubs()
    saved = sys.modules.get("pomegranate")
    sys.modules["pomegranate"] = backend
    try:
        importlib.import_module("advntr.profile_hmm")
        spec = importlib.util.spec_from_file_location(
            name, os.path.join(REF_ROOT, "advntr", "hmm_utils.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    finally:
        if saved is not None:
            sys.modules["pomegranate"] = saved
        else:
            sys.modules.pop("pomegranate", None)
    return mod


def _install_io_stubs():
    """Import-time stand-ins for the IO / ML packages ``advntr/vntr_finder.py`` pulls in at module
    level (pysam, keras, Biopython's Seq/SeqIO/pairwise2).  None of them is on the hot path; only
    ``Seq(...).reverse_complement()`` is actually exercised (``vntr_finder.py:241``)."""
    _install_bio_stubs()
    bio = sys.modules["Bio"]
    if hasattr(bio, "Seq"):
        return
    comp = str.maketrans("ACGTacgt", "TGCAtgca")

    class Seq(str):
        def reverse_complement(self):
            return Seq(str(self).translate(comp)[::-1])

    class SeqRecord(object):
        def __init__(self, seq=None, id=None, **kw):
            self.seq, self.id = seq, id

    seqmod = _stub("Bio.Seq", Seq=Seq)
    seqio = _stub("Bio.SeqIO", SeqRecord=SeqRecord, parse=lambda *a, **k: iter(()), write=lambda *a, **k: 0)
    pw = _stub("Bio.pairwise2", align=None)
    recmod = _stub("Bio.SeqRecord", SeqRecord=SeqRecord)
    bio.Seq, bio.SeqIO, bio.pairwise2, bio.SeqRecord = seqmod, seqio, pw, recmod
    sys.modules.update({"Bio.Seq": seqmod, "Bio.SeqIO": seqio, "Bio.pairwise2": pw, "Bio.SeqRecord": recmod})
    if "pysam" not in sys.modules:
        sys.modules["pysam"] = _stub("pysam", AlignmentFile=None)
    if "keras" not in sys.modules:
        keras = _stub("keras")
        km = _stub("keras.models", Sequential=None, load_model=None)
        kl = _stub("keras.layers", Dense=None, Activation=None)
        keras.models, keras.layers = km, kl
        sys.modules.update({"keras": keras, "keras.models": km, "keras.layers": kl})


def reference_vntr_finder(backend, tag: str):
    """The reference's ``advntr/vntr_finder.py`` (THIS container only) bound to ``backend``."""
    name = "_advntr_ref_vntr_finder_" + tag
    if name in sys.modules:
        return sys.modules[name]
    reference_settings()
    _install_io_stubs()
    saved = sys.modules.get("pomegranate")
    saved_hu = sys.modules.pop("advntr.hmm_utils", None)
    sys.modules["pomegranate"] = backend
    try:
        spec = importlib.util.spec_from_file_location(
            name, os.path.join(REF_ROOT, "advntr", "vntr_finder.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    finally:
        sys.modules.pop("advntr.hmm_utils", None)
        if saved_hu is not None:
            sys.modules["advntr.hmm_utils"] = saved_hu
        if saved is not None:
            sys.modules["pomegranate"] = saved
        else:
            sys.modules.pop("pomegranate", None)
    return mod
