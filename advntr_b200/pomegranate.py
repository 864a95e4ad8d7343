"""Drop-in model-builder surface for adVNTR's hot path, backed by the B200 engine.

This module exposes the class surface adVNTR imports from its vendored pomegranate
(``/root/reference/advntr/hmm_utils.py:4-5``, ``vntr_finder.py:24``)::

    from advntr_b200.pomegranate import DiscreteDistribution, State, HiddenMarkovModel

so that ``hmm_utils.py`` / ``vntr_finder.py`` run with only that import swapped.  The
graph container and ``bake()`` are host-side Python (model compilation is not the hot
path); ``viterbi`` / ``log_probability`` and their batched forms run on the GPU through
the C-ABI library (``advntr_b200/engine.py`` -> ``libadvhmm.so``).  There is NO CPU
fallback: without the CUDA library the decoding calls raise.

Reference semantics restated here (citations into ``/root/reference``):

* ``State`` -- ``pomegranate/base.pyx:362-412`` (silent <=> ``distribution is None``;
  identity hashing).
* ``DiscreteDistribution`` -- ``pomegranate/distributions.pyx:1270-1288, 1366-1406``
  (``log_dist[c] = log(p)`` or -inf for p <= 0).
* ``HiddenMarkovModel.add_transition`` -- ``pomegranate/hmm.pyx:392-434`` (stores
  ``log(p)``; re-adding an edge updates it in place, keeping its position).
* ``bake(merge=None)`` -- ``hmm.pyx:844-1123``: emitting states sorted by name, silent
  states sorted by name then put in networkx-1.11 DFS topological order
  (SURVEY.md appendix A), CSR in-edge lists filled by walking edges in node-insertion
  x successor-insertion order.
* ``dense_transition_matrix`` -- ``hmm.pyx:492-514`` (``numpy.exp`` of the log matrix).
* ``from_matrix`` -- ``hmm.pyx:3147-3238`` including its wiring of the LAST state of the
  list (not the indexed one) to the new end state (``:3233-3235``).
* ``concatenate`` -- ``hmm.pyx:584-615``.
* ``viterbi`` / ``log_probability`` error behaviour -- ``hmm.pyx:57-81, 1911-1967,
  1258-1298``.
"""
from __future__ import annotations

import json
import math
import uuid

import numpy as np

NEGINF = float("-inf")

__all__ = ["DiscreteDistribution", "State", "HiddenMarkovModel"]


def _log(x) -> float:
    """log(x), or -inf for x <= 0 (``pomegranate/utils.pyx:64-70``; libm ``log``)."""
    x = float(x)
    return math.log(x) if x > 0 else NEGINF


class DiscreteDistribution(object):
    """Symbol -> probability table; only the log table is used on the hot path."""

    name = "DiscreteDistribution"
    d = 1

    def __init__(self, characters, frozen=False):
        if not isinstance(characters, dict):
            raise TypeError("DiscreteDistribution needs a dict of symbol -> probability")
        self.frozen = bool(frozen)
        self.dist = dict(characters)
        self.log_dist = {k: _log(v) for k, v in characters.items()}

    @property
    def parameters(self):
        return [self.dist]

    def keys(self):
        return tuple(self.dist.keys())

    def items(self):
        return tuple(self.dist.items())

    def values(self):
        return tuple(self.dist.values())

    def __len__(self):
        return len(self.dist)

    def log_probability(self, symbol):
        return self.log_dist.get(symbol, NEGINF)

    def probability(self, symbol):
        return math.exp(self.log_probability(symbol))

    def copy(self):
        return DiscreteDistribution(self.dist, self.frozen)

    def to_json(self, separators=(",", " : "), indent=4):
        return json.dumps({"class": "Distribution", "name": self.name,
                           "parameters": [self.dist], "frozen": self.frozen},
                          separators=separators, indent=indent)

    @classmethod
    def from_json(cls, s):
        d = json.loads(s)
        return cls(d["parameters"][0], d.get("frozen", False))


class State(object):
    """A node of the HMM graph.  Silent iff it has no distribution."""

    def __init__(self, distribution, name=None, weight=None):
        self.distribution = distribution
        self.name = name or str(uuid.uuid4())
        self.weight = weight or 1.0

    def is_silent(self):
        return self.distribution is None

    def tie(self, state):
        state.distribution = self.distribution

    def tied_copy(self):
        return State(distribution=self.distribution, name=self.name + "-tied")

    def copy(self):
        return State(distribution=self.distribution.copy(), name=self.name)

    def to_json(self, separators=(",", " : "), indent=4):
        return json.dumps({"class": "State",
                           "distribution": None if self.is_silent()
                           else json.loads(self.distribution.to_json()),
                           "name": self.name, "weight": self.weight},
                          separators=separators, indent=indent)

    @classmethod
    def from_json(cls, s):
        d = json.loads(s)
        if d["class"] != "State":
            raise IOError("State object attempting to decode {} object".format(d["class"]))
        if d["distribution"] is None:
            return cls(None, str(d["name"]), d["weight"])
        return cls(DiscreteDistribution.from_json(json.dumps(d["distribution"])),
                   str(d["name"]), d["weight"])

    def __repr__(self):
        return "State(%r)" % (self.name,)


class _OrderedDiGraph(object):
    """Insertion-ordered adjacency container (node -> {successor -> log p}).

    Iteration orders are part of the parity contract: they fix the in-edge order
    of every state after ``bake`` and hence Viterbi tie-breaking.
    """

    __slots__ = ("succ",)

    def __init__(self):
        self.succ = {}

    def add_node(self, n):
        if n not in self.succ:
            self.succ[n] = {}

    def add_edge(self, a, b, logp):
        self.add_node(a)
        self.add_node(b)
        self.succ[a][b] = logp  # existing key keeps its slot

    def nodes(self):
        return list(self.succ)

    def edges(self):
        for a, nbrs in self.succ.items():
            for b, w in nbrs.items():
                yield a, b, w

    def number_of_edges(self):
        return sum(len(v) for v in self.succ.values())

    def remove_node(self, n):
        del self.succ[n]
        for nbrs in self.succ.values():
            nbrs.pop(n, None)

    def remove_edge(self, a, b):
        del self.succ[a][b]

    @staticmethod
    def disjoint_union(g, h):
        """G's nodes, G's edges, then H's nodes, H's edges (networkx-1.11 ``union``)."""
        r = _OrderedDiGraph()
        for part in (g, h):
            for n in part.succ:
                if n in r.succ and part is h:
                    raise ValueError("The node sets of G and H are not disjoint.")
                r.add_node(n)
            for a, b, w in part.edges():
                r.add_edge(a, b, w)
        return r


def _dfs_topological_order(nodes, succ):
    """Topological order of ``nodes`` exactly as networkx 1.11 produces it.

    Iterative DFS seeded in ``nodes`` order; every unexplored successor is pushed
    (so the last one is visited first); a node is emitted when it has no
    unexplored successor; the reversed post-order is returned.
    """
    members = set(nodes)
    explored, seen, post = set(), set(), []
    for v in nodes:
        if v in explored:
            continue
        stack = [v]
        while stack:
            w = stack[-1]
            if w in explored:
                stack.pop()
                continue
            seen.add(w)
            fresh = []
            for n in succ[w]:
                if n in members and n not in explored:
                    if n in seen:
                        raise ValueError("Graph contains a cycle of silent states.")
                    fresh.append(n)
            if fresh:
                stack.extend(fresh)
            else:
                explored.add(w)
                post.append(w)
                stack.pop()
    post.reverse()
    return post


class HiddenMarkovModel(object):
    """Graph container + ``bake`` on the host, decoding on the B200 engine."""

    model = "HiddenMarkovModel"

    def __init__(self, name=None, start=None, end=None):
        self.name = str(name) or str(id(self))
        self.graph = _OrderedDiGraph()
        self.start = start or State(None, name=self.name + "-start")
        self.end = end or State(None, name=self.name + "-end")
        self.graph.add_node(self.start)
        self.graph.add_node(self.end)
        self.d = 0
        self.n_states = 0
        self.n_edges = 0
        self.discrete = 0
        self.multivariate = 0
        self.states = []
        self.start_index = 0
        self.end_index = 0
        self.silent_start = 0
        self.finite = 0
        self.keymap = []
        self._state_names = set()
        self._pseudo = {}      # (a, b) -> edge pseudocount (hmm.pyx:432-434); only to_json reads it
        self._baked = None     # dict of numpy arrays (see bake)
        self._engine = None    # lazily created device model handle

    # ------------------------------------------------------------------ building
    def add_state(self, state):
        if state.name in self._state_names:
            raise ValueError("A state with name '{}' already exists".format(state.name))
        self.graph.add_node(state)
        self._state_names.add(state.name)

    def add_states(self, *states):
        for s in states:
            if isinstance(s, (list, tuple)):
                for x in s:
                    self.add_state(x)
            else:
                self.add_state(s)

    def add_transition(self, a, b, probability, pseudocount=None, group=None):
        self._pseudo[(a, b)] = pseudocount or probability
        self.graph.add_edge(a, b, _log(probability))

    def add_transitions(self, a, b, probabilities, pseudocounts=None, groups=None):
        if isinstance(a, list) and isinstance(b, list):
            for x, y, p in zip(a, b, probabilities):
                self.add_transition(x, y, p)
        elif isinstance(a, list):
            for x, p in zip(a, probabilities):
                self.add_transition(x, b, p)
        else:
            for y, p in zip(b, probabilities):
                self.add_transition(a, y, p)

    def add_model(self, other):
        self.graph = _OrderedDiGraph.disjoint_union(self.graph, other.graph)
        self._pseudo.update(other._pseudo)

    def concatenate(self, other, suffix="", prefix=""):
        other.name = "{}{}{}".format(prefix, other.name, suffix)
        for s in other.states:
            s.name = "{}{}{}".format(prefix, s.name, suffix)
        self.graph = _OrderedDiGraph.disjoint_union(self.graph, other.graph)
        self._pseudo.update(other._pseudo)
        self.add_transition(self.end, other.start, 1.00)
        self.end = other.end

    def state_count(self):
        return len(self.states)

    def edge_count(self):
        return self.n_edges

    def is_infinite(self):
        return self.finite == 0

    # --------------------------------------------------------------------- bake
    def _merge_pass(self, merge):
        """What ``bake`` does before ordering the states when ``merge`` is 'all' / 'partial'
        (``hmm.pyx:720-838``).  Every read-matcher model is baked with ``merge=None``; this path serves
        ``build_reference_repeat_finder_hmm`` (``hmm_utils.py:674``) and ``from_json`` (``:3143``).

        * 'all': states other than start / end without in-edges or without out-edges are removed,
          repeatedly.  The reference allocates its two degree counters ONCE and keeps adding to
          them while it re-indexes the surviving states every round (``:720-746``), so from the
          second round on a state is tested against stale sums; reproduced as is.
        * rows whose probabilities do not sum to 1 (8 decimals) are renormalised (``:766-784``);
        * a silent state with a probability-1 out-edge is merged into its successor ('partial': only
          into silent successors): its in-edges are re-pointed, the state disappears (``:789-828``)."""
        g = self.graph
        n0 = len(g.nodes())
        in_count = np.zeros(n0, dtype=np.int32)
        out_count = np.zeros(n0, dtype=np.int32)
        while merge == "all":
            removed = 0
            pre = g.nodes()
            idx = {s: i for i, s in enumerate(pre)}
            for a, b, _ in list(g.edges()):
                out_count[idx[a]] += 1
                in_count[idx[b]] += 1
            for i, s in enumerate(pre):
                if s is self.start or s is self.end:
                    continue
                if in_count[i] == 0 or out_count[i] == 0:
                    removed += 1
                    g.remove_node(s)
            if removed == 0:
                break
        for s in g.nodes():
            total = round(sum(math.e ** w for w in g.succ[s].values()), 8)
            if total != 1.0 and s is not self.end:
                for b in g.succ[s]:
                    g.succ[s][b] = g.succ[s][b] - _log(total)
        while True:
            merged = 0
            for a, b, w in list(g.edges()):
                if a not in g.succ or b not in g.succ:
                    continue
                if a is self.start or b is self.end:
                    continue
                if w == 0.0 and a.is_silent() and (merge == "all" or b.is_silent()):
                    for x, y, d in list(g.edges()):
                        if y is a:
                            merged += 1
                            g.remove_edge(x, y)
                            g.add_edge(x, b, d)
                            self._pseudo[(x, b)] = max(self._pseudo.get((a, b), 0.0), self._pseudo.get((x, y), 0.0))
                    g.remove_node(a)
            if merged == 0:
                break

    def bake(self, verbose=False, merge="All"):
        merge = merge.lower() if merge else None
        self._release_engine()
        g = self.graph
        if merge in ("all", "partial"):
            self._merge_pass(merge)
        nodes = g.nodes()
        emitting = sorted((s for s in nodes if not s.is_silent()), key=lambda s: s.name)
        silent = sorted((s for s in nodes if s.is_silent()), key=lambda s: s.name)
        silent = _dfs_topological_order(silent, g.succ)
        self.states = emitting + silent
        self.silent_start = len(emitting)
        m = len(self.states)
        index = {s: i for i, s in enumerate(self.states)}

        src, dst, wts = [], [], []
        for a, b, w in g.edges():
            src.append(index[a])
            dst.append(index[b])
            wts.append(w)
        E = len(src)
        src = np.asarray(src, dtype=np.int32).reshape(E)
        dst = np.asarray(dst, dtype=np.int32).reshape(E)
        wts = np.asarray(wts, dtype=np.float64).reshape(E)
        self.n_states, self.n_edges = m, E

        # CSR by target, stable in edge-walk order  (hmm.pyx:970-1011)
        order_in = np.argsort(dst, kind="stable")
        in_off = np.zeros(m + 1, dtype=np.int32)
        np.cumsum(np.bincount(dst, minlength=m), out=in_off[1:])
        # CSR by source, same walk order           (hmm.pyx:1013-1023)
        order_out = np.argsort(src, kind="stable")
        out_off = np.zeros(m + 1, dtype=np.int32)
        np.cumsum(np.bincount(src, minlength=m), out=out_off[1:])

        try:
            self.start_index = index[self.start]
            self.end_index = index[self.end]
        except KeyError:
            raise SyntaxError("Model.start / Model.end has been deleted from the model.")
        self.finite = 1 if in_off[self.end_index + 1] - in_off[self.end_index] > 0 else 0

        # emissions: one row per emitting state over the model alphabet (hmm.pyx:1072-1080)
        keys = []
        for s in emitting:
            for k in s.distribution.keys():
                if k not in keys:
                    keys.append(k)
        if set(keys) <= set("ACGT"):
            keys = list("ACGT")     # fixed device symbol codes A,C,G,T = 0..3
        self.keymap = [{k: i for i, k in enumerate(keys)}]
        emis = np.full((len(emitting), max(len(keys), 1)), NEGINF, dtype=np.float64)
        for i, s in enumerate(emitting):
            ld = s.distribution.log_dist
            w = _log(s.weight)   # state weight, log(1) = 0 (hmm.pyx:928-930, 1993-1997)
            for k, j in self.keymap[0].items():
                emis[i, j] = ld.get(k, NEGINF) + w
        self.d = 1 if emitting else 0
        self.discrete = 1
        self._baked = {
            "n_states": m, "silent_start": self.silent_start,
            "start_index": self.start_index, "end_index": self.end_index,
            "finite": self.finite,
            "in_off": in_off, "in_src": src[order_in].copy(), "in_logp": wts[order_in].copy(),
            "out_off": out_off, "out_dst": dst[order_out].copy(), "out_logp": wts[order_out].copy(),
            "emis": emis, "alphabet": "".join(keys) if all(isinstance(k, str) and len(k) == 1
                                                            for k in keys) else None,
        }

    @property
    def baked(self):
        """The compiled tables (numpy): CSR in/out transition lists and emission rows."""
        if self._baked is None:
            raise ValueError("must bake model first")
        return self._baked

    # ---------------------------------------------------------- matrix round trips
    def dense_transition_matrix(self):
        b = self.baked
        m = b["n_states"]
        logm = np.zeros((m, m)) + NEGINF
        rows = np.repeat(np.arange(m), np.diff(b["out_off"]))
        logm[rows, b["out_dst"]] = b["out_logp"]
        return np.exp(logm)

    @classmethod
    def from_matrix(cls, transition_probabilities, distributions, starts, ends=None,
                    state_names=None, name=None, verbose=False, merge="All"):
        model = cls(name=name)
        n = len(distributions)
        names = state_names or ["s{}".format(i) for i in range(n)]
        states = [State(d, name=nm) for nm, d in zip(names, distributions)]
        for s in states:
            model.add_state(s)
        for i, p in enumerate(starts):
            if p != 0:
                model.add_transition(model.start, states[i], p)
        tp = np.asarray(transition_probabilities, dtype=np.float64)
        for i, j in zip(*np.nonzero(tp)):           # row-major walk of non-zero cells
            model.add_transition(states[i], states[j], tp[i, j])
        if ends is not None:
            # The reference connects ``states[j]`` with ``j`` left over from the loop above,
            # i.e. always the LAST state of the list, whatever index ``ends`` marks
            # (hmm.pyx:3231-3235).  Reproduced on purpose: it shapes every adVNTR model.
            tail = states[tp.shape[1] - 1] if tp.ndim == 2 and tp.shape[1] else None
            for p in ends:
                if p != 0:
                    model.add_transition(tail, model.end, p)
        model.bake(verbose=verbose, merge=merge)
        return model

    def sparse_transition_matrix(self):
        """The non-zero cells of ``dense_transition_matrix()`` as ``{(i, j): probability}``.
        Same values (``numpy.exp`` of the stored logs), without the O(m^2) array."""
        b = self.baked
        rows = np.repeat(np.arange(b["n_states"]), np.diff(b["out_off"]))
        probs = np.exp(b["out_logp"])
        return {(int(i), int(j)): p for i, j, p in zip(rows, b["out_dst"], probs) if p != 0}

    @classmethod
    def from_sparse(cls, cells, distributions, starts, ends=None, state_names=None, name=None,
                    verbose=False, merge="All"):
        """``from_matrix`` for a matrix given as ``{(i, j): probability}``: identical model (the
        cells are walked row-major like the dense loop, hmm.pyx:3226-3229)."""
        model = cls(name=name)
        n = len(distributions)
        names = state_names or ["s{}".format(i) for i in range(n)]
        states = [State(d, name=nm) for nm, d in zip(names, distributions)]
        for s in states:
            model.add_state(s)
        for i, p in enumerate(starts):
            if p != 0:
                model.add_transition(model.start, states[i], p)
        for (i, j) in sorted(cells):
            p = cells[(i, j)]
            if p != 0:
                model.add_transition(states[i], states[j], p)
        if ends is not None:
            tail = states[n - 1] if n else None      # see from_matrix: always the LAST state
            for p in ends:
                if p != 0:
                    model.add_transition(tail, model.end, p)
        model.bake(verbose=verbose, merge=merge)
        return model

    # ------------------------------------------------------------------ decoding
    def _release_engine(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    def _device_model(self):
        if self._baked is None or self.d == 0:
            raise ValueError("must bake model before using Viterbi algorithm")
        if self._engine is None:
            from . import engine, path_utils
            self._engine = engine.DeviceModel.from_baked(self._baked)
            self._engine.set_state_classes(path_utils.state_classes([s.name for s in self.states],
                                                                    self._baked["emis"]))
        return self._engine

    def _encode(self, sequence):
        """str -> uint8 codes; same ValueError as ``_check_input`` (hmm.pyx:57-81)."""
        from . import engine
        km = self.keymap[0]
        if self._baked["alphabet"] == "ACGT" and isinstance(sequence, (str, bytes)):
            codes, bad = engine.encode_acgt(sequence)
            if bad >= 0:
                raise ValueError("Symbol '{}' is not defined in a distribution".format(
                    sequence[bad] if isinstance(sequence, str) else chr(sequence[bad])))
            return codes
        out = np.empty(len(sequence), dtype=np.uint8)
        for i, c in enumerate(sequence):
            try:
                out[i] = km[c]
            except KeyError:
                raise ValueError("Symbol '{}' is not defined in a distribution".format(c))
        return out

    def viterbi(self, sequence):
        """``(logp, [(state_index, State), ...])`` or ``(-inf, None)`` (hmm.pyx:1911-1967)."""
        if self.d == 0:
            raise ValueError("must bake model before using Viterbi algorithm")
        res = self.viterbi_batch([sequence])
        logp = float(res.logp[0])
        if not logp > NEGINF:
            return logp, None
        st = self.states
        return logp, [(int(i), st[i]) for i in res.path(0)]

    def viterbi_batch(self, sequences, both_strands=False, want_path=True, precision="fp64",
                      want_summary=False):
        """Decode every read of a locus in ONE device call (the batched call site).
        ``want_summary``: also return the on-device path reducers (``result.summaries``)."""
        if self.d == 0:
            raise ValueError("must bake model before using Viterbi algorithm")
        dm = self._device_model()
        if self._baked["alphabet"] == "ACGT" and all(isinstance(s, str) for s in sequences):
            # one encode call for the whole batch instead of one per read
            from . import engine
            seqs, off = engine.encode_batch(sequences)
            return dm.viterbi_packed(seqs, off, both_strands=both_strands, want_path=want_path,
                                     precision=precision, want_summary=want_summary)
        codes = [self._encode(s) for s in sequences]
        return dm.viterbi(codes, both_strands=both_strands, want_path=want_path,
                          precision=precision, want_summary=want_summary)

    def log_probability(self, sequence, check_input=True):
        """Forward log-likelihood (hmm.pyx:1258-1313)."""
        if self.d == 0:
            raise ValueError("must bake model before computing probability")
        return float(self.log_probability_batch([sequence])[0])

    def log_probability_batch(self, sequences):
        if self.d == 0:
            raise ValueError("must bake model before computing probability")
        codes = [self._encode(s) for s in sequences]
        return self._device_model().log_probability(codes)

    # ------------------------------------------------------------- serialisation
    def to_json(self, separators=(",", " : "), indent=4):
        """Same document layout as ``hmm.pyx:3023-3095`` (edges as probabilities)."""
        idx = {s: i for i, s in enumerate(self.states)}
        edges = [(idx[a], idx[b], math.e ** w, self._pseudo.get((a, b), math.e ** w), None)
                 for a, b, w in self.graph.edges()]
        groups = {}                                   # states sharing one distribution OBJECT are tied
        for i, st in enumerate(self.states[:self.silent_start]):
            groups.setdefault(id(st.distribution), []).append(i)
        ties = [(i, j) for i, st in enumerate(self.states[:self.silent_start])
                for j in groups[id(st.distribution)] if j != i]          # hmm.pyx:895-925, :284-291
        return json.dumps({
            "class": "HiddenMarkovModel", "name": self.name,
            "start": json.loads(self.start.to_json()), "end": json.loads(self.end.to_json()),
            "states": [json.loads(s.to_json()) for s in self.states],
            "end_index": self.end_index, "start_index": self.start_index,
            "silent_index": self.silent_start, "edges": edges, "distribution ties": ties,
        }, separators=separators, indent=indent)

    @classmethod
    def from_json(cls, s, verbose=False):
        """``hmm.pyx:3098-3144`` (how ``get_vntr_matcher_hmm`` loads a stored model,
        ``vntr_finder.py:125-129``): a JSON string or the name of a JSON file.  As in the reference the
        new model keeps its own fresh ``<name>-start`` / ``<name>-end`` states next to the stored ones
        and is baked with the DEFAULT ``merge='All'``, which removes those two orphans again and
        merges silent states with a probability-1 out-edge -- so a stored read matcher comes back
        with fewer silent states than it was built with, in the reference and here alike."""
        try:
            d = json.loads(s)
        except Exception:
            try:
                with open(s, "r") as infile:
                    d = json.load(infile)
            except Exception:
                raise IOError("String must be properly formatted JSON or filename of properly formatted JSON.")
        model = cls(str(d["name"]))
        states = [State.from_json(json.dumps(j)) for j in d["states"]]
        for i, j in d["distribution ties"]:
            states[i].tie(states[j])
        model.add_states(states)
        model.start = states[d["start_index"]]
        model.end = states[d["end_index"]]
        for start, end, probability, pseudocount, group in d["edges"]:
            model.add_transition(states[start], states[end], probability, pseudocount, group)
        model.bake(verbose=verbose)
        return model

    def __del__(self):
        try:
            self._release_engine()
        except Exception:
            pass
