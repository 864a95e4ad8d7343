"""networkx-1.11 behavioural shim for the compiled reference engine (TEST INFRASTRUCTURE).

The reference pins ``networkx==1.11`` (``/root/reference/setup.py:18``,
``requirements.txt:6``) and networkx decides two things that matter for
bit-exact Viterbi parity: the topological order of silent states
(``pomegranate/hmm.pyx:870-874``) and the in-edge order of every state
(``hmm.pyx:970,994`` walk ``edges_iter()``).  networkx 1.11 is not installed here
(3.x is, and it removed ``topological_sort(nbunch=)``, ``edges_iter`` and
``graph.edge``), so this module restates the published 1.11 semantics of exactly
the calls ``hmm.pyx`` makes -- nothing else:

* ``DiGraph``: dict-of-dicts with insertion-ordered ``node`` / ``succ`` / ``pred``;
  ``add_edge`` on an existing edge updates the attribute dict in place.
* ``edges_iter``: for u in node order, for v in succ[u] order.
* ``subgraph(nbunch)``: nodes in nbunch order, successors in the parent's order.
* ``union(G, H)``: G nodes, G edges, H nodes, H edges (attribute dicts copied).
* ``topological_sort(G, nbunch)``: 1.11's iterative DFS -- all unexplored
  successors are pushed, the LAST pushed is explored first, a node is appended
  when it has no unexplored successor, and the reversed post-order is returned.

It is put on ``sys.path`` (ahead of site-packages) only by ``oracle/refenv.py``.
"""

__version__ = "1.11-shim"


class NetworkXError(Exception):
    pass


class NetworkXUnfeasible(NetworkXError):
    pass


class DiGraph(object):
    def __init__(self, data=None, **attr):
        self.graph = {}
        self.node = {}
        self.adj = {}
        self.pred = {}
        self.succ = self.adj
        self.edge = self.adj
        self.graph.update(attr)

    # -- container protocol -------------------------------------------------
    @property
    def name(self):
        return self.graph.get("name", "")

    @name.setter
    def name(self, s):
        self.graph["name"] = s

    def __iter__(self):
        return iter(self.node)

    def __contains__(self, n):
        try:
            return n in self.node
        except TypeError:
            return False

    def __len__(self):
        return len(self.node)

    def __getitem__(self, n):
        return self.adj[n]

    def is_multigraph(self):
        return False

    def is_directed(self):
        return True

    # -- nodes ----------------------------------------------------------------
    def add_node(self, n, attr_dict=None, **attr):
        if attr_dict is None:
            attr_dict = attr
        else:
            attr_dict.update(attr)
        if n not in self.succ:
            self.succ[n] = {}
            self.pred[n] = {}
            self.node[n] = attr_dict
        else:
            self.node[n].update(attr_dict)

    def add_nodes_from(self, nodes, **attr):
        for n in nodes:
            if n not in self.succ:
                self.succ[n] = {}
                self.pred[n] = {}
                self.node[n] = attr.copy()
            else:
                self.node[n].update(attr)

    def remove_node(self, n):
        try:
            nbrs = self.succ[n]
            del self.node[n]
        except KeyError:
            raise NetworkXError("The node %s is not in the digraph." % (n,))
        for u in nbrs:
            del self.pred[u][n]
        del self.succ[n]
        for u in self.pred[n]:
            del self.succ[u][n]
        del self.pred[n]

    def nodes(self, data=False):
        return list(self.nodes_iter(data))

    def nodes_iter(self, data=False):
        if data:
            return iter(self.node.items())
        return iter(self.node)

    def number_of_nodes(self):
        return len(self.node)

    # -- edges ----------------------------------------------------------------
    def add_edge(self, u, v, attr_dict=None, **attr):
        if attr_dict is None:
            attr_dict = attr
        else:
            attr_dict.update(attr)
        if u not in self.succ:
            self.succ[u] = {}
            self.pred[u] = {}
            self.node[u] = {}
        if v not in self.succ:
            self.succ[v] = {}
            self.pred[v] = {}
            self.node[v] = {}
        datadict = self.adj[u].get(v, {})
        datadict.update(attr_dict)
        self.succ[u][v] = datadict
        self.pred[v][u] = datadict

    def add_edges_from(self, ebunch, attr_dict=None, **attr):
        if attr_dict is None:
            attr_dict = attr
        else:
            attr_dict.update(attr)
        for e in ebunch:
            if len(e) == 3:
                u, v, dd = e
            else:
                u, v = e
                dd = {}
            if u not in self.succ:
                self.succ[u] = {}
                self.pred[u] = {}
                self.node[u] = {}
            if v not in self.succ:
                self.succ[v] = {}
                self.pred[v] = {}
                self.node[v] = {}
            datadict = self.adj[u].get(v, {})
            datadict.update(attr_dict)
            datadict.update(dd)
            self.succ[u][v] = datadict
            self.pred[v][u] = datadict

    def remove_edge(self, u, v):
        try:
            del self.succ[u][v]
            del self.pred[v][u]
        except KeyError:
            raise NetworkXError("The edge %s-%s not in graph." % (u, v))

    def has_edge(self, u, v):
        try:
            return v in self.adj[u]
        except KeyError:
            return False

    def edges_iter(self, nbunch=None, data=False, default=None):
        if nbunch is None:
            nodes_nbrs = self.adj.items()
        else:
            nodes_nbrs = ((n, self.adj[n]) for n in self.nbunch_iter(nbunch))
        if data is True:
            for n, nbrs in nodes_nbrs:
                for nbr, ddict in nbrs.items():
                    yield (n, nbr, ddict)
        elif data is not False:
            for n, nbrs in nodes_nbrs:
                for nbr, ddict in nbrs.items():
                    yield (n, nbr, ddict[data] if data in ddict else default)
        else:
            for n, nbrs in nodes_nbrs:
                for nbr in nbrs:
                    yield (n, nbr)

    def edges(self, nbunch=None, data=False, default=None):
        return list(self.edges_iter(nbunch, data, default))

    out_edges = edges
    out_edges_iter = edges_iter

    def in_edges_iter(self, nbunch=None, data=False):
        if nbunch is None:
            nodes_nbrs = self.pred.items()
        else:
            nodes_nbrs = ((n, self.pred[n]) for n in self.nbunch_iter(nbunch))
        for n, nbrs in nodes_nbrs:
            for nbr, ddict in nbrs.items():
                yield (nbr, n, ddict) if data else (nbr, n)

    def in_edges(self, nbunch=None, data=False):
        return list(self.in_edges_iter(nbunch, data))

    def successors(self, n):
        return list(self.succ[n])

    def predecessors(self, n):
        return list(self.pred[n])

    def number_of_edges(self):
        return sum(len(nbrs) for nbrs in self.succ.values())

    def nbunch_iter(self, nbunch=None):
        if nbunch is None:
            return iter(self.adj)
        if nbunch in self:
            return iter([nbunch])
        adj = self.adj
        return (n for n in nbunch if n in adj)

    # -- views ------------------------------------------------------------------
    def subgraph(self, nbunch):
        bunch = self.nbunch_iter(nbunch)
        H = self.__class__()
        for n in bunch:
            H.node[n] = self.node[n]
        for n in H.node:
            H.succ[n] = {}
            H.pred[n] = {}
        for u in H.succ:
            Hnbrs = H.succ[u]
            for v, datadict in self.succ[u].items():
                if v in H.succ:
                    Hnbrs[v] = datadict
                    H.pred[v][u] = datadict
        H.graph = self.graph
        return H


def union(G, H, rename=(None, None), name=None):
    R = G.__class__()
    R.name = name if name is not None else "union( %s, %s )" % (G.name, H.name)
    if set(G) & set(H):
        raise NetworkXError("The node sets of G and H are not disjoint.")
    R.add_nodes_from(G)
    R.add_edges_from(e for e in G.edges_iter(data=True))
    R.add_nodes_from(H)
    R.add_edges_from(e for e in H.edges_iter(data=True))
    R.node.update(G.node)
    R.node.update(H.node)
    R.graph.update(G.graph)
    R.graph.update(H.graph)
    return R


def topological_sort(G, nbunch=None, reverse=False):
    seen = set()
    order = []
    explored = set()
    if nbunch is None:
        nbunch = G.nodes_iter()
    for v in nbunch:
        if v in explored:
            continue
        fringe = [v]
        while fringe:
            w = fringe[-1]
            if w in explored:
                fringe.pop()
                continue
            seen.add(w)
            new_nodes = []
            for n in G[w]:
                if n not in explored:
                    if n in seen:
                        raise NetworkXUnfeasible("Graph contains a cycle.")
                    new_nodes.append(n)
            if new_nodes:
                fringe.extend(new_nodes)
            else:
                explored.add(w)
                order.append(w)
                fringe.pop()
    if reverse:
        return order
    return list(reversed(order))
