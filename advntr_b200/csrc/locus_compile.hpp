// locus_compile.hpp -- native compiler for the per-locus read-matcher HMM (host side, pure C++17).
//
// What the reference does per locus in Python (/root/reference/advntr/hmm_utils.py:289-595,
// profile_hmm.py:13-161, pomegranate/hmm.pyx bake / from_matrix / concatenate): build three sub-models,
// bake eight times, go through two dense m x m matrix round trips.  For a given SHAPE -- flank lengths,
// repeat-unit match columns R, unrolled copies C -- every locus has the same states and the same edges
// in the same order; only the numbers differ.  This file
//
//   1. builds the STRUCTURE of a shape symbolically (rm::build_shape): the same sequence of graph
//      operations as the reference -- insertion-ordered graphs, bake(merge=None) = name sort + the
//      networkx-1.11 DFS topological order + CSR in edge-walk order, the sparse form of the matrix
//      round trips, concatenate -- on edges that carry a symbolic label (which probability, how many
//      exp -> log round trips it has been through, whether its row was rescaled) instead of a number;
//   2. evaluates the repeat-unit profile of a locus (rm::repeat_profile, profile_hmm.py:13-161) and
//      the few hundred parameters of its shape through the SAME chain of float operations
//      (libm log, the caller's vector exp = numpy.exp for the reference, hmm.pyx:433, :514);
//   3. fills the device tables of the banded kernel (shared-memory image, first-row tables, row-0
//      closure) for that locus directly from the shape's scatter maps (rm::fill_lean), without the
//      per-model graph analysis of model_compile.hpp (that analysis runs once per shape).
//
// Nothing here is assumed exact: tests/test_native_compile.py compares structure and bit patterns
// with the literal Python build (itself bit-equal to the compiled reference) over a grid of shapes.
#pragma once

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <exception>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <tuple>

#ifdef __linux__
#include <sched.h>
#endif

#include "model_compile.hpp"

namespace advhmm {
namespace rm {

// ---------------------------------------------------------------------------------------------
// symbolic labels
// ---------------------------------------------------------------------------------------------
enum Sym : int32_t {
    SYM_ONE = 0,         // probability 1 (glue edges): log 0, stays 0 through every round trip
    SYM_IE,              // flank insert      error_rate * 2 / 5            (hmm_utils.py:317, :384)
    SYM_DE,              // flank delete      error_rate * 1 / 5
    SYM_ADV,             // 1 - ie - de
    SYM_ADV_M01,         // prefix match -> next match: adv - 0.01          (hmm_utils.py:344)
    SYM_ONE_MINUS_IE,    // last flank column -> gate: 1 - ie
    SYM_ADV_OVER_L,      // suffix gate -> every match column: adv / L_left (hmm_utils.py:388-389)
    SYM_C001,            // prefix match -> gate early exit 0.01            (hmm_utils.py:346)
    SYM_C05,             // unit_end -> next unit / end of repeats 0.5      (hmm_utils.py:530-536)
    SYM_C03,             // read starts in the left flank 0.3               (hmm_utils.py:574)
    SYM_FIRST_COPY,      // ... or at any match column of the first copy: 0.7 / R   (:575-576)
    SYM_TO_END,          // repeat match -> model end: (0.7 / N) / (1 + 0.7 / N)    (:584)
    SYM_PROFILE0         // + profile index, see prof_index()
};

struct Lab {
    int32_t sym = SYM_ONE;
    uint8_t trips = 0;   // exp -> log round trips after the first log (dense_transition_matrix + from_matrix)
    uint8_t div = 0;     // the probability is divided by (1 + 0.7 / N) inside the last round trip (:578-583)
    bool operator<(const Lab& o) const { return std::tie(sym, trips, div) < std::tie(o.sym, o.trips, o.div); }
    bool operator==(const Lab& o) const { return sym == o.sym && trips == o.trips && div == o.div; }
};

// repeat-unit profile transitions: source (kind, column i) -> target kind
enum PKind : int { PK_START = 0, PK_I = 1, PK_M = 2, PK_D = 3 };
inline int prof_index(int R, int src_kind, int i, int dst_kind /* 0 I, 1 M, 2 D, 3 unit_end */)
{
    return ((src_kind * (R + 1) + i) * 4) + dst_kind;
}
inline int prof_size(int R) { return 16 * (R + 1); }

enum NodeKind : uint8_t { NK_OTHER = 0, NK_M = 1, NK_I = 2, NK_D = 3, NK_UNIT_START = 4, NK_UNIT_END = 5 };
enum NodePart : uint8_t { NP_NONE = 0, NP_SUFFIX = 1, NP_PREFIX = 2, NP_REPEAT = 3 };

struct Node {
    std::string name;
    bool silent = true;
    uint8_t kind = NK_OTHER, part = NP_NONE;
    int32_t idx = 0;     // column number in the name (M{idx}_..)
};

// insertion-ordered directed graph (networkx-1.11 DiGraph on ordered dicts)
struct Graph {
    std::vector<Node> nodes;
    std::vector<std::vector<std::pair<int32_t, Lab>>> succ;
    int start = -1, end = -1;
    int add_node(Node n)
    {
        nodes.push_back(std::move(n));
        succ.emplace_back();
        return (int)nodes.size() - 1;
    }
    // re-adding an edge updates its label in place and keeps its position (hmm.pyx:433)
    void add_edge(int a, int b, Lab l)
    {
        for (auto& e : succ[a])
            if (e.first == b) { e.second = l; return; }
        succ[a].emplace_back(b, l);
    }
    // networkx union + the glue edge of HiddenMarkovModel.concatenate (hmm.pyx:584-615)
    void concatenate(const Graph& o)
    {
        const int off = (int)nodes.size();
        for (size_t i = 0; i < o.nodes.size(); ++i) {
            nodes.push_back(o.nodes[i]);
            succ.push_back(o.succ[i]);
            for (auto& e : succ.back()) e.first += off;
        }
        add_edge(end, o.start + off, Lab{});
        end = o.end + off;
    }
};

struct Baked {
    int m = 0, S = 0, start_index = 0, end_index = 0;
    std::vector<int32_t> order;   // state -> node
    std::vector<int32_t> index;   // node -> state
    std::vector<int32_t> in_off, in_src;
    std::vector<Lab> in_lab;
    std::vector<int32_t> out_off, out_dst;
    std::vector<Lab> out_lab;
};

// HiddenMarkovModel.bake(merge=None), hmm.pyx:844-1023
inline bool bake(const Graph& g, Baked& b, std::string& err)
{
    const int n = (int)g.nodes.size();
    std::vector<int32_t> emitting, silent;
    for (int i = 0; i < n; ++i) (g.nodes[i].silent ? silent : emitting).push_back(i);
    auto by_name = [&](int32_t x, int32_t y) { return g.nodes[x].name < g.nodes[y].name; };
    std::stable_sort(emitting.begin(), emitting.end(), by_name);
    std::stable_sort(silent.begin(), silent.end(), by_name);
    // networkx-1.11 topological_sort(subgraph(silent), nbunch=silent): iterative DFS seeded in nbunch
    // order, every unexplored successor pushed (the last one is visited first), reversed post-order
    std::vector<char> explored(n, 0), seen(n, 0);
    std::vector<int32_t> post, stack, fresh;
    post.reserve(silent.size());
    for (int32_t v : silent) {
        if (explored[v]) continue;
        stack.assign(1, v);
        while (!stack.empty()) {
            const int32_t w = stack.back();
            if (explored[w]) { stack.pop_back(); continue; }
            seen[w] = 1;
            fresh.clear();
            for (const auto& e : g.succ[w]) {
                const int32_t t = e.first;
                if (!g.nodes[t].silent || explored[t]) continue;
                if (seen[t]) { err = "cycle of silent states"; return false; }
                fresh.push_back(t);
            }
            if (!fresh.empty()) stack.insert(stack.end(), fresh.begin(), fresh.end());
            else { explored[w] = 1; post.push_back(w); stack.pop_back(); }
        }
    }
    b.order = emitting;
    b.order.insert(b.order.end(), post.rbegin(), post.rend());
    b.S = (int)emitting.size();
    b.m = n;
    b.index.assign(n, -1);
    for (int i = 0; i < n; ++i) b.index[b.order[i]] = i;
    b.start_index = b.index[g.start];
    b.end_index = b.index[g.end];
    // edges in walk order (node insertion x successor insertion), CSR by target and by source, stable
    size_t E = 0;
    for (const auto& s : g.succ) E += s.size();
    b.in_off.assign(n + 1, 0);
    b.out_off.assign(n + 1, 0);
    for (int a = 0; a < n; ++a)
        for (const auto& e : g.succ[a]) { b.in_off[b.index[e.first] + 1]++; b.out_off[b.index[a] + 1]++; }
    for (int i = 0; i < n; ++i) { b.in_off[i + 1] += b.in_off[i]; b.out_off[i + 1] += b.out_off[i]; }
    b.in_src.assign(E, 0); b.in_lab.assign(E, Lab{});
    b.out_dst.assign(E, 0); b.out_lab.assign(E, Lab{});
    std::vector<int32_t> fi(b.in_off.begin(), b.in_off.end() - 1), fo(b.out_off.begin(), b.out_off.end() - 1);
    for (int a = 0; a < n; ++a)
        for (const auto& e : g.succ[a]) {
            const int s = b.index[a], d = b.index[e.first];
            b.in_src[fi[d]] = s; b.in_lab[fi[d]++] = e.second;
            b.out_dst[fo[s]] = d; b.out_lab[fo[s]++] = e.second;
        }
    return true;
}

// cells of sparse_transition_matrix() keyed (row state, column state); every surviving edge has been
// through one more exp -> log round trip (hmm.pyx:514, then :433 in from_matrix)
using Cells = std::map<std::pair<int32_t, int32_t>, Lab>;

inline Cells cells_of(const Baked& b)
{
    Cells c;
    for (int i = 0; i < b.m; ++i)
        for (int k = b.out_off[i]; k < b.out_off[i + 1]; ++k) {
            Lab l = b.out_lab[k];
            if (l.sym != SYM_ONE) l.trips++;          // exp(0) = 1, log(1) = 0: glue edges stay 0
            c[{i, b.out_dst[k]}] = l;
        }
    return c;
}

// HiddenMarkovModel.from_matrix on the non-zero cells (hmm.pyx:3147-3238): new start / end states,
// the old states in old index order, start -> states[start_state], row-major cells, and the LAST
// state of the list wired to the new end (the stale-j wiring, hmm.pyx:3231-3235)
inline Graph from_cells(const std::string& name, const std::vector<Node>& states, const Cells& cells, int start_state)
{
    Graph g;
    Node s; s.name = name + "-start";
    Node e; e.name = name + "-end";
    g.start = g.add_node(s);
    g.end = g.add_node(e);
    for (const Node& st : states) g.add_node(st);
    g.add_edge(g.start, 2 + start_state, Lab{});
    for (const auto& kv : cells) g.add_edge(2 + kv.first.first, 2 + kv.first.second, kv.second);
    g.add_edge(2 + (int)states.size() - 1, g.end, Lab{});
    return g;
}

inline std::vector<Node> states_of(const Graph& g, const Baked& b)
{
    std::vector<Node> out;
    out.reserve(b.m);
    for (int i = 0; i < b.m; ++i) out.push_back(g.nodes[b.order[i]]);
    return out;
}

// hmm_utils.py:357-420 (suffix, left flank) / :290-353 (prefix, right flank)
inline Graph flank_matcher(bool suffix, int L)
{
    const char* tag = suffix ? "suffix" : "prefix";
    const std::string title = suffix ? "Suffix Matcher HMM Model" : "Prefix Matcher HMM Model";
    const uint8_t part = suffix ? NP_SUFFIX : NP_PREFIX;
    Graph g;
    Node s; s.name = title + "-start";
    Node e; e.name = title + "-end";
    g.start = g.add_node(s);
    g.end = g.add_node(e);
    std::vector<int> ins(L + 1), mat(L), del(L);
    auto mk = [&](char c, int i, bool silent, uint8_t kind) {
        Node n;
        n.name = std::string(1, c) + std::to_string(i) + "_" + tag;
        n.silent = silent; n.kind = kind; n.part = part; n.idx = i;
        return g.add_node(n);
    };
    for (int i = 0; i <= L; ++i) ins[i] = mk('I', i, false, NK_I);
    for (int i = 0; i < L; ++i) mat[i] = mk('M', i + 1, false, NK_M);
    for (int i = 0; i < L; ++i) del[i] = mk('D', i + 1, true, NK_D);
    Node gi; gi.name = std::string(tag) + "_start_" + tag;
    Node go; go.name = std::string(tag) + "_end_" + tag;
    const int gate_in = g.add_node(gi), gate_out = g.add_node(go);
    auto T = [&](int a, int b, int32_t sym) { Lab l; l.sym = sym; g.add_edge(a, b, l); };
    T(g.start, gate_in, SYM_ONE);
    T(gate_out, g.end, SYM_ONE);
    if (suffix) {
        T(gate_in, del[0], SYM_DE);
        T(gate_in, ins[0], SYM_IE);
        for (int k = 0; k < L; ++k) T(gate_in, mat[k], SYM_ADV_OVER_L);
    } else {
        T(gate_in, mat[0], SYM_ADV);
        T(gate_in, del[0], SYM_DE);
        T(gate_in, ins[0], SYM_IE);
    }
    T(ins[0], ins[0], SYM_IE);
    T(ins[0], del[0], SYM_DE);
    T(ins[0], mat[0], SYM_ADV);
    const int z = L - 1;
    T(del[z], gate_out, SYM_ONE_MINUS_IE);
    T(del[z], ins[z + 1], SYM_IE);
    T(mat[z], gate_out, SYM_ONE_MINUS_IE);
    T(mat[z], ins[z + 1], SYM_IE);
    T(ins[z + 1], ins[z + 1], SYM_IE);
    T(ins[z + 1], gate_out, SYM_ONE_MINUS_IE);
    for (int k = 0; k < L; ++k) {
        T(mat[k], ins[k + 1], SYM_IE);
        T(del[k], ins[k + 1], SYM_IE);
        T(ins[k + 1], ins[k + 1], SYM_IE);
        if (k < z) {
            T(ins[k + 1], mat[k + 1], SYM_ADV);
            T(ins[k + 1], del[k + 1], SYM_DE);
            if (suffix) {
                T(mat[k], mat[k + 1], SYM_ADV);
                T(mat[k], del[k + 1], SYM_DE);
            } else {
                T(mat[k], mat[k + 1], SYM_ADV_M01);
                T(mat[k], del[k + 1], SYM_DE);
                T(mat[k], gate_out, SYM_C001);
            }
            T(del[k], del[k + 1], SYM_DE);
            T(del[k], mat[k + 1], SYM_ADV);
        }
    }
    return g;
}

// hmm_utils.py:424-497: C unrolled copies of the repeat-unit profile
inline Graph constant_repeats(int R, int C)
{
    Graph g;
    Node s; s.name = "Repeating Pattern Matcher HMM Model-start";
    Node e; e.name = "Repeating Pattern Matcher HMM Model-end";
    g.start = g.add_node(s);
    g.end = g.add_node(e);
    auto P = [&](int a, int b, int sk, int i, int dk) { Lab l; l.sym = SYM_PROFILE0 + prof_index(R, sk, i, dk); g.add_edge(a, b, l); };
    int prev_out = -1;
    for (int k = 0; k < C; ++k) {
        const std::string suf = "_" + std::to_string(k);
        std::vector<int> ins(R + 1), mat(R), del(R);
        auto mk = [&](char c, int i, bool silent, uint8_t kind) {
            Node n;
            n.name = std::string(1, c) + std::to_string(i) + suf;
            n.silent = silent; n.kind = kind; n.part = NP_REPEAT; n.idx = i;
            return g.add_node(n);
        };
        for (int i = 0; i <= R; ++i) ins[i] = mk('I', i, false, NK_I);
        for (int i = 1; i <= R; ++i) mat[i - 1] = mk('M', i, false, NK_M);
        for (int i = 1; i <= R; ++i) del[i - 1] = mk('D', i, true, NK_D);
        Node gi; gi.name = "unit_start" + suf; gi.kind = NK_UNIT_START;
        Node go; go.name = "unit_end" + suf; go.kind = NK_UNIT_END;
        const int gate_in = g.add_node(gi), gate_out = g.add_node(go);
        g.add_edge(k ? prev_out : g.start, gate_in, Lab{});
        if (k == C - 1) g.add_edge(gate_out, g.end, Lab{});
        // targets: 0 I, 1 M, 2 D, 3 unit_end
        P(gate_in, mat[0], PK_START, 0, 1);
        P(gate_in, del[0], PK_START, 0, 2);
        P(gate_in, ins[0], PK_START, 0, 0);
        P(ins[0], ins[0], PK_I, 0, 0);
        P(ins[0], del[0], PK_I, 0, 2);
        P(ins[0], mat[0], PK_I, 0, 1);
        P(del[R - 1], gate_out, PK_D, R, 3);
        P(del[R - 1], ins[R], PK_D, R, 0);
        P(mat[R - 1], gate_out, PK_M, R, 3);
        P(mat[R - 1], ins[R], PK_M, R, 0);
        P(ins[R], ins[R], PK_I, R, 0);
        P(ins[R], gate_out, PK_I, R, 3);
        for (int i = 1; i <= R; ++i) {
            P(mat[i - 1], ins[i], PK_M, i, 0);
            P(del[i - 1], ins[i], PK_D, i, 0);
            P(ins[i], ins[i], PK_I, i, 0);
            if (i < R) {
                P(ins[i], mat[i], PK_I, i, 1);
                P(ins[i], del[i], PK_I, i, 2);
                P(mat[i - 1], mat[i], PK_M, i, 1);
                P(mat[i - 1], del[i], PK_M, i, 2);
                P(del[i - 1], mat[i], PK_D, i, 1);
                P(del[i - 1], del[i], PK_D, i, 2);
            }
        }
        prev_out = gate_out;
    }
    return g;
}

// ---------------------------------------------------------------------------------------------
// the structure of one shape
// ---------------------------------------------------------------------------------------------
struct ShapeKey {
    int Ll, Lr, R, C;
    bool operator<(const ShapeKey& o) const { return std::tie(Ll, Lr, R, C) < std::tie(o.Ll, o.Lr, o.R, o.C); }
};

struct ShapeStructure {
    ShapeKey key{};
    // the baked arrays every locus of the shape shares (reference order)
    int m = 0, S = 0, start = 0, end = 0, finite = 0;
    std::vector<Node> states;
    std::vector<int32_t> in_off, in_src;
    std::vector<int32_t> edge_slot;        // [E] parameter slot of every edge
    std::vector<Lab> slots;                // slot -> label
    std::vector<int32_t> emis_row;         // [S] >= 0: row of the locus emission table; < 0: flank match, -(1 + position in left|right)
    std::vector<uint8_t> base_class;       // [m] class byte without the flank base (include/advhmm.h)
    // kernel-side analysis, run ONCE per shape (model_compile.hpp on a tagged copy)
    CompiledModel cm;                      // structural members only are meaningful
    std::vector<int32_t> g_slot;           // [g edges] parameter slot of every evaluated edge (g order)
    std::vector<std::pair<int32_t, int32_t>> img_w;   // (index into w10 doubles, g edge)
    std::vector<int32_t> fin_edge;         // [fin edges] -> g edge
    std::vector<unsigned char> image0;     // 208 * P bytes: every entry -inf
    int image_bytes = 0;
    // emitting states: first-row candidates come from silent sources only (row 0 is -inf elsewhere)
    std::vector<int32_t> r1_off, r1_edge;  // [S+1], g edges whose source is silent, in candidate order
    bool banded = false;
    std::string why;
};

inline bool build_shape(const ShapeKey& key, ShapeStructure& sh, std::string& err)
{
    const int Ll = key.Ll, Lr = key.Lr, R = key.R, C = key.C;
    if (Ll < 1 || Lr < 1 || R < 1 || C < 1) { err = "shape needs flanks, match columns and copies >= 1"; return false; }
    sh.key = key;
    Baked b;
    // ---- repeat block: constant copies -> variable number of copies (hmm_utils.py:501-549) ----
    Graph rep = constant_repeats(R, C);
    if (!bake(rep, b, err)) return false;
    std::vector<Node> st = states_of(rep, b);
    Cells cells = cells_of(b);
    {
        Node enter; enter.name = "start_repeating_pattern_match";
        Node leave; leave.name = "end_repeating_pattern_match";
        const int m0 = b.m, i_enter = m0, i_leave = m0 + 1;
        std::vector<int32_t> last_col(m0, -1);
        for (const auto& kv : cells) last_col[kv.first.first] = std::max(last_col[kv.first.first], kv.first.second);
        const int first_unit = last_col[b.start_index];
        if (first_unit < 0) { err = "repeat model without a first unit"; return false; }
        cells.erase({b.start_index, first_unit});
        cells[{b.start_index, i_enter}] = Lab{};
        cells[{i_enter, first_unit}] = Lab{};
        for (int i = 0; i < m0; ++i)
            if (st[i].kind == NK_UNIT_END) {
                Lab half; half.sym = SYM_C05;
                cells[{i, last_col[i]}] = half;
                cells[{i, i_leave}] = half;
            }
        cells[{i_leave, b.end_index}] = Lab{};
        st.push_back(enter);
        st.push_back(leave);
    }
    Graph var = from_cells("Repeat Matcher HMM Model", st, cells, b.start_index);
    // ---- suffix + repeats + prefix (hmm_utils.py:555-559) ---------------------------------------
    Graph all = flank_matcher(true, Ll);
    all.concatenate(var);
    all.concatenate(flank_matcher(false, Lr));
    if (!bake(all, b, err)) return false;
    st = states_of(all, b);
    cells = cells_of(b);
    {
        int entry = -1;
        std::vector<int> first_copy, matches;
        for (int i = 0; i < b.m; ++i) {
            const Node& n = st[i];
            if (n.kind == NK_M && n.part == NP_REPEAT) {
                matches.push_back(i);
                if (n.name.size() >= 2 && n.name.compare(n.name.size() - 2, 2, "_0") == 0) first_copy.push_back(i);
            }
            if (n.name == "suffix_start_suffix") entry = i;
        }
        if (entry < 0 || (int)first_copy.size() != R || (int)matches.size() != R * C) { err = "unexpected read-matcher layout"; return false; }
        Lab l03; l03.sym = SYM_C03;
        cells[{b.start_index, entry}] = l03;
        Lab lfc; lfc.sym = SYM_FIRST_COPY;
        for (int i : first_copy) cells[{b.start_index, i}] = lfc;
        std::vector<char> is_match(b.m, 0);
        for (int i : matches) is_match[i] = 1;
        for (auto& kv : cells)
            if (is_match[kv.first.first]) kv.second.div = 1;
        Lab lte; lte.sym = SYM_TO_END;
        for (int i : matches) cells[{i, b.end_index}] = lte;
    }
    Graph fin = from_cells("Read Matcher", st, cells, b.start_index);
    if (!bake(fin, b, err)) return false;

    sh.m = b.m; sh.S = b.S; sh.start = b.start_index; sh.end = b.end_index;
    sh.finite = (b.in_off[b.end_index + 1] - b.in_off[b.end_index]) > 0 ? 1 : 0;
    sh.states = states_of(fin, b);
    sh.in_off = b.in_off;
    sh.in_src = b.in_src;
    std::map<Lab, int32_t> slot_of;
    sh.edge_slot.resize(b.in_src.size());
    for (size_t k = 0; k < b.in_src.size(); ++k) {
        Lab l = b.in_lab[k];
        if (l.sym == SYM_ONE) { l.trips = 0; l.div = 0; }
        auto it = slot_of.find(l);
        if (it == slot_of.end()) { it = slot_of.emplace(l, (int32_t)sh.slots.size()).first; sh.slots.push_back(l); }
        sh.edge_slot[k] = it->second;
    }
    // emissions and class bytes from the node attributes
    sh.emis_row.assign(sh.S, 0);
    sh.base_class.assign(sh.m, 0);
    for (int i = 0; i < sh.m; ++i) {
        const Node& n = sh.states[i];
        const int part = (n.kind == NK_M || n.kind == NK_I || n.kind == NK_D) ? n.part : 0;
        sh.base_class[i] = (uint8_t)(n.kind | (part << 3));
        if (i >= sh.S) continue;
        if (n.silent) { err = "silent state among the emitting ones"; return false; }
        if (n.part == NP_SUFFIX || n.part == NP_PREFIX) {
            if (n.kind == NK_I) sh.emis_row[i] = 4;
            else sh.emis_row[i] = -(1 + (n.part == NP_SUFFIX ? n.idx - 1 : Ll + n.idx - 1));
        } else {
            sh.emis_row[i] = 5 + (n.kind == NK_M ? n.idx - 1 : R + n.idx);
        }
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// repeat-unit profile of a locus (profile_hmm.py:13-161), same float operations in the same order
// ---------------------------------------------------------------------------------------------
struct Profile {
    int R = 0;
    std::vector<double> trans;     // [prof_size(R)] probabilities, 0 where the pair does not exist
    std::vector<double> emis;      // [(2R + 1) * 4] rows M1..MR then I0..IR, probabilities over A,C,G,T
};

inline int acgt_code(char c)
{
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; }
}

// alignment: n_seq rows of `width` characters over ACGT- (row-major).  false: not a valid alignment.
inline bool repeat_profile(const char* aln, int n_seq, int width, double error_rate, Profile& out, std::string& err)
{
    if (n_seq < 1 || width < 1) { err = "empty alignment"; return false; }
    const double pseudo = ((double)n_seq / 4.0) * (error_rate / 10);
    const double gap_limit = 0.5 * (double)n_seq;
    std::vector<char> insert_col(width, 0);
    int R = 0;
    for (int j = 0; j < width; ++j) {
        double gaps = 0;
        for (int r = 0; r < n_seq; ++r) {
            const char ch = aln[(size_t)r * width + j];
            if (ch == '-') gaps += 1.0;
            else if (acgt_code(ch) < 0) { err = "alignment holds a symbol outside ACGT-"; return false; }
        }
        insert_col[j] = gaps >= gap_limit;
        if (!insert_col[j]) ++R;
    }
    if (R < 1) { err = "alignment without a match column"; return false; }
    out.R = R;
    // label ids: kinds I (col 0..R), M (1..R), D (1..R)
    auto IL = [&](int i) { return i; };                 // I0..IR -> 0..R
    auto ML = [&](int i) { return (R + 1) + (i - 1); }; // M1..MR
    auto DL = [&](int i) { return (2 * R + 1) + (i - 1); };
    const int END = 3 * R + 1, START = 3 * R + 2, NL = 3 * R + 3;
    std::vector<int32_t> ecount((size_t)(2 * R + 1) * 4, 0);        // emission counts of I / M labels (label id < 2R+1)
    // transition counts: every label has at most these targets: I (same column index), M / D of the next
    // column, unit_end; keyed [label][0 I, 1 M, 2 D, 3 END]
    std::vector<int32_t> tcount((size_t)NL * 4, 0);
    std::vector<char> seen_label(NL, 0);
    std::vector<int32_t> walk;
    auto target_slot = [&](int from, int to) -> int {
        // which of the four target kinds `to` is, seen from `from`
        if (to == END) return 3;
        if (to <= R) return 0;
        if (to < 2 * R + 1) return 1;
        return 2;
    };
    (void)START;
    for (int r = 0; r < n_seq; ++r) {
        walk.clear();
        int col = 1;
        for (int j = 0; j < width; ++j) {
            const char ch = aln[(size_t)r * width + j];
            if (insert_col[j]) {
                if (ch != '-') {
                    const int lab = IL(col - 1);
                    walk.push_back(lab);
                    ecount[(size_t)lab * 4 + acgt_code(ch)]++;
                }
            } else {
                if (ch == '-') walk.push_back(DL(col));
                else {
                    const int lab = ML(col);
                    walk.push_back(lab);
                    ecount[(size_t)lab * 4 + acgt_code(ch)]++;
                }
                ++col;
            }
        }
        if (walk.empty()) { err = "alignment row without a state"; return false; }
        tcount[(size_t)START * 4 + target_slot(START, walk[0])]++;
        for (size_t k = 0; k + 1 < walk.size(); ++k) {
            tcount[(size_t)walk[k] * 4 + target_slot(walk[k], walk[k + 1])]++;
            seen_label[walk[k]] = 1;
        }
        tcount[(size_t)walk.back() * 4 + 3]++;
        seen_label[walk.back()] = 1;
    }
    // emissions (profile_hmm.py:55-71): count / seen + pseudo, renormalised; 1/4 without observations
    out.emis.assign((size_t)(2 * R + 1) * 4, 0.0);
    auto emis_row = [&](int lab, double* dst) {
        int seen = 0;
        for (int x = 0; x < 4; ++x) seen += ecount[(size_t)lab * 4 + x];
        if (seen > 0) {
            double v[4], norm = 0;
            for (int x = 0; x < 4; ++x) {
                v[x] = (1.0 * ecount[(size_t)lab * 4 + x]) / seen + pseudo;
                norm += 1.0 * v[x];
            }
            for (int x = 0; x < 4; ++x) dst[x] = v[x] / norm;
        } else {
            for (int x = 0; x < 4; ++x) dst[x] = 1.0 / 4;
        }
    };
    for (int i = 1; i <= R; ++i) emis_row(ML(i), &out.emis[(size_t)(i - 1) * 4]);
    for (int i = 0; i <= R; ++i) emis_row(IL(i), &out.emis[(size_t)(R + i) * 4]);
    // transitions (profile_hmm.py:72-149).  The row of a label holds exactly: unit_start, I0 -> {I0, D1, M1};
    // column i < R -> {I_i, D_{i+1}, M_{i+1}}; column R -> {I_R, unit_end}.
    out.trans.assign(prof_size(R), 0.0);
    auto row = [&](int lab, int src_kind, int i) {
        const bool last = (src_kind != PK_START) && (i == R) && !(src_kind == PK_I && R == 0);
        const int n_keys = last ? 2 : 3;
        int seen = 0;
        for (int t = 0; t < 4; ++t) seen += tcount[(size_t)lab * 4 + t];
        for (int t = 0; t < 4; ++t) {
            const bool exists = last ? (t == 0 || t == 3) : (t != 3);
            if (!exists) continue;
            double p;
            if (seen > 0) {
                p = 1.0 * tcount[(size_t)lab * 4 + t] / seen;
                p = (p + pseudo) / (1 + pseudo * n_keys);
            } else {
                p = last ? 1.0 / 2 : 1.0 / 3;
            }
            out.trans[prof_index(R, src_kind, i, t)] = p;
        }
    };
    row(START, PK_START, 0);
    for (int i = 0; i <= R; ++i) row(IL(i), PK_I, i);
    for (int i = 1; i <= R; ++i) { row(ML(i), PK_M, i); row(DL(i), PK_D, i); }
    return true;
}

// ---------------------------------------------------------------------------------------------
// kernel-side analysis of a shape: run model_compile.hpp ONCE on a copy whose weights are tags
// (edge k carries the value k), and read off where every edge lands in the device tables
// ---------------------------------------------------------------------------------------------
constexpr int kImageBytesPerCol = 208;   // = kImgBytesPerCol of kernels_banded.cuh: w10 | e2 | v1
constexpr int kImageE = 80, kImageV1 = 144;

inline bool analyse_shape(ShapeStructure& sh, std::string& err)
{
    const size_t E = sh.in_src.size();
    std::vector<double> tag_w(E), tag_e((size_t)sh.S * 4, 0.0);
    for (size_t k = 0; k < E; ++k) tag_w[k] = (double)k;
    advhmm_model_desc d{};
    d.n_states = sh.m; d.silent_start = sh.S; d.start_index = sh.start; d.end_index = sh.end;
    d.finite = sh.finite; d.n_symbols = 4;
    d.in_off = sh.in_off.data(); d.in_src = sh.in_src.data(); d.in_logp = tag_w.data(); d.emis = tag_e.data();
    if (!compile_model(d, sh.cm, err)) return false;
    const GenericTables& g = sh.cm.g;
    const BandedTables& b = sh.cm.b;
    sh.banded = b.valid;
    sh.why = b.why;
    sh.g_slot.resize(g.in_w.size());
    for (size_t k = 0; k < g.in_w.size(); ++k) sh.g_slot[k] = sh.edge_slot[(size_t)g.in_w[k]];
    // first-row candidates of the emitting states: silent sources only (row 0 is -inf on emitting states)
    sh.r1_off.assign(sh.S + 1, 0);
    sh.r1_edge.clear();
    for (int l = 0; l < sh.S; ++l) {
        for (int k = g.in_off[l]; k < g.in_off[l + 1]; ++k)
            if (g.in_src[k] >= sh.S) sh.r1_edge.push_back(k);
        sh.r1_off[l + 1] = (int32_t)sh.r1_edge.size();
    }
    if (!b.valid) return true;
    // tag -> g edge (g.in_w holds the tags in g order)
    std::vector<int32_t> g_of_tag(E, -1);
    for (size_t k = 0; k < g.in_w.size(); ++k) g_of_tag[(size_t)g.in_w[k]] = (int32_t)k;
    const size_t P = (size_t)b.NCpad;
    sh.img_w.clear();
    for (int ts = 0; ts < 9; ++ts)
        for (size_t c = 0; c < P; ++c) {
            const double v = b.w[(size_t)ts * P + c];
            if (v > kNegInf) sh.img_w.emplace_back((int32_t)(c * 10 + ts), g_of_tag[(size_t)v]);
        }
    for (size_t c = 0; c < P; ++c)
        if (b.accw[c] > kNegInf) sh.img_w.emplace_back((int32_t)(c * 10 + 9), g_of_tag[(size_t)b.accw[c]]);
    sh.fin_edge.resize(b.fin_w.size());
    for (size_t k = 0; k < b.fin_w.size(); ++k) sh.fin_edge[k] = g_of_tag[(size_t)b.fin_w[k]];
    sh.image_bytes = (int)(kImageBytesPerCol * P);
    sh.image0.assign((size_t)sh.image_bytes, 0);
    double* img = reinterpret_cast<double*>(sh.image0.data());
    for (size_t i = 0; i < (size_t)sh.image_bytes / 8; ++i) img[i] = kNegInf;
    return true;
}

// process-wide cache of shapes
struct ShapeCache {
    std::mutex mu;
    std::map<ShapeKey, std::shared_ptr<const ShapeStructure>> map;
    static ShapeCache& instance() { static ShapeCache c; return c; }
};

inline std::shared_ptr<const ShapeStructure> get_shape(const ShapeKey& key, std::string& err)
{
    ShapeCache& c = ShapeCache::instance();
    {
        std::lock_guard<std::mutex> lock(c.mu);
        auto it = c.map.find(key);
        if (it != c.map.end()) return it->second;
    }
    auto sh = std::make_shared<ShapeStructure>();
    if (!build_shape(key, *sh, err) || !analyse_shape(*sh, err)) return nullptr;
    std::lock_guard<std::mutex> lock(c.mu);
    return c.map.emplace(key, sh).first->second;   // a racing builder made the same structure: keep the first
}

// forget every shape (cold-start measurements; models that exist keep theirs alive)
inline void clear_shape_cache()
{
    ShapeCache& c = ShapeCache::instance();
    std::lock_guard<std::mutex> lock(c.mu);
    c.map.clear();
}

// ---------------------------------------------------------------------------------------------
// per-locus values
// ---------------------------------------------------------------------------------------------
// The vector exp the chains go through.  The reference applies numpy.exp (hmm.pyx:514), whose SIMD
// implementation differs from libm's exp in the last bit for ~3 % of the arguments, so the caller
// that wants the reference's bits passes numpy's (advhmm_set_vexp); the default is libm.
using VexpFn = void (*)(const double* in, double* out, int64_t n, void* user);

struct LocusInput {
    const uint8_t* left = nullptr;  int left_len = 0;      // the flank bases that enter the model (codes 0..3)
    const uint8_t* right = nullptr; int right_len = 0;
    const char* aln = nullptr; int n_seq = 0, width = 0;   // aligned repeat segments, row-major, ACGT-
    int copies = 0;
    double error_rate = 0.05;
};

struct LocusValues {
    std::shared_ptr<const ShapeStructure> shape;
    std::vector<double> slot_log;     // [n_slots] final log-probability of every parameter slot
    std::vector<double> emis_tab;     // [(5 + 2R + 1) * 4] log emission rows: 0..3 flank match on A,C,G,T; 4 uniform; M1..MR; I0..IR
    std::vector<uint8_t> flank;       // left codes then right codes
};

// base probability of a slot (before any log), hmm_utils.py / profile_hmm.py
inline double slot_probability(const Lab& l, const ShapeKey& k, double error_rate, const Profile& pr)
{
    const double p_ins = error_rate * 2 / 5;
    const double p_del = error_rate * 1 / 5;
    const double p_adv = 1 - p_ins - p_del;
    switch (l.sym) {
        case SYM_ONE: return 1.0;
        case SYM_IE: return p_ins;
        case SYM_DE: return p_del;
        case SYM_ADV: return p_adv;
        case SYM_ADV_M01: return p_adv - 0.01;
        case SYM_ONE_MINUS_IE: return 1 - p_ins;
        case SYM_ADV_OVER_L: return p_adv / k.Ll;
        case SYM_C001: return 0.01;
        case SYM_C05: return 0.5;
        case SYM_C03: return 0.3;
        case SYM_FIRST_COPY: return 0.7 / k.R;
        case SYM_TO_END: {
            const double to_end = 0.7 / (k.C * k.R);
            return to_end / (1 + to_end);
        }
        default: return pr.trans[(size_t)(l.sym - SYM_PROFILE0)];
    }
}

inline double log_or_neginf(double p) { return p > 0 ? std::log(p) : kNegInf; }

// Evaluate the slots of MANY loci with two vector-exp calls in total (numpy.exp costs a Python
// callback, so it is called on everything at once).  Values that are equal (same probability, same
// chain) are evaluated once per locus: a profile has a handful of distinct probabilities.
struct ChainBatch {
    struct Item { double p; uint8_t trips, div; double total; double out; };
    std::vector<Item> items;
    void run(VexpFn vexp, void* user)
    {
        const size_t n = items.size();
        std::vector<double> a(n), b(n);
        for (size_t i = 0; i < n; ++i) a[i] = log_or_neginf(items[i].p);             // hmm.pyx:433
        auto do_exp = [&](std::vector<double>& in, std::vector<double>& out) {
            if (vexp) vexp(in.data(), out.data(), (int64_t)n, user);
            else for (size_t i = 0; i < n; ++i) out[i] = std::exp(in[i]);
        };
        // first round trip (everything with trips >= 1; the others are computed along and ignored)
        do_exp(a, b);                                                                  // hmm.pyx:514
        std::vector<double> l2(n);
        for (size_t i = 0; i < n; ++i) l2[i] = log_or_neginf(b[i]);
        do_exp(l2, b);
        for (size_t i = 0; i < n; ++i) {
            Item& it = items[i];
            if (it.trips == 0) it.out = a[i];
            else if (it.trips == 1) it.out = l2[i];
            else {
                double e2 = b[i];
                if (it.div) e2 = e2 / it.total;                                        // hmm_utils.py:578-583
                it.out = log_or_neginf(e2);
            }
        }
    }
};

// Everything of a locus that does not need the exp: shape key, profile, emission log table, the
// distinct (probability, chain) items.  slot_item[s] = index (relative to `first_item`) of slot s.
struct LocusPrep {
    ShapeKey key{};
    Profile prof;
    std::vector<double> emis_tab;
    std::vector<uint8_t> flank;
    double error_rate = 0.05;
    std::string err;
    bool ok = false;
};

inline void prepare_locus(const LocusInput& in, LocusPrep& out)
{
    out.ok = false;
    out.error_rate = in.error_rate;
    if (!in.left || !in.right || in.left_len < 1 || in.right_len < 1 || !in.aln || in.copies < 1 || !(in.error_rate > 0)) {
        out.err = "locus needs flanks, aligned repeat segments, copies >= 1 and an error rate > 0";
        return;
    }
    if (!repeat_profile(in.aln, in.n_seq, in.width, in.error_rate, out.prof, out.err)) return;
    out.key = ShapeKey{in.left_len, in.right_len, out.prof.R, in.copies};
    out.flank.resize((size_t)in.left_len + in.right_len);
    for (int i = 0; i < in.left_len; ++i) out.flank[i] = in.left[i];
    for (int i = 0; i < in.right_len; ++i) out.flank[(size_t)in.left_len + i] = in.right[i];
    for (uint8_t c : out.flank)
        if (c > 3) { out.err = "flank holds a code outside 0..3"; return; }
    const int R = out.prof.R;
    out.emis_tab.assign((size_t)(5 + 2 * R + 1) * 4, 0.0);
    const double l97 = std::log(0.97), l01 = std::log(0.01), l25 = std::log(0.25);     // hmm_utils.py:359, :368-370
    for (int r = 0; r < 4; ++r)
        for (int x = 0; x < 4; ++x) out.emis_tab[(size_t)r * 4 + x] = (r == x) ? l97 : l01;
    for (int x = 0; x < 4; ++x) out.emis_tab[16 + x] = l25;
    for (size_t i = 0; i < out.prof.emis.size(); ++i) out.emis_tab[20 + i] = log_or_neginf(out.prof.emis[i]);
    out.ok = true;
}

// ---------------------------------------------------------------------------------------------
// the device tables of one locus (what the banded kernels read), written straight into the
// caller's buffers (pinned staging on the way to the device)
// ---------------------------------------------------------------------------------------------
struct LeanLayout {      // byte offsets inside the per-model device blob, 256-aligned sections
    size_t o_desc = 0, o_image = 0, o_tb1 = 0, o_tb0 = 0, o_finw = 0, o_cls = 0, bytes = 0;
};

inline LeanLayout lean_layout(const ShapeStructure& sh, size_t desc_bytes)
{
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    LeanLayout L;
    L.o_desc = 0;
    L.o_image = al(desc_bytes);
    L.o_tb1 = L.o_image + al((size_t)sh.image_bytes);
    L.o_tb0 = L.o_tb1 + al((size_t)4 * sh.S * sizeof(int32_t));
    L.o_finw = L.o_tb0 + al((size_t)sh.m * sizeof(int32_t));
    L.o_cls = L.o_finw + al(sh.cm.b.fin_w.size() * sizeof(double));
    L.bytes = L.o_cls + al((size_t)sh.m);
    return L;
}

struct LeanScratch { std::vector<double> in_w, v0; };

// -> logp of the empty read (v0[end]).  image / tb1 / tb0 / fin_w / classes as laid out by lean_layout.
inline double fill_lean(const LocusValues& lv, LeanScratch& sc, unsigned char* image, int32_t* tb1, int32_t* tb0,
                        double* fin_w, uint8_t* classes)
{
    const ShapeStructure& sh = *lv.shape;
    const GenericTables& g = sh.cm.g;
    const BandedTables& b = sh.cm.b;
    const int m = sh.m, S = sh.S;
    const size_t P = (size_t)b.NCpad;
    const size_t GE = sh.g_slot.size();
    sc.in_w.resize(GE);
    for (size_t k = 0; k < GE; ++k) sc.in_w[k] = lv.slot_log[sh.g_slot[k]];
    const double* in_w = sc.in_w.data();
    // row 0 (hmm.pyx:1999-2023): silent closure in index order
    sc.v0.assign(m, kNegInf);
    double* v0 = sc.v0.data();
    for (int l = 0; l < m; ++l) tb0[l] = -1;
    v0[g.start] = 0.0;
    for (int l = S; l < m; ++l) {
        if (l == g.start) continue;
        for (int k = g.in_off[l]; k < g.in_off[l + 1]; ++k) {
            const int ki = g.in_src[k];
            if (ki < S) continue;
            const double cand = v0[ki] + in_w[k];
            if (cand > v0[l]) { v0[l] = cand; tb0[l] = ki; }
        }
    }
    // image: weights
    memcpy(image, sh.image0.data(), (size_t)sh.image_bytes);
    double* w10 = reinterpret_cast<double*>(image);
    double* e2 = reinterpret_cast<double*>(image + (size_t)kImageE * P);
    double* v12 = reinterpret_cast<double*>(image + (size_t)kImageV1 * P);
    for (const auto& pe : sh.img_w) w10[pe.first] = in_w[pe.second];
    // emissions and first-row values of the emitting states
    const double* tab = lv.emis_tab.data();
    for (int x = 0; x < 4; ++x)
        for (int l = 0; l < S; ++l) tb1[(size_t)x * S + l] = -1;
    for (int l = 0; l < S; ++l) {
        const int row = sh.emis_row[l] >= 0 ? sh.emis_row[l] : (int)lv.flank[(size_t)(-sh.emis_row[l] - 1)];
        const double* e = tab + (size_t)row * 4;
        const size_t sl = b.slot_of[l] == SLOT_I ? 0 : 1, c = (size_t)b.col_of[l];
        for (int x = 0; x < 4; ++x) {
            e2[((size_t)x * P + c) * 2 + sl] = e[x];
            double best = kNegInf;
            int arg = -1;
            for (int q = sh.r1_off[l]; q < sh.r1_off[l + 1]; ++q) {
                const int k = sh.r1_edge[q];
                const double cand = v0[g.in_src[k]] + in_w[k] + e[x];
                if (cand > best) { best = cand; arg = g.in_src[k]; }
            }
            v12[((size_t)x * P + c) * 2 + sl] = best;
            tb1[(size_t)x * S + l] = arg;
        }
    }
    for (size_t k = 0; k < sh.fin_edge.size(); ++k) fin_w[k] = in_w[sh.fin_edge[k]];
    for (int i = 0; i < m; ++i) {
        uint8_t c = sh.base_class[i];
        if (i < S && sh.emis_row[i] < 0) c |= (uint8_t)(lv.flank[(size_t)(-sh.emis_row[i] - 1)] << 5);
        classes[i] = c;
    }
    return v0[g.end];
}

// the baked arrays of a locus in the reference's order (tests, lazily built full models)
inline void baked_values(const LocusValues& lv, std::vector<double>& in_logp, std::vector<double>& emis)
{
    const ShapeStructure& sh = *lv.shape;
    in_logp.resize(sh.edge_slot.size());
    for (size_t k = 0; k < in_logp.size(); ++k) in_logp[k] = lv.slot_log[sh.edge_slot[k]];
    emis.resize((size_t)sh.S * 4);
    for (int l = 0; l < sh.S; ++l) {
        const int row = sh.emis_row[l] >= 0 ? sh.emis_row[l] : (int)lv.flank[(size_t)(-sh.emis_row[l] - 1)];
        for (int x = 0; x < 4; ++x) emis[(size_t)l * 4 + x] = lv.emis_tab[(size_t)row * 4 + x];
    }
}

// ---------------------------------------------------------------------------------------------
// a small thread pool: run(n, fn) calls fn(i, worker) for i in [0, n) on up to n_threads workers
// ---------------------------------------------------------------------------------------------
inline int default_threads()
{
    int n = 0;
#ifdef __linux__
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) n = CPU_COUNT(&set);
#endif
    if (n <= 0) n = (int)std::thread::hardware_concurrency();
    return std::max(1, n);
}

template <typename Fn>
inline void parallel_for(size_t n, int n_threads, Fn&& fn)
{
    if (n == 0) return;
    const int nt = (int)std::min<size_t>((size_t)std::max(1, n_threads), n);
    std::atomic<size_t> next{0};
    std::exception_ptr first;
    std::atomic<bool> failed{false};
    auto work = [&](int worker) {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= n || failed.load()) return;
            try { fn(i, worker); }
            catch (...) { if (!failed.exchange(true)) first = std::current_exception(); }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    if (failed.load()) std::rethrow_exception(first);
}

}  // namespace rm
}  // namespace advhmm
