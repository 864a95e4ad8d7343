// model_compile.hpp -- host-side compilation of a baked HMM into device tables.
//
// Input: the arrays pomegranate's bake() produces (/root/reference/pomegranate/hmm.pyx:844-1123),
// passed through the C-ABI as `advhmm_model_desc`.  Output:
//
//   GenericTables  -- any model.  In-edge CSR with the reference's candidate order made
//                     explicit, the row-0 closure (hmm.pyx:1999-2023) evaluated once, and a
//                     level schedule of the silent states (states of one level are independent).
//   BandedTables   -- profile-shaped models (what hmm_utils.get_read_matcher_model builds,
//                     /root/reference/advntr/hmm_utils.py:553-595): every live silent state is
//                     one column; a column holds up to three slots I (emitting, self loop),
//                     M (emitting) and D (silent) with in-edges only from the previous column
//                     (M, D) or the same column (I).  Everything that does not fit is either
//                     provably irrelevant after the first row (sources that no emitting state
//                     can reach), evaluated only on the last row (silent states from which no
//                     emitting state is reachable) or a single "collector" silent state fed by
//                     D slots of earlier columns (end_repeating_pattern_match).  If a model
//                     does not validate, `banded.valid` is false and the generic kernel is used.
//
// All arithmetic here (row 0, first-row tables) is IEEE-754 double add / strict compare in
// the reference's order, so it is bit-identical to what the reference computes at run time.
// Pure C++17, no CUDA: also compiled into the test-only host mirror (tests/host_mirror.cpp).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <string>
#include <vector>

#include "../../include/advhmm.h"

namespace advhmm {

static constexpr double kNegInf = -std::numeric_limits<double>::infinity();

enum Slot : int { SLOT_I = 0, SLOT_M = 1, SLOT_D = 2, SLOT_FINAL = 3, SLOT_NONE = 4 };

struct GenericTables {
    int m = 0, S = 0, K = 0, start = 0, end = 0, finite = 0;
    std::vector<int32_t> in_off;     // [m+1]  candidate lists, reference order made explicit
    std::vector<int32_t> in_src;     // [E]
    std::vector<double>  in_w;       // [E]
    std::vector<double>  emis;       // [S*K]
    std::vector<double>  v0;         // [m]   row-0 Viterbi values
    std::vector<int32_t> tb0;        // [m]   row-0 predecessor state, -1 = none
    std::vector<int32_t> lvl_off;    // [n_levels+1]
    std::vector<int32_t> lvl_state;  // [m-S] silent states grouped by level
    int n_levels = 0;
    int max_in_degree = 0;
};

struct BandedTables {
    bool valid = false;
    std::string why;                 // reason the model is not banded (diagnostics)
    int NC = 0;                      // columns
    int NCpad = 0;                   // padded to a multiple of 16 (bank-conflict-free rows)
    int S = 0, K = 4;
    std::vector<int32_t> st[3];      // [NC] state index of slot I/M/D, -1 if absent
    std::vector<int32_t> col_of;     // [m]  column of a state (-1 for final-only states)
    std::vector<int8_t>  slot_of;    // [m]  Slot
    // transition weights, SoA: w[(t*3+s)*NCpad + c] = log p(source slot s -> target slot t of column c)
    // sources: target M and D read column c-1, target I reads column c.  -inf where no edge.
    std::vector<double>  w;          // [9*NCpad]
    std::vector<double>  e;          // [(slot(I=0,M=1)*K + sym)*NCpad + c] emission log-probs
    std::vector<double>  v0d;        // [NCpad] row-0 value of the D slot of each column
    std::vector<double>  v1;         // [(slot*K + sym)*NCpad + c] first-row values of I/M slots
    std::vector<int32_t> tb1;        // [K*S] first-row predecessor state of emitting state l given sym
    // collector: silent state at column acc_col fed by D slots of earlier columns
    int acc_col = -1;
    std::vector<double>  accw;       // [NCpad] weight D(c) -> collector, -inf if none
    std::vector<int32_t> acc_src_col;// source columns in candidate order
    // final-only silent states (topological order) and their candidate lists.
    // source code >= 0: slot*NCpad + col of a banded state;  < 0: -(ordinal+1) of a final state
    std::vector<int32_t> fin_state;
    std::vector<int32_t> fin_off, fin_src;
    std::vector<double>  fin_w;
    int end_final = -1;              // ordinal of the end state among the final states
    bool nonpositive = false;        // every table entry <= 0: DP values can be ordered as integers
};

// fp32 twin of the banded tables (ADVHMM_FP32): every table rounded to float once, and the two
// host-evaluated pieces (row-0 closure, first-row tables) re-evaluated in float arithmetic with
// the same operation order, so the device fp32 DP equals a float restatement of the reference.
struct BandedF32 {
    std::vector<float>   w;          // [9*NCpad]  as BandedTables::w
    std::vector<float>   accw;       // [NCpad]
    std::vector<float>   e;          // [2*K*NCpad]
    std::vector<float>   v1;         // [2*K*NCpad]
    std::vector<int32_t> tb1;        // [K*S]
    std::vector<float>   v0;         // [m]  row-0 values in float
    std::vector<int32_t> tb0;        // [m]
    std::vector<float>   fin_w;
};

struct CompiledModel {
    GenericTables g;
    BandedTables b;
    BandedF32 f;
};

namespace detail {

inline bool build_generic(const advhmm_model_desc& d, GenericTables& g, std::string& err)
{
    const int m = d.n_states, S = d.silent_start, K = d.n_symbols;
    if (m <= 0 || S < 0 || S > m || !d.in_off || d.start_index < 0 ||
        d.start_index >= m || d.end_index < 0 || d.end_index >= m) {
        err = "malformed model descriptor";
        return false;
    }
    // reads are packed 2 bits per symbol on the device: alphabets of up to four symbols only
    if (K < 1 || K > 4) {
        err = "n_symbols must be 1..4 (reads are packed 2 bits per symbol)";
        return false;
    }
    if (d.in_off[0] != 0) {
        err = "in_off[0] must be 0";
        return false;
    }
    const int E = d.in_off[m];
    if (E < 0 || (E > 0 && (!d.in_src || !d.in_logp)) || (S > 0 && !d.emis)) {
        err = "malformed model descriptor (edge arrays)";
        return false;
    }
    g.m = m; g.S = S; g.K = K; g.start = d.start_index; g.end = d.end_index; g.finite = d.finite;
    g.emis.assign(d.emis, d.emis + (size_t)S * K);
    g.in_off.assign(m + 1, 0);
    g.in_src.clear(); g.in_w.clear();
    g.in_src.reserve(E); g.in_w.reserve(E);
    for (int l = 0; l < m; ++l) {
        const int a = d.in_off[l], b = d.in_off[l + 1];
        if (a > b || b > E) { err = "in_off is not monotone"; return false; }
        for (int k = a; k < b; ++k)
            if (d.in_src[k] < 0 || d.in_src[k] >= m) { err = "in_src out of range"; return false; }
        if (l < S) {
            // emitting target: one pass in in-edge order (hmm.pyx:2026-2042)
            for (int k = a; k < b; ++k) { g.in_src.push_back(d.in_src[k]); g.in_w.push_back(d.in_logp[k]); }
        } else {
            // silent target: emitting sources first (pass 1, hmm.pyx:2044-2063), then silent
            // sources with a smaller index (pass 2, :2065-2083); silent sources >= l are never read.
            for (int k = a; k < b; ++k)
                if (d.in_src[k] < S) { g.in_src.push_back(d.in_src[k]); g.in_w.push_back(d.in_logp[k]); }
            for (int k = a; k < b; ++k)
                if (d.in_src[k] >= S && d.in_src[k] < l) { g.in_src.push_back(d.in_src[k]); g.in_w.push_back(d.in_logp[k]); }
        }
        g.in_off[l + 1] = (int32_t)g.in_src.size();
        g.max_in_degree = std::max(g.max_in_degree, g.in_off[l + 1] - g.in_off[l]);
    }
    // row 0 (hmm.pyx:1999-2023): start = 0, silent closure in index order, start itself skipped
    g.v0.assign(m, kNegInf);
    g.tb0.assign(m, -1);
    g.v0[g.start] = 0.0;
    for (int l = S; l < m; ++l) {
        if (l == g.start) continue;
        for (int k = g.in_off[l]; k < g.in_off[l + 1]; ++k) {
            const int ki = g.in_src[k];
            if (ki < S) continue;
            const double cand = g.v0[ki] + g.in_w[k];
            if (cand > g.v0[l]) { g.v0[l] = cand; g.tb0[l] = ki; }
        }
    }
    // silent levels
    std::vector<int> level(m, 0);
    int nl = 0;
    for (int l = S; l < m; ++l) {
        int lv = 0;
        for (int k = g.in_off[l]; k < g.in_off[l + 1]; ++k)
            if (g.in_src[k] >= S) lv = std::max(lv, level[g.in_src[k]] + 1);
        level[l] = lv;
        nl = std::max(nl, lv + 1);
    }
    if (m == S) nl = 0;
    g.n_levels = nl;
    g.lvl_off.assign(nl + 1, 0);
    for (int l = S; l < m; ++l) g.lvl_off[level[l] + 1]++;
    for (int i = 0; i < nl; ++i) g.lvl_off[i + 1] += g.lvl_off[i];
    g.lvl_state.assign(m - S, 0);
    std::vector<int> fill(g.lvl_off.begin(), g.lvl_off.end());
    for (int l = S; l < m; ++l) g.lvl_state[fill[level[l]]++] = l;
    return true;
}

inline void build_banded(const GenericTables& g, BandedTables& b)
{
    const int m = g.m, S = g.S, K = g.K;
    auto fail = [&](const std::string& why) { b.valid = false; b.why = why; };
    b.valid = false;
    if (K != 4) return fail("alphabet is not 4 symbols");
    if (!g.finite) return fail("model has no end state in-edges (infinite)");
    if (S == 0) return fail("no emitting states");

    // out-adjacency over the evaluated edges
    std::vector<int> out_off(m + 1, 0);
    for (int l = 0; l < m; ++l)
        for (int k = g.in_off[l]; k < g.in_off[l + 1]; ++k) out_off[g.in_src[k] + 1]++;
    for (int i = 0; i < m; ++i) out_off[i + 1] += out_off[i];
    std::vector<int> out_dst(out_off[m]);
    {
        std::vector<int> f(out_off.begin(), out_off.end() - 1);
        for (int l = 0; l < m; ++l)
            for (int k = g.in_off[l]; k < g.in_off[l + 1]; ++k) out_dst[f[g.in_src[k]]++] = l;
    }
    // fromEmit[s]: s can hold a finite value on some row >= 1  (reachable from an emitting state)
    std::vector<char> fromEmit(m, 0), toEmit(m, 0);
    std::vector<int> stack;
    for (int l = 0; l < S; ++l) { fromEmit[l] = 1; stack.push_back(l); }
    while (!stack.empty()) {
        int s = stack.back(); stack.pop_back();
        for (int k = out_off[s]; k < out_off[s + 1]; ++k)
            if (!fromEmit[out_dst[k]]) { fromEmit[out_dst[k]] = 1; stack.push_back(out_dst[k]); }
    }
    // toEmit[s]: some emitting state is reachable from s (its value matters before the last row)
    for (int l = 0; l < S; ++l) { toEmit[l] = 1; stack.push_back(l); }
    while (!stack.empty()) {
        int s = stack.back(); stack.pop_back();
        for (int k = g.in_off[s]; k < g.in_off[s + 1]; ++k)
            if (!toEmit[g.in_src[k]]) { toEmit[g.in_src[k]] = 1; stack.push_back(g.in_src[k]); }
    }
    auto isFinal = [&](int s) { return s >= S && !toEmit[s]; };
    auto isRow0Only = [&](int s) { return s >= S && !fromEmit[s]; };
    if (!isFinal(g.end)) return fail("end state feeds an emitting state");

    // columns: one per non-final silent state, in topological (baked) order
    b.col_of.assign(m, -1);
    b.slot_of.assign(m, (int8_t)SLOT_NONE);
    int NC = 0;
    for (int s = S; s < m; ++s) {
        if (isFinal(s)) { b.slot_of[s] = SLOT_FINAL; continue; }
        b.col_of[s] = NC++;
        b.slot_of[s] = SLOT_D;
    }
    if (NC == 0) return fail("no live silent states");
    // emitting states: I (self loop) sits in the column of its silent source, M one to the right
    for (int l = 0; l < S; ++l) {
        bool self = false;
        int best_live = -1, best_r0 = -1;
        for (int k = g.in_off[l]; k < g.in_off[l + 1]; ++k) {
            const int a = g.in_src[k];
            if (a == l) self = true;
            if (a >= S) {
                if (isFinal(a)) return fail("final-only state feeds an emitting state");
                if (isRow0Only(a)) best_r0 = std::max(best_r0, b.col_of[a]);
                else best_live = std::max(best_live, b.col_of[a]);
            }
        }
        const int sc = best_live >= 0 ? best_live : best_r0;
        if (sc < 0) return fail("emitting state without a silent source");
        b.slot_of[l] = self ? SLOT_I : SLOT_M;
        b.col_of[l] = self ? sc : sc + 1;
        if (b.col_of[l] >= NC) return fail("match state beyond the last column");
    }
    b.NC = NC;
    b.NCpad = (NC + 15) / 16 * 16;
    b.S = S; b.K = K;
    for (int t = 0; t < 3; ++t) b.st[t].assign(NC, -1);
    for (int s = 0; s < m; ++s) {
        if (b.slot_of[s] > SLOT_D) continue;
        int32_t& cell = b.st[b.slot_of[s]][b.col_of[s]];
        if (cell != -1) return fail("two states share a column slot");
        cell = s;
    }
    const size_t P = (size_t)b.NCpad;
    b.w.assign(9 * P, kNegInf);
    b.e.assign(2 * K * P, kNegInf);
    b.v0d.assign(P, kNegInf);
    b.accw.assign(P, kNegInf);
    b.acc_col = -1;
    b.acc_src_col.clear();

    // validate every candidate list of a banded state and scatter its weights
    for (int l = 0; l < m; ++l) {
        const int t = b.slot_of[l];
        if (t > SLOT_D) continue;
        const int c = b.col_of[l];
        const int want_col = (t == SLOT_I) ? c : c - 1;
        // does this silent target collect from far columns?
        bool collector = false;
        if (t == SLOT_D)
            for (int k = g.in_off[l]; k < g.in_off[l + 1]; ++k) {
                const int a = g.in_src[k];
                if (!isRow0Only(a) && b.col_of[a] != want_col) collector = true;
            }
        if (collector) {
            if (b.acc_col != -1) return fail("more than one collector state");
            b.acc_col = c;
            int last = -1;
            for (int k = g.in_off[l]; k < g.in_off[l + 1]; ++k) {
                const int a = g.in_src[k];
                if (isRow0Only(a)) continue;  // -inf after row 0; row 0 is precomputed
                if (b.slot_of[a] != SLOT_D) return fail("collector fed by a non-silent state");
                if (b.col_of[a] <= last || b.col_of[a] >= c) return fail("collector sources out of order");
                last = b.col_of[a];
                b.accw[last] = g.in_w[k];
                b.acc_src_col.push_back(last);
            }
            continue;
        }
        int last_slot = -1;
        for (int k = g.in_off[l]; k < g.in_off[l + 1]; ++k) {
            const int a = g.in_src[k];
            const bool regular = b.slot_of[a] <= SLOT_D && b.col_of[a] == want_col;
            if (!regular) {
                if (isRow0Only(a)) continue;   // far source that is -inf on every row >= 1
                return fail("edge outside the band");
            }
            const int s = b.slot_of[a];
            if (s <= last_slot) {
                // candidate order differs from the kernel's fixed I, M, D order.  Harmless only
                // if the out-of-order source can never be finite after row 0.
                if (!isRow0Only(a)) return fail("candidate order is not I, M, D");
            }
            last_slot = std::max(last_slot, s);
            b.w[(size_t)(t * 3 + s) * P + c] = g.in_w[k];
        }
    }
    // a row-0-only source listed out of order inside the band must not be finite with a live one:
    // it is -inf on rows >= 1 (where the kernel evaluates), so the fixed order is exact.

    for (int l = 0; l < S; ++l)
        for (int x = 0; x < K; ++x)
            b.e[((size_t)(b.slot_of[l] == SLOT_I ? 0 : 1) * K + x) * P + b.col_of[l]] = g.emis[(size_t)l * K + x];
    for (int c = 0; c < NC; ++c)
        if (b.st[SLOT_D][c] >= 0) b.v0d[c] = g.v0[b.st[SLOT_D][c]];

    // first-row tables: emitting states of row 1 from the full row 0 (any source, true order)
    b.v1.assign(2 * K * P, kNegInf);
    b.tb1.assign((size_t)K * S, -1);
    for (int l = 0; l < S; ++l)
        for (int x = 0; x < K; ++x) {
            double best = kNegInf; int arg = -1;
            const double e = g.emis[(size_t)l * K + x];
            for (int k = g.in_off[l]; k < g.in_off[l + 1]; ++k) {
                const double cand = g.v0[g.in_src[k]] + g.in_w[k] + e;
                if (cand > best) { best = cand; arg = g.in_src[k]; }
            }
            b.v1[((size_t)(b.slot_of[l] == SLOT_I ? 0 : 1) * K + x) * P + b.col_of[l]] = best;
            b.tb1[(size_t)x * S + l] = arg;
        }

    // final-only states
    b.fin_state.clear(); b.fin_off.assign(1, 0); b.fin_src.clear(); b.fin_w.clear();
    std::vector<int> fin_ord(m, -1);
    for (int s = S; s < m; ++s)
        if (isFinal(s)) { fin_ord[s] = (int)b.fin_state.size(); b.fin_state.push_back(s); }
    for (int s : b.fin_state) {
        for (int k = g.in_off[s]; k < g.in_off[s + 1]; ++k) {
            const int a = g.in_src[k];
            if (fin_ord[a] >= 0) b.fin_src.push_back(-(fin_ord[a] + 1));
            else b.fin_src.push_back((int32_t)(b.slot_of[a] * P + b.col_of[a]));
            b.fin_w.push_back(g.in_w[k]);
        }
        b.fin_off.push_back((int32_t)b.fin_src.size());
    }
    b.end_final = fin_ord[g.end];
    if (b.fin_state.size() > 32) return fail("too many final-only states");
    if (b.acc_src_col.size() > 65535) return fail("collector has too many sources");
    auto le0 = [](const std::vector<double>& v) { for (double x : v) if (x > 0.0 || (x == 0.0 && std::signbit(x))) return false; return true; };
    b.nonpositive = le0(b.w) && le0(b.e) && le0(b.v1) && le0(b.accw) && le0(b.fin_w) && le0(g.v0);
    b.valid = true;
}

}  // namespace detail

namespace detail {
inline void build_banded_f32(const GenericTables& g, const BandedTables& b, BandedF32& f)
{
    const int m = g.m, S = g.S, K = g.K;
    const size_t P = (size_t)b.NCpad;
    auto cast = [](const std::vector<double>& v) { return std::vector<float>(v.begin(), v.end()); };
    f.w = cast(b.w); f.accw = cast(b.accw); f.e = cast(b.e); f.fin_w = cast(b.fin_w);
    const float ninf = -std::numeric_limits<float>::infinity();
    f.v0.assign(m, ninf);
    f.tb0.assign(m, -1);
    f.v0[g.start] = 0.0f;
    for (int l = S; l < m; ++l) {
        if (l == g.start) continue;
        for (int k = g.in_off[l]; k < g.in_off[l + 1]; ++k) {
            const int ki = g.in_src[k];
            if (ki < S) continue;
            const float cand = f.v0[ki] + (float)g.in_w[k];
            if (cand > f.v0[l]) { f.v0[l] = cand; f.tb0[l] = ki; }
        }
    }
    f.v1.assign(2 * K * P, ninf);
    f.tb1.assign((size_t)K * S, -1);
    for (int l = 0; l < S; ++l)
        for (int x = 0; x < K; ++x) {
            float best = ninf; int arg = -1;
            const float e = (float)g.emis[(size_t)l * K + x];
            for (int k = g.in_off[l]; k < g.in_off[l + 1]; ++k) {
                const float cand = (f.v0[g.in_src[k]] + (float)g.in_w[k]) + e;
                if (cand > best) { best = cand; arg = g.in_src[k]; }
            }
            f.v1[((size_t)(b.slot_of[l] == SLOT_I ? 0 : 1) * K + x) * P + b.col_of[l]] = best;
            f.tb1[(size_t)x * S + l] = arg;
        }
}
}  // namespace detail

inline bool compile_model(const advhmm_model_desc& d, CompiledModel& out, std::string& err)
{
    if (!detail::build_generic(d, out.g, err)) return false;
    detail::build_banded(out.g, out.b);
    if (out.b.valid) detail::build_banded_f32(out.g, out.b, out.f);
    return true;
}

}  // namespace advhmm
