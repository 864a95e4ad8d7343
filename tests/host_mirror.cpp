// host_mirror.cpp -- TEST-ONLY sequential mirror of the device kernels' table-driven schedules.
//
// Not part of the product: libadvhmm.so contains no CPU DP.  This file is compiled by
// tests/conftest.py into tests/_host_mirror.so and exists so that the host-side model
// compiler (advntr_b200/csrc/model_compile.hpp) -- column assignment, first-row tables,
// collector, final-only states, traceback encoding -- can be checked against the oracle in
// a container without a GPU (test plan item 3).  It walks exactly the
// tables the CUDA kernels read, in plain loops instead of lanes.
#include <cstring>
#include <string>
#include <vector>

#include "../advntr_b200/csrc/model_compile.hpp"

using namespace advhmm;

namespace {

struct Best {
    double v = kNegInf;
    int arg = 0;
    inline void take(double cand, int a) { if (cand > v) { v = cand; arg = a; } }
};

double generic_viterbi(const GenericTables& g, const uint8_t* seq, int n, std::vector<int32_t>& path)
{
    const int m = g.m, S = g.S, K = g.K;
    std::vector<double> prev(g.v0), cur(m);
    std::vector<uint16_t> tb((size_t)(n + 1) * m, 0);
    for (int i = 0; i < n; ++i) {
        const int x = seq[i];
        for (int l = 0; l < S; ++l) {
            Best b; const double e = g.emis[(size_t)l * K + x];
            for (int k = g.in_off[l]; k < g.in_off[l + 1]; ++k) b.take(prev[g.in_src[k]] + g.in_w[k] + e, k - g.in_off[l]);
            cur[l] = b.v; tb[(size_t)(i + 1) * m + l] = (uint16_t)b.arg;
        }
        for (int L = 0; L < g.n_levels; ++L)
            for (int q = g.lvl_off[L]; q < g.lvl_off[L + 1]; ++q) {
                const int l = g.lvl_state[q];
                Best b;
                for (int k = g.in_off[l]; k < g.in_off[l + 1]; ++k) b.take(cur[g.in_src[k]] + g.in_w[k], k - g.in_off[l]);
                cur[l] = b.v; tb[(size_t)(i + 1) * m + l] = (uint16_t)b.arg;
            }
        prev.swap(cur);
    }
    double logp; int end;
    if (g.finite) { end = g.end; logp = prev[end]; }
    else { end = -1; logp = kNegInf; for (int l = 0; l < m; ++l) if (prev[l] > logp) { logp = prev[l]; end = l; } }
    path.clear();
    if (logp == kNegInf) return logp;
    int px = n, py = end;
    while (px > 0) {
        path.push_back(py);
        const int src = g.in_src[g.in_off[py] + tb[(size_t)px * m + py]];
        if (py < S) --px;
        py = src;
    }
    while (py != g.start) { path.push_back(py); py = g.tb0[py]; }
    path.push_back(py);
    std::reverse(path.begin(), path.end());
    return logp;
}

double banded_viterbi(const GenericTables& g, const BandedTables& b, const uint8_t* seq, int n,
                      std::vector<int32_t>& path)
{
    const int NC = b.NC, K = b.K;
    const size_t P = b.NCpad;
    path.clear();
    auto W = [&](int t, int s, int c) { return b.w[(size_t)(t * 3 + s) * P + c]; };
    // DP rows 1..n over (slot, column); traceback 6 bits per (row, column)
    std::vector<double> prev(3 * P, kNegInf), cur(3 * P, kNegInf);
    std::vector<uint8_t> tb((size_t)(n + 1) * P, 0);
    std::vector<uint16_t> acc_tb(n + 1, 0);
    for (int r = 1; r <= n; ++r) {
        const int x = seq[r - 1];
        Best acc;
        for (int c = 0; c < NC; ++c) {
            Best M, I, D;
            if (r == 1) {
                M.v = b.v1[((size_t)1 * K + x) * P + c];
                I.v = b.v1[((size_t)0 * K + x) * P + c];
            } else {
                const double eM = b.e[((size_t)1 * K + x) * P + c], eI = b.e[((size_t)0 * K + x) * P + c];
                if (c > 0)
                    for (int s = 0; s < 3; ++s) M.take(prev[s * P + c - 1] + W(SLOT_M, s, c) + eM, s);
                for (int s = 0; s < 3; ++s) I.take(prev[s * P + c] + W(SLOT_I, s, c) + eI, s);
            }
            if (c == b.acc_col) { D.v = acc.v; acc_tb[r] = (uint16_t)acc.arg; }
            else if (c > 0)
                for (int s = 0; s < 3; ++s) D.take(cur[s * P + c - 1] + W(SLOT_D, s, c), s);
            cur[SLOT_M * P + c] = M.v; cur[SLOT_I * P + c] = I.v; cur[SLOT_D * P + c] = D.v;
            tb[(size_t)r * P + c] = (uint8_t)(I.arg | (M.arg << 2) | (D.arg << 4));
            if (b.accw[c] != kNegInf) {
                // ordinal of this source among the collector's sources
                int ord = 0; while (b.acc_src_col[ord] != c) ++ord;
                acc.take(D.v + b.accw[c], ord);
            }
        }
        prev.swap(cur);
    }
    if (n == 0) {
        const double logp = g.v0[g.end];
        if (logp == kNegInf) return logp;
        int py = g.end;
        while (py != g.start) { path.push_back(py); py = g.tb0[py]; }
        path.push_back(py);
        std::reverse(path.begin(), path.end());
        return logp;
    }
    // final-only states on the last row
    const int NF = (int)b.fin_state.size();
    std::vector<double> fv(NF, kNegInf);
    std::vector<int32_t> ftb(NF, 0);
    for (int j = 0; j < NF; ++j) {
        Best f;
        for (int k = b.fin_off[j]; k < b.fin_off[j + 1]; ++k) {
            const int code = b.fin_src[k];
            const double sv = code < 0 ? fv[-(code + 1)] : prev[code];
            if (sv + b.fin_w[k] > f.v) { f.v = sv + b.fin_w[k]; f.arg = code; }
        }
        fv[j] = f.v; ftb[j] = f.arg;
    }
    const double logp = fv[b.end_final];
    if (logp == kNegInf) return logp;
    // backtrack
    int j = b.end_final, code;
    for (;;) { path.push_back(b.fin_state[j]); code = ftb[j]; if (code >= 0) break; j = -(code + 1); }
    int slot = code / (int)P, c = code % (int)P, r = n, state = -1;
    while (r >= 1) {
        const int s = b.st[slot][c];
        path.push_back(s);
        const uint8_t t = tb[(size_t)r * P + c];
        if (slot == SLOT_D) {
            if (c == b.acc_col) c = b.acc_src_col[acc_tb[r]];
            else { slot = (t >> 4) & 3; c -= 1; }
        } else if (r == 1) {
            state = b.tb1[(size_t)seq[0] * b.S + s]; r = 0;
        } else if (slot == SLOT_M) { slot = (t >> 2) & 3; c -= 1; r -= 1; }
        else { slot = t & 3; r -= 1; }
    }
    while (state != g.start) { path.push_back(state); state = g.tb0[state]; }
    path.push_back(state);
    std::reverse(path.begin(), path.end());
    return logp;
}

}  // namespace

extern "C" {

// kind: 0 generic tables, 1 banded tables.  Returns 0, or -1 if the model is not banded /
// malformed (why[] gets the reason).  path needs n + n_states entries.
int mirror_viterbi(const advhmm_model_desc* d, int kind, const uint8_t* seqs, const int64_t* seq_off,
                   int n_reads, double* logp, int32_t* path_len, int32_t* paths, int64_t stride,
                   char* why, int why_cap)
{
    CompiledModel cm; std::string err;
    if (!compile_model(*d, cm, err)) { strncpy(why, err.c_str(), why_cap - 1); return -1; }
    if (kind == 1 && !cm.b.valid) { strncpy(why, cm.b.why.c_str(), why_cap - 1); return -1; }
    std::vector<int32_t> path;
    for (int r = 0; r < n_reads; ++r) {
        const uint8_t* s = seqs + seq_off[r];
        const int n = (int)(seq_off[r + 1] - seq_off[r]);
        logp[r] = kind == 1 ? banded_viterbi(cm.g, cm.b, s, n, path) : generic_viterbi(cm.g, s, n, path);
        path_len[r] = logp[r] == kNegInf ? -1 : (int)path.size();
        if (path_len[r] > 0) memcpy(paths + (size_t)r * stride, path.data(), sizeof(int32_t) * path.size());
    }
    return 0;
}

// Shape of the compiled model: {valid, NC, n_final, acc_col, n_levels, n_edges}
int mirror_info(const advhmm_model_desc* d, int32_t* out6, char* why, int why_cap)
{
    CompiledModel cm; std::string err;
    if (!compile_model(*d, cm, err)) { strncpy(why, err.c_str(), why_cap - 1); return -1; }
    out6[0] = cm.b.valid; out6[1] = cm.b.NC; out6[2] = (int)cm.b.fin_state.size();
    out6[3] = cm.b.acc_col; out6[4] = cm.g.n_levels; out6[5] = cm.g.in_off[cm.g.m];
    strncpy(why, cm.b.why.c_str(), why_cap - 1);
    return 0;
}
}
