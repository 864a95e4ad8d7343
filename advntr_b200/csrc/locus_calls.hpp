// locus_calls.hpp: from the per-read Viterbi results of a locus to its genotype call -- host C++, no CUDA;
// part of libadvhmm.so (see advhmm.cu for the overview).
//
// What the reference does with the paths of a locus's reads, for whole batches of loci and on all host
// threads (it is what follows the decode in GenomeAnalyzer's per-locus loop, genome_analyzer.py:280):
//   recruit_read                                   vntr_finder.py:179-190
//   the better strand of a filtered unmapped read  vntr_finder.py:235-254
//   spanning test                                  vntr_finder.py:311-322
//   spanning + flanking counts -> observed list    vntr_finder.py:846-875
//   find_genotype_based_on_observed_repeats        vntr_finder.py:486-532 (+ get_conditional_likelihood, :466-484)
// The per-read inputs are the on-device path reducers' records (advhmm_read_summary), so no state path is
// needed.  max_prob is the reference's float: the same likelihoods are multiplied in the same order (ranked
// counts, stable), raised with libm's pow like Python's float.__pow__, summed left to right, and the first
// strictly larger posterior wins.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

#include "../../include/advhmm.h"

namespace advhmm {
namespace calls {

constexpr double kSequencingError = 0.03;        // r, vntr_finder.py:498
constexpr int kMinSupport = 3;                   // settings.ACCURACY_FILTER_SR_MIN_SUPPORT

// P(observing count ck | genotype (ci, cj)), vntr_finder.py:466-484
inline double conditional_likelihood(int ck, int ci, int cj, double r, double r_e)
{
    if (ck == ci && ci == cj) return 1 - r;
    if (cj == 0) return 0.5 * (1 - r);
    if (ck == ci) return 0.5 * ((1 - r) + std::pow(r_e, (double)std::abs(ck - cj)));
    if (ck == cj) return 0.5 * ((1 - r) + std::pow(r_e, (double)std::abs(ck - ci)));
    return 0.5 * (std::pow(r_e, (double)std::abs(ck - ci)) + std::pow(r_e, (double)std::abs(ck - cj)));
}

struct Genotype {
    bool found = false;
    int c1 = 0, c2 = 0;
    double max_prob = 1e-20;
};

// Counts of the distinct values of `v` in order of first appearance (a Python dict filled in a loop).
inline std::vector<std::pair<int, int>> tally(const int32_t* v, size_t n)
{
    std::vector<std::pair<int, int>> counts;
    for (size_t k = 0; k < n; ++k) {
        size_t p = 0;
        while (p < counts.size() && counts[p].first != v[k]) ++p;
        if (p == counts.size()) counts.emplace_back(v[k], 1);
        else ++counts[p].second;
    }
    return counts;
}

inline Genotype genotype_from_observed(const int32_t* observed, size_t n, bool haploid)
{
    std::vector<std::pair<int, int>> counts = tally(observed, n);
    double prior;
    if (counts.size() < 2) {
        prior = 0.5;
        size_t p = 0;                                  // counts[0] = 1: overwrites an observed 0, appends otherwise
        while (p < counts.size() && counts[p].first != 0) ++p;
        if (p == counts.size()) counts.emplace_back(0, 1);
        else counts[p].second = 1;
    } else {
        prior = 1.0 / ((double)(counts.size() * (counts.size() - 1)) / 2);
    }
    std::stable_sort(counts.begin(), counts.end(),
                     [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.second > b.second; });
    const double r = kSequencingError, r_e = r / (2 + r);
    const size_t m = counts.size();
    Genotype g;
    // posterior of every pair (i, j >= i) in the reference's key order; total first, then the first maximum
    std::vector<double> post;
    std::vector<std::pair<int, int>> keys;
    bool any_observed = false;
    for (size_t k = 0; k < m; ++k) any_observed |= counts[k].first != 0;
    if (!any_observed) return g;
    for (size_t i = 0; i < m; ++i)
        for (size_t j = i; j < m; ++j) {
            if (haploid && i != j) continue;
            const int ci = counts[i].first, cj = counts[j].first;
            double prod = 0.0;
            bool first = true;
            for (size_t k = 0; k < m; ++k) {
                const int ck = counts[k].first;
                if (ck == 0) continue;
                const double f = std::pow(conditional_likelihood(ck, ci, cj, r, r_e), (double)counts[k].second);
                prod = first ? f : prod * f;
                first = false;
            }
            post.push_back(prod * prior);
            keys.emplace_back(ci, cj);
        }
    double total = 0.0;
    for (size_t p = 0; p < post.size(); ++p) total = p ? total + post[p] : post[p];
    for (size_t p = 0; p < post.size(); ++p) {
        const double q = post[p] / total;
        if (q > g.max_prob) { g.max_prob = q; g.found = true; g.c1 = keys[p].first; g.c2 = keys[p].second; }
    }
    return g;
}

// Counter(counts).most_common() with the values below `min_support` dropped (vntr_finder.py:859-866)
inline std::vector<int32_t> drop_unsupported(const std::vector<int32_t>& v)
{
    std::vector<std::pair<int, int>> counts = tally(v.data(), v.size());
    std::stable_sort(counts.begin(), counts.end(),
                     [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.second > b.second; });
    std::vector<int32_t> kept;
    for (const auto& c : counts)
        if (c.second >= kMinSupport) kept.insert(kept.end(), (size_t)c.second, c.first);
    return kept;
}

// spanning counts + (when at least five flanking reads agree on the largest lower bound and it is not below
// the largest spanning count) that lower bound, vntr_finder.py:850-879
inline Genotype genotype_from_illumina_counts(std::vector<int32_t> covered, std::vector<int32_t> flanking,
                                              bool accuracy_filter, bool haploid)
{
    std::sort(flanking.begin(), flanking.end());
    const int floor_ = covered.empty() ? 0 : *std::max_element(covered.begin(), covered.end());
    std::vector<int32_t> top;
    if (!flanking.empty()) {
        const int mx = flanking.back();
        if (mx >= floor_)
            for (int32_t x : flanking)
                if (x == mx) top.push_back(x);
    }
    if (top.size() < 5) top.clear();
    if (accuracy_filter) {
        covered = drop_unsupported(covered);
        top.clear();
    }
    covered.insert(covered.end(), top.begin(), top.end());
    return genotype_from_observed(covered.data(), covered.size(), haploid);
}

struct ReadView {
    const double* logp;
    const advhmm_read_summary* S;
    const int32_t* path_len;
    const int64_t* seq_off;
};

inline double flank_rate(const advhmm_read_summary& s)
{
    const double right = s.right_bp > 0 ? (double)s.right_hits / (double)s.right_bp : 1.0;
    const double left = s.left_bp > 0 ? (double)s.left_hits / (double)s.left_bp : 1.0;
    return right < left ? right : left;
}

// recruit_read (vntr_finder.py:179-190); `score` NaN = no minimum Viterbi score known for the locus
inline bool recruit_read(const ReadView& R, int64_t i, double score, double rate)
{
    const bool possible = R.path_len[i] >= 0;
    if (score == score) return R.logp[i] > score && rate >= 0.9 && possible;
    const double len = (double)(R.seq_off[i + 1] - R.seq_off[i]);
    return (double)R.S[i].n_match >= 0.9 * len && R.logp[i] > -len && rate >= 0.9 && possible;
}

// One locus: reads [a, a + n_mapped) are its mapped reads, then both strands of n_unm filtered unmapped reads.
// `score` NaN = no minimum Viterbi score known for the locus (recruit_read's fallback rule).
inline void call_locus(const ReadView& R, int64_t a, int32_t n_mapped, int32_t n_unm, double score, bool accuracy_filter,
                       bool haploid, int32_t min_repeat_bp, advhmm_locus_call& out, uint8_t* read_class)
{
    std::vector<int32_t> covered, flanking;
    int32_t recruited = 0;
    auto consider = [&](int64_t i) {
        const advhmm_read_summary& s = R.S[i];
        const double rate = flank_rate(s);
        if (!recruit_read(R, i, score, rate)) return;
        ++recruited;
        const bool spanning = rate >= 0.95 && s.left_bp > 5 && s.right_bp > 5;
        (spanning ? covered : flanking).push_back(s.repeats);
        if (read_class) read_class[i] = spanning ? 1 : 2;
    };
    for (int64_t i = a; i < a + n_mapped; ++i) consider(i);
    for (int32_t u = 0; u < n_unm; ++u) {             // the better strand of every unmapped read
        const int64_t f = a + n_mapped + 2 * (int64_t)u;
        const int64_t best = R.logp[f] < R.logp[f + 1] ? f + 1 : f;
        if (R.S[best].repeat_bp > min_repeat_bp) consider(best);
    }
    if (accuracy_filter) flanking.clear();
    const int32_t n_spanning = (int32_t)(accuracy_filter ? drop_unsupported(covered).size() : covered.size());
    const int32_t n_flanking = (int32_t)flanking.size();
    const Genotype g = genotype_from_illumina_counts(std::move(covered), std::move(flanking), accuracy_filter, haploid);
    out.has_call = g.found ? 1 : 0;
    out.c1 = g.c1;
    out.c2 = g.c2;
    out.recruited = recruited;
    out.spanning = n_spanning;
    out.flanking = n_flanking;
    out.max_prob = g.max_prob;
}

// ---------------------------------------------------------------------------------------------
// --frameshift mode: find_frameshift_from_selected_reads up to its binomial test (vntr_finder.py:265-300)
// on the full state paths of a locus's recruited reads.  The reference reads state NAMES; here a state is
// its class byte (include/advhmm.h: kind / part, as for the on-device reducers) plus, for the insert and
// delete states of the repeat units, the number in its name ("I7_2" -> 7).
// ---------------------------------------------------------------------------------------------
struct Mutation {
    int kind, column, base, count;                  // kind 2 = I, 3 = D (class-byte kinds); base -1 for deletions
};

struct PathView {
    const int32_t* path;                            // the whole path, model start / end states included
    int32_t len;
    const uint8_t* cls;                             // class byte per state of the model
    const int32_t* label;                           // number in the name of a repeat-unit I / D state
    const uint8_t* seq;                             // the read as it was decoded, codes 0..3
};

// adds the frame-shifting indel states of one read to `mut` (insertion order kept: a Python dict)
inline void frameshift_mutations_of_read(const PathView& v, int pattern_len, std::vector<Mutation>& mut,
                                         std::vector<int32_t>& lengths, std::vector<std::pair<int32_t, int32_t>>& first_visit)
{
    lengths.clear();
    first_visit.clear();
    const int32_t a = 1, b = v.len - 1;             // vpath[1:-1]
    // lengths of the repeat units on the path (get_repeating_pattern_lengths, hmm_utils.py:122-141)
    {
        int32_t bp = 0, open_at = -1;
        for (int32_t k = a; k < b; ++k) {
            const uint8_t c = v.cls[v.path[k]];
            const int kind = c & 7;
            if (kind == 1 || kind == 2) ++bp;
            if (kind == 5 && open_at >= 0) lengths.push_back(bp - open_at);
            if (kind == 4) open_at = bp;
        }
    }
    int32_t unit = -1, pos = 0;                     // pos: read bases emitted before the current state
    for (int32_t k = a; k < b; ++k) {
        const int32_t st = v.path[k];
        const uint8_t c = v.cls[st];
        const int kind = c & 7, part = (c >> 3) & 3;
        const int32_t here = pos;
        if (kind == 1 || kind == 2) ++pos;
        if (kind == 2 && part == 3) {               // first visit of an insert state: the base it emitted then
            size_t f = 0;
            while (f < first_visit.size() && first_visit[f].first != st) ++f;
            if (f == first_visit.size()) first_visit.emplace_back(st, here);
        }
        if (kind == 4) ++unit;
        if (kind != 2 && kind != 3) continue;
        if (part != 3) continue;                    // flank states: names end with "fix"
        if (unit < 0 || unit >= (int32_t)lengths.size()) continue;
        const int32_t len = lengths[unit];
        if (len == pattern_len || std::abs(len - pattern_len) > 2) continue;
        int base = -1;
        if (kind == 2) {
            size_t f = 0;
            while (first_visit[f].first != st) ++f;
            base = v.seq[first_visit[f].second];
        }
        const int column = v.label[st];
        size_t m = 0;
        while (m < mut.size() && !(mut[m].kind == kind && mut[m].column == column && mut[m].base == base)) ++m;
        if (m == mut.size()) mut.push_back(Mutation{kind, column, base, 1});
        else ++mut[m].count;
    }
}

// sorted(mutations.items(), key=count)[-1]: the largest count, the LAST inserted among equals
inline const Mutation* frameshift_candidate(const std::vector<Mutation>& mut)
{
    const Mutation* best = nullptr;
    for (const Mutation& m : mut)
        if (!best || m.count >= best->count) best = &m;
    return best;
}

}  // namespace calls
}  // namespace advhmm
