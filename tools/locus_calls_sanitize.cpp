// Sanitizer harness for the native step after the decode (advntr_b200/csrc/locus_calls.hpp is pure C++17):
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off \
//       -o /tmp/calls_asan tools/locus_calls_sanitize.cpp && /tmp/calls_asan
// Random per-read results (empty loci, impossible reads, zero-length flanks, absurd repeat counts) go through
// call_locus with every flag combination; random state paths over random class tables (paths of length 0..2,
// unit_end before any unit_start, insert states revisited, labels of every size) go through the frameshift walk.
#include <cstdio>
#include <random>

#include "../advntr_b200/csrc/locus_calls.hpp"

using namespace advhmm;

int main()
{
    std::mt19937 rng(4242);
    auto uni = [&](int lo, int hi) { return lo + (int)(rng() % (unsigned)(hi - lo + 1)); };
    long long calls_made = 0, with_call = 0, mutations = 0;
    for (int rep = 0; rep < 4000; ++rep) {
        const int n_mapped = rep % 17 == 0 ? 0 : uni(0, 40), n_unm = rep % 13 == 0 ? 0 : uni(0, 12);
        const int R = n_mapped + 2 * n_unm;
        std::vector<double> logp(R);
        std::vector<advhmm_read_summary> S(R);
        std::vector<int32_t> plen(R);
        std::vector<int64_t> off(R + 1, 0);
        for (int i = 0; i < R; ++i) {
            const int len = uni(0, 300);
            off[i + 1] = off[i] + len;
            logp[i] = rng() % 9 == 0 ? -INFINITY : -(double)len * (0.1 + (rng() % 100) / 80.0);
            plen[i] = rng() % 11 == 0 ? -1 - (int)(rng() % 2) : len + 5;
            S[i].repeats = rng() % 23 == 0 ? -1 : uni(0, rep % 7 == 0 ? 2000000 : 9);
            S[i].n_match = uni(0, len);
            S[i].repeat_bp = uni(0, 40);
            S[i].left_bp = rng() % 4 == 0 ? 0 : uni(0, 80);
            S[i].right_bp = rng() % 4 == 0 ? 0 : uni(0, 80);
            S[i].left_hits = uni(0, S[i].left_bp);
            S[i].right_hits = uni(0, S[i].right_bp);
        }
        std::vector<uint8_t> cls(R + 1, 7);
        const calls::ReadView view{logp.data(), S.data(), plen.data(), off.data()};
        for (int flags = 0; flags < 4; ++flags) {
            advhmm_locus_call out{};
            const double score = rep % 3 == 0 ? NAN : -(double)uni(10, 200);
            calls::call_locus(view, 0, n_mapped, n_unm, score, flags & 1, flags & 2, uni(0, 5), out, cls.data());
            ++calls_made;
            with_call += out.has_call;
            if (out.recruited < out.spanning + (flags & 1 ? 0 : out.flanking) && !(flags & 1)) { std::puts("count mismatch"); return 1; }
        }
        // frameshift walk on a random path
        const int n_states = uni(1, 60);
        std::vector<uint8_t> sc(n_states);
        std::vector<int32_t> label(n_states);
        for (int s = 0; s < n_states; ++s) {
            const int kind = uni(0, 5), part = kind >= 1 && kind <= 3 ? uni(1, 3) : 0;
            sc[s] = (uint8_t)(kind | (part << 3) | (uni(0, 3) << 5));
            label[s] = uni(-1, 5000);
        }
        const int plen_ = uni(0, 400);
        std::vector<int32_t> path(plen_);
        int emitting = 0;
        for (int k = 0; k < plen_; ++k) {
            path[k] = k && rng() % 5 == 0 ? path[k - 1] : uni(0, n_states - 1);     // self loops too
            if (k >= 1 && k < plen_ - 1) { const int kind = sc[path[k]] & 7; emitting += kind == 1 || kind == 2; }
        }
        std::vector<uint8_t> seq(emitting + 1);
        for (auto& b : seq) b = (uint8_t)uni(0, 3);
        std::vector<calls::Mutation> mut;
        std::vector<int32_t> lengths;
        std::vector<std::pair<int32_t, int32_t>> first_visit;
        for (int pattern = 1; pattern < 40; pattern += 7)
            calls::frameshift_mutations_of_read(calls::PathView{path.data(), plen_, sc.data(), label.data(), seq.data()}, pattern, mut,
                                                lengths, first_visit);
        mutations += (long long)mut.size();
        if (const calls::Mutation* m = calls::frameshift_candidate(mut)) (void)m->count;
        // count lists
        std::vector<int32_t> obs(uni(0, 60));
        for (auto& o : obs) o = uni(0, rep % 5 == 0 ? 3 : 40);
        for (int hap = 0; hap < 2; ++hap) {
            const calls::Genotype g = calls::genotype_from_observed(obs.data(), obs.size(), hap);
            if (g.found && !(g.max_prob > 0)) { std::puts("bad posterior"); return 1; }
            const std::vector<int32_t> kept = calls::drop_unsupported(obs);
            (void)calls::genotype_from_observed(kept.data(), kept.size(), hap);
        }
    }
    std::printf("locus_calls: %lld calls (%lld with a genotype), %lld mutation keys, no sanitizer report\n", calls_made, with_call,
                mutations);
    return 0;
}
