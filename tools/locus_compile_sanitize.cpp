// Sanitizer harness for the host-side locus compiler (advntr_b200/csrc/locus_compile.hpp is pure C++17):
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off \
//       -o /tmp/lc_asan tools/locus_compile_sanitize.cpp -lpthread && /tmp/lc_asan
//   g++ -std=c++17 -O1 -g -fsanitize=thread -ffp-contract=off -o /tmp/lc_tsan tools/locus_compile_sanitize.cpp -lpthread && /tmp/lc_tsan
// Shapes (degenerate ones included) are built and analysed from several threads through the shared cache,
// random gapped alignments go through the profile code, and the device tables of every locus are filled
// into exactly-sized heap buffers (an out-of-bounds scatter would trip ASan).
#include <cstdio>
#include <random>

#include "../advntr_b200/csrc/locus_compile.hpp"

using namespace advhmm;

int main()
{
    std::mt19937 rng(12345);
    auto dna = [&](int n) { std::string s(n, 'A'); for (char& c : s) c = "ACGT"[rng() % 4]; return s; };
    struct Case { std::string left, right, aln; int n_seq, width, copies; double eps; };
    std::vector<Case> cases;
    const int shapes[][4] = {{1, 1, 1, 1}, {1, 2, 1, 2}, {2, 1, 2, 1}, {3, 3, 1, 3}, {5, 4, 2, 2}, {10, 7, 3, 12},
                             {20, 20, 11, 3}, {7, 30, 6, 11}, {150, 150, 30, 6}, {150, 150, 6, 26}, {100, 100, 60, 20}};
    for (auto& sh : shapes)
        for (int rep = 0; rep < 6; ++rep) {
            Case c;
            c.n_seq = 1 + rng() % 6;
            c.width = sh[2] + (rep % 3);                         // up to two extra (possibly insert) columns
            for (int r = 0; r < c.n_seq; ++r) {
                std::string row = dna(c.width);
                if (rep % 3)
                    for (int j = 0; j < c.width; ++j)
                        if (rng() % 5 == 0) row[j] = '-';
                if (row.find_first_not_of('-') == std::string::npos) row[0] = 'C';
                c.aln += row;
            }
            c.left = dna(sh[0]); c.right = dna(sh[1]); c.copies = sh[3]; c.eps = rep % 2 ? 0.3 : 0.05;
            cases.push_back(c);
        }
    std::atomic<int> done{0}, skipped{0};
    rm::parallel_for(cases.size(), 8, [&](size_t i, int) {
        const Case& c = cases[i];
        std::vector<uint8_t> l(c.left.size()), r(c.right.size());
        for (size_t k = 0; k < l.size(); ++k) l[k] = (uint8_t)rm::acgt_code(c.left[k]);
        for (size_t k = 0; k < r.size(); ++k) r[k] = (uint8_t)rm::acgt_code(c.right[k]);
        rm::LocusInput in;
        in.left = l.data(); in.left_len = (int)l.size(); in.right = r.data(); in.right_len = (int)r.size();
        in.aln = c.aln.data(); in.n_seq = c.n_seq; in.width = c.width; in.copies = c.copies; in.error_rate = c.eps;
        rm::LocusPrep prep;
        rm::prepare_locus(in, prep);
        if (!prep.ok) { ++skipped; return; }                    // e.g. no match column left
        std::string err;
        auto shape = rm::get_shape(prep.key, err);
        if (!shape) { fprintf(stderr, "shape failed: %s\n", err.c_str()); abort(); }
        if (!shape->banded) { fprintf(stderr, "shape not banded: %s\n", shape->why.c_str()); abort(); }
        rm::ChainBatch chain;
        std::vector<int32_t> slot_item(shape->slots.size());
        const double to_end = 0.7 / (shape->key.C * shape->key.R);
        for (size_t s = 0; s < shape->slots.size(); ++s) {
            const rm::Lab& lab = shape->slots[s];
            chain.items.push_back({rm::slot_probability(lab, shape->key, c.eps, prep.prof), lab.trips, lab.div, 1 + to_end, 0.0});
            slot_item[s] = (int32_t)s;
        }
        chain.run(nullptr, nullptr);
        rm::LocusValues lv;
        lv.shape = shape;
        lv.slot_log.resize(slot_item.size());
        for (size_t s = 0; s < slot_item.size(); ++s) lv.slot_log[s] = chain.items[s].out;
        lv.emis_tab = prep.emis_tab;
        lv.flank = prep.flank;
        const rm::LeanLayout ly = rm::lean_layout(*shape, 256);
        std::vector<unsigned char> image((size_t)shape->image_bytes);
        std::vector<int32_t> tb1((size_t)4 * shape->S), tb0((size_t)shape->m);
        std::vector<double> fin_w(shape->cm.b.fin_w.size());
        std::vector<uint8_t> cls((size_t)shape->m);
        rm::LeanScratch sc;
        const double e = rm::fill_lean(lv, sc, image.data(), tb1.data(), tb0.data(), fin_w.data(), cls.data());
        if (!(e < 0) || ly.bytes < (size_t)shape->image_bytes) abort();
        std::vector<double> w, em;
        rm::baked_values(lv, w, em);
        ++done;
    });
    rm::clear_shape_cache();
    printf("locus compiler sanitizer harness: %d loci compiled, %d degenerate alignments refused\n", done.load(), skipped.load());
    return 0;
}
