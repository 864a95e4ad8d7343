// Robustness harness for libadvbam: every entry point over (possibly damaged) BAM files given as arguments.
//   g++ -O1 -g -fsanitize=address,undefined -std=c++17 -o /tmp/bam_asan tools/bam_asan_harness.cpp advntr_b200/csrc/bam_ingest.cpp -lz -lpthread
// Round 1: 301 files with random damage inside valid BGZF blocks and in the index: no report.
#include "../include/advbam.h"
#include <cstdio>
#include <vector>
#include <string>
int main(int argc, char** argv) {
    int ok = 0, refused = 0;
    for (int a = 1; a < argc; ++a) {
        advbam_file* f;
        if (advbam_open(argv[a], nullptr, &f)) { ++refused; continue; }
        bool bad = false;
        for (int threads : {1, 4}) {
            advbam_reads* r;
            if (advbam_scan(f, 0, 0, threads, &r)) { bad = true; continue; }
            advbam_view v; advbam_reads_view(r, &v);
            advbam_reads_to_fastq_orientation(r);
            advbam_reads_free(r);
        }
        long starts[5] = {40000, 1000020, 16383990, 134217700, 5000};
        for (int t = 0; t < advbam_n_references(f) && t < 3; ++t)
            for (long s : starts) {
                advbam_reads* r;
                if (advbam_fetch(f, t, s - 300, s + 400, &r)) { bad = true; continue; }
                advbam_view v; advbam_reads_view(r, &v);
                std::vector<uint8_t> d(v.n + 1); int64_t bp;
                advbam_illumina_params p{s, s + 80, 150, 135, 0, 20, 0.1};
                advbam_select_illumina(r, &p, d.data(), &bp);
                std::vector<int64_t> a(v.n + 1), b(v.n + 1); std::vector<int32_t> l(v.n + 1), rr(v.n + 1);
                advbam_spanning_segments(r, s, s + 80, 100, 10, a.data(), b.data(), l.data(), rr.data());
                int64_t ns, nc; advbam_gather_codes(r, d.data(), nullptr, nullptr, nullptr, &ns, &nc);
                std::vector<uint8_t> codes(nc + 1); std::vector<int64_t> off(ns + 1), idx(ns + 1);
                advbam_gather_codes(r, d.data(), codes.data(), off.data(), idx.data(), &ns, &nc);
                advbam_reads_free(r);
            }
        advbam_reads* h; if (!advbam_head(f, 5, &h)) advbam_reads_free(h); else bad = true;
        advbam_close(f);
        bad ? ++refused : ++ok;
    }
    printf("files ok %d, with refusals %d\n", ok, refused);
}
