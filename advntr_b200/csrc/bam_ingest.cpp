// bam_ingest.cpp -- libadvbam.so: BAM region fetch / whole-file scan -> columns -> Viterbi batches.
//
// Host side of SURVEY.md section 8f rank 4 ("batched read ingest").  The reference goes through pysam
// record by record (vntr_finder.py:709-753, :453-462) and through `samtools view -f4 | samtools bam2fq`
// for the unmapped reads (sam_utils.py:9-23); here the same records are produced in bulk from the
// memory-mapped file (hts-specs SAMv1: section 4 BAM records, 4.1 BGZF blocks, 5.2 BAI bins) and the
// reference's per-read tests run over the columns.  See include/advbam.h for the interface each entry
// point replaces.  No CUDA here: inflate is the bound (zlib, one block per task, all cores for scans).
#include "../../include/advbam.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

thread_local std::string g_error;

struct Failure : std::runtime_error {
    int code;
    Failure(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint64_t le64(const uint8_t* p) { return (uint64_t)le32(p) | ((uint64_t)le32(p + 4) << 32); }

constexpr int kMaxBlock = 65536;                       // BGZF: at most 64 KiB before and after compression
const char kNibble[] = "=ACMGRSVTWYHKDBN";             // SAMv1 4.2.3
// complement in the same nibble code (A<->T, C<->G, IUPAC sets complemented, = and N fixed)
const uint8_t kNibbleComplement[16] = {0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15};

struct Chunk {
    uint64_t beg, end;
};
struct RefIndex {
    std::unordered_map<uint32_t, std::vector<Chunk>> bins;
    std::vector<uint64_t> linear;
};

}  // namespace

struct advbam_file {
    int fd = -1;
    const uint8_t* data = nullptr;
    size_t size = 0;
    std::string path;
    std::vector<std::string> ref_names;
    std::vector<int64_t> ref_lens;
    uint64_t first_record = 0;                          // virtual offset of the first alignment
    bool has_index = false;
    std::string index_error;
    std::vector<RefIndex> index;
};

struct advbam_reads {
    std::vector<uint16_t> flag;
    std::vector<uint8_t> mapq, has_qual;
    std::vector<int32_t> tid, pos, ref_end;
    std::vector<int64_t> seq_off{0}, name_off{0}, cigar_off{0};
    std::vector<char> seq, names;
    std::vector<uint8_t> qual;
    std::vector<uint32_t> cigar;
    int64_t n() const { return (int64_t)flag.size(); }
};

namespace {

// ---- BGZF --------------------------------------------------------------------------------------
// total size of the block starting at p (BSIZE + 1), from the BC extra subfield
int64_t bgzf_block_size(const uint8_t* p, size_t avail) {
    if (avail < 18 || p[0] != 31 || p[1] != 139 || p[2] != 8 || !(p[3] & 4)) return -1;
    size_t xlen = le16(p + 10), q = 12, end = 12 + xlen;
    if (end + 8 > avail) return -1;
    while (q + 4 <= end) {
        unsigned slen = le16(p + q + 2);
        if (p[q] == 66 && p[q + 1] == 67 && slen == 2) {
            int64_t bs = (int64_t)le16(p + q + 4) + 1;
            return bs <= (int64_t)avail && bs >= (int64_t)end + 8 ? bs : -1;
        }
        q += 4 + slen;
    }
    return -1;
}

// inflate one block into out[cap]; returns the uncompressed size, -1 on a damaged block
int bgzf_inflate(const uint8_t* p, int64_t bsize, uint8_t* out, size_t cap = kMaxBlock) {
    size_t xlen = le16(p + 10);
    const uint8_t* cdata = p + 12 + xlen;
    int64_t clen = bsize - (int64_t)xlen - 20;
    uint32_t crc = le32(p + bsize - 8), isize = le32(p + bsize - 4);
    if (clen < 0 || isize > (uint32_t)kMaxBlock || isize > cap) return -1;
    if (isize == 0) return 0;                           // the end-of-file marker block
    z_stream zs;
    std::memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) return -1;
    zs.next_in = const_cast<Bytef*>(cdata);
    zs.avail_in = (uInt)clen;
    zs.next_out = out;
    zs.avail_out = (uInt)std::min<size_t>(cap, kMaxBlock);
    int rc = inflate(&zs, Z_FINISH);
    uLong produced = zs.total_out;
    inflateEnd(&zs);
    if (rc != Z_STREAM_END || produced != isize) return -1;
    if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), out, (uInt)isize) != crc) return -1;
    return (int)isize;
}

// sequential reader addressed by BGZF virtual offsets (compressed offset << 16 | offset in block)
struct Cursor {
    const advbam_file* f;
    int64_t coff = 0, next = 0;
    int uoff = 0, blen = 0;
    std::vector<uint8_t> buf;
    explicit Cursor(const advbam_file* file) : f(file), buf(kMaxBlock) {}

    bool load(int64_t c) {
        coff = next = c;
        uoff = blen = 0;
        if (c >= (int64_t)f->size) return false;
        int64_t bs = bgzf_block_size(f->data + c, f->size - (size_t)c);
        if (bs < 0) throw Failure(ADVBAM_E_FORMAT, "not a BGZF block at byte " + std::to_string(c) + " of " + f->path);
        int got = bgzf_inflate(f->data + c, bs, buf.data());
        if (got < 0) throw Failure(ADVBAM_E_FORMAT, "damaged BGZF block at byte " + std::to_string(c) + " of " + f->path);
        blen = got;
        next = c + bs;
        return true;
    }
    void seek(uint64_t v) {
        load((int64_t)(v >> 16));
        uoff = (int)(v & 0xffff);
        if (uoff > blen) throw Failure(ADVBAM_E_FORMAT, "virtual offset past the end of its block in " + f->path);
    }
    // a position at the very end of a block is reported as the start of the next one, which is how
    // writers record it in the index
    uint64_t tell() const { return uoff == blen ? (uint64_t)next << 16 : ((uint64_t)coff << 16) | (uint64_t)uoff; }
    bool read(uint8_t* dst, size_t n) {
        while (n) {
            if (uoff == blen) {
                if (!load(next)) return false;
                continue;
            }
            size_t k = std::min(n, (size_t)(blen - uoff));
            std::memcpy(dst, buf.data() + uoff, k);
            dst += k;
            uoff += (int)k;
            n -= k;
        }
        return true;
    }
};

// ---- records -----------------------------------------------------------------------------------
inline bool consumes_reference(unsigned op) { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }
inline bool is_aligned_op(unsigned op) { return op == 0 || op == 7 || op == 8; }

// reference length of a CIGAR (bam_cigar2rlen)
int64_t cigar_reference_length(const uint32_t* c, size_t n) {
    int64_t len = 0;
    for (size_t i = 0; i < n; ++i)
        if (consumes_reference(c[i] & 15)) len += c[i] >> 4;
    return len;
}

// the CG:B,I tag that carries CIGARs of more than 65535 operations (SAMv1 4.2.2); null when absent
const uint8_t* find_long_cigar(const uint8_t* p, const uint8_t* end, uint32_t* n_ops) {
    while (p + 3 <= end) {
        char t0 = (char)p[0], t1 = (char)p[1], ty = (char)p[2];
        p += 3;
        size_t skip;
        switch (ty) {
            case 'A': case 'c': case 'C': skip = 1; break;
            case 's': case 'S': skip = 2; break;
            case 'i': case 'I': case 'f': skip = 4; break;
            case 'Z': case 'H': {
                const uint8_t* z = (const uint8_t*)std::memchr(p, 0, (size_t)(end - p));
                if (!z) return nullptr;
                skip = (size_t)(z - p) + 1;
                break;
            }
            case 'B': {
                if (p + 5 > end) return nullptr;
                char sub = (char)p[0];
                uint32_t cnt = le32(p + 1);
                size_t w = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                if (t0 == 'C' && t1 == 'G' && sub == 'I' && p + 5 + (size_t)cnt * 4 <= end) {
                    *n_ops = cnt;
                    return p + 5;
                }
                skip = 5 + (size_t)cnt * w;
                break;
            }
            default: return nullptr;
        }
        if (skip > (size_t)(end - p)) return nullptr;
        p += skip;
    }
    return nullptr;
}

struct RecordHead {
    int32_t tid, pos, l_seq;
    uint32_t l_name, n_cigar;
    uint16_t flag;
    uint8_t mapq;
};
inline RecordHead parse_head(const uint8_t* b) {
    RecordHead h;
    h.tid = (int32_t)le32(b);
    h.pos = (int32_t)le32(b + 4);
    h.l_name = b[8];
    h.mapq = b[9];
    h.n_cigar = le16(b + 12);
    h.flag = le16(b + 14);
    h.l_seq = (int32_t)le32(b + 16);
    return h;
}
inline bool record_is_sane(const RecordHead& h, size_t size) {
    return h.l_seq >= 0 && 32 + (size_t)h.l_name + 4 * (size_t)h.n_cigar + ((size_t)h.l_seq + 1) / 2 + (size_t)h.l_seq <= size;
}
// bam_endpos: pos + reference length of the CIGAR, at least one base; unmapped records span one base
inline int64_t record_endpos(const RecordHead& h, const uint8_t* b) {
    int64_t rlen = 0;
    if (!(h.flag & 4)) {
        const uint8_t* c = b + 32 + h.l_name;
        for (uint32_t i = 0; i < h.n_cigar; ++i) {
            uint32_t w = le32(c + 4 * i);
            if (consumes_reference(w & 15)) rlen += w >> 4;
        }
    }
    return (int64_t)h.pos + (rlen ? rlen : 1);
}

void append_record(advbam_reads& r, const RecordHead& h, const uint8_t* b, size_t size) {
    const uint8_t* name = b + 32;
    const uint8_t* cig = name + h.l_name;
    const uint8_t* sq = cig + 4 * (size_t)h.n_cigar;
    const uint8_t* ql = sq + ((size_t)h.l_seq + 1) / 2;
    const uint8_t* tags = ql + h.l_seq;
    r.flag.push_back(h.flag);
    r.mapq.push_back(h.mapq);
    r.tid.push_back(h.tid);
    r.pos.push_back(h.pos);
    size_t nlen = h.l_name ? strnlen((const char*)name, h.l_name) : 0;
    r.names.insert(r.names.end(), name, name + nlen);
    r.name_off.push_back((int64_t)r.names.size());
    // CIGAR; a placeholder "<l_seq>S<n>N" stands for a long one kept in the CG tag
    size_t c0 = r.cigar.size();
    uint32_t n_long = 0;
    const uint8_t* lc = nullptr;
    if (h.n_cigar == 2 && (le32(cig) & 15) == 4 && (int64_t)(le32(cig) >> 4) == h.l_seq && (le32(cig + 4) & 15) == 3)
        lc = find_long_cigar(tags, b + size, &n_long);
    if (lc) {
        for (uint32_t i = 0; i < n_long; ++i) r.cigar.push_back(le32(lc + 4 * i));
    } else {
        for (uint32_t i = 0; i < h.n_cigar; ++i) r.cigar.push_back(le32(cig + 4 * i));
    }
    r.cigar_off.push_back((int64_t)r.cigar.size());
    size_t nc = r.cigar.size() - c0;
    // read.reference_end: None when unmapped or without CIGAR, else bam_endpos
    if ((h.flag & 4) || nc == 0) {
        r.ref_end.push_back(-1);
    } else {
        int64_t rl = cigar_reference_length(r.cigar.data() + c0, nc);
        r.ref_end.push_back((int32_t)(h.pos + (rl ? rl : 1)));
    }
    size_t s0 = r.seq.size();
    r.seq.resize(s0 + (size_t)h.l_seq);
    char* dst = r.seq.data() + s0;
    int32_t pairs = h.l_seq >> 1;
    for (int32_t i = 0; i < pairs; ++i) {               // two bases per packed byte
        dst[2 * i] = kNibble[sq[i] >> 4];
        dst[2 * i + 1] = kNibble[sq[i] & 15];
    }
    if (h.l_seq & 1) dst[h.l_seq - 1] = kNibble[sq[pairs] >> 4];
    r.seq_off.push_back((int64_t)r.seq.size());
    bool hq = h.l_seq > 0 && ql[0] != 0xff;
    r.has_qual.push_back(hq ? 1 : 0);
    r.qual.insert(r.qual.end(), ql, ql + h.l_seq);
}

// one record from the cursor into `rec` (without its 4-byte length); false at end of file
bool next_record(Cursor& cur, std::vector<uint8_t>& rec) {
    uint8_t lenb[4];
    if (!cur.read(lenb, 4)) return false;
    uint32_t bs = le32(lenb);
    if (bs < 32 || bs > (1u << 30)) throw Failure(ADVBAM_E_FORMAT, "implausible BAM record length in " + cur.f->path);
    rec.resize(bs);
    if (!cur.read(rec.data(), bs)) throw Failure(ADVBAM_E_FORMAT, "truncated BAM record in " + cur.f->path);
    return true;
}

// ---- header and index --------------------------------------------------------------------------
void read_header(advbam_file* f) {
    Cursor cur(f);
    cur.load(0);
    uint8_t b[8];
    if (!cur.read(b, 8) || std::memcmp(b, "BAM\1", 4) != 0) throw Failure(ADVBAM_E_FORMAT, f->path + " is not a BAM file");
    uint32_t l_text = le32(b + 4);
    std::vector<uint8_t> text(l_text);
    if (l_text && !cur.read(text.data(), l_text)) throw Failure(ADVBAM_E_FORMAT, "truncated BAM header in " + f->path);
    if (!cur.read(b, 4)) throw Failure(ADVBAM_E_FORMAT, "truncated BAM header in " + f->path);
    uint32_t n_ref = le32(b);
    for (uint32_t i = 0; i < n_ref; ++i) {
        if (!cur.read(b, 4)) throw Failure(ADVBAM_E_FORMAT, "truncated reference list in " + f->path);
        uint32_t l_name = le32(b);
        if (l_name == 0 || l_name > (1u << 20)) throw Failure(ADVBAM_E_FORMAT, "bad reference name in " + f->path);
        std::vector<uint8_t> nm(l_name);
        if (!cur.read(nm.data(), l_name) || !cur.read(b, 4)) throw Failure(ADVBAM_E_FORMAT, "truncated reference list in " + f->path);
        f->ref_names.emplace_back((const char*)nm.data(), strnlen((const char*)nm.data(), l_name));
        f->ref_lens.push_back((int64_t)le32(b));
    }
    f->first_record = cur.tell();
}

bool slurp(const std::string& path, std::vector<uint8_t>& out) {
    FILE* fp = std::fopen(path.c_str(), "rb");
    if (!fp) return false;
    std::fseek(fp, 0, SEEK_END);
    long n = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    bool ok = out.empty() || std::fread(out.data(), 1, out.size(), fp) == out.size();
    std::fclose(fp);
    return ok;
}

void read_index(advbam_file* f, const char* bai_path) {
    std::vector<std::string> candidates;
    if (bai_path && *bai_path) {
        candidates.push_back(bai_path);
    } else {
        candidates.push_back(f->path + ".bai");
        size_t dot = f->path.rfind('.');
        if (dot != std::string::npos) candidates.push_back(f->path.substr(0, dot) + ".bai");
    }
    std::vector<uint8_t> d;
    std::string used;
    for (auto& c : candidates)
        if (slurp(c, d)) {
            used = c;
            break;
        }
    if (used.empty()) {
        f->index_error = "no index found for " + f->path + " (looked for " + candidates[0] + ")";
        return;
    }
    auto bad = [&]() { return Failure(ADVBAM_E_INDEX, used + " is not a valid BAI index"); };
    size_t q = 0;
    auto need = [&](size_t k) {
        if (q + k > d.size()) throw bad();
    };
    need(8);
    if (std::memcmp(d.data(), "BAI\1", 4) != 0) throw bad();
    uint32_t n_ref = le32(d.data() + 4);
    q = 8;
    if (n_ref != f->ref_names.size()) throw Failure(ADVBAM_E_INDEX, used + " indexes a different set of references than " + f->path);
    f->index.resize(n_ref);
    for (uint32_t r = 0; r < n_ref; ++r) {
        need(4);
        uint32_t n_bin = le32(d.data() + q);
        q += 4;
        for (uint32_t b = 0; b < n_bin; ++b) {
            need(8);
            uint32_t bin = le32(d.data() + q), n_chunk = le32(d.data() + q + 4);
            q += 8;
            need((size_t)n_chunk * 16);
            if (bin != 37450) {                         // 37450: metadata pseudo-bin
                auto& v = f->index[r].bins[bin];
                for (uint32_t c = 0; c < n_chunk; ++c) v.push_back({le64(d.data() + q + 16 * c), le64(d.data() + q + 16 * c + 8)});
            }
            q += (size_t)n_chunk * 16;
        }
        need(4);
        uint32_t n_intv = le32(d.data() + q);
        q += 4;
        need((size_t)n_intv * 8);
        f->index[r].linear.resize(n_intv);
        for (uint32_t i = 0; i < n_intv; ++i) f->index[r].linear[i] = le64(d.data() + q + 8 * i);
        q += (size_t)n_intv * 8;
    }
    f->has_index = true;
}

// bins that may hold records overlapping [beg, end) (SAMv1 5.3)
void region_bins(int64_t beg, int64_t end, std::vector<uint32_t>& out) {
    --end;
    out.push_back(0);
    for (int64_t k = 1 + (beg >> 26); k <= 1 + (end >> 26); ++k) out.push_back((uint32_t)k);
    for (int64_t k = 9 + (beg >> 23); k <= 9 + (end >> 23); ++k) out.push_back((uint32_t)k);
    for (int64_t k = 73 + (beg >> 20); k <= 73 + (end >> 20); ++k) out.push_back((uint32_t)k);
    for (int64_t k = 585 + (beg >> 17); k <= 585 + (end >> 17); ++k) out.push_back((uint32_t)k);
    for (int64_t k = 4681 + (beg >> 14); k <= 4681 + (end >> 14); ++k) out.push_back((uint32_t)k);
}

// ---- fetch / head / scan -----------------------------------------------------------------------
void fetch_region(advbam_file* f, int32_t tid, int64_t beg, int64_t end, advbam_reads& out) {
    if (!f->has_index) throw Failure(ADVBAM_E_INDEX, f->index_error.empty() ? "no index loaded" : f->index_error);
    if (tid < 0 || tid >= (int32_t)f->ref_names.size()) throw Failure(ADVBAM_E_ARG, "reference id out of range");
    beg = std::max<int64_t>(beg, 0);
    end = std::min<int64_t>(end, (int64_t)1 << 29);
    if (beg >= end) return;
    const RefIndex& ri = f->index[tid];
    uint64_t min_off = 0;
    if (!ri.linear.empty()) {
        size_t w = (size_t)(beg >> 14);
        min_off = ri.linear[std::min(w, ri.linear.size() - 1)];
        if (w >= ri.linear.size()) min_off = ri.linear.back();
    }
    std::vector<uint32_t> bins;
    region_bins(beg, end, bins);
    std::vector<Chunk> chunks;
    for (uint32_t b : bins) {
        auto it = ri.bins.find(b);
        if (it == ri.bins.end()) continue;
        for (const Chunk& c : it->second)
            if (c.end > min_off) chunks.push_back(c);
    }
    if (chunks.empty()) return;
    std::sort(chunks.begin(), chunks.end(), [](const Chunk& a, const Chunk& b) { return a.beg < b.beg; });
    std::vector<Chunk> merged;
    for (const Chunk& c : chunks) {
        if (!merged.empty() && c.beg <= merged.back().end)
            merged.back().end = std::max(merged.back().end, c.end);
        else
            merged.push_back(c);
    }
    Cursor cur(f);
    std::vector<uint8_t> rec;
    for (const Chunk& c : merged) {
        cur.seek(c.beg);
        while (cur.tell() < c.end) {
            if (!next_record(cur, rec)) return;
            RecordHead h = parse_head(rec.data());
            if (!record_is_sane(h, rec.size())) throw Failure(ADVBAM_E_FORMAT, "damaged BAM record in " + f->path);
            if (h.tid != tid) {
                if (h.tid > tid || h.tid < 0) return;   // sorted file: past the reference
                continue;
            }
            if ((int64_t)h.pos >= end) return;          // sorted by position: nothing further overlaps
            if (record_endpos(h, rec.data()) > beg) append_record(out, h, rec.data(), rec.size());
        }
    }
}

void head_records(advbam_file* f, int32_t n, advbam_reads& out) {
    Cursor cur(f);
    cur.seek(f->first_record);
    std::vector<uint8_t> rec;
    for (int32_t i = 0; i < n && next_record(cur, rec); ++i) {
        RecordHead h = parse_head(rec.data());
        if (!record_is_sane(h, rec.size())) throw Failure(ADVBAM_E_FORMAT, "damaged BAM record in " + f->path);
        append_record(out, h, rec.data(), rec.size());
    }
}

void append_batch(advbam_reads& d, const advbam_reads& s) {
    auto cat = [](auto& a, const auto& b) { a.insert(a.end(), b.begin(), b.end()); };
    auto cat_off = [](std::vector<int64_t>& a, const std::vector<int64_t>& b) {
        int64_t base = a.back();
        for (size_t i = 1; i < b.size(); ++i) a.push_back(base + b[i]);
    };
    cat(d.flag, s.flag);
    cat(d.mapq, s.mapq);
    cat(d.has_qual, s.has_qual);
    cat(d.tid, s.tid);
    cat(d.pos, s.pos);
    cat(d.ref_end, s.ref_end);
    cat(d.seq, s.seq);
    cat(d.qual, s.qual);
    cat(d.names, s.names);
    cat(d.cigar, s.cigar);
    cat_off(d.seq_off, s.seq_off);
    cat_off(d.name_off, s.name_off);
    cat_off(d.cigar_off, s.cigar_off);
}

int resolve_threads(int n_threads) {
    if (n_threads <= 0) {
        const char* e = std::getenv("ADVBAM_THREADS");
        n_threads = e ? std::atoi(e) : (int)std::thread::hardware_concurrency();
    }
    return std::max(1, std::min(n_threads, 64));
}

// run body(i) for i in [0, n) on up to n_threads threads; the first exception is rethrown
template <class Body>
void parallel_for(size_t n, int n_threads, Body&& body) {
    std::atomic<size_t> next{0};
    std::exception_ptr first;                           // the first worker failure, rethrown unchanged
    std::atomic<bool> failed{false};
    auto work = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= n || failed.load()) return;
            try {
                body(i);
            } catch (...) {
                if (!failed.exchange(true)) first = std::current_exception();
            }
        }
    };
    int nt = (int)std::min<size_t>((size_t)n_threads, n);
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    if (failed.load()) std::rethrow_exception(first);   // a Failure keeps its code, bad_alloc stays bad_alloc
}

void scan_windowed(advbam_file* f, uint64_t start, uint32_t require, uint32_t exclude, int n_threads, advbam_reads& out);

// Whole file with an index: the linear index holds virtual offsets of record STARTS all over the
// file, so the file is cut there into independent ranges; every thread inflates and parses its own
// ranges (no sequential stage), the batches are joined in file order.  The part behind the last cut
// (the unplaced reads at the end of the file) has no known record starts: scan_windowed.
bool scan_file_by_index(advbam_file* f, uint32_t require, uint32_t exclude, int n_threads, advbam_reads& out) {
    std::vector<uint64_t> cuts;
    for (const RefIndex& ri : f->index)
        for (uint64_t v : ri.linear)
            if (v > f->first_record) cuts.push_back(v);
    std::sort(cuts.begin(), cuts.end());
    cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
    if (cuts.size() < 2) return false;
    // thin the cut points to about 8 ranges per thread, evenly spaced in the compressed file
    std::vector<uint64_t> bounds{f->first_record};
    size_t want = (size_t)n_threads * 8;
    uint64_t span = (uint64_t)f->size / want + 1, next_at = (f->first_record >> 16) + span;
    for (uint64_t v : cuts)
        if ((v >> 16) >= next_at) {
            bounds.push_back(v);
            next_at = (v >> 16) + span;
        }
    size_t n_ranges = bounds.size() - 1;                // the last bound starts the tail
    std::vector<advbam_reads> parts(n_ranges);
    parallel_for(n_ranges, n_threads, [&](size_t i) {
        std::vector<uint8_t> rec;
        Cursor cur(f);
        cur.seek(bounds[i]);
        while (cur.tell() < bounds[i + 1] && next_record(cur, rec)) {
            RecordHead h = parse_head(rec.data());
            if (!record_is_sane(h, rec.size())) throw Failure(ADVBAM_E_FORMAT, "damaged BAM record in " + f->path);
            if ((h.flag & require) == require && (h.flag & exclude) == 0) append_record(parts[i], h, rec.data(), rec.size());
        }
    });
    for (const advbam_reads& p : parts) append_batch(out, p);
    scan_windowed(f, bounds.back(), require, exclude, n_threads, out);
    return true;
}

// From virtual offset `start` to the end of the file, without knowing record starts: windows of
// blocks are inflated in place by all threads into one stream (each block's slice is known from its
// ISIZE trailer), a walk over the 4-byte record lengths finds the record starts, and the records are
// parsed in parallel in runs of kRun.
void scan_windowed(advbam_file* f, uint64_t start, uint32_t require, uint32_t exclude, int n_threads, advbam_reads& out) {
    const size_t kWindow = 2048;                        // blocks per window: <= 128 MiB inflated
    const size_t kRun = 8192;                           // records per parsing task
    int64_t coff = (int64_t)(start >> 16);
    size_t skip = (size_t)(start & 0xffff);
    std::vector<uint8_t> stream;                        // unparsed tail of the last window + this window
    std::vector<int64_t> boff, bsz;
    std::vector<size_t> dst, rec_at;
    while (coff < (int64_t)f->size) {
        boff.clear();
        bsz.clear();
        dst.clear();
        size_t total = stream.size();
        while (boff.size() < kWindow && coff < (int64_t)f->size) {
            int64_t bs = bgzf_block_size(f->data + coff, f->size - (size_t)coff);
            if (bs < 0) throw Failure(ADVBAM_E_FORMAT, "not a BGZF block at byte " + std::to_string(coff) + " of " + f->path);
            uint32_t isize = le32(f->data + coff + bs - 4);
            if (isize > (uint32_t)kMaxBlock) throw Failure(ADVBAM_E_FORMAT, "damaged BGZF block at byte " + std::to_string(coff) + " of " + f->path);
            boff.push_back(coff);
            bsz.push_back(bs);
            dst.push_back(total);
            total += isize;
            coff += bs;
        }
        dst.push_back(total);
        stream.resize(total);
        parallel_for(boff.size(), n_threads, [&](size_t i) {
            size_t cap = dst[i + 1] - dst[i];
            if (bgzf_inflate(f->data + boff[i], bsz[i], stream.data() + dst[i], cap) != (int)cap)
                throw Failure(ADVBAM_E_FORMAT, "damaged BGZF block at byte " + std::to_string(boff[i]) + " of " + f->path);
        });
        size_t q = std::min(skip, stream.size());       // bytes before `start` in its block
        skip -= q;
        rec_at.clear();
        while (stream.size() - q >= 4) {
            uint32_t bs = le32(stream.data() + q);
            if (bs < 32 || bs > (1u << 30)) throw Failure(ADVBAM_E_FORMAT, "implausible BAM record length in " + f->path);
            if (stream.size() - q - 4 < bs) break;
            rec_at.push_back(q);
            q += 4 + (size_t)bs;
        }
        size_t n_runs = (rec_at.size() + kRun - 1) / kRun;
        std::vector<advbam_reads> parts(n_runs);
        parallel_for(n_runs, n_threads, [&](size_t run) {
            size_t hi = std::min(rec_at.size(), (run + 1) * kRun);
            for (size_t k = run * kRun; k < hi; ++k) {
                const uint8_t* b = stream.data() + rec_at[k] + 4;
                size_t bs = le32(b - 4);
                RecordHead h = parse_head(b);
                if (!record_is_sane(h, bs)) throw Failure(ADVBAM_E_FORMAT, "damaged BAM record in " + f->path);
                if ((h.flag & require) == require && (h.flag & exclude) == 0) append_record(parts[run], h, b, bs);
            }
        });
        for (const advbam_reads& p : parts) append_batch(out, p);
        stream.erase(stream.begin(), stream.begin() + (std::ptrdiff_t)q);
    }
    if (!stream.empty()) throw Failure(ADVBAM_E_FORMAT, "truncated BAM record at the end of " + f->path);
}

void scan_file(advbam_file* f, uint32_t require, uint32_t exclude, int n_threads, advbam_reads& out) {
    n_threads = resolve_threads(n_threads);
    if (n_threads > 1 && f->has_index && scan_file_by_index(f, require, exclude, n_threads, out)) return;
    scan_windowed(f, f->first_record, require, exclude, n_threads, out);
}

template <class F>
int guarded(F&& body) {
    try {
        body();
        return ADVBAM_OK;
    } catch (const Failure& e) {
        g_error = e.what();
        return e.code;
    } catch (const std::bad_alloc&) {
        g_error = "out of memory";
        return ADVBAM_E_IO;
    } catch (const std::exception& e) {
        g_error = e.what();
        return ADVBAM_E_FORMAT;
    }
}

}  // namespace

extern "C" {

const char* advbam_last_error(void) { return g_error.c_str(); }

int advbam_open(const char* bam_path, const char* bai_path, advbam_file** out) {
    if (!bam_path || !out) {
        g_error = "advbam_open: null argument";
        return ADVBAM_E_ARG;
    }
    *out = nullptr;
    advbam_file* f = new advbam_file;
    int rc = guarded([&]() {
        f->path = bam_path;
        f->fd = ::open(bam_path, O_RDONLY);
        if (f->fd < 0) throw Failure(ADVBAM_E_IO, std::string("cannot open ") + bam_path);
        struct stat st;
        if (fstat(f->fd, &st) != 0 || st.st_size <= 0) throw Failure(ADVBAM_E_IO, std::string("cannot stat ") + bam_path);
        f->size = (size_t)st.st_size;
        void* m = mmap(nullptr, f->size, PROT_READ, MAP_PRIVATE, f->fd, 0);
        if (m == MAP_FAILED) throw Failure(ADVBAM_E_IO, std::string("cannot map ") + bam_path);
        f->data = (const uint8_t*)m;
        read_header(f);
        read_index(f, bai_path);
    });
    if (rc != ADVBAM_OK) {
        advbam_close(f);
        return rc;
    }
    *out = f;
    return ADVBAM_OK;
}

void advbam_close(advbam_file* f) {
    if (!f) return;
    if (f->data) munmap(const_cast<uint8_t*>(f->data), f->size);
    if (f->fd >= 0) ::close(f->fd);
    delete f;
}

int32_t advbam_n_references(const advbam_file* f) { return f ? (int32_t)f->ref_names.size() : 0; }
const char* advbam_reference_name(const advbam_file* f, int32_t tid) {
    return f && tid >= 0 && tid < (int32_t)f->ref_names.size() ? f->ref_names[tid].c_str() : nullptr;
}
int64_t advbam_reference_length(const advbam_file* f, int32_t tid) {
    return f && tid >= 0 && tid < (int32_t)f->ref_lens.size() ? f->ref_lens[tid] : -1;
}
int32_t advbam_reference_id(const advbam_file* f, const char* name) {
    if (!f || !name) return -1;
    for (size_t i = 0; i < f->ref_names.size(); ++i)
        if (f->ref_names[i] == name) return (int32_t)i;
    return -1;
}

int advbam_head(advbam_file* f, int32_t n, advbam_reads** out) {
    if (!f || !out) {
        g_error = "advbam_head: null argument";
        return ADVBAM_E_ARG;
    }
    advbam_reads* r = new advbam_reads;
    int rc = guarded([&]() { head_records(f, n, *r); });
    if (rc != ADVBAM_OK) {
        delete r;
        r = nullptr;
    }
    *out = r;
    return rc;
}

int advbam_fetch(advbam_file* f, int32_t tid, int64_t beg, int64_t end, advbam_reads** out) {
    if (!f || !out) {
        g_error = "advbam_fetch: null argument";
        return ADVBAM_E_ARG;
    }
    advbam_reads* r = new advbam_reads;
    int rc = guarded([&]() { fetch_region(f, tid, beg, end, *r); });
    if (rc != ADVBAM_OK) {
        delete r;
        r = nullptr;
    }
    *out = r;
    return rc;
}

int advbam_scan(advbam_file* f, uint32_t require_flags, uint32_t exclude_flags, int32_t n_threads, advbam_reads** out) {
    if (!f || !out) {
        g_error = "advbam_scan: null argument";
        return ADVBAM_E_ARG;
    }
    advbam_reads* r = new advbam_reads;
    int rc = guarded([&]() { scan_file(f, require_flags, exclude_flags, n_threads, *r); });
    if (rc != ADVBAM_OK) {
        delete r;
        r = nullptr;
    }
    *out = r;
    return rc;
}

void advbam_reads_free(advbam_reads* r) { delete r; }

int advbam_reads_view(const advbam_reads* r, advbam_view* v) {
    if (!r || !v) {
        g_error = "advbam_reads_view: null argument";
        return ADVBAM_E_ARG;
    }
    v->n = r->n();
    v->flag = r->flag.data();
    v->mapq = r->mapq.data();
    v->tid = r->tid.data();
    v->pos = r->pos.data();
    v->ref_end = r->ref_end.data();
    v->has_qual = r->has_qual.data();
    v->seq_off = r->seq_off.data();
    v->seq = r->seq.data();
    v->qual = r->qual.data();
    v->name_off = r->name_off.data();
    v->names = r->names.data();
    v->cigar_off = r->cigar_off.data();
    v->cigar = r->cigar.data();
    return ADVBAM_OK;
}

int advbam_reads_to_fastq_orientation(advbam_reads* r) {
    if (!r) {
        g_error = "advbam_reads_to_fastq_orientation: null argument";
        return ADVBAM_E_ARG;
    }
    static uint8_t comp[256];
    static bool ready = false;
    if (!ready) {
        for (int c = 0; c < 256; ++c) comp[c] = (uint8_t)c;
        for (int k = 0; k < 16; ++k) comp[(uint8_t)kNibble[k]] = (uint8_t)kNibble[kNibbleComplement[k]];
        ready = true;
    }
    int64_t n = r->n();
    std::vector<char> names;
    names.reserve(r->names.size() + 2 * (size_t)n);
    std::vector<int64_t> off{0};
    off.reserve((size_t)n + 1);
    for (int64_t i = 0; i < n; ++i) {
        uint16_t fl = r->flag[i];
        if (fl & 0x10) {
            int64_t a = r->seq_off[i], b = r->seq_off[i + 1];
            std::reverse(r->seq.begin() + a, r->seq.begin() + b);
            for (int64_t k = a; k < b; ++k) r->seq[k] = (char)comp[(uint8_t)r->seq[k]];
            if (r->has_qual[i]) std::reverse(r->qual.begin() + a, r->qual.begin() + b);
        }
        names.insert(names.end(), r->names.begin() + r->name_off[i], r->names.begin() + r->name_off[i + 1]);
        bool r1 = fl & 0x40, r2 = fl & 0x80;
        if (r1 != r2) {                                 // both or neither: no suffix (samtools fastq)
            names.push_back('/');
            names.push_back(r1 ? '1' : '2');
        }
        off.push_back((int64_t)names.size());
    }
    r->names.swap(names);
    r->name_off.swap(off);
    return ADVBAM_OK;
}

int advbam_select_illumina(const advbam_reads* r, const advbam_illumina_params* p, uint8_t* decision, int64_t* vntr_bp) {
    if (!r || !p || !decision) {
        g_error = "advbam_select_illumina: null argument";
        return ADVBAM_E_ARG;
    }
    int64_t bp = 0, n = r->n();
    for (int64_t i = 0; i < n; ++i) {
        uint16_t fl = r->flag[i];
        if ((fl & 0x4) || (fl & 0x400)) {               // vntr_finder.py:728
            decision[i] = ADVBAM_SKIP_FLAGS;
            continue;
        }
        int64_t a = r->seq_off[i], len = r->seq_off[i + 1] - a;
        if (len < p->min_read_length) {                 // :731
            decision[i] = ADVBAM_SKIP_SHORT;
            continue;
        }
        int64_t start = r->pos[i];
        int64_t rend = r->ref_end[i] > 0 ? r->ref_end[i] : start + len;   // :734 (None and 0 are both falsy)
        bool inside = (p->vntr_start - p->read_length < start && start < p->vntr_end) ||
                      (p->vntr_start < rend && rend < p->vntr_end);       // :735
        if (!inside) {
            decision[i] = ADVBAM_SKIP_REGION;
            continue;
        }
        bp += std::min(rend, p->vntr_end) - std::max(start, p->vntr_start);   // :751-753
        const char* s = r->seq.data() + a;
        bool has_n = false, other = false;
        for (int64_t k = 0; k < len; ++k) {
            char c = s[k];
            has_n |= c == 'N';
            other |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N');
        }
        if (has_n) {                                    // :736
            decision[i] = ADVBAM_SKIP_N;
            continue;
        }
        if (other) {
            decision[i] = ADVBAM_BAD_SYMBOL;
            continue;
        }
        // is_low_quality_read, utils.py:20-38
        if ((int)r->mapq[i] <= p->mapq_cutoff) {
            decision[i] = ADVBAM_SKIP_LOW_QUALITY;
            continue;
        }
        if (!r->has_qual[i]) {
            decision[i] = ADVBAM_NO_QUALITIES;
            continue;
        }
        const uint8_t* q = r->qual.data() + a;
        int64_t n_low = 0;
        for (int64_t k = 0; k < len; ++k) n_low += q[k] < p->quality_cutoff;
        bool low = (double)n_low >= p->low_quality_fraction * (double)len;       // :25
        if (!low && n_low) {
            // :28-37: a low-quality base followed by max_run - 1 further low-quality bases inside the read
            int64_t max_run = (int64_t)(p->low_quality_fraction * (double)len / 4);
            int64_t need = std::max<int64_t>(max_run, 1), run = 0;
            for (int64_t k = 0; k < len && !low; ++k) {
                run = q[k] < p->quality_cutoff ? run + 1 : 0;
                low = run >= need;
            }
        }
        decision[i] = low ? ADVBAM_SKIP_LOW_QUALITY : ADVBAM_DECODE;
    }
    if (vntr_bp) *vntr_bp = bp;
    return ADVBAM_OK;
}

int advbam_gather_codes(const advbam_reads* r, const uint8_t* decision, uint8_t* codes, int64_t* off, int64_t* index,
                        int64_t* n_selected, int64_t* n_codes) {
    if (!r) {
        g_error = "advbam_gather_codes: null argument";
        return ADVBAM_E_ARG;
    }
    int64_t n = r->n(), ns = 0, nc = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (decision && decision[i] != ADVBAM_DECODE) continue;
        int64_t a = r->seq_off[i], b = r->seq_off[i + 1];
        if (codes) {
            for (int64_t k = a; k < b; ++k) {
                char c = r->seq[k];
                uint8_t v = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 255;
                if (v == 255) {
                    g_error = "advbam_gather_codes: record " + std::to_string(i) + " holds a base outside ACGT";
                    return ADVBAM_E_ARG;
                }
                codes[nc + (k - a)] = v;
            }
        }
        if (off) off[ns] = nc;
        if (index) index[ns] = i;
        nc += b - a;
        ++ns;
    }
    if (off) off[ns] = nc;
    if (n_selected) *n_selected = ns;
    if (n_codes) *n_codes = nc;
    return ADVBAM_OK;
}

int advbam_spanning_segments(const advbam_reads* r, int64_t vntr_start, int64_t vntr_end, int32_t hmm_flank,
                             int32_t min_flank_bp, int64_t* seg_start, int64_t* seg_end, int32_t* left_bp, int32_t* right_bp) {
    if (!r || !seg_start || !seg_end || !left_bp || !right_bp) {
        g_error = "advbam_spanning_segments: null argument";
        return ADVBAM_E_ARG;
    }
    const int64_t region_start = vntr_start - hmm_flank, region_end = vntr_end + hmm_flank;
    int64_t n = r->n();
    for (int64_t i = 0; i < n; ++i) {
        seg_start[i] = seg_end[i] = -1;
        left_bp[i] = right_bp[i] = 0;
        if (r->flag[i] & 4) continue;                   // pysam: no aligned pairs for unmapped records
        const uint32_t* c = r->cigar.data() + r->cigar_off[i];
        size_t nc = (size_t)(r->cigar_off[i + 1] - r->cigar_off[i]);
        // first and last aligned reference position (get_reference_positions()[0], [-1])
        int64_t ref = r->pos[i], first = -1, last = -1;
        for (size_t k = 0; k < nc; ++k) {
            unsigned op = c[k] & 15;
            int64_t len = c[k] >> 4;
            if (is_aligned_op(op) && len) {
                if (first < 0) first = ref;
                last = ref + len - 1;
            }
            if (consumes_reference(op)) ref += len;
        }
        if (first < 0) continue;                        // vntr_finder.py:458
        if (!(first <= vntr_start - min_flank_bp && vntr_end + min_flank_bp < last)) continue;   // :382
        int64_t rrs = -1, rre = -1, read_pos = 0, left = 0, right = 0;
        ref = r->pos[i];
        bool done = false;
        for (size_t k = 0; k < nc && !done; ++k) {
            unsigned op = c[k] & 15;
            int64_t len = c[k] >> 4;
            if (is_aligned_op(op)) {
                // only the part of the run inside [region_start, region_end] matters (:391-392 stops beyond it)
                for (int64_t j = 0; j < len; ++j) {
                    int64_t rp = ref + j;
                    if (rp > region_end) {
                        done = true;
                        break;
                    }
                    if (rp >= region_start && rp < region_end) {
                        if (rp < vntr_start) {
                            if (rrs < 0) rrs = read_pos + j;
                            ++left;
                        } else if (rp >= vntr_end) {
                            if (rre < 0) rre = read_pos + j;
                            ++right;
                        }
                    }
                }
                ref += len;
                read_pos += len;
            } else if (op == 1 || op == 4) {            // insertion, soft clip: read only
                read_pos += len;
            } else if (op == 2 || op == 3) {            // deletion, skip: reference only
                ref += len;
            }                                           // hard clip, padding: neither
        }
        if (left < min_flank_bp || right < min_flank_bp) continue;           // :408
        int64_t l_seq = r->seq_off[i + 1] - r->seq_off[i];
        if (rrs < 0 || rre < 0 || l_seq == 0) continue;                       // :413 (read.seq None when empty)
        seg_start[i] = std::min(rrs, l_seq);
        seg_end[i] = std::max(seg_start[i], std::min(rre + right, l_seq));    // python slice clamps
        left_bp[i] = (int32_t)left;
        right_bp[i] = (int32_t)right;
    }
    return ADVBAM_OK;
}

}  // extern "C"
