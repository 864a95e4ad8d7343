// advhmm.cu -- B200 (sm_100a) Viterbi / forward engine for adVNTR's profile HMMs + its C-ABI.
//
// Replaces the vendored pomegranate's `_viterbi` / `_forward`
// (/root/reference/pomegranate/hmm.pyx:1970-2136, 1371-1484) for batches of reads.
// See DESIGN.md for the data layout and the roofline of each kernel.
//
// Files: engine_types.cuh (handles, device descriptors), kernels_common.cuh (TMA / mbarrier /
// shuffle helpers, pack kernel), kernels_banded.cuh, kernels_generic.cuh, model_compile.hpp
// (host-side graph analysis of one baked model), locus_compile.hpp (native compiler of the
// read-matcher models of whole batches of loci: shape structures, parameter chains, device tables),
// locus_calls.hpp (host: per-read results of many loci -> recruitment, spanning test, genotype calls),
// this file (model upload, batch planning, launches, C-ABI).
//
// Kernels (all hand-written, no tensor cores -- this is max-plus DP, not a contraction):
//   pack_reads_kernel        byte codes -> 2-bit packed reads (+ reverse complement, validation)
//   banded_fill_kernel<RPL>  profile-shaped models: one warp per read, lanes own blocks of RPL
//                            read positions, columns sweep as a register wavefront (skew 1
//                            column per lane, 3 shuffles per step); model tables staged into
//                            shared memory with one TMA bulk copy per CTA; 6-bit traceback per
//                            (position, column) packed into one word per lane and step.  The column
//                            loop is phased (ramp-up / steady / the 32 steps around the collector's
//                            column / steady / ramp-down) with warp-uniform bounds, so the steady
//                            phase carries neither range tests nor the collector select.
//   banded_fill_f32_kernel   the same schedule in float (optional ADVHMM_FP32 mode)
//   banded_long_kernel<T,FWD> long reads / models larger than shared memory: 160-position
//                            stripes dealt to 1..8 warps per read, carry through HBM with
//                            release / acquire progress flags, the model image streamed through a
//                            per-warp ring in shared memory by TMA bulk copies; instantiated for
//                            double (Viterbi), float (ADVHMM_FP32) and double forward
//   banded_backtrack_kernel  device backtrack to the state path + on-device path reducers
//   generic_fill_kernel<FWD> any baked model: row-synchronous CSR kernel, silent states by level;
//                            FWD = log_probability (sum-product with the reference's pair_lse)
//   generic_backtrack_kernel
//
// Exactness: every DP value is produced by the same IEEE-754 double operations in the same
// order as the reference ((v + t) + e per edge, strict '>' in candidate order), so paths and
// log-probabilities are bit-identical; see model_compile.hpp for why the banded schedule may
// skip / reorder the candidates it does.
#include "kernels_generic.cuh"
#include "kernels_kfilter.cuh"
#include "locus_calls.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <set>
#include <unordered_map>

namespace {

// =============================================================================================
// host side: model upload
// =============================================================================================
struct BlobBuilder {
    std::vector<unsigned char> bytes;
    size_t add(const void* src, size_t n)
    {
        size_t off = (bytes.size() + 255) / 256 * 256;
        bytes.resize(off + n);
        if (n && src) memcpy(bytes.data() + off, src, n);
        return off;
    }
    template <typename T> size_t add(const std::vector<T>& v) { return add(v.data(), v.size() * sizeof(T)); }
};

// forward-algorithm row 0 (hmm.pyx:1402-1424): same closure as Viterbi's with pair_lse
double host_pair_lse(double x, double y)       // utils.pyx:72-90
{
    if (x == kNegInf) return y;
    if (y == kNegInf) return x;
    if (x > y) return x + std::log(std::exp(y - x) + 1.0);
    return y + std::log(std::exp(x - y) + 1.0);
}

std::vector<double> forward_row0(const GenericTables& g)
{
    auto lse = host_pair_lse;
    std::vector<double> f(g.m, kNegInf);
    f[g.start] = 0.0;
    for (int l = g.S; l < g.m; ++l) {
        if (l == g.start) continue;
        double acc = kNegInf;
        for (int k = g.in_off[l]; k < g.in_off[l + 1]; ++k)
            if (g.in_src[k] >= g.S) acc = lse(acc, f[g.in_src[k]] + g.in_w[k]);
        f[l] = acc;
    }
    return f;
}

}  // namespace

namespace {

// the shared-memory image of a banded model (layout: kernels_banded.cuh)
void pack_banded_image(const BandedTables& b, std::vector<unsigned char>& image)
{
    const size_t P = b.NCpad;
    image.assign((size_t)kImgBytesPerCol * P, 0);
    double* w10 = reinterpret_cast<double*>(image.data() + (size_t)kImgW * P);
    double* e2 = reinterpret_cast<double*>(image.data() + (size_t)kImgE * P);
    double* v12 = reinterpret_cast<double*>(image.data() + (size_t)kImgV1 * P);
    for (size_t c = 0; c < P; ++c) {
        for (int k = 0; k < 9; ++k) w10[c * 10 + k] = b.w[(size_t)k * P + c];
        w10[c * 10 + 9] = b.accw[c];
        for (int x = 0; x < 4; ++x)
            for (int sl = 0; sl < 2; ++sl) {       // sl: 0 = I slot, 1 = M slot
                e2[((size_t)x * P + c) * 2 + sl] = b.e[((size_t)sl * 4 + x) * P + c];
                v12[((size_t)x * P + c) * 2 + sl] = b.v1[((size_t)sl * 4 + x) * P + c];
            }
    }
}

int upload_model(advhmm_model* mod)
{
    advhmm_context* ctx = mod->ctx;
    const GenericTables& g = mod->cm.g;
    const BandedTables& b = mod->cm.b;
    BlobBuilder bb;
    // generic tables
    const size_t o_in_off = bb.add(g.in_off), o_in_src = bb.add(g.in_src), o_in_w = bb.add(g.in_w);
    const size_t o_emis = bb.add(g.emis), o_v0 = bb.add(g.v0), o_tb0 = bb.add(g.tb0);
    const size_t o_lvl_off = bb.add(g.lvl_off), o_lvl_state = bb.add(g.lvl_state);
    const std::vector<double> f0 = forward_row0(g);
    const size_t o_f0 = bb.add(f0);
    // banded tables
    size_t o_image = 0, o_st = 0, o_tb1 = 0, o_acc = 0, o_fs = 0, o_fo = 0, o_fsrc = 0, o_fw = 0;
    size_t o_image_f = 0, o_tb1f = 0, o_tb0f = 0, o_fwf = 0, o_f1 = 0;
    int image_bytes = 0, image_f_bytes = 0;
    if (b.valid) {
        const size_t P = b.NCpad;
        std::vector<unsigned char> image;
        pack_banded_image(b, image);
        image_bytes = (int)image.size();
        o_image = bb.add(image);
        // forward first-row table: emitting states of row 1 from the forward row 0 (hmm.pyx:1427-1444)
        std::vector<double> f1(8 * P, kNegInf);
        for (int l = 0; l < g.S; ++l) {
            double acc = kNegInf;
            for (int k = g.in_off[l]; k < g.in_off[l + 1]; ++k) acc = host_pair_lse(acc, f0[g.in_src[k]] + g.in_w[k]);
            const int sl = b.slot_of[l] == SLOT_I ? 0 : 1;
            for (int x = 0; x < 4; ++x) f1[((size_t)x * P + b.col_of[l]) * 2 + sl] = acc + g.emis[(size_t)l * 4 + x];
        }
        o_f1 = bb.add(f1);
        std::vector<int32_t> st(3 * (size_t)b.NC);
        for (int t = 0; t < 3; ++t) memcpy(st.data() + (size_t)t * b.NC, b.st[t].data(), sizeof(int32_t) * b.NC);
        o_st = bb.add(st); o_tb1 = bb.add(b.tb1); o_acc = bb.add(b.acc_src_col);
        o_fs = bb.add(b.fin_state); o_fo = bb.add(b.fin_off); o_fsrc = bb.add(b.fin_src); o_fw = bb.add(b.fin_w);
        // fp32 twin
        const BandedF32& f = mod->cm.f;
        std::vector<unsigned char> imf((size_t)kImgFBytesPerCol * P, 0);
        float* w12 = reinterpret_cast<float*>(imf.data());
        float* e2f = reinterpret_cast<float*>(imf.data() + (size_t)kImgFE * P);
        float* v1f = reinterpret_cast<float*>(imf.data() + (size_t)kImgFV1 * P);
        for (size_t c = 0; c < P; ++c) {
            for (int k = 0; k < 9; ++k) w12[c * 12 + k] = f.w[(size_t)k * P + c];
            w12[c * 12 + 9] = f.accw[c];
            for (int x = 0; x < 4; ++x)
                for (int sl = 0; sl < 2; ++sl) {
                    e2f[((size_t)x * P + c) * 2 + sl] = f.e[((size_t)sl * 4 + x) * P + c];
                    v1f[((size_t)x * P + c) * 2 + sl] = f.v1[((size_t)sl * 4 + x) * P + c];
                }
        }
        image_f_bytes = (int)imf.size();
        o_image_f = bb.add(imf); o_tb1f = bb.add(f.tb1); o_tb0f = bb.add(f.tb0); o_fwf = bb.add(f.fin_w);
    }
    const std::vector<uint8_t> zero_classes((size_t)g.m, 0);
    const size_t o_cls = bb.add(zero_classes);
    const size_t o_dg = bb.add(nullptr, sizeof(DevGeneric));
    const size_t o_dgf = bb.add(nullptr, sizeof(DevGeneric));
    const size_t o_db = bb.add(nullptr, sizeof(DevBanded));

    mod->K = g.K; mod->m = g.m; mod->NCpad = b.valid ? b.NCpad : 0;
    mod->info.kind = ADVHMM_KIND_GENERIC;
    mod->info.n_states = g.m;
    mod->info.n_edges = g.in_off[g.m];
    mod->info.n_columns = g.n_levels;
    mod->info.n_final_states = 0;
    mod->info.smem_bytes = 0;
    mod->info.max_in_degree = g.max_in_degree;
    const size_t smem_limit = ctx->device >= 0 ? ctx->smem_optin : (size_t)232448;
    if (b.valid) {
        mod->info.kind = ADVHMM_KIND_BANDED;
        mod->info.n_columns = b.NC;
        mod->info.n_final_states = (int)b.fin_state.size();
        mod->info.smem_bytes = image_bytes;
        // 0: the image does not fit in shared memory -> every read of this model takes the
        // long-read kernel, which streams the tables from global memory
        mod->banded_smem = ((size_t)image_bytes + 4096 <= smem_limit) ? image_bytes : 0;
    }
    if (ctx->device < 0) return ADVHMM_OK;   // host-only context: analysis only

    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(mod->blob.ensure(bb.bytes.size()));
    unsigned char* base = mod->blob.as<unsigned char>();
    auto P8 = [&](size_t off) { return base + off; };
    DevGeneric dg{};
    dg.m = g.m; dg.S = g.S; dg.K = g.K; dg.start = g.start; dg.end = g.end; dg.finite = g.finite;
    dg.n_levels = g.n_levels; dg.max_in_degree = g.max_in_degree;
    dg.in_off = (const int32_t*)P8(o_in_off); dg.in_src = (const int32_t*)P8(o_in_src);
    dg.in_w = (const double*)P8(o_in_w); dg.emis = (const double*)P8(o_emis);
    dg.v0 = (const double*)P8(o_v0); dg.tb0 = (const int32_t*)P8(o_tb0);
    dg.lvl_off = (const int32_t*)P8(o_lvl_off); dg.lvl_state = (const int32_t*)P8(o_lvl_state);
    dg.classes = (const uint8_t*)P8(o_cls);
    mod->d_classes = (uint8_t*)P8(o_cls);
    memcpy(bb.bytes.data() + o_dg, &dg, sizeof dg);
    DevGeneric dgf = dg;
    dgf.v0 = (const double*)P8(o_f0);
    memcpy(bb.bytes.data() + o_dgf, &dgf, sizeof dgf);
    if (b.valid) {
        DevBanded db{};
        db.NC = b.NC; db.P = b.NCpad; db.S = b.S; db.m = g.m; db.NF = (int)b.fin_state.size();
        db.end_final = b.end_final; db.acc_col = b.acc_col; db.n_acc = (int)b.acc_src_col.size();
        db.start = g.start; db.end = g.end; db.image_bytes = image_bytes;
        db.logp_empty = g.v0[g.end];
        db.image = P8(o_image); db.st = (const int32_t*)P8(o_st); db.tb1 = (const int32_t*)P8(o_tb1);
        db.acc_src_col = (const int32_t*)P8(o_acc); db.fin_state = (const int32_t*)P8(o_fs);
        db.fin_off = (const int32_t*)P8(o_fo); db.fin_src = (const int32_t*)P8(o_fsrc);
        db.fin_w = (const double*)P8(o_fw); db.tb0 = (const int32_t*)P8(o_tb0);
        db.image_f = P8(o_image_f); db.image_f_bytes = image_f_bytes;
        db.logp_empty_f = mod->cm.f.v0[g.end];
        db.tb1_f = (const int32_t*)P8(o_tb1f); db.tb0_f = (const int32_t*)P8(o_tb0f);
        db.fin_w_f = (const float*)P8(o_fwf);
        db.classes = (const uint8_t*)P8(o_cls);
        db.f1 = (const double*)P8(o_f1);
        db.logp_empty_fwd = f0[g.end];
        memcpy(bb.bytes.data() + o_db, &db, sizeof db);
    }
    CU_TRY(cudaMemcpyAsync(base, bb.bytes.data(), bb.bytes.size(), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    mod->d_generic = reinterpret_cast<DevGeneric*>(P8(o_dg));
    mod->d_generic_fwd = reinterpret_cast<DevGeneric*>(P8(o_dgf));
    mod->d_banded = b.valid ? reinterpret_cast<DevBanded*>(P8(o_db)) : nullptr;
    return ADVHMM_OK;
}

// =============================================================================================
// locus models: many read-matcher models per call, compiled natively (locus_compile.hpp)
// =============================================================================================
static_assert(rm::kImageBytesPerCol == kImgBytesPerCol && rm::kImageE == kImgE && rm::kImageV1 == kImgV1,
              "locus_compile.hpp and kernels_banded.cuh must agree on the image layout");

rm::VexpFn g_vexp = nullptr;
void* g_vexp_user = nullptr;

// A locus model that is asked for something outside the banded Viterbi path (generic kernel, forward,
// fp32) gets the full set of tables of a descriptor-made model: same arrays, same analysis.
int ensure_full_model(advhmm_model* mod)
{
    if (!mod->lean) return ADVHMM_OK;
    const rm::ShapeStructure& sh = *mod->locus.shape;
    std::vector<double> in_logp, emis;
    rm::baked_values(mod->locus, in_logp, emis);
    advhmm_model_desc d{};
    d.n_states = sh.m; d.silent_start = sh.S; d.start_index = sh.start; d.end_index = sh.end;
    d.finite = sh.finite; d.n_symbols = 4;
    d.in_off = sh.in_off.data(); d.in_src = sh.in_src.data(); d.in_logp = in_logp.data(); d.emis = emis.data();
    std::string err;
    if (!compile_model(d, mod->cm, err)) return set_error(ADVHMM_EINVAL, "%s", err.c_str());
    if (int rc = upload_model(mod)) return rc;
    mod->lean = false;
    if (mod->ctx->device >= 0) {
        std::vector<uint8_t> cls((size_t)sh.m);
        for (int i = 0; i < sh.m; ++i) {
            uint8_t c = sh.base_class[i];
            if (i < sh.S && sh.emis_row[i] < 0) c |= (uint8_t)(mod->locus.flank[(size_t)(-sh.emis_row[i] - 1)] << 5);
            cls[i] = c;
        }
        CU_TRY(cudaMemcpyAsync(mod->d_classes, cls.data(), cls.size(), cudaMemcpyHostToDevice, mod->ctx->stream));
        CU_TRY(cudaStreamSynchronize(mod->ctx->stream));
    }
    return ADVHMM_OK;
}

int ensure_shape_on_device(advhmm_context* ctx, const std::shared_ptr<const rm::ShapeStructure>& shp, DevShape** out)
{
    const rm::ShapeStructure* sh = shp.get();
    auto it = ctx->shape_dev.find(sh);
    if (it != ctx->shape_dev.end()) { *out = it->second.get(); return ADVHMM_OK; }
    const BandedTables& b = sh->cm.b;
    BlobBuilder bb;
    std::vector<int32_t> st(3 * (size_t)b.NC);
    for (int t = 0; t < 3; ++t) memcpy(st.data() + (size_t)t * b.NC, b.st[t].data(), sizeof(int32_t) * b.NC);
    const size_t o_st = bb.add(st), o_acc = bb.add(b.acc_src_col), o_fs = bb.add(b.fin_state);
    const size_t o_fo = bb.add(b.fin_off), o_fsrc = bb.add(b.fin_src);
    std::unique_ptr<DevShape> ds(new DevShape);
    ds->shape = shp;            // (a structure dropped from the cache must not give its address to another shape)
    CU_TRY(ds->blob.ensure(bb.bytes.size() + 256));
    unsigned char* base = ds->blob.as<unsigned char>();
    CU_TRY(cudaMemcpyAsync(base, bb.bytes.data(), bb.bytes.size(), cudaMemcpyHostToDevice, ctx->upload_stream));
    CU_TRY(cudaStreamSynchronize(ctx->upload_stream));
    ds->st = (const int32_t*)(base + o_st); ds->acc_src_col = (const int32_t*)(base + o_acc);
    ds->fin_state = (const int32_t*)(base + o_fs); ds->fin_off = (const int32_t*)(base + o_fo);
    ds->fin_src = (const int32_t*)(base + o_fsrc);
    *out = ds.get();
    ctx->shape_dev[sh] = std::move(ds);
    return ADVHMM_OK;
}

void lean_model_info(advhmm_model* mod, size_t smem_limit)
{
    const rm::ShapeStructure& sh = *mod->locus.shape;
    const BandedTables& b = sh.cm.b;
    mod->K = 4; mod->m = sh.m; mod->NCpad = b.NCpad;
    mod->info.kind = ADVHMM_KIND_BANDED;
    mod->info.n_states = sh.m;
    mod->info.n_edges = sh.cm.g.in_off[sh.m];
    mod->info.n_columns = b.NC;
    mod->info.n_final_states = (int)b.fin_state.size();
    mod->info.smem_bytes = sh.image_bytes;
    mod->info.max_in_degree = sh.cm.g.max_in_degree;
    mod->banded_smem = ((size_t)sh.image_bytes + 4096 <= smem_limit) ? sh.image_bytes : 0;
}

int create_models_for_loci(advhmm_context* ctx, const advhmm_loci* L, int n_threads, advhmm_model** out)
{
    const int N = L->n_loci;
    const int nt = n_threads > 0 ? n_threads : rm::default_threads();
    for (int i = 0; i < N; ++i) out[i] = nullptr;
    static const bool dbg = getenv("ADVHMM_DEBUG_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_mark = now();
    auto lap = [&](const char* what) {
        if (!dbg) return;
        const double t = now();
        fprintf(stderr, "[advhmm] create_for_loci: %-28s %8.2f ms\n", what, t - t_mark);
        t_mark = t;
    };
    // ---- 1. per locus, in parallel: repeat-unit profile, emission logs, shape key ---------------
    std::vector<rm::LocusPrep> prep((size_t)N);
    rm::parallel_for((size_t)N, nt, [&](size_t i, int) {
        rm::LocusInput in;
        in.left = L->left + L->left_off[i];   in.left_len = (int)(L->left_off[i + 1] - L->left_off[i]);
        in.right = L->right + L->right_off[i]; in.right_len = (int)(L->right_off[i + 1] - L->right_off[i]);
        in.aln = L->segments + L->seg_off[i];
        in.n_seq = L->n_segments[i];
        const int64_t chars = L->seg_off[i + 1] - L->seg_off[i];
        in.width = in.n_seq > 0 ? (int)(chars / in.n_seq) : 0;
        in.copies = L->copies[i];
        in.error_rate = L->error_rate[i];
        if (in.n_seq < 1 || (int64_t)in.width * in.n_seq != chars) {
            prep[i].err = "the aligned repeat segments of a locus must have one width";
            return;
        }
        rm::prepare_locus(in, prep[i]);
    });
    for (int i = 0; i < N; ++i)
        if (!prep[i].ok) return set_error(ADVHMM_EINVAL, "locus %d: %s", i, prep[i].err.c_str());
    lap("profiles");
    // ---- 2. shapes: the distinct ones are built (or found in the cache) in parallel ----------------
    std::map<rm::ShapeKey, std::shared_ptr<const rm::ShapeStructure>> shapes;
    for (int i = 0; i < N; ++i) shapes.emplace(prep[i].key, nullptr);
    {
        std::vector<rm::ShapeKey> keys;
        for (auto& kv : shapes) keys.push_back(kv.first);
        std::vector<std::shared_ptr<const rm::ShapeStructure>> built(keys.size());
        std::vector<std::string> errs(keys.size());
        rm::parallel_for(keys.size(), nt, [&](size_t k, int) { built[k] = rm::get_shape(keys[k], errs[k]); });
        for (size_t k = 0; k < keys.size(); ++k) {
            if (!built[k]) return set_error(ADVHMM_EINVAL, "shape (%d, %d, %d, %d): %s", keys[k].Ll, keys[k].Lr, keys[k].R,
                                            keys[k].C, errs[k].c_str());
            shapes[keys[k]] = built[k];
        }
    }
    lap("shapes");
    // ---- 3. parameter chains: distinct (probability, chain) items per locus, two vector exps in all ----
    std::vector<std::vector<rm::ChainBatch::Item>> items((size_t)N);
    std::vector<std::vector<int32_t>> slot_item((size_t)N);
    std::atomic<int> bad{-1};
    rm::parallel_for((size_t)N, nt, [&](size_t i, int) {
        const rm::ShapeStructure& sh = *shapes.find(prep[i].key)->second;
        const double to_end = 0.7 / (sh.key.C * sh.key.R);
        const double total = 1 + to_end;
        auto& its = items[i];
        auto& si = slot_item[i];
        si.resize(sh.slots.size());
        for (size_t s = 0; s < sh.slots.size(); ++s) {
            const rm::Lab& l = sh.slots[s];
            const double p = rm::slot_probability(l, sh.key, prep[i].error_rate, prep[i].prof);
            if (!(p > 0)) { bad.store((int)i); return; }
            size_t j = 0;
            for (; j < its.size(); ++j)
                if (its[j].p == p && its[j].trips == l.trips && its[j].div == l.div) break;
            if (j == its.size()) its.push_back(rm::ChainBatch::Item{p, l.trips, l.div, total, 0.0});
            si[s] = (int32_t)j;
        }
    });
    if (bad.load() >= 0)
        return set_error(ADVHMM_EUNSUPPORTED, "locus %d has a zero-probability transition: its structure differs from its "
                                              "shape's, build it through advhmm_model_create", bad.load());
    lap("chain items");
    rm::ChainBatch chain;
    std::vector<size_t> first((size_t)N + 1, 0);
    for (int i = 0; i < N; ++i) first[i + 1] = first[i] + items[i].size();
    chain.items.reserve(first[N]);
    for (int i = 0; i < N; ++i) chain.items.insert(chain.items.end(), items[i].begin(), items[i].end());
    chain.run(g_vexp, g_vexp_user);
    lap("chains (log / exp)");
    // ---- 4. handles --------------------------------------------------------------------------------
    std::vector<std::unique_ptr<advhmm_model>> mods((size_t)N);
    const size_t smem_limit = ctx->device >= 0 ? ctx->smem_optin : (size_t)232448;
    rm::parallel_for((size_t)N, nt, [&](size_t i, int) {
        std::unique_ptr<advhmm_model> mod(new advhmm_model);
        mod->ctx = ctx;
        mod->lean = true;
        mod->locus.shape = shapes.find(prep[i].key)->second;
        mod->locus.slot_log.resize(slot_item[i].size());
        for (size_t s = 0; s < slot_item[i].size(); ++s) mod->locus.slot_log[s] = chain.items[first[i] + slot_item[i][s]].out;
        mod->locus.emis_tab = std::move(prep[i].emis_tab);
        mod->locus.flank = std::move(prep[i].flank);
        lean_model_info(mod.get(), smem_limit);
        mods[i] = std::move(mod);
    });
    lap("handles");
    // shapes the banded kernel cannot run (none of the read-matcher family so far): full models
    for (int i = 0; i < N; ++i)
        if (!mods[i]->locus.shape->banded)
            if (int rc = ensure_full_model(mods[i].get())) return rc;
    if (ctx->device >= 0) {
        // ---- 5. device: one arena for the batch, tables written by all threads into pinned staging
        //         chunks, one H2D copy per chunk (double-buffered) ----------------------------------
        CU_TRY(cudaSetDevice(ctx->device));
        if (!ctx->upload_stream) {
            CU_TRY(cudaStreamCreateWithFlags(&ctx->upload_stream, cudaStreamNonBlocking));
            CU_TRY(cudaEventCreateWithFlags(&ctx->upload_done, cudaEventDisableTiming));
            // keep freed arenas in the pool: the next batch takes the same memory without a trip to the driver
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
        }
        std::vector<rm::LeanLayout> lay((size_t)N);
        std::vector<size_t> off((size_t)N + 1, 0);
        std::vector<DevShape*> dshape((size_t)N, nullptr);
        for (int i = 0; i < N; ++i) {
            if (!mods[i]->lean) { off[i + 1] = off[i]; continue; }
            lay[i] = rm::lean_layout(*mods[i]->locus.shape, sizeof(DevBanded));
            off[i + 1] = off[i] + lay[i].bytes;
            if (int rc = ensure_shape_on_device(ctx, mods[i]->locus.shape, &dshape[i])) return rc;
        }
        auto arena = std::make_shared<DeviceArena>();
        arena->device = ctx->device;
        arena->bytes = off[N];
        arena->free_stream = ctx->stream;
        arena->ctx_alive = ctx->alive;
        if (arena->bytes) CU_TRY(cudaMallocAsync(&arena->p, arena->bytes, ctx->upload_stream));
        unsigned char* dbase = static_cast<unsigned char*>(arena->p);
        size_t stage_bytes = (size_t)64 << 20;
        for (int i = 0; i < N; ++i) stage_bytes = std::max(stage_bytes, lay[i].bytes);
        for (int k = 0; k < 2; ++k) {
            CU_TRY(ctx->h_stage[k].ensure(stage_bytes));
            if (!ctx->stage_done[k]) CU_TRY(cudaEventCreateWithFlags(&ctx->stage_done[k], cudaEventDisableTiming));
        }
        std::vector<rm::LeanScratch> scratch((size_t)nt);
        int chunk = 0;
        for (int lo = 0; lo < N;) {
            int hi = lo;
            while (hi < N && off[hi + 1] - off[lo] <= stage_bytes) ++hi;
            if (hi == lo) ++hi;
            const int k = chunk & 1;
            CU_TRY(cudaEventSynchronize(ctx->stage_done[k]));           // the copy that last read this buffer
            unsigned char* hbase = static_cast<unsigned char*>(ctx->h_stage[k].p);
            rm::parallel_for((size_t)(hi - lo), nt, [&](size_t j, int worker) {
                const int i = lo + (int)j;
                advhmm_model* mod = mods[i].get();
                if (!mod->lean) return;
                const rm::ShapeStructure& sh = *mod->locus.shape;
                const BandedTables& b = sh.cm.b;
                const rm::LeanLayout& ly = lay[i];
                unsigned char* h = hbase + (off[i] - off[lo]);
                unsigned char* dv = dbase + off[i];
                memset(h + ly.o_tb1, 0, ly.o_tb0 - ly.o_tb1);            // (alignment padding stays defined)
                const double logp_empty = rm::fill_lean(mod->locus, scratch[worker], h + ly.o_image,
                                                        reinterpret_cast<int32_t*>(h + ly.o_tb1),
                                                        reinterpret_cast<int32_t*>(h + ly.o_tb0),
                                                        reinterpret_cast<double*>(h + ly.o_finw), h + ly.o_cls);
                DevBanded db{};
                db.NC = b.NC; db.P = b.NCpad; db.S = b.S; db.m = sh.m; db.NF = (int)b.fin_state.size();
                db.end_final = b.end_final; db.acc_col = b.acc_col; db.n_acc = (int)b.acc_src_col.size();
                db.start = sh.start; db.end = sh.end; db.image_bytes = sh.image_bytes;
                db.logp_empty = logp_empty;
                db.image = dv + ly.o_image;
                db.st = dshape[i]->st; db.acc_src_col = dshape[i]->acc_src_col;
                db.fin_state = dshape[i]->fin_state; db.fin_off = dshape[i]->fin_off; db.fin_src = dshape[i]->fin_src;
                db.tb1 = reinterpret_cast<const int32_t*>(dv + ly.o_tb1);
                db.tb0 = reinterpret_cast<const int32_t*>(dv + ly.o_tb0);
                db.fin_w = reinterpret_cast<const double*>(dv + ly.o_finw);
                db.classes = dv + ly.o_cls;
                memcpy(h + ly.o_desc, &db, sizeof db);
                mod->d_banded = reinterpret_cast<DevBanded*>(dv + ly.o_desc);
                mod->d_classes = dv + ly.o_cls;
                mod->arena = arena;
            });
            if (off[hi] > off[lo])
                CU_TRY(cudaMemcpyAsync(dbase + off[lo], hbase, off[hi] - off[lo], cudaMemcpyHostToDevice, ctx->upload_stream));
            CU_TRY(cudaEventRecord(ctx->stage_done[k], ctx->upload_stream));
            lo = hi;
            ++chunk;
        }
        // whatever is queued on the compute stream from here on sees the tables; the compute stream itself
        // is NOT waited for: a batch that is being decoded keeps running while this one is compiled
        CU_TRY(cudaEventRecord(ctx->upload_done, ctx->upload_stream));
        CU_TRY(cudaStreamWaitEvent(ctx->stream, ctx->upload_done, 0));
        CU_TRY(cudaStreamSynchronize(ctx->upload_stream));
        lap("tables + upload");
    }
    for (int i = 0; i < N; ++i) out[i] = mods[i].release();
    return ADVHMM_OK;
}

// =============================================================================================
// host side: batch planning and launches
// =============================================================================================
// work items (result reads) of one kernel family, in launch order
struct Family {
    std::vector<int32_t> items;      // result-read ids
    std::vector<int32_t> model;      // model index of every item
    int first_item = 0;              // position of the family inside Plan::order
    int max_len = 0;
};

struct Plan {
    std::vector<int64_t> pk_off;        // [n_out]
    int64_t pk_words = 0;
    std::vector<int32_t> order;         // short-banded items, then long-banded, then generic
    std::vector<int32_t> item_tile;     // tile of every item
    std::vector<Tile> tiles;
    int max_P_short = 0, max_smem_short = 0, max_P_long = 0, max_m_generic = 0;
};

template <typename T> size_t vec_bytes(const std::vector<T>& v) { return v.size() * sizeof(T); }

// event pair bracketing one kernel launch when profiling is on (kind: 0 banded fill, 1 backtrack)
struct ProfScope {
    advhmm_context* ctx; int kind; cudaEvent_t stop = nullptr; cudaStream_t stream;
    ProfScope(advhmm_context* c, int k, cudaStream_t st = nullptr) : ctx(c), kind(k), stream(st ? st : c->stream)
    {
        if (!ctx->profile) return;
        auto& ev = ctx->prof_events[kind];
        if (ctx->prof_used[kind] == ev.size()) {
            cudaEvent_t a, b;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
            ev.emplace_back(a, b);
        }
        auto& pr = ev[ctx->prof_used[kind]++];
        cudaEventRecord(pr.first, stream);
        stop = pr.second;
    }
    ~ProfScope() { if (stop) cudaEventRecord(stop, stream); }
};

// Dynamic shared memory above 48 KB needs an opt-in that is a property of the KERNEL (per device),
// not of a context: raise it once per (kernel, device) to the device maximum.  (Tracking the value
// per context let a second context lower the limit under the first one's feet.)
template <typename Kernel>
int allow_max_dynamic_smem(advhmm_context* ctx, Kernel* kernel)
{
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    std::lock_guard<std::mutex> lock(mu);
    const std::pair<const void*, int> key(reinterpret_cast<const void*>(kernel), ctx->device);
    if (done.count(key)) return ADVHMM_OK;
    cudaFuncAttributes attr{};
    CU_TRY(cudaFuncGetAttributes(&attr, kernel));
    const int room = (int)ctx->smem_optin - (int)attr.sharedSizeBytes;      // static shared memory counts too
    CU_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, room));
    done.insert(key);
    return ADVHMM_OK;
}

template <int RPL, int WPB>
int launch_banded_variant(advhmm_context* ctx, int grid, int smem, const BandedArgs& args)
{
    if (int rc = allow_max_dynamic_smem(ctx, banded_fill_kernel<RPL, WPB>)) return rc;
    banded_fill_kernel<RPL, WPB><<<grid, WPB * 32, smem, ctx->stream>>>(args);
    return ADVHMM_OK;
}

// Two CTAs of 8 warps per SM (126 registers/thread) as long as two model images fit in shared memory;
// a larger image (250 bp reads: 250-base flanks, ~770 columns, 160 KB) leaves room for one CTA only, which
// then carries 16 warps so that the SM still has 4 warps per scheduler.  (10 / 12 warps per CTA need
// 96 / 80 registers, spill and measured 25-40 % slower; profiles/r1_variants.md.)
int banded_warps_for(const advhmm_context* ctx, int smem_image)
{
    static const bool no_wide = getenv("ADVHMM_NO_WIDE_CTA") != nullptr;     // measurement knob
    const size_t per_cta = (size_t)smem_image + 4096;          // + static shared memory + the driver's 1 KB
    return (2 * per_cta <= ctx->smem_optin + 1024 || no_wide) ? ctx->banded_warps : kBandedWarpsMax;
}

int launch_banded_chunk(advhmm_context* ctx, int rpl, int grid, int smem, const BandedArgs& args, int warps)
{
    ProfScope prof(ctx, 0);
    int rc = ADVHMM_OK;
#define ADV_CASE(R)                                                                              \
    case R: rc = warps == kBandedWarpsMax ? launch_banded_variant<R, kBandedWarpsMax>(ctx, grid, smem, args) \
                                          : launch_banded_variant<R, 8>(ctx, grid, smem, args); break;
    switch (rpl) {
        ADV_CASE(1) ADV_CASE(2) ADV_CASE(3) ADV_CASE(4) ADV_CASE(5)
        ADV_CASE(6) ADV_CASE(7) ADV_CASE(8) ADV_CASE(9) ADV_CASE(10)
        default: return set_error(ADVHMM_EINVAL, "unsupported rows-per-lane %d", rpl);
    }
#undef ADV_CASE
    if (rc) return rc;
    CU_TRY(cudaGetLastError());
    ctx->launches++;
    return ADVHMM_OK;
}

int launch_banded_fwd_chunk(advhmm_context* ctx, int rpl, int grid, int smem, const BandedArgs& args)
{
    ProfScope prof(ctx, 0);
#define ADV_CASE(R)                                                                                   \
    case R:                                                                                           \
        if (int rc = allow_max_dynamic_smem(ctx, banded_forward_kernel<R, 8>)) return rc;             \
        banded_forward_kernel<R, 8><<<grid, 8 * 32, smem, ctx->stream>>>(args);                       \
        break;
    switch (rpl) {
        ADV_CASE(1) ADV_CASE(2) ADV_CASE(3) ADV_CASE(4) ADV_CASE(5)
        ADV_CASE(6) ADV_CASE(7) ADV_CASE(8) ADV_CASE(9) ADV_CASE(10)
        default: return set_error(ADVHMM_EINVAL, "unsupported rows-per-lane %d", rpl);
    }
#undef ADV_CASE
    CU_TRY(cudaGetLastError());
    ctx->launches++;
    return ADVHMM_OK;
}

int launch_banded_f32_chunk(advhmm_context* ctx, int rpl, int grid, int smem, const BandedArgs& args)
{
    ProfScope prof(ctx, 0);
    // the fp32 image is smaller than the fp64 one, so the fp64 size is a safe dynamic-smem request
#define ADV_CASE(R)                                                                                   \
    case R:                                                                                           \
        if (int rc = allow_max_dynamic_smem(ctx, banded_fill_f32_kernel<R>)) return rc;               \
        banded_fill_f32_kernel<R><<<grid, 8 * 32, smem, ctx->stream>>>(args);                         \
        break;
    switch (rpl) {
        ADV_CASE(1) ADV_CASE(2) ADV_CASE(3) ADV_CASE(4) ADV_CASE(5)
        ADV_CASE(6) ADV_CASE(7) ADV_CASE(8) ADV_CASE(9) ADV_CASE(10)
        default: return set_error(ADVHMM_EINVAL, "unsupported rows-per-lane %d", rpl);
    }
#undef ADV_CASE
    CU_TRY(cudaGetLastError());
    ctx->launches++;
    return ADVHMM_OK;
}

constexpr size_t kMaxChunkMarks = 4096;

// after a backtrack launch: snapshot the path cursor and record an event, so that the host-buffer
// front end can start copying this chunk's paths while the next chunk runs
int mark_chunk(advhmm_context* ctx, const unsigned long long* d_cursor, cudaStream_t stream = nullptr)
{
    if (!stream) stream = ctx->stream;
    if (!ctx->mark_chunks || ctx->n_marks >= kMaxChunkMarks) return ADVHMM_OK;
    const size_t i = ctx->n_marks;
    if (i == ctx->chunk_events.size()) {
        cudaEvent_t ev;
        CU_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->chunk_events.push_back(ev);
    }
    CU_TRY(cudaMemcpyAsync(static_cast<unsigned long long*>(ctx->h_cursors.p) + i, d_cursor, sizeof(unsigned long long),
                           cudaMemcpyDeviceToHost, stream));
    CU_TRY(cudaEventRecord(ctx->chunk_events[i], stream));
    ctx->n_marks = i + 1;
    return ADVHMM_OK;
}

struct OutPtrs {
    double* logp; int32_t* path_len; int64_t* path_off; int32_t* path; int64_t path_cap;
    unsigned long long* cursor;
    advhmm_read_summary* summaries;   // NULL unless ADVHMM_WANT_SUMMARY
};

// Runs the whole batch on the context's stream.  All pointers in `out` and d_seqs are device
// pointers.  seq_off / group_off are HOST arrays (planning metadata).
// read_base / continue_cursor: the host-buffer front end feeds a large batch as several sub-batches so
// that the planning of one overlaps the decoding of the previous one; they share the path cursor.
int run_batch(advhmm_context* ctx, advhmm_model* const* models, int n_models, const int64_t* group_off,
              const uint8_t* d_seqs, const int64_t* seq_off, int n_reads, uint32_t flags,
              const OutPtrs& out, bool forward, int32_t* d_bad, int read_base = 0, bool continue_cursor = false)
{
    const bool want_path = (flags & ADVHMM_WANT_PATH) && !forward;
    const bool want_walk = want_path || (out.summaries && !forward);   // backtrack kernel needed
    const int strands = (flags & ADVHMM_BOTH_STRANDS) ? 2 : 1;
    const int n_out = n_reads * strands;
    if (n_out == 0) return ADVHMM_OK;
    const bool fp32 = (flags & ADVHMM_FP32) != 0;
    if (fp32 && forward) return set_error(ADVHMM_EUNSUPPORTED, "fp32 mode exists for Viterbi only");
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };

    // ---- plan: packed-read offsets, kernel family of every result read, tiles ----------------
    Plan pl;
    pl.pk_off.resize(n_out);
    int64_t words = 0;
    for (int r = 0; r < n_reads; ++r) {
        const int64_t len = seq_off[r + 1] - seq_off[r];
        if (len < 0 || len > 0x3fffffff) return set_error(ADVHMM_EINVAL, "bad seq_off at read %d", r);
        for (int s = 0; s < strands; ++s) {
            pl.pk_off[(size_t)r * strands + s] = words;
            words += (len + 15) / 16 + 1;
        }
    }
    pl.pk_words = words;

    Family fam_short, fam_long, fam_generic;
    fam_short.items.reserve(n_out);                 // the common case: every read takes the short-read kernel
    fam_short.model.reserve(n_out);
    // every read belongs to exactly one group: reads outside every group would never be decoded and
    // their result slots would be returned uninitialised
    if (group_off[0] != 0 || group_off[n_models] != n_reads)
        return set_error(ADVHMM_EINVAL, "group_off must start at 0 and end at n_reads (%lld .. %lld, n_reads %d)",
                         (long long)group_off[0], (long long)group_off[n_models], n_reads);
    for (int gi = 0; gi < n_models; ++gi) {
        advhmm_model* mod = models[gi];
        if (!mod || mod->ctx != ctx) return set_error(ADVHMM_EINVAL, "model %d does not belong to this context", gi);
        // reads are packed 2 bits per symbol and validated against ONE alphabet size per call
        if (mod->K != models[0]->K)
            return set_error(ADVHMM_EINVAL, "model %d has %d symbols, model 0 has %d: one alphabet per call", gi,
                             mod->K, models[0]->K);
        // a locus model carries the banded Viterbi tables only: the generic kernel, forward and the
        // fp32 twin need the full set, built on first use
        if (mod->lean && (forward || (flags & (ADVHMM_FORCE_GENERIC | ADVHMM_FP32))))
            if (int rc = ensure_full_model(mod)) return rc;
        const int64_t r0 = group_off[gi], r1 = group_off[gi + 1];
        if (r0 < 0 || r1 < r0 || r1 > n_reads) return set_error(ADVHMM_EINVAL, "bad group_off at model %d", gi);
        const bool banded = mod->d_banded && !(flags & ADVHMM_FORCE_GENERIC);
        for (int64_t r = r0; r < r1; ++r) {
            const int len = (int)(seq_off[r + 1] - seq_off[r]);
            Family* f = &fam_generic;
            if (banded) {
                if (mod->banded_smem > 0 && len <= (forward ? 32 * kMaxRPL : ctx->short_max_len)) {
                    f = &fam_short;
                    pl.max_P_short = std::max(pl.max_P_short, mod->NCpad);
                    pl.max_smem_short = std::max(pl.max_smem_short, mod->banded_smem);
                } else {
                    f = &fam_long;                  // (forward too: the long-read kernel's log-sum-exp instantiation)
                    pl.max_P_long = std::max(pl.max_P_long, mod->NCpad);
                }
            }
            if (f == &fam_generic) pl.max_m_generic = std::max(pl.max_m_generic, mod->m);
            f->max_len = std::max(f->max_len, len);
            for (int s = 0; s < strands; ++s) {
                f->items.push_back((int32_t)(r * strands + s));
                f->model.push_back(gi);
            }
        }
    }
    if (fp32 && !fam_generic.items.empty())
        return set_error(ADVHMM_EUNSUPPORTED, "fp32 mode is implemented for the banded kernels (profile-shaped models) only");

    // generic launch geometry: as many warps per CTA as DP rows fit in shared memory
    int gwarps = kGenericWarpsMax, rows_in_smem = 1;
    if (!fam_generic.items.empty()) {
        const size_t per_warp = (size_t)pl.max_m_generic * 2 * sizeof(double);
        const size_t budget = ctx->smem_optin > 8192 ? ctx->smem_optin - 8192 : 0;
        gwarps = (int)std::min<size_t>(kGenericWarpsMax, per_warp ? budget / per_warp : kGenericWarpsMax);
        if (gwarps < 1) { gwarps = 4; rows_in_smem = 0; }
    }
    pl.order.reserve(n_out);
    pl.item_tile.reserve(n_out);
    auto make_tiles = [&](Family& f, int per_tile, int kind) {
        f.first_item = (int)pl.order.size();
        int in_tile = 0, last_model = -1;
        for (size_t i = 0; i < f.items.size(); ++i) {
            const int gi = f.model[i];
            if (in_tile == 0 || in_tile == per_tile || gi != last_model) {
                const void* dm = kind == 2 ? (forward ? (const void*)models[gi]->d_generic_fwd
                                                      : (const void*)models[gi]->d_generic)
                                           : (const void*)models[gi]->d_banded;
                pl.tiles.push_back(Tile{dm, (int32_t)pl.order.size(), 0});
                in_tile = 0;
            }
            last_model = gi;
            pl.tiles.back().cnt = ++in_tile;
            pl.item_tile.push_back((int32_t)pl.tiles.size() - 1);
            pl.order.push_back(f.items[i]);
        }
    };
    const int rpl_short = std::max(1, (fam_short.max_len + 31) / 32);
    const int short_warps = (forward || fp32 || rpl_short >= 8) ? 8 : banded_warps_for(ctx, pl.max_smem_short);
    make_tiles(fam_short, short_warps, 0);
    // long reads differ several-fold in length: within a model, longest first, so that the tail of a
    // launch is made of the cheap reads (CTAs start in order); results go back by read id anyway
    if (fam_long.items.size() > 1) {
        std::vector<int32_t> idx(fam_long.items.size());
        for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int32_t)i;
        auto len_of = [&](int32_t i) { const int64_t r = fam_long.items[i] / strands; return seq_off[r + 1] - seq_off[r]; };
        std::stable_sort(idx.begin(), idx.end(), [&](int32_t x, int32_t y) {
            if (fam_long.model[x] != fam_long.model[y]) return fam_long.model[x] < fam_long.model[y];
            return len_of(x) > len_of(y);
        });
        std::vector<int32_t> items(idx.size()), model(idx.size());
        for (size_t i = 0; i < idx.size(); ++i) { items[i] = fam_long.items[idx[i]]; model[i] = fam_long.model[idx[i]]; }
        fam_long.items.swap(items);
        fam_long.model.swap(model);
    }
    // long reads: `long_wpr` warps of a CTA share one read (its stripes are dealt round-robin).  One warp per
    // read while the traceback of two CTAs of reads per SM fits in the workspace and the reads are short;
    // more warps per read when a read has many stripes or few reads fit (a 20 kb read takes ~100 MB).
    int long_wpr = 1;
    if (!fam_long.items.empty()) {
        const size_t stripes = ((size_t)std::max(fam_long.max_len, 1) + 32 * kLongRPL - 1) / (32 * kLongRPL);
        const size_t per_read = stripes * 32 * (size_t)pl.max_P_long * 4;                    // traceback words
        const size_t sms = (size_t)std::max(ctx->sm_count, 1);
        // (a) two CTAs of reads per SM must fit in the workspace -- with room to spare: a launch of several
        // waves of CTAs (reads are ordered longest-first) back-fills the SMs whose reads finish early, a
        // launch of one wave lasts as long as its longest read (measured on 10-20 kb reads, 1,184 per step:
        // 8 warps per read 1,107 GCUPS, 4: 1,039, 2: 662; profiles/r2_long_kernel.md), hence also (b): reads
        // of many stripes always get more warps; (c) few reads (a PacBio locus has tens of spanning reads):
        // more warps per read fill the warp slots the missing reads leave empty
        while (long_wpr < kLongWarps &&
               (per_read * (2 * kLongWarps / long_wpr) * sms > ctx->workspace_budget_long ||
                stripes >= (size_t)16 * long_wpr ||
                (fam_long.items.size() * (size_t)long_wpr < 2 * kLongWarps * sms && stripes >= (size_t)4 * long_wpr)))
            long_wpr *= 2;
        if (const char* env = getenv("ADVHMM_LONG_WPR")) {
            const int v = atoi(env);
            if (v == 1 || v == 2 || v == 4 || v == 8) long_wpr = v;
        }
    }
    make_tiles(fam_long, kLongWarps / long_wpr, 1);
    make_tiles(fam_generic, gwarps, 2);
    const int n_short = (int)fam_short.items.size(), n_long = (int)fam_long.items.size();
    const int n_generic = (int)fam_generic.items.size();

    // ---- metadata upload (pinned staging, one H2D) -----------------------------------------
    const size_t b_seq_off = (size_t)(n_reads + 1) * sizeof(int64_t);
    const size_t b_pk_off = vec_bytes(pl.pk_off), b_order = vec_bytes(pl.order);
    const size_t b_item_tile = vec_bytes(pl.item_tile), b_tiles = vec_bytes(pl.tiles);
    const size_t o_seq_off = 0, o_pk_off = o_seq_off + al(b_seq_off), o_order = o_pk_off + al(b_pk_off);
    const size_t o_item_tile = o_order + al(b_order), o_tiles = o_item_tile + al(b_item_tile);
    const size_t meta_bytes = o_tiles + al(b_tiles);
    if (ctx->meta_done) CU_TRY(cudaEventSynchronize(ctx->meta_done));
    CU_TRY(ctx->h_meta.ensure(meta_bytes));
    CU_TRY(ctx->d_meta.ensure(meta_bytes));
    unsigned char* hm = static_cast<unsigned char*>(ctx->h_meta.p);
    memcpy(hm + o_seq_off, seq_off, b_seq_off);
    memcpy(hm + o_pk_off, pl.pk_off.data(), b_pk_off);
    memcpy(hm + o_order, pl.order.data(), b_order);
    memcpy(hm + o_item_tile, pl.item_tile.data(), b_item_tile);
    memcpy(hm + o_tiles, pl.tiles.data(), b_tiles);
    CU_TRY(cudaMemcpyAsync(ctx->d_meta.p, hm, meta_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (!ctx->meta_done) CU_TRY(cudaEventCreateWithFlags(&ctx->meta_done, cudaEventDisableTiming));
    CU_TRY(cudaEventRecord(ctx->meta_done, ctx->stream));
    unsigned char* dm = ctx->d_meta.as<unsigned char>();
    const int64_t* d_seq_off = reinterpret_cast<const int64_t*>(dm + o_seq_off);
    const int64_t* d_pk_off = reinterpret_cast<const int64_t*>(dm + o_pk_off);
    const int32_t* d_order = reinterpret_cast<const int32_t*>(dm + o_order);
    const int32_t* d_item_tile = reinterpret_cast<const int32_t*>(dm + o_item_tile);
    const Tile* d_tiles = reinterpret_cast<const Tile*>(dm + o_tiles);

    // ---- pack ------------------------------------------------------------------------------
    const size_t pk_words_al = (size_t)(pl.pk_words + 63) / 64 * 64;
    CU_TRY(ctx->d_pk.ensure(pk_words_al * sizeof(uint32_t) + (size_t)n_out * sizeof(int32_t)));
    uint32_t* d_pk = ctx->d_pk.as<uint32_t>();
    int32_t* d_rlen = reinterpret_cast<int32_t*>(d_pk + pk_words_al);
    {
        PackArgs pa{d_seqs, d_seq_off, d_pk_off, d_pk, d_rlen, d_bad, n_out, strands, 0, read_base};
        pa.n_symbols = models[0]->K;
        pack_reads_kernel<<<n_out, 32, 0, ctx->stream>>>(pa);
        CU_TRY(cudaGetLastError());
        ctx->launches++;
    }
    if (want_path && !continue_cursor) CU_TRY(cudaMemsetAsync(out.cursor, 0, sizeof(unsigned long long), ctx->stream));

    // ---- workspace: sized once for all families (they run one after the other on the stream),
    //      chunks end on tile boundaries --------------------------------------------------------
    const int rpl = std::max(1, (fam_short.max_len + 31) / 32);
    const size_t Ps = (size_t)pl.max_P_short, Pl = (size_t)pl.max_P_long;
    const size_t nw_short = rpl > 5 ? 2 : 1;
    const size_t s_per_item = n_short ? 32 * Ps * nw_short * 4 + 3 * Ps * 8 + (size_t)32 * rpl * 2 + 32 * 4 : 0;
    const size_t stripes_max = n_long ? ((size_t)std::max(fam_long.max_len, 1) + 32 * kLongRPL - 1) / (32 * kLongRPL) : 0;
    const size_t l_tbw_words = stripes_max * 32 * Pl;
    const size_t l_acc = stripes_max * 32 * kLongRPL;
    // (+ one path-sized scratch region per read: the backtrack of a long read walks once, not twice)
    const size_t l_scratch = (want_path && n_long) ? ((size_t)fam_long.max_len + 3 * Pl + 64) : 0;
    const size_t l_per_item = n_long ? l_tbw_words * 4 + 6 * Pl * 8 + l_acc * 2 + 32 * 4 + l_scratch * 4 : 0;
    const size_t gm = (size_t)pl.max_m_generic;
    const size_t g_tb_per = (want_walk && n_generic) ? (size_t)std::max(fam_generic.max_len, 1) * gm * sizeof(uint16_t) : 0;
    const size_t g_rows_per = (n_generic && !rows_in_smem) ? 2 * gm * sizeof(double) : 0;
    const size_t g_per_item = n_generic ? g_tb_per + g_rows_per + 8 : 0;
    auto chunk_of = [&](size_t per_item, int n_items, size_t floor_items, size_t budget = 0) {
        if (!n_items) return (size_t)0;
        size_t c = std::max<size_t>((budget ? budget : ctx->workspace_budget) / per_item, floor_items);
        return std::min<size_t>(c, (size_t)n_items);
    };
    size_t s_chunk = chunk_of(s_per_item, n_short, (size_t)kBandedWarpsMax * ctx->sm_count);
    // host-buffer calls that return state paths: run_host asks for host_chunks chunks (8 for a single batch,
    // 3 for the last sub-batch of a split one), so that the paths of chunk i travel to the host while chunk
    // i+1 is decoded
    if (ctx->mark_chunks && ctx->host_chunks > 1 && n_short >= ctx->host_chunks * 16384)
        s_chunk = std::min<size_t>(s_chunk, ((size_t)n_short + ctx->host_chunks - 1) / ctx->host_chunks);
    size_t l_chunk = chunk_of(l_per_item, n_long, (size_t)(kLongWarps / long_wpr), ctx->workspace_budget_long);
    // Long reads that do not fit in one chunk: the workspace is cut in two halves, the backtrack of chunk k
    // (a serial pointer chase per read, ~10 % of the fill time) runs on a second stream while chunk k+1 is
    // filled into the other half.
    const bool l_double = want_walk && l_chunk < (size_t)n_long && ctx->device >= 0;
    if (l_double) l_chunk = std::max<size_t>(chunk_of(2 * l_per_item, n_long, (size_t)(kLongWarps / long_wpr), ctx->workspace_budget_long), 1);
    {   // whole waves: two CTAs of kLongWarps / long_wpr reads per SM
        const size_t wave = (size_t)2 * (kLongWarps / long_wpr) * std::max(ctx->sm_count, 1);
        if (l_chunk > wave && l_chunk < (size_t)n_long) l_chunk = l_chunk / wave * wave;
    }
    const size_t g_chunk = chunk_of(g_per_item, n_generic, (size_t)gwarps);
    // short layout
    const size_t so_tbw = 0, so_vfin = al(s_chunk * 32 * Ps * nw_short * 4);
    const size_t so_acc = so_vfin + al(s_chunk * 3 * Ps * 8), so_ftb = so_acc + al(s_chunk * 32 * rpl * 2);
    const size_t s_bytes = so_ftb + al(s_chunk * 32 * 4);
    // long layout
    const size_t lo_tbw = 0, lo_vfin = al(l_chunk * l_tbw_words * 4), lo_carry = lo_vfin + al(l_chunk * 3 * Pl * 8);
    const size_t lo_acc = lo_carry + al(l_chunk * 3 * Pl * 8), lo_ftb = lo_acc + al(l_chunk * l_acc * 2);
    const size_t lo_scratch = lo_ftb + al(l_chunk * 32 * 4);
    const size_t l_half = lo_scratch + al(l_chunk * l_scratch * 4);
    const size_t l_bytes = l_double ? 2 * l_half : l_half;
    // generic layout
    const size_t go_tb = 0, go_rows = al(g_chunk * g_tb_per);
    const size_t g_bytes = go_rows + al(g_chunk * g_rows_per);
    const size_t o_end = std::max(s_bytes, std::max(l_bytes, g_bytes));   // end_state[n_out], whole batch
    CU_TRY(ctx->d_work.ensure(o_end + al((size_t)n_out * sizeof(int32_t))));
    unsigned char* w = ctx->d_work.as<unsigned char>();

    // [lo, hi) item ranges that start and end on tile boundaries and hold <= cap items
    auto next_chunk = [&](int lo, int end, size_t cap) {
        int hi = (int)std::min<size_t>((size_t)end, (size_t)lo + cap);
        if (hi < end)
            while (hi > lo && pl.item_tile[hi] == pl.item_tile[hi - 1]) --hi;
        if (hi == lo) {
            hi = lo + 1;
            while (hi < end && pl.item_tile[hi] == pl.item_tile[hi - 1]) ++hi;
        }
        return hi;
    };
    auto launch_backtrack = [&](int lo, int items, int bt_rpl, const uint32_t* tbw, size_t tbw_stride,
                                const uint16_t* acc, size_t acc_stride, const int32_t* ftb,
                                int32_t* scratch = nullptr, int64_t scratch_stride = 0, cudaStream_t bt_stream = nullptr) -> int {
        if (!bt_stream) bt_stream = ctx->stream;
        BandedBtArgs ba{};
        ba.scratch = scratch; ba.scratch_stride = scratch_stride;
        ba.tiles = d_tiles; ba.order = d_order; ba.chunk_base = lo; ba.n_items = items; ba.rpl = bt_rpl;
        ba.fp32 = fp32 ? 1 : 0;
        ba.pk = d_pk; ba.pk_off = d_pk_off; ba.rlen = d_rlen; ba.logp = out.logp;
        ba.tbw = tbw; ba.tbw_stride = tbw_stride; ba.acc_tb = acc; ba.acc_stride = acc_stride; ba.ftb = ftb;
        ba.item_tile = d_item_tile;
        ba.path_len = out.path_len; ba.path_off = out.path_off; ba.path = want_path ? out.path : nullptr;
        ba.path_cap = out.path_cap; ba.cursor = out.cursor; ba.summaries = out.summaries;
        {
            ProfScope prof(ctx, 1, bt_stream);
            banded_backtrack_kernel<<<(items + 127) / 128, 128, 0, bt_stream>>>(ba);
        }
        CU_TRY(cudaGetLastError());
        ctx->launches++;
        return ADVHMM_OK;
    };

    // ---- short banded reads (the Illumina path) ---------------------------------------------
    for (int lo = fam_short.first_item, end = lo + n_short; lo < end;) {
        const int hi = next_chunk(lo, end, s_chunk);
        const int tile0 = pl.item_tile[lo], tile1 = pl.item_tile[hi - 1] + 1;
        BandedArgs fa{};
        fa.tiles = d_tiles + tile0; fa.order = d_order; fa.chunk_base = lo;
        fa.pk = d_pk; fa.pk_off = d_pk_off; fa.rlen = d_rlen; fa.logp = out.logp;
        fa.tbw = reinterpret_cast<uint32_t*>(w + so_tbw); fa.tbw_stride = 32 * Ps * nw_short;
        fa.acc_tb = reinterpret_cast<uint16_t*>(w + so_acc); fa.acc_stride = 32 * rpl;
        fa.vfin = reinterpret_cast<double*>(w + so_vfin); fa.vfin_stride = 3 * Ps;
        fa.ftb = reinterpret_cast<int32_t*>(w + so_ftb);
        int rc = forward ? launch_banded_fwd_chunk(ctx, rpl, tile1 - tile0, pl.max_smem_short, fa)
                 : fp32  ? launch_banded_f32_chunk(ctx, rpl, tile1 - tile0, pl.max_smem_short, fa)
                         : launch_banded_chunk(ctx, rpl, tile1 - tile0, pl.max_smem_short, fa, short_warps);
        if (rc) return rc;
        if (want_walk) {
            rc = launch_backtrack(lo, hi - lo, rpl, fa.tbw, fa.tbw_stride, fa.acc_tb, (size_t)fa.acc_stride, fa.ftb);
            if (rc) return rc;
            if (want_path && (rc = mark_chunk(ctx, out.cursor))) return rc;
        }
        lo = hi;
    }

    // ---- long banded reads / models too large for shared memory ------------------------------
    if (l_double && !ctx->bt_stream) {
        CU_TRY(cudaStreamCreateWithFlags(&ctx->bt_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            CU_TRY(cudaEventCreateWithFlags(&ctx->fill_done[k], cudaEventDisableTiming));
            CU_TRY(cudaEventCreateWithFlags(&ctx->bt_done[k], cudaEventDisableTiming));
        }
    }
    int l_index = 0;
    for (int lo = fam_long.first_item, end = lo + n_long; lo < end; ++l_index) {
        const int hi = next_chunk(lo, end, l_chunk);
        const int tile0 = pl.item_tile[lo], tile1 = pl.item_tile[hi - 1] + 1;
        const int half = l_double ? (l_index & 1) : 0;
        unsigned char* wl = w + (size_t)half * l_half;
        if (l_double && l_index >= 2) CU_TRY(cudaStreamWaitEvent(ctx->stream, ctx->bt_done[half], 0));   // its backtrack is done
        LongArgs la{};
        la.tiles = d_tiles + tile0; la.order = d_order; la.chunk_base = lo;
        la.pk = d_pk; la.pk_off = d_pk_off; la.rlen = d_rlen; la.logp = out.logp;
        la.tbw = reinterpret_cast<uint32_t*>(wl + lo_tbw); la.tbw_stride = l_tbw_words;
        la.acc_tb = reinterpret_cast<uint16_t*>(wl + lo_acc); la.acc_stride = l_acc;
        la.vfin = reinterpret_cast<double*>(wl + lo_vfin); la.vfin_stride = 3 * Pl;
        la.carry = reinterpret_cast<double*>(wl + lo_carry); la.carry_stride = 3 * Pl;
        la.ftb = reinterpret_cast<int32_t*>(wl + lo_ftb);
        la.wpr = long_wpr;
        if (forward) {
            if (int rc = allow_max_dynamic_smem(ctx, banded_long_kernel<double, true>)) return rc;
            ProfScope prof(ctx, 0);
            banded_long_kernel<double, true><<<tile1 - tile0, kLongWarps * 32, kLongWarps * sizeof(LongRing<double>), ctx->stream>>>(la);
        } else if (fp32) {
            if (int rc = allow_max_dynamic_smem(ctx, banded_long_kernel<float>)) return rc;
            ProfScope prof(ctx, 0);
            banded_long_kernel<float><<<tile1 - tile0, kLongWarps * 32, kLongWarps * sizeof(LongRing<float>), ctx->stream>>>(la);
        } else {
            if (int rc = allow_max_dynamic_smem(ctx, banded_long_kernel<double>)) return rc;
            ProfScope prof(ctx, 0);
            banded_long_kernel<double><<<tile1 - tile0, kLongWarps * 32, kLongWarps * sizeof(LongRing<double>), ctx->stream>>>(la);
        }
        CU_TRY(cudaGetLastError());
        ctx->launches++;
        if (want_walk) {
            cudaStream_t bts = l_double ? ctx->bt_stream : ctx->stream;
            if (l_double) {
                CU_TRY(cudaEventRecord(ctx->fill_done[half], ctx->stream));
                CU_TRY(cudaStreamWaitEvent(bts, ctx->fill_done[half], 0));
            }
            int rc = launch_backtrack(lo, hi - lo, kLongRPL, la.tbw, la.tbw_stride, la.acc_tb, la.acc_stride, la.ftb,
                                      l_scratch ? reinterpret_cast<int32_t*>(wl + lo_scratch) : nullptr, (int64_t)l_scratch, bts);
            if (rc) return rc;
            if (want_path && (rc = mark_chunk(ctx, out.cursor, bts))) return rc;
            if (l_double) CU_TRY(cudaEventRecord(ctx->bt_done[half], bts));
        }
        lo = hi;
    }
    if (l_double && l_index > 0) {
        // whatever follows on the context's stream (the generic family in the same workspace, result copies)
        // comes after the last backtracks
        CU_TRY(cudaStreamWaitEvent(ctx->stream, ctx->bt_done[0], 0));
        if (l_index > 1) CU_TRY(cudaStreamWaitEvent(ctx->stream, ctx->bt_done[1], 0));
    }

    // ---- everything else: generic kernel ------------------------------------------------------
    if (n_generic > 0) {
        const int smem = rows_in_smem ? (int)(gwarps * 2 * gm * sizeof(double)) : 0;
        if (smem > 48 * 1024) {
            if (int rc = allow_max_dynamic_smem(ctx, generic_fill_kernel<false>)) return rc;
            if (int rc = allow_max_dynamic_smem(ctx, generic_fill_kernel<true>)) return rc;
        }
        for (int lo = fam_generic.first_item, end = lo + n_generic; lo < end;) {
            const int hi = next_chunk(lo, end, g_chunk);
            const int items = hi - lo;
            const int tile0 = pl.item_tile[lo], tile1 = pl.item_tile[hi - 1] + 1;
            GenericArgs ga{};
            ga.tiles = d_tiles + tile0; ga.order = d_order; ga.chunk_base = lo;
            ga.pk = d_pk; ga.pk_off = d_pk_off; ga.rlen = d_rlen; ga.logp = out.logp;
            ga.end_state = reinterpret_cast<int32_t*>(w + o_end);
            ga.tb = reinterpret_cast<uint16_t*>(w + go_tb); ga.tb_stride = g_tb_per / sizeof(uint16_t);
            ga.rows = reinterpret_cast<double*>(w + go_rows); ga.rows_stride = 2 * gm;
            ga.warps = gwarps; ga.rows_in_smem = rows_in_smem;
            if (forward) generic_fill_kernel<true><<<tile1 - tile0, gwarps * 32, smem, ctx->stream>>>(ga);
            else generic_fill_kernel<false><<<tile1 - tile0, gwarps * 32, smem, ctx->stream>>>(ga);
            CU_TRY(cudaGetLastError());
            ctx->launches++;
            if (want_walk) {
                GenericBtArgs ba{};
                ba.tiles = d_tiles; ba.order = d_order; ba.chunk_base = lo; ba.n_items = items;
                ba.rlen = d_rlen; ba.logp = out.logp; ba.end_state = ga.end_state;
                ba.tb = ga.tb; ba.tb_stride = ga.tb_stride; ba.item_tile = d_item_tile;
                ba.path_len = out.path_len; ba.path_off = out.path_off; ba.path = want_path ? out.path : nullptr;
                ba.path_cap = out.path_cap; ba.cursor = out.cursor;
                ba.pk = d_pk; ba.pk_off = d_pk_off; ba.summaries = out.summaries;
                generic_backtrack_kernel<<<(items + 127) / 128, 128, 0, ctx->stream>>>(ba);
                CU_TRY(cudaGetLastError());
                ctx->launches++;
                if (want_path) { int rc = mark_chunk(ctx, out.cursor); if (rc) return rc; }
            }
            lo = hi;
        }
    }
    return ADVHMM_OK;
}

// host-buffer front end shared by the three public decoding calls
int run_host(advhmm_context* ctx, advhmm_model* const* models, int n_models, const int64_t* group_off,
             const uint8_t* seqs, const int64_t* seq_off, int n_reads, uint32_t flags, bool forward,
             double* logp, int32_t* path_len, int64_t* path_off, int32_t* path, int64_t path_cap,
             int64_t* path_total, advhmm_read_summary* summaries = nullptr)
{
    if (!ctx || ctx->device < 0) return set_error(ADVHMM_ECUDA, "this context has no CUDA device (host-only analysis context)");
    if (n_reads < 0 || !seq_off || (n_reads > 0 && (!logp || !models || n_models <= 0)))
        return set_error(ADVHMM_EINVAL, "null or negative argument");
    const bool want_path = (flags & ADVHMM_WANT_PATH) && !forward;
    const bool want_sum = (flags & ADVHMM_WANT_SUMMARY) && !forward;
    if (want_path && (!path_len || !path_off || !path_total || (path_cap > 0 && !path)))
        return set_error(ADVHMM_EINVAL, "ADVHMM_WANT_PATH needs path_len, path_off, path and path_total");
    if (want_sum && !summaries) return set_error(ADVHMM_EINVAL, "ADVHMM_WANT_SUMMARY needs a summaries buffer");
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU_TRY(cudaSetDevice(ctx->device));
    const int strands = (flags & ADVHMM_BOTH_STRANDS) ? 2 : 1;
    const int n_out = n_reads * strands;
    if (path_total) *path_total = 0;
    if (n_out == 0) return ADVHMM_OK;
    const int64_t n_bases = seq_off[n_reads];
    if (seq_off[0] != 0 || n_bases < 0) return set_error(ADVHMM_EINVAL, "seq_off must start at 0 and be non-decreasing");
    // (run_batch checks the same, but it sees the sub-batches below with re-based offsets)
    if (!group_off || group_off[0] != 0 || group_off[n_models] != n_reads)
        return set_error(ADVHMM_EINVAL, "group_off must start at 0 and end at n_reads: reads outside every group "
                                        "would not be decoded");
    for (int g = 0; g < n_models; ++g)
        if (group_off[g + 1] < group_off[g]) return set_error(ADVHMM_EINVAL, "group_off must be non-decreasing (model %d)", g);
    if (n_bases > 0 && !seqs) return set_error(ADVHMM_EINVAL, "seqs is null");

    CU_TRY(ctx->d_seqs.ensure((size_t)n_bases + 16));
    // outputs: logp | path_len | path_off | cursor | bad
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t o_logp = 0, o_plen = al((size_t)n_out * 8), o_poff = o_plen + al((size_t)n_out * 4);
    const size_t o_cursor = o_poff + al((size_t)n_out * 8), o_bad = o_cursor + 256;
    const size_t o_sum = o_bad + 256;
    const size_t out_bytes = o_sum + (want_sum ? al((size_t)n_out * sizeof(advhmm_read_summary)) : 0);
    CU_TRY(ctx->d_out.ensure(out_bytes));
    unsigned char* d = ctx->d_out.as<unsigned char>();
    int64_t cap = want_path ? path_cap : 0;
    if (want_path) CU_TRY(ctx->d_paths.ensure((size_t)std::max<int64_t>(cap, 1) * sizeof(int32_t)));
    OutPtrs op{reinterpret_cast<double*>(d + o_logp), reinterpret_cast<int32_t*>(d + o_plen),
               reinterpret_cast<int64_t*>(d + o_poff), ctx->d_paths.as<int32_t>(), cap,
               reinterpret_cast<unsigned long long*>(d + o_cursor),
               want_sum ? reinterpret_cast<advhmm_read_summary*>(d + o_sum) : nullptr};
    int32_t* d_bad = reinterpret_cast<int32_t*>(d + o_bad);
    CU_TRY(cudaMemsetAsync(d_bad, 0x7f, sizeof(int32_t), ctx->stream));
    ctx->n_marks = 0;
    ctx->mark_chunks = false;
    static const bool dbg = getenv("ADVHMM_DEBUG_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = now();
    if (want_path && cap > 0 && !getenv("ADVHMM_NO_OVERLAP")) {
        if (!ctx->copy_stream) CU_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CU_TRY(ctx->h_cursors.ensure(kMaxChunkMarks * sizeof(unsigned long long)));
        ctx->mark_chunks = true;
    }
    // Large batches of many models go to the device as up to five sub-batches (cut at model boundaries):
    // planning sub-batch k+1 on the host overlaps decoding sub-batch k on the device.
    // The first sub-batch is small (1/16 of the reads): nothing runs on the device while it is planned.
    const int n_sub = (n_out >= 262144 && n_models >= 8) ? 5 : 1;
    std::vector<int> cut{0};                            // model index where sub-batch k starts
    for (int k = 0; k + 1 < n_sub; ++k) {
        const int64_t want = (int64_t)n_reads * (1 + 4 * k) / 16;         // reads before the cut: 1/16, 5/16, 9/16, 13/16
        const int g = (int)(std::lower_bound(group_off + cut.back(), group_off + n_models, want) - group_off);
        if (g > cut.back() && g < n_models) cut.push_back(g);
    }
    cut.push_back(n_models);
    const int n_cut = (int)cut.size() - 1;
    // Sequences: the first sub-batch's bases travel on the compute stream, the others on the copy stream
    // while the first is being decoded; every later sub-batch waits for its own bases only.
    const bool split_h2d = n_cut > 1;
    if (split_h2d && !ctx->copy_stream) CU_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < n_cut && n_bases; ++k) {
        const int64_t b0 = k == 0 ? 0 : seq_off[group_off[cut[k]]];
        const int64_t b1 = k + 1 == n_cut ? n_bases : seq_off[group_off[cut[k + 1]]];
        cudaStream_t st = (k == 0 || !split_h2d) ? ctx->stream : ctx->copy_stream;
        if (b1 > b0)
            CU_TRY(cudaMemcpyAsync(ctx->d_seqs.as<uint8_t>() + b0, seqs + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, st));
        if (k > 0 && split_h2d) {
            while ((int)ctx->h2d_events.size() < n_cut) {
                cudaEvent_t ev;
                CU_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                ctx->h2d_events.push_back(ev);
            }
            CU_TRY(cudaEventRecord(ctx->h2d_events[k], ctx->copy_stream));
        }
    }
    int rc = ADVHMM_OK;
    std::vector<int64_t> rebased;
    if (want_path) CU_TRY(cudaMemsetAsync(op.cursor, 0, sizeof(unsigned long long), ctx->stream));
    for (int k = 0; k < n_cut && rc == ADVHMM_OK; ++k) {
        const int g0 = cut[k], g1 = cut[k + 1];
        // state paths leave in chunks while the next chunk is decoded; the last sub-batch is cut finer,
        // because its last chunk is the one copy nothing overlaps
        ctx->host_chunks = n_cut > 1 ? (k + 1 == n_cut ? 3 : 1) : 8;       // every extra launch costs ~0.2 ms of tail
        if (k > 0 && split_h2d) CU_TRY(cudaStreamWaitEvent(ctx->stream, ctx->h2d_events[k], 0));
        const int64_t r0 = group_off[g0], r1 = group_off[g1];
        rebased.assign(group_off + g0, group_off + g1 + 1);
        for (int64_t& v : rebased) v -= r0;
        OutPtrs sub = op;
        sub.logp += r0 * strands; sub.path_len += r0 * strands; sub.path_off += r0 * strands;
        if (sub.summaries) sub.summaries += r0 * strands;
        rc = run_batch(ctx, models + g0, g1 - g0, rebased.data(), ctx->d_seqs.as<uint8_t>(), seq_off + r0, (int)(r1 - r0),
                       flags, sub, forward, d_bad, (int)r0, /*continue_cursor=*/true);
    }
    ctx->mark_chunks = false;
    if (rc) { cudaStreamSynchronize(ctx->stream); return rc; }
    const double t_enq = now();
    // state paths: every finished chunk goes home on the copy stream while the next one is decoded
    // (chunks run one after the other and allocate from one cursor, so chunk i owns
    // [cursor after chunk i-1, cursor after chunk i) of the path buffer)
    unsigned long long copied = 0;
    for (size_t i = 0; i < ctx->n_marks; ++i) {
        CU_TRY(cudaEventSynchronize(ctx->chunk_events[i]));
        const unsigned long long upto = std::min<unsigned long long>(
            static_cast<const unsigned long long*>(ctx->h_cursors.p)[i], (unsigned long long)cap);
        if (upto > copied) {
            CU_TRY(cudaMemcpyAsync(path + copied, ctx->d_paths.as<int32_t>() + copied, (size_t)(upto - copied) * sizeof(int32_t),
                                   cudaMemcpyDeviceToHost, ctx->copy_stream));
            copied = upto;
        }
    }
    const double t_marks = now();
    // results back: straight into the caller's arrays when those are page-locked, else through the
    // context's pinned staging buffer and a host copy
    auto pinned = [](const void* p) {
        if (!p) return true;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return at.type == cudaMemoryTypeHost;
    };
    const bool direct = pinned(logp) && pinned(want_path || want_sum ? path_len : nullptr) &&
                        pinned(want_path ? path_off : nullptr) && pinned(want_sum ? summaries : nullptr);
    CU_TRY(ctx->h_out.ensure(out_bytes));
    if (direct) {
        CU_TRY(cudaMemcpyAsync(logp, d + o_logp, (size_t)n_out * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if ((want_path || want_sum) && path_len)
            CU_TRY(cudaMemcpyAsync(path_len, d + o_plen, (size_t)n_out * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (want_path) CU_TRY(cudaMemcpyAsync(path_off, d + o_poff, (size_t)n_out * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (want_sum)
            CU_TRY(cudaMemcpyAsync(summaries, d + o_sum, (size_t)n_out * sizeof(advhmm_read_summary), cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(cudaMemcpyAsync(static_cast<unsigned char*>(ctx->h_out.p) + o_cursor, d + o_cursor, 512, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        CU_TRY(cudaMemcpyAsync(ctx->h_out.p, d, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    const double t_sync = now();
    const unsigned char* h = static_cast<const unsigned char*>(ctx->h_out.p);
    int32_t bad;
    memcpy(&bad, h + o_bad, sizeof bad);
    if (bad != 0x7f7f7f7f) return set_error(ADVHMM_ESYMBOL, "read %d contains a symbol code outside the model alphabet", bad);
    if (!direct) {
        memcpy(logp, h + o_logp, (size_t)n_out * 8);
        if (want_sum) {
            memcpy(summaries, h + o_sum, (size_t)n_out * sizeof(advhmm_read_summary));
            if (path_len && !want_path) memcpy(path_len, h + o_plen, (size_t)n_out * 4);
        }
    }
    if (want_path) {
        if (!direct) {
            memcpy(path_len, h + o_plen, (size_t)n_out * 4);
            memcpy(path_off, h + o_poff, (size_t)n_out * 8);
        }
        unsigned long long total;
        memcpy(&total, h + o_cursor, sizeof total);
        *path_total = (int64_t)total;
        if (ctx->copy_stream) CU_TRY(cudaStreamSynchronize(ctx->copy_stream));
        if ((int64_t)total > path_cap)
            return set_error(ADVHMM_ECAPACITY, "path buffer too small: need %lld entries, have %lld",
                             (long long)total, (long long)path_cap);
        if (dbg) fprintf(stderr, "[advhmm] host call: enqueue %.2f ms, chunk marks (%zu) %.2f ms, outputs+sync %.2f ms, copy-stream drain %.2f ms, copied %llu of %llu\n",
                         t_enq - t_start, ctx->n_marks, t_marks - t_enq, t_sync - t_marks, now() - t_sync, copied, total);
        if (total > copied) {               // whatever the chunk marks did not cover
            CU_TRY(cudaMemcpyAsync(path + copied, ctx->d_paths.as<int32_t>() + copied, (size_t)(total - copied) * sizeof(int32_t),
                                   cudaMemcpyDeviceToHost, ctx->stream));
            CU_TRY(cudaStreamSynchronize(ctx->stream));
        }
    }
    return ADVHMM_OK;
}

}  // namespace

// =============================================================================================
// fp64 add/compare issue-rate microbenchmark (the roofline denominator of the fill kernel)
// =============================================================================================
namespace {
// Each thread runs 8 independent dependent-chains of DADD; with 1024 threads per SM resident the
// fp64 pipe is saturated.  ops = threads * iters * 8.
__global__ void __launch_bounds__(256) fp64_add_peak_kernel(double* out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double d = seed * 1e-9;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a0 = __dadd_rn(a0, d); a1 = __dadd_rn(a1, d); a2 = __dadd_rn(a2, d); a3 = __dadd_rn(a3, d);
            a4 = __dadd_rn(a4, d); a5 = __dadd_rn(a5, d); a6 = __dadd_rn(a6, d); a7 = __dadd_rn(a7, d);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
}  // namespace

// =============================================================================================
// keyword pre-filter: host side
// =============================================================================================
struct advhmm_kfilter {
    advhmm_context* ctx = nullptr;
    DevKFilter dev{};
    DevBuf tables, seqs, meta, tiles, counters, hits;
    int64_t n_keywords = 0, n_unique = 0;
};

namespace {

unsigned long long next_pow2(unsigned long long x)
{
    unsigned long long p = 1;
    while (p < x) p <<= 1;
    return p;
}

struct KfWord {                         // one distinct (length class, keyword)
    int cls;
    std::string codes;                  // symbol codes 0..4
    std::vector<int32_t> loci;
};

int kfilter_build(advhmm_kfilter* kf, int64_t n, const char* words, const int64_t* word_off, const int32_t* locus)
{
    advhmm_context* ctx = kf->ctx;
    // length classes
    std::vector<int> lengths;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t k = word_off[i + 1] - word_off[i];
        if (k < 1 || k > kKfMaxK)
            return set_error(k < 1 ? ADVHMM_EINVAL : ADVHMM_EUNSUPPORTED,
                             "keyword %lld has length %lld: the device filter handles 1..%d", (long long)i, (long long)k, kKfMaxK);
        if (std::find(lengths.begin(), lengths.end(), (int)k) == lengths.end()) lengths.push_back((int)k);
    }
    std::sort(lengths.begin(), lengths.end());
    if ((int)lengths.size() > kKfMaxClasses)
        return set_error(ADVHMM_EUNSUPPORTED, "%d distinct keyword lengths: the device filter handles %d per filter",
                         (int)lengths.size(), kKfMaxClasses);
    DevKFilter& d = kf->dev;
    d = DevKFilter{};
    d.n_classes = (int)lengths.size();
    int kmax = 1;
    for (int c = 0; c < d.n_classes; ++c) {
        KfClass& cl = d.cls[c];
        cl.k = lengths[c];
        cl.exact = cl.k <= 21;
        cl.key_mask = (3 * cl.k >= 64) ? ~0ull : ((1ull << (3 * cl.k)) - 1);
        cl.salt = 0xD6E8FEB86659FD93ull * (unsigned long long)(c + 1);
        cl.bk = 1;
        for (int j = 0; j < cl.k; ++j) cl.bk *= kKfBase;
        kmax = std::max(kmax, cl.k);
    }
    d.halo = (kmax - 1 + 15) / 16 * 16;
    // distinct keywords per class, each with the loci that own it (in the caller's order)
    std::unordered_map<std::string, size_t> index;
    std::vector<KfWord> uniq;
    index.reserve((size_t)n * 2);
    for (int64_t i = 0; i < n; ++i) {
        const int k = (int)(word_off[i + 1] - word_off[i]);
        const int c = (int)(std::lower_bound(lengths.begin(), lengths.end(), k) - lengths.begin());
        std::string codes((size_t)k, '\0');
        for (int j = 0; j < k; ++j) codes[j] = (char)kf_code((unsigned char)words[word_off[i] + j]);
        auto it = index.find(codes);       // the length is part of the string, hence of the class
        if (it == index.end()) {
            it = index.emplace(codes, uniq.size()).first;
            uniq.push_back(KfWord{c, codes, {}});
        }
        uniq[it->second].loci.push_back(locus[i]);
    }
    const unsigned long long nu = uniq.size();
    const unsigned long long tsize = next_pow2(std::max<unsigned long long>(64, nu * 2));
    const unsigned long long bwords = std::min<unsigned long long>(1ull << 28, next_pow2(std::max<unsigned long long>(1 << 10, nu * 2)));
    int bbits = 0;
    while ((1ull << bbits) < bwords) ++bbits;
    std::vector<KfEntry> table(tsize, KfEntry{0, 0, 0, 0, 0, 0});
    std::vector<uint32_t> bloom(bwords, 0);
    std::vector<uint32_t> l0(kKfL0Bits / 32, 0);
    std::vector<int32_t> loci;
    std::vector<uint8_t> text;
    loci.reserve((size_t)n);
    for (const KfWord& w : uniq) {
        const KfClass& cl = d.cls[w.cls];
        unsigned long long key = 0;
        if (cl.exact) for (char ch : w.codes) key = (key << 3) | (unsigned long long)ch;
        else          for (char ch : w.codes) key = key * kKfBase + ((unsigned long long)ch + 1u);
        uint32_t a;
        const unsigned long long h = kf_hash(key, cl.salt, a);
        bloom[a >> (32 - bbits)] |= 1u << kf_bloom_bit(h);
        l0[(a & (kKfL0Bits - 1)) >> 5] |= 1u << (a & 31);
        unsigned long long slot = kf_slot(h, tsize - 1);
        while (table[slot].loci_cnt) slot = (slot + 1) & (tsize - 1);
        KfEntry& e = table[slot];
        e.key = key;
        e.cls = (uint32_t)w.cls;
        e.loci_off = (uint32_t)loci.size();
        e.loci_cnt = (uint32_t)w.loci.size();
        e.text_off = (uint32_t)text.size();
        loci.insert(loci.end(), w.loci.begin(), w.loci.end());
        if (!cl.exact) text.insert(text.end(), w.codes.begin(), w.codes.end());
    }
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t o_table = 0, o_bloom = al(table.size() * sizeof(KfEntry));
    const size_t o_loci = o_bloom + al(bloom.size() * 4), o_text = o_loci + al(loci.size() * 4 + 4);
    const size_t o_l0 = o_text + al(text.size() + 16);
    const size_t total = o_l0 + al(l0.size() * 4);
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(kf->tables.ensure(total));
    unsigned char* dp = kf->tables.as<unsigned char>();
    CU_TRY(cudaMemcpyAsync(dp + o_table, table.data(), table.size() * sizeof(KfEntry), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaMemcpyAsync(dp + o_bloom, bloom.data(), bloom.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (!loci.empty())
        CU_TRY(cudaMemcpyAsync(dp + o_loci, loci.data(), loci.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (!text.empty())
        CU_TRY(cudaMemcpyAsync(dp + o_text, text.data(), text.size(), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaMemcpyAsync(dp + o_l0, l0.data(), l0.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    d.table = reinterpret_cast<const KfEntry*>(dp + o_table);
    d.table_mask = tsize - 1;
    d.bloom = reinterpret_cast<const uint32_t*>(dp + o_bloom);
    d.l0 = reinterpret_cast<const uint32_t*>(dp + o_l0);
    d.bloom_shift = (uint32_t)(32 - bbits);
    d.loci = reinterpret_cast<const int32_t*>(dp + o_loci);
    d.text = dp + o_text;
    kf->n_keywords = n;
    kf->n_unique = (int64_t)nu;
    return ADVHMM_OK;
}

// d_seqs, d_off and the hit arrays are device pointers; d_n_hits a device counter.  Queues tile
// index + scan + compaction on the context's stream; the only host synchronisation is the
// read-back of the "counter table full" flag.
int kfilter_scan_device(advhmm_kfilter* kf, const unsigned char* d_seqs, const int64_t* d_off, int n_reads,
                        int64_t n_bases, int min_matches, int32_t* d_hit_read, int32_t* d_hit_locus,
                        int32_t* d_hit_count, int64_t hit_cap, unsigned long long* d_n_hits)
{
    advhmm_context* ctx = kf->ctx;
    CU_TRY(cudaMemsetAsync(d_n_hits, 0, sizeof(unsigned long long), ctx->stream));
    if (n_bases <= 0 || kf->dev.n_classes == 0) return ADVHMM_OK;
    if (reinterpret_cast<uintptr_t>(d_seqs) % 16)
        return set_error(ADVHMM_EINVAL, "keyword filter: the device read buffer must be 16-byte aligned");
    // which read owns the first byte of every tile
    const int64_t n_tiles = (n_bases + kKfTile - 1) / kKfTile;
    CU_TRY(kf->tiles.ensure((size_t)(n_tiles + 1) * 4));
    int32_t* d_tile = kf->tiles.as<int32_t>();
    kfilter_tile_index_kernel<<<(unsigned)((n_tiles + 1 + 255) / 256), 256, 0, ctx->stream>>>(d_off, n_reads, n_tiles, d_tile);
    CU_TRY(cudaGetLastError());
    ctx->launches++;
    // warps per CTA: as many as fit next to the first-level bitmap with two (halo + tile) buffers each
    const size_t per_warp = 2 * ((size_t)kf->dev.halo + kKfTile) + (size_t)kKfQueue * 4;
    const size_t budget = (ctx->smem_optin > 4096 ? ctx->smem_optin - 4096 : 0);
    int n_warps = budget > (size_t)kKfL0Bytes ? (int)((budget - kKfL0Bytes) / per_warp) : 0;
    n_warps = std::min(n_warps, kKfMaxWarps);
    if (n_warps < 1) return set_error(ADVHMM_EUNSUPPORTED, "keyword filter: not enough shared memory for a warp pipeline");
    const size_t smem = (size_t)kKfL0Bytes + (size_t)n_warps * per_warp;
    if (int rc = allow_max_dynamic_smem(ctx, kfilter_scan_kernel)) return rc;
    const unsigned grid = (unsigned)std::min<int64_t>((n_tiles + n_warps - 1) / n_warps, std::max(ctx->sm_count, 1));
    unsigned long long cap = next_pow2(std::max<unsigned long long>(1 << 16, (unsigned long long)n_reads / 2));
    for (int attempt = 0; attempt < 10; ++attempt, cap <<= 2) {
        const size_t bytes = (size_t)cap * 12 + 256;
        CU_TRY(kf->counters.ensure(bytes));
        unsigned long long* ck = kf->counters.as<unsigned long long>();
        uint32_t* cv = reinterpret_cast<uint32_t*>(ck + cap);
        int32_t* ovf = reinterpret_cast<int32_t*>(cv + cap);
        CU_TRY(cudaMemsetAsync(ck, 0xff, (size_t)cap * 8, ctx->stream));
        CU_TRY(cudaMemsetAsync(cv, 0, (size_t)cap * 4 + 64, ctx->stream));
        KfScanArgs sa{kf->dev, d_seqs, d_off, d_tile, n_bases, n_reads, (int32_t)n_tiles, ck, cv, cap - 1, ovf};
        {
            ProfScope prof(ctx, 0);
            kfilter_scan_kernel<<<grid, n_warps * 32, smem, ctx->stream>>>(sa);
        }
        CU_TRY(cudaGetLastError());
        ctx->launches++;
        int32_t overflow = 0;
        CU_TRY(cudaMemcpyAsync(&overflow, ovf, sizeof overflow, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(cudaStreamSynchronize(ctx->stream));
        if (overflow) continue;
        KfCompactArgs ca{ck, cv, cap, min_matches, d_hit_read, d_hit_locus, d_hit_count, (long long)hit_cap, d_n_hits};
        kfilter_compact_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, ctx->stream>>>(ca);
        CU_TRY(cudaGetLastError());
        ctx->launches++;
        return ADVHMM_OK;
    }
    return set_error(ADVHMM_ENOMEM, "keyword filter: occurrence counter table kept overflowing");
}

}  // namespace

// =============================================================================================
// C-ABI
// =============================================================================================
extern "C" {

const char* advhmm_last_error(void) { return g_last_error.c_str(); }
int advhmm_abi_version(void) { return ADVHMM_ABI_VERSION; }

int64_t advhmm_encode_acgt(const char* ascii, int64_t n, uint8_t* codes)
{
    static const struct Lut {
        uint8_t t[256];
        Lut() { memset(t, 255, sizeof t); t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3; }
    } lut;
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t c = lut.t[(unsigned char)ascii[i]];
        if (c == 255) return i;
        codes[i] = c;
    }
    return -1;
}

int advhmm_context_create(int device, void* stream, advhmm_context** out)
{
    if (!out) return set_error(ADVHMM_EINVAL, "out is null");
    *out = nullptr;
    std::unique_ptr<advhmm_context> ctx(new (std::nothrow) advhmm_context);
    if (!ctx) return set_error(ADVHMM_ENOMEM, "out of host memory");
    ctx->device = device;
    if (device >= 0) {
        int count = 0;
        CU_TRY(cudaGetDeviceCount(&count));
        if (device >= count) return set_error(ADVHMM_EINVAL, "device %d does not exist (%d visible)", device, count);
        CU_TRY(cudaSetDevice(device));
        cudaDeviceProp prop;
        CU_TRY(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) return set_error(ADVHMM_EUNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        ctx->sm_count = prop.multiProcessorCount;
        ctx->smem_optin = prop.sharedMemPerBlockOptin;
        if (stream) { ctx->stream = static_cast<cudaStream_t>(stream); }
        else { CU_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->owns_stream = true; }
        size_t free_b = 0, total_b = 0;
        CU_TRY(cudaMemGetInfo(&free_b, &total_b));
        // traceback workspace (only what a batch needs is allocated): short reads take up to half of the free
        // device memory (capped at 96 GB); long reads -- ~100 MB per 20 kb read, and a launch of several waves
        // of CTAs back-fills the SMs whose reads finish early -- up to three quarters (capped at 140 GB);
        // ADVHMM_WORKSPACE_MB overrides both
        ctx->workspace_budget = std::min<size_t>((size_t)96 << 30, free_b / 2);
        ctx->workspace_budget_long = std::min<size_t>((size_t)140 << 30, free_b / 4 * 3);
        const char* env = getenv("ADVHMM_WORKSPACE_MB");
        if (env && atoll(env) > 0) ctx->workspace_budget = ctx->workspace_budget_long = (size_t)atoll(env) << 20;
        env = getenv("ADVHMM_SHORT_MAX_LEN");
        if (env && atoi(env) > 0) ctx->short_max_len = std::min(atoi(env), 32 * kMaxRPL);
    }
    *out = ctx.release();
    return ADVHMM_OK;
}

void advhmm_context_destroy(advhmm_context* ctx)
{
    if (!ctx) return;
    if (ctx->device >= 0) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        for (DevBuf* b : {&ctx->d_seqs, &ctx->d_seq_off, &ctx->d_pk, &ctx->d_meta, &ctx->d_work, &ctx->d_out, &ctx->d_paths, &ctx->d_flags, &ctx->d_badflag}) b->release();
        ctx->h_meta.release(); ctx->h_out.release(); ctx->h_cursors.release();
        for (int k = 0; k < 2; ++k) {
            ctx->h_stage[k].release();
            if (ctx->stage_done[k]) cudaEventDestroy(ctx->stage_done[k]);
        }
        for (auto& kv : ctx->shape_dev) kv.second->blob.release();
        ctx->alive->store(false);
        if (ctx->bt_stream) {
            cudaStreamSynchronize(ctx->bt_stream);
            cudaStreamDestroy(ctx->bt_stream);
            for (int k = 0; k < 2; ++k) { cudaEventDestroy(ctx->fill_done[k]); cudaEventDestroy(ctx->bt_done[k]); }
        }
        if (ctx->upload_stream) { cudaStreamSynchronize(ctx->upload_stream); cudaStreamDestroy(ctx->upload_stream); }
        if (ctx->upload_done) cudaEventDestroy(ctx->upload_done);
        for (cudaEvent_t ev : ctx->chunk_events) cudaEventDestroy(ev);
        for (cudaEvent_t ev : ctx->h2d_events) cudaEventDestroy(ev);
        if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
        if (ctx->meta_done) cudaEventDestroy(ctx->meta_done);
        for (auto& v : ctx->prof_events)
            for (auto& pr : v) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
        if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    }
    delete ctx;
}

int advhmm_context_synchronize(advhmm_context* ctx)
{
    if (!ctx || ctx->device < 0) return ADVHMM_OK;
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    return ADVHMM_OK;
}

void* advhmm_context_stream(advhmm_context* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int advhmm_context_profile(advhmm_context* ctx, int enable)
{
    if (!ctx) return set_error(ADVHMM_EINVAL, "ctx is null");
    ctx->profile = enable != 0;
    ctx->prof_used[0] = ctx->prof_used[1] = 0;
    return ADVHMM_OK;
}

int advhmm_context_profile_read(advhmm_context* ctx, double* fill_ms, int64_t* fill_launches,
                                double* backtrack_ms, int64_t* backtrack_launches)
{
    if (!ctx || ctx->device < 0) return set_error(ADVHMM_EINVAL, "no device context");
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    double ms[2] = {0, 0};
    for (int k = 0; k < 2; ++k)
        for (size_t i = 0; i < ctx->prof_used[k]; ++i) {
            float t = 0;
            CU_TRY(cudaEventElapsedTime(&t, ctx->prof_events[k][i].first, ctx->prof_events[k][i].second));
            ms[k] += t;
        }
    if (fill_ms) *fill_ms = ms[0];
    if (fill_launches) *fill_launches = (int64_t)ctx->prof_used[0];
    if (backtrack_ms) *backtrack_ms = ms[1];
    if (backtrack_launches) *backtrack_launches = (int64_t)ctx->prof_used[1];
    ctx->prof_used[0] = ctx->prof_used[1] = 0;
    return ADVHMM_OK;
}

int advhmm_fp64_add_peak(advhmm_context* ctx, double* gops)
{
    if (!ctx || ctx->device < 0 || !gops) return set_error(ADVHMM_EINVAL, "no device context");
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU_TRY(cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
    CU_TRY(ctx->d_flags.ensure((size_t)blocks * threads * sizeof(double) + 256));
    double* out = reinterpret_cast<double*>(ctx->d_flags.as<unsigned char>() + 256);
    cudaEvent_t a, b;
    CU_TRY(cudaEventCreate(&a));
    CU_TRY(cudaEventCreate(&b));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CU_TRY(cudaEventRecord(a, ctx->stream));
        fp64_add_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(out, iters, 1.0 + rep);
        CU_TRY(cudaEventRecord(b, ctx->stream));
        CU_TRY(cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        CU_TRY(cudaEventElapsedTime(&ms, a, b));
        const double ops = (double)blocks * threads * iters * 64.0;
        if (rep > 0) best = std::max(best, ops / (ms * 1e-3) / 1e9);
        ctx->launches++;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    *gops = best;
    return ADVHMM_OK;
}
int64_t advhmm_context_launch_count(advhmm_context* ctx) { return ctx ? ctx->launches : 0; }

int advhmm_model_create(advhmm_context* ctx, const advhmm_model_desc* desc, advhmm_model** out)
{
    if (!ctx || !desc || !out) return set_error(ADVHMM_EINVAL, "null argument");
    *out = nullptr;
    std::unique_ptr<advhmm_model> mod(new (std::nothrow) advhmm_model);
    if (!mod) return set_error(ADVHMM_ENOMEM, "out of host memory");
    mod->ctx = ctx;
    std::string err;
    if (!compile_model(*desc, mod->cm, err)) return set_error(ADVHMM_EINVAL, "%s", err.c_str());
    if (mod->cm.g.max_in_degree > 65535) return set_error(ADVHMM_EUNSUPPORTED, "in-degree %d exceeds the 16-bit traceback slot", mod->cm.g.max_in_degree);
    std::lock_guard<std::mutex> lock(ctx->mu);
    int rc = upload_model(mod.get());
    if (rc) return rc;
    *out = mod.release();
    return ADVHMM_OK;
}

void advhmm_model_destroy(advhmm_model* model)
{
    if (!model) return;
    if (model->ctx && model->ctx->device >= 0 && model->blob.p) {
        cudaSetDevice(model->ctx->device);
        cudaStreamSynchronize(model->ctx->stream);
        model->blob.release();
    }
    // (a locus model's tables live in its batch's arena, freed in stream order behind the kernels that read them)
    delete model;
}

void advhmm_shape_cache_clear(void) { rm::clear_shape_cache(); }

int advhmm_set_vexp(advhmm_vexp_fn fn, void* user)
{
    g_vexp = fn;
    g_vexp_user = user;
    return ADVHMM_OK;
}

int advhmm_models_create_for_loci(advhmm_context* ctx, const advhmm_loci* loci, int32_t n_threads, advhmm_model** out)
{
    if (!ctx || !loci || !out || loci->n_loci < 0) return set_error(ADVHMM_EINVAL, "null argument");
    if (loci->n_loci == 0) return ADVHMM_OK;
    if (!loci->left || !loci->left_off || !loci->right || !loci->right_off || !loci->segments || !loci->seg_off ||
        !loci->n_segments || !loci->copies || !loci->error_rate)
        return set_error(ADVHMM_EINVAL, "advhmm_loci has a null column");
    std::lock_guard<std::mutex> lock(ctx->mu);
    try {
        int rc = create_models_for_loci(ctx, loci, n_threads, out);
        if (rc != ADVHMM_OK)
            for (int i = 0; i < loci->n_loci; ++i) out[i] = nullptr;
        return rc;
    } catch (const std::bad_alloc&) {
        return set_error(ADVHMM_ENOMEM, "out of host memory");
    } catch (const std::exception& e) {
        return set_error(ADVHMM_EINVAL, "%s", e.what());
    }
}

int advhmm_model_dims_get(const advhmm_model* model, advhmm_model_dims* out)
{
    if (!model || !out) return set_error(ADVHMM_EINVAL, "null argument");
    if (!model->locus.shape) return set_error(ADVHMM_EUNSUPPORTED, "only models made by advhmm_models_create_for_loci keep their tables");
    const rm::ShapeStructure& sh = *model->locus.shape;
    out->n_states = sh.m; out->silent_start = sh.S; out->start_index = sh.start; out->end_index = sh.end;
    out->finite = sh.finite; out->n_symbols = 4;
    out->n_edges = (int64_t)sh.in_src.size();
    int64_t chars = 0;
    for (const rm::Node& n : sh.states) chars += (int64_t)n.name.size() + 1;
    out->names_bytes = chars;
    out->shape[0] = sh.key.Ll; out->shape[1] = sh.key.Lr; out->shape[2] = sh.key.R; out->shape[3] = sh.key.C;
    return ADVHMM_OK;
}

int advhmm_model_tables_get(const advhmm_model* model, int32_t* in_off, int32_t* in_src, double* in_logp, double* emis,
                            char* names)
{
    if (!model) return set_error(ADVHMM_EINVAL, "null argument");
    if (!model->locus.shape) return set_error(ADVHMM_EUNSUPPORTED, "only models made by advhmm_models_create_for_loci keep their tables");
    const rm::ShapeStructure& sh = *model->locus.shape;
    if (in_off) memcpy(in_off, sh.in_off.data(), sh.in_off.size() * sizeof(int32_t));
    if (in_src) memcpy(in_src, sh.in_src.data(), sh.in_src.size() * sizeof(int32_t));
    if (in_logp || emis) {
        std::vector<double> w, e;
        rm::baked_values(model->locus, w, e);
        if (in_logp) memcpy(in_logp, w.data(), w.size() * sizeof(double));
        if (emis) memcpy(emis, e.data(), e.size() * sizeof(double));
    }
    if (names) {
        char* p = names;
        for (const rm::Node& n : sh.states) { memcpy(p, n.name.data(), n.name.size()); p += n.name.size(); *p++ = '\n'; }
    }
    return ADVHMM_OK;
}

int advhmm_model_banded_tables_get(advhmm_model* model, unsigned char* image, int64_t image_cap, int32_t* tb1, int32_t* tb0,
                                   double* fin_w, uint8_t* classes, double* logp_empty, int64_t* image_bytes)
{
    if (!model || !image_bytes) return set_error(ADVHMM_EINVAL, "null argument");
    if (model->lean) {
        const rm::ShapeStructure& sh = *model->locus.shape;
        *image_bytes = sh.image_bytes;
        if (!image) return ADVHMM_OK;
        if (image_cap < sh.image_bytes) return set_error(ADVHMM_ECAPACITY, "image buffer too small");
        rm::LeanScratch sc;
        const double e = rm::fill_lean(model->locus, sc, image, tb1, tb0, fin_w, classes);
        if (logp_empty) *logp_empty = e;
        return ADVHMM_OK;
    }
    const BandedTables& b = model->cm.b;
    if (!b.valid) return set_error(ADVHMM_EUNSUPPORTED, "not a banded model");
    std::vector<unsigned char> img;
    pack_banded_image(b, img);
    *image_bytes = (int64_t)img.size();
    if (!image) return ADVHMM_OK;
    if (image_cap < (int64_t)img.size()) return set_error(ADVHMM_ECAPACITY, "image buffer too small");
    memcpy(image, img.data(), img.size());
    if (tb1) memcpy(tb1, b.tb1.data(), b.tb1.size() * sizeof(int32_t));
    if (tb0) memcpy(tb0, model->cm.g.tb0.data(), model->cm.g.tb0.size() * sizeof(int32_t));
    if (fin_w) memcpy(fin_w, b.fin_w.data(), b.fin_w.size() * sizeof(double));
    if (classes) memset(classes, 0, (size_t)model->cm.g.m);
    if (logp_empty) *logp_empty = model->cm.g.v0[model->cm.g.end];
    return ADVHMM_OK;
}

int advhmm_model_info_get(const advhmm_model* model, advhmm_model_info* out)
{
    if (!model || !out) return set_error(ADVHMM_EINVAL, "null argument");
    *out = model->info;
    return ADVHMM_OK;
}

int advhmm_viterbi_batch(advhmm_model* model, const uint8_t* seqs, const int64_t* seq_off, int32_t n_reads,
                         uint32_t flags, double* logp, int32_t* path_len, int64_t* path_off,
                         int32_t* path, int64_t path_cap, int64_t* path_total)
{
    if (!model) return set_error(ADVHMM_EINVAL, "model is null");
    const int64_t group_off[2] = {0, n_reads};
    advhmm_model* models[1] = {model};
    return run_host(model->ctx, models, 1, group_off, seqs, seq_off, n_reads, flags & ~ADVHMM_DEVICE_BUFFERS, false,
                    logp, path_len, path_off, path, path_cap, path_total);
}

int advhmm_log_probability_batch(advhmm_model* model, const uint8_t* seqs, const int64_t* seq_off,
                                 int32_t n_reads, uint32_t flags, double* logp)
{
    if (!model) return set_error(ADVHMM_EINVAL, "model is null");
    const int64_t group_off[2] = {0, n_reads};
    advhmm_model* models[1] = {model};
    return run_host(model->ctx, models, 1, group_off, seqs, seq_off, n_reads,
                    flags & ~(ADVHMM_DEVICE_BUFFERS | ADVHMM_WANT_PATH | ADVHMM_BOTH_STRANDS), true,
                    logp, nullptr, nullptr, nullptr, 0, nullptr);
}

int advhmm_viterbi_multi(advhmm_context* ctx, advhmm_model* const* models, int32_t n_models,
                         const int64_t* group_off, const uint8_t* seqs, const int64_t* seq_off,
                         int32_t n_reads, uint32_t flags, double* logp, int32_t* path_len,
                         int64_t* path_off, int32_t* path, int64_t path_cap, int64_t* path_total)
{
    return advhmm_viterbi_multi_summary(ctx, models, n_models, group_off, seqs, seq_off, n_reads,
                                        flags & ~ADVHMM_WANT_SUMMARY, logp, path_len, path_off, path, path_cap,
                                        path_total, nullptr);
}

int advhmm_model_set_state_classes(advhmm_model* model, const uint8_t* classes)
{
    if (!model || !classes) return set_error(ADVHMM_EINVAL, "null argument");
    if (!model->ctx || model->ctx->device < 0) return ADVHMM_OK;      // host-only context: nothing to upload
    std::lock_guard<std::mutex> lock(model->ctx->mu);
    CU_TRY(cudaSetDevice(model->ctx->device));
    CU_TRY(cudaMemcpyAsync(model->d_classes, classes, (size_t)model->m, cudaMemcpyHostToDevice, model->ctx->stream));
    CU_TRY(cudaStreamSynchronize(model->ctx->stream));
    return ADVHMM_OK;
}

int advhmm_viterbi_multi_summary(advhmm_context* ctx, advhmm_model* const* models, int32_t n_models,
                                 const int64_t* group_off, const uint8_t* seqs, const int64_t* seq_off,
                                 int32_t n_reads, uint32_t flags, double* logp, int32_t* path_len,
                                 int64_t* path_off, int32_t* path, int64_t path_cap, int64_t* path_total,
                                 advhmm_read_summary* summaries)
{
    if (!ctx || !models || !group_off || n_models <= 0) return set_error(ADVHMM_EINVAL, "null argument");
    if (!(flags & ADVHMM_DEVICE_BUFFERS))
        return run_host(ctx, models, n_models, group_off, seqs, seq_off, n_reads, flags, false,
                        logp, path_len, path_off, path, path_cap, path_total, summaries);
    // device-resident buffers: asynchronous on the context's stream, nothing is copied back
    if (ctx->device < 0) return set_error(ADVHMM_ECUDA, "this context has no CUDA device");
    const bool want_path = flags & ADVHMM_WANT_PATH;
    const bool want_sum = (flags & ADVHMM_WANT_SUMMARY) != 0;
    if (!seq_off || !logp || (want_path && (!path_len || !path_off || !path_total)) ||
        (want_sum && (!summaries || !path_len || !path_off)))
        return set_error(ADVHMM_EINVAL, "null argument");
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(ctx->d_badflag.ensure(256));
    int32_t* d_bad = ctx->d_badflag.as<int32_t>();
    CU_TRY(cudaMemsetAsync(d_bad, 0x7f, sizeof(int32_t), ctx->stream));
    OutPtrs op{logp, path_len, path_off, path, want_path ? path_cap : 0,
               reinterpret_cast<unsigned long long*>(path_total), want_sum ? summaries : nullptr};
    return run_batch(ctx, models, n_models, group_off, seqs, seq_off, n_reads, flags & ~ADVHMM_DEVICE_BUFFERS, op,
                     false, d_bad);
}

int advhmm_context_bad_symbol(advhmm_context* ctx, int32_t* first_bad_read)
{
    if (!ctx || ctx->device < 0 || !first_bad_read) return set_error(ADVHMM_EINVAL, "no device context");
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU_TRY(cudaSetDevice(ctx->device));
    *first_bad_read = -1;
    if (!ctx->d_badflag.p) return ADVHMM_OK;               // no device-buffer call was made yet
    int32_t bad = 0x7f7f7f7f;
    CU_TRY(cudaMemcpyAsync(&bad, ctx->d_badflag.p, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    if (bad != 0x7f7f7f7f) *first_bad_read = bad;
    return ADVHMM_OK;
}

int advhmm_kfilter_create(advhmm_context* ctx, int64_t n_keywords, const char* keywords, const int64_t* keyword_off,
                          const int32_t* keyword_locus, advhmm_kfilter** out)
{
    if (!ctx || !out || n_keywords < 0 || (n_keywords > 0 && (!keywords || !keyword_off || !keyword_locus)))
        return set_error(ADVHMM_EINVAL, "null or negative argument");
    *out = nullptr;
    if (ctx->device < 0) return set_error(ADVHMM_ECUDA, "this context has no CUDA device");
    std::unique_ptr<advhmm_kfilter> kf(new (std::nothrow) advhmm_kfilter);
    if (!kf) return set_error(ADVHMM_ENOMEM, "out of host memory");
    kf->ctx = ctx;
    std::lock_guard<std::mutex> lock(ctx->mu);
    int rc = kfilter_build(kf.get(), n_keywords, keywords, keyword_off, keyword_locus);
    if (rc) {
        kf->tables.release();
        return rc;
    }
    *out = kf.release();
    return ADVHMM_OK;
}

void advhmm_kfilter_destroy(advhmm_kfilter* kf)
{
    if (!kf) return;
    if (kf->ctx && kf->ctx->device >= 0) {
        cudaSetDevice(kf->ctx->device);
        cudaStreamSynchronize(kf->ctx->stream);
        for (DevBuf* b : {&kf->tables, &kf->seqs, &kf->meta, &kf->tiles, &kf->counters, &kf->hits}) b->release();
    }
    delete kf;
}

int advhmm_kfilter_scan(advhmm_kfilter* kf, const char* seqs, const int64_t* seq_off, int32_t n_reads,
                        int32_t min_matches, uint32_t flags, int32_t* hit_read, int32_t* hit_locus,
                        int32_t* hit_count, int64_t hit_cap, int64_t* n_hits)
{
    if (!kf || (n_reads > 0 && !seq_off) || n_reads < 0 || !n_hits || hit_cap < 0 ||
        (hit_cap > 0 && (!hit_read || !hit_locus || !hit_count)))
        return set_error(ADVHMM_EINVAL, "null or negative argument");
    if ((flags & ADVHMM_DEVICE_OFFSETS) && !(flags & ADVHMM_DEVICE_BUFFERS))
        return set_error(ADVHMM_EINVAL, "ADVHMM_DEVICE_OFFSETS needs ADVHMM_DEVICE_BUFFERS");
    advhmm_context* ctx = kf->ctx;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CU_TRY(cudaSetDevice(ctx->device));
    const int64_t* d_off = nullptr;
    int64_t n_bases = 0;
    if (flags & ADVHMM_DEVICE_OFFSETS) {
        d_off = seq_off;
        if (n_reads > 0) {
            CU_TRY(cudaMemcpyAsync(&n_bases, seq_off + n_reads, sizeof n_bases, cudaMemcpyDeviceToHost, ctx->stream));
            CU_TRY(cudaStreamSynchronize(ctx->stream));
        }
    } else if (n_reads > 0) {
        if (seq_off[0] != 0) return set_error(ADVHMM_EINVAL, "seq_off[0] must be 0");
        for (int32_t r = 0; r < n_reads; ++r)
            if (seq_off[r + 1] < seq_off[r]) return set_error(ADVHMM_EINVAL, "seq_off must be non-decreasing");
        n_bases = seq_off[n_reads];
        if (n_bases > 0 && !seqs) return set_error(ADVHMM_EINVAL, "null read buffer");
        CU_TRY(kf->meta.ensure((size_t)(n_reads + 1) * 8));
        CU_TRY(cudaMemcpyAsync(kf->meta.p, seq_off, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        d_off = kf->meta.as<int64_t>();
    }
    if (flags & ADVHMM_DEVICE_BUFFERS)
        // seqs, hit_* and n_hits are device pointers
        return kfilter_scan_device(kf, reinterpret_cast<const unsigned char*>(seqs), d_off, n_reads, n_bases, min_matches,
                                   hit_read, hit_locus, hit_count, hit_cap, reinterpret_cast<unsigned long long*>(n_hits));
    *n_hits = 0;
    if (n_reads == 0) return ADVHMM_OK;
    CU_TRY(kf->seqs.ensure((size_t)n_bases + 16));
    if (n_bases) CU_TRY(cudaMemcpyAsync(kf->seqs.p, seqs, (size_t)n_bases, cudaMemcpyHostToDevice, ctx->stream));
    const size_t hb = ((size_t)std::max<int64_t>(hit_cap, 1) * 4 + 255) / 256 * 256;
    CU_TRY(kf->hits.ensure(3 * hb + 256));
    unsigned char* dh = kf->hits.as<unsigned char>();
    int32_t* d_r = reinterpret_cast<int32_t*>(dh);
    int32_t* d_l = reinterpret_cast<int32_t*>(dh + hb);
    int32_t* d_c = reinterpret_cast<int32_t*>(dh + 2 * hb);
    unsigned long long* d_n = reinterpret_cast<unsigned long long*>(dh + 3 * hb);
    int rc = kfilter_scan_device(kf, kf->seqs.as<unsigned char>(), d_off, n_reads, n_bases, min_matches,
                                 d_r, d_l, d_c, hit_cap, d_n);
    if (rc) return rc;
    unsigned long long total = 0;
    CU_TRY(cudaMemcpyAsync(&total, d_n, sizeof total, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    *n_hits = (int64_t)total;
    if ((int64_t)total > hit_cap)
        return set_error(ADVHMM_ECAPACITY, "hit buffers too small: need %lld entries, have %lld", (long long)total, (long long)hit_cap);
    if (total) {
        CU_TRY(cudaMemcpyAsync(hit_read, d_r, (size_t)total * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(cudaMemcpyAsync(hit_locus, d_l, (size_t)total * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(cudaMemcpyAsync(hit_count, d_c, (size_t)total * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return ADVHMM_OK;
}

// ---- the step after the path: per-read results of many loci -> genotype calls (host, all threads) ----
int advhmm_genotypes_from_summaries(int64_t n_loci, const int64_t* group_off, const int32_t* n_mapped,
                                    const int32_t* n_unmapped, const double* min_score,
                                    const double* logp, const advhmm_read_summary* summaries, const int32_t* path_len,
                                    const int64_t* seq_off, uint32_t flags, int32_t min_repeat_bp, int32_t n_threads,
                                    advhmm_locus_call* out, uint8_t* read_class)
{
    if (n_loci < 0 || (n_loci > 0 && (!group_off || !n_mapped || !n_unmapped || !logp || !summaries || !path_len || !seq_off || !out)))
        return set_error(ADVHMM_EINVAL, "null argument");
    if (flags & ~(ADVHMM_CALL_ACCURACY_FILTER | ADVHMM_CALL_HAPLOID)) return set_error(ADVHMM_EINVAL, "unknown flag");
    for (int64_t g = 0; g < n_loci; ++g) {
        if (n_mapped[g] < 0 || n_unmapped[g] < 0 || group_off[g] < 0 ||
            group_off[g + 1] - group_off[g] != (int64_t)n_mapped[g] + 2 * (int64_t)n_unmapped[g])
            return set_error(ADVHMM_EINVAL, "locus %lld: group of %lld reads does not hold %d mapped reads + 2 x %d unmapped reads",
                             (long long)g, (long long)(group_off[g + 1] - group_off[g]), n_mapped[g], n_unmapped[g]);
    }
    try {
        if (read_class && n_loci) memset(read_class + group_off[0], 0, (size_t)(group_off[n_loci] - group_off[0]));
        const calls::ReadView R{logp, summaries, path_len, seq_off};
        const bool acc = flags & ADVHMM_CALL_ACCURACY_FILTER, hap = flags & ADVHMM_CALL_HAPLOID;
        const int nt = n_threads > 0 ? n_threads : rm::default_threads();
        // loci in blocks so that a worker takes a few hundred microseconds of work per grab
        const int64_t block = 64, n_blocks = (n_loci + block - 1) / block;
        rm::parallel_for((size_t)n_blocks, nt, [&](size_t b, int) {
            for (int64_t g = (int64_t)b * block; g < std::min<int64_t>(n_loci, ((int64_t)b + 1) * block); ++g)
                calls::call_locus(R, group_off[g], n_mapped[g], n_unmapped[g], min_score ? min_score[g] : NAN, acc, hap,
                                  min_repeat_bp, out[g], read_class);
        });
    } catch (const std::bad_alloc&) {
        return set_error(ADVHMM_ENOMEM, "out of host memory");
    } catch (const std::exception& e) {
        return set_error(ADVHMM_EINVAL, "%s", e.what());
    }
    return ADVHMM_OK;
}

int advhmm_genotypes_from_counts(int64_t n_lists, const int32_t* observed, const int64_t* obs_off, uint32_t flags,
                                 advhmm_locus_call* out)
{
    if (n_lists < 0 || (n_lists > 0 && (!obs_off || !out))) return set_error(ADVHMM_EINVAL, "null argument");
    if (flags & ~(ADVHMM_CALL_ACCURACY_FILTER | ADVHMM_CALL_HAPLOID)) return set_error(ADVHMM_EINVAL, "unknown flag");
    try {
        for (int64_t i = 0; i < n_lists; ++i) {
            if (obs_off[i + 1] < obs_off[i]) return set_error(ADVHMM_EINVAL, "obs_off must not decrease");
            std::vector<int32_t> v(observed + obs_off[i], observed + obs_off[i + 1]);
            if (flags & ADVHMM_CALL_ACCURACY_FILTER) v = calls::drop_unsupported(v);
            const calls::Genotype g = calls::genotype_from_observed(v.data(), v.size(), flags & ADVHMM_CALL_HAPLOID);
            out[i] = advhmm_locus_call{g.found ? 1 : 0, g.c1, g.c2, (int32_t)(obs_off[i + 1] - obs_off[i]), (int32_t)v.size(), 0,
                                       g.max_prob};
        }
    } catch (const std::bad_alloc&) {
        return set_error(ADVHMM_ENOMEM, "out of host memory");
    }
    return ADVHMM_OK;
}

int advhmm_frameshift_candidates(int64_t n_loci, const int64_t* group_off, const int32_t* pattern_len, const double* min_score,
                                 const int64_t* state_off, const uint8_t* state_class, const int32_t* state_label,
                                 const double* logp, const advhmm_read_summary* summaries, const int32_t* path_len,
                                 const int64_t* path_off, const int32_t* path, const uint8_t* seqs, const int64_t* seq_off,
                                 int32_t n_threads, advhmm_frameshift_call* out)
{
    if (n_loci < 0 || (n_loci > 0 && (!group_off || !pattern_len || !state_off || !state_class || !state_label || !logp ||
                                      !summaries || !path_len || !path_off || !path || !seqs || !seq_off || !out)))
        return set_error(ADVHMM_EINVAL, "null argument");
    for (int64_t g = 0; g < n_loci; ++g)
        if (group_off[g] < 0 || group_off[g + 1] < group_off[g] || state_off[g + 1] < state_off[g])
            return set_error(ADVHMM_EINVAL, "locus %lld: offsets must not decrease", (long long)g);
    std::atomic<int64_t> bad{-1};
    try {
        const calls::ReadView R{logp, summaries, path_len, seq_off};
        const int nt = n_threads > 0 ? n_threads : rm::default_threads();
        const int64_t block = 16, n_blocks = (n_loci + block - 1) / block;
        rm::parallel_for((size_t)n_blocks, nt, [&](size_t b, int) {
            std::vector<calls::Mutation> mut;
            std::vector<int32_t> lengths;
            std::vector<std::pair<int32_t, int32_t>> first_visit;
            for (int64_t g = (int64_t)b * block; g < std::min<int64_t>(n_loci, ((int64_t)b + 1) * block); ++g) {
                mut.clear();
                const int64_t n_states = state_off[g + 1] - state_off[g];
                advhmm_frameshift_call c{};
                c.base = -1;
                for (int64_t i = group_off[g]; i < group_off[g + 1]; ++i) {
                    if (!calls::recruit_read(R, i, min_score ? min_score[g] : NAN, calls::flank_rate(summaries[i]))) continue;
                    ++c.selected;
                    c.repeat_bp += summaries[i].repeat_bp;
                    const int32_t* p = path + path_off[i];
                    int64_t emitted = 0;                               // a path must be one of ITS read on ITS model
                    for (int32_t k = 0; k < path_len[i]; ++k) {
                        if (p[k] < 0 || p[k] >= n_states) { bad.store(i); return; }
                        const int kind = state_class[state_off[g] + p[k]] & 7;
                        emitted += (k > 0 && k + 1 < path_len[i]) && (kind == 1 || kind == 2);
                    }
                    if (emitted != seq_off[i + 1] - seq_off[i]) { bad.store(i); return; }
                    calls::frameshift_mutations_of_read(
                        calls::PathView{p, path_len[i], state_class + state_off[g], state_label + state_off[g], seqs + seq_off[i]},
                        pattern_len[g], mut, lengths, first_visit);
                }
                if (const calls::Mutation* m = calls::frameshift_candidate(mut)) {
                    c.kind = m->kind; c.column = m->column; c.base = m->base; c.count = m->count;
                }
                out[g] = c;
            }
        });
    } catch (const std::bad_alloc&) {
        return set_error(ADVHMM_ENOMEM, "out of host memory");
    } catch (const std::exception& e) {
        return set_error(ADVHMM_EINVAL, "%s", e.what());
    }
    if (bad.load() >= 0) return set_error(ADVHMM_EINVAL, "read %lld: its path is not a path of this read on its locus's model (state index or emitted length)", (long long)bad.load());
    return ADVHMM_OK;
}

}  // extern "C"
