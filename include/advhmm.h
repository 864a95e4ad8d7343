/*
 * advhmm.h -- C-ABI of the B200 Viterbi / forward engine for adVNTR's profile HMMs.
 *
 * This is the drop-in boundary of the hot path.  The reference has no C-level plugin
 * ABI: its engine is the private Cython methods of the vendored pomegranate
 * (/root/reference/pomegranate/hmm.pyx).  Each entry point below names the reference
 * interface it replaces; INTEGRATION.md shows the ctypes stub a reference maintainer
 * would add to pomegranate's `HiddenMarkovModel` to bind them.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every host buffer it passes;
 *   - the library owns all device memory tied to a context / model handle; per-context
 *     workspaces grow, they are never allocated per call;
 *   - every function returns ADVHMM_OK (0) or a negative ADVHMM_E* code;
 *     advhmm_last_error() gives the thread-local message of the last failure;
 *   - a handle may be used from one host thread at a time; one CUDA stream per context;
 *   - multi-GPU = one context (and one copy of each model) per device, no collective.
 *   - symbols are small integer codes 0..n_symbols-1 (A,C,G,T = 0,1,2,3 for DNA), one
 *     byte per base in host/device buffers; the engine packs them 2-bit internally.
 */
#ifndef ADVHMM_H_
#define ADVHMM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADVHMM_ABI_VERSION 1

/* error codes */
#define ADVHMM_OK             0
#define ADVHMM_EINVAL        -1  /* bad argument / malformed model                        */
#define ADVHMM_ECUDA         -2  /* CUDA runtime failure (message has the CUDA error)     */
#define ADVHMM_ENOMEM        -3  /* host or device allocation failed                      */
#define ADVHMM_ESYMBOL       -4  /* a read contains a code >= n_symbols                   */
#define ADVHMM_ECAPACITY     -5  /* caller's path buffer too small (see path_total)       */
#define ADVHMM_EUNSUPPORTED  -6  /* feature not available for this model / precision      */

/* flags for the decoding calls */
#define ADVHMM_WANT_PATH      0x1u  /* backtrack on device and return state paths          */
#define ADVHMM_BOTH_STRANDS   0x2u  /* also decode the reverse complement of every read;   */
                                    /* results are interleaved: 2*i forward, 2*i+1 revcomp */
#define ADVHMM_FP32           0x4u  /* fp32 DP arithmetic (tolerance documented in DESIGN) */
#define ADVHMM_WANT_SUMMARY   0x10u /* fill advhmm_read_summary per read on the device         */
#define ADVHMM_FORCE_GENERIC  0x8u  /* use the generic CSR kernel even if the model is     */
                                    /* banded (testing / cross-checking)                   */

/* kernel families a model can be compiled to (advhmm_model_info.kind) */
#define ADVHMM_KIND_GENERIC   0     /* any baked model: row-synchronous CSR kernel         */
#define ADVHMM_KIND_BANDED    1     /* profile-shaped model: register wavefront kernel     */

typedef struct advhmm_context advhmm_context;
typedef struct advhmm_model   advhmm_model;

/*
 * A baked model, exactly the arrays pomegranate's bake() produces
 * (hmm.pyx:844-1123): states 0..silent_start-1 emit, silent_start..n_states-1 are
 * silent and topologically ordered; in-edges of state l are
 * in_src/in_logp[in_off[l] .. in_off[l+1]) in the reference's in-edge order (it
 * decides Viterbi tie-breaking); emis[l*n_symbols + code] is the emission
 * log-probability (+ log state weight) of emitting state l.
 */
typedef struct advhmm_model_desc {
    int32_t        n_states;
    int32_t        silent_start;
    int32_t        start_index;
    int32_t        end_index;
    int32_t        finite;       /* 1: logp = v[n][end_index]; 0: best state (hmm.pyx:2089-2098) */
    int32_t        n_symbols;    /* 1..4: reads are packed 2 bits per symbol; one alphabet size per call */
    const int32_t* in_off;       /* [n_states + 1] */
    const int32_t* in_src;       /* [in_off[n_states]] */
    const double*  in_logp;      /* [in_off[n_states]] */
    const double*  emis;         /* [silent_start * n_symbols] */
} advhmm_model_desc;

typedef struct advhmm_model_info {
    int32_t kind;            /* ADVHMM_KIND_*                                           */
    int32_t n_states;
    int32_t n_edges;         /* edges the DP evaluates (dead silent->silent edges dropped) */
    int32_t n_columns;       /* banded: profile columns; generic: silent levels         */
    int32_t n_final_states;  /* banded: silent states evaluated only on the last row    */
    int32_t smem_bytes;      /* shared memory the model tables occupy per CTA           */
    int32_t max_in_degree;
    int32_t reserved;
} advhmm_model_info;

/* ---- contexts ------------------------------------------------------------------------- */

/* Create an engine context on CUDA device `device`.  `stream` is a cudaStream_t to launch
 * on (e.g. torch's current stream) or NULL to let the context create its own. */
int  advhmm_context_create(int device, void* stream, advhmm_context** out);
void advhmm_context_destroy(advhmm_context* ctx);
/* Block until everything queued on the context's stream has finished. */
int  advhmm_context_synchronize(advhmm_context* ctx);
/* The cudaStream_t the context launches on (for CUDA-event timing by the caller). */
void* advhmm_context_stream(advhmm_context* ctx);
/* Number of kernels this context has launched since creation (bench `gpu_launches`). */
int64_t advhmm_context_launch_count(advhmm_context* ctx);

/* Per-kernel device timing for roofline reports: when enabled, every banded fill / backtrack
 * launch is bracketed by CUDA events on the context's stream.  profile_read() synchronises,
 * returns the summed durations (ms) and launch counts since the last read, and resets. */
int  advhmm_context_profile(advhmm_context* ctx, int enable);
int  advhmm_context_profile_read(advhmm_context* ctx, double* fill_ms, int64_t* fill_launches,
                                 double* backtrack_ms, int64_t* backtrack_launches);
/* Measured fp64 add issue rate of this device in Gop/s (lane-operations): the denominator of
 * the fill kernel's compute roofline (the DP is fp64 add + compare, no FMA, no tensor cores). */
int  advhmm_fp64_add_peak(advhmm_context* ctx, double* gops);

/* ---- models ---------------------------------------------------------------------------
 * Replaces the tail of HiddenMarkovModel.bake() (hmm.pyx:932-1023: building the C arrays
 * the DP kernels read) -- the host analyses the graph once (silent levels / profile
 * columns, row-0 closure, first-row tables) and uploads the tables. */
int  advhmm_model_create(advhmm_context* ctx, const advhmm_model_desc* desc, advhmm_model** out);
void advhmm_model_destroy(advhmm_model* model);
int  advhmm_model_info_get(const advhmm_model* model, advhmm_model_info* out);

/* ---- models of many loci in one call ------------------------------------------------------
 * Replaces, for a batch of loci, what the reference does per locus before it can decode a read:
 * VNTRFinder.get_vntr_matcher_hmm -> build_vntr_matcher_hmm -> hmm_utils.get_read_matcher_model
 * (vntr_finder.py:108-138, hmm_utils.py:289-595, profile_hmm.py:13-161: three sub-models, eight
 * bake() calls and two dense matrix round trips in Python) plus the tail of bake().  The library
 * evaluates the repeat-unit profile of every locus, takes the structure of its shape (flank
 * lengths, match columns, copies) from a per-process cache (built once by the same sequence of
 * graph operations as the reference), evaluates the parameters through the reference's chain of
 * log / exp calls and writes the device tables of all loci with all host threads; one upload.
 * Tables are bit-identical to a model of the reference's builders handed to advhmm_model_create
 * (tests/test_native_compile.py) PROVIDED the vector exp the reference uses is set: the reference
 * applies numpy.exp (hmm.pyx:514), whose SIMD kernels differ from libm's exp in the last bit of
 * ~3 % of the values; advhmm_set_vexp(NULL) = libm (the Python binding installs numpy.exp).
 *
 * Columns of advhmm_loci (locus i):
 *   left / right     flank bases that enter the model, codes 0..3: left[left_off[i] .. left_off[i+1])
 *                    = the last flank_size bases before the repeats, right = the first ones after
 *                    (vntr_finder.py:111-112, :131)
 *   segments         the aligned repeat segments, n_segments[i] rows of equal width over "ACGT-",
 *                    row-major from segments[seg_off[i]] (equal-length segments are their own
 *                    alignment; profile_hmm.py:165-171 runs MUSCLE otherwise -- pass its output)
 *   copies           unrolled copies of the repeat unit (vntr_finder.py:98-99)
 *   error_rate       settings.MAX_ERROR_RATE (0.05 Illumina, 0.3 PacBio / nanopore)
 * out[n_loci] receives the handles (destroy each with advhmm_model_destroy).  n_threads <= 0: all
 * cores of the calling process.  State classes for the on-device path reducers are set. */
typedef struct advhmm_loci {
    int32_t        n_loci;
    const uint8_t* left;        const int64_t* left_off;    /* [n_loci + 1] */
    const uint8_t* right;       const int64_t* right_off;   /* [n_loci + 1] */
    const char*    segments;    const int64_t* seg_off;     /* [n_loci + 1] */
    const int32_t* n_segments;  /* [n_loci] */
    const int32_t* copies;      /* [n_loci] */
    const double*  error_rate;  /* [n_loci] */
} advhmm_loci;

typedef void (*advhmm_vexp_fn)(const double* in, double* out, int64_t n, void* user);
int advhmm_set_vexp(advhmm_vexp_fn fn, void* user);
int advhmm_models_create_for_loci(advhmm_context* ctx, const advhmm_loci* loci, int32_t n_threads, advhmm_model** out);
/* Forget the cached shape structures (cold-start measurements). */
void advhmm_shape_cache_clear(void);

/* The baked arrays of a model made by advhmm_models_create_for_loci, in the layout of
 * advhmm_model_desc (what pomegranate's bake() would hold for it), and its state names ('\n' after
 * each, names_bytes in all): the Python wrapper builds its State list from them, the tests compare
 * them with the reference's builders.  Any output pointer may be NULL. */
typedef struct advhmm_model_dims {
    int32_t n_states, silent_start, start_index, end_index, finite, n_symbols;
    int64_t n_edges, names_bytes;
    int32_t shape[4];        /* left flank, right flank, match columns, copies */
} advhmm_model_dims;
int advhmm_model_dims_get(const advhmm_model* model, advhmm_model_dims* out);
int advhmm_model_tables_get(const advhmm_model* model, int32_t* in_off, int32_t* in_src, double* in_logp, double* emis,
                            char* names);
/* What the banded kernels read for this model, as host copies (tests: a locus model and a
 * descriptor-made model of the same tables must agree byte for byte).  image == NULL: only
 * *image_bytes is set.  tb1[4 * silent_start], tb0[n_states], classes[n_states], fin_w as many as the
 * model has final-state edges (<= n_edges). */
int advhmm_model_banded_tables_get(advhmm_model* model, unsigned char* image, int64_t image_cap, int32_t* tb1, int32_t* tb0,
                                   double* fin_w, uint8_t* classes, double* logp_empty, int64_t* image_bytes);

/* ---- decoding, host buffers -------------------------------------------------------------
 * Replaces HiddenMarkovModel.viterbi / _viterbi (hmm.pyx:1911-2136) for a batch of reads
 * of ONE model.  seqs holds the reads back to back, read r = seqs[seq_off[r] .. seq_off[r+1]).
 * n_out = n_reads * (BOTH_STRANDS ? 2 : 1) results are produced:
 *   logp[n_out]      Viterbi log-probability, -inf for an impossible read (hmm.pyx:1967)
 *   path_len[n_out]  number of states on the path (silent ones included), -1 if impossible
 *   path_off[n_out]  start of that read's path inside `path`
 *   path[path_cap]   state indices in the baked order, concatenated (order unspecified)
 *   *path_total      total entries written (or needed, with ADVHMM_ECAPACITY)
 * path*, path_total may be NULL when ADVHMM_WANT_PATH is not set. */
int advhmm_viterbi_batch(advhmm_model* model,
                         const uint8_t* seqs, const int64_t* seq_off, int32_t n_reads,
                         uint32_t flags,
                         double* logp, int32_t* path_len, int64_t* path_off,
                         int32_t* path, int64_t path_cap, int64_t* path_total);

/* Replaces HiddenMarkovModel.log_probability / _forward (hmm.pyx:1258-1313, 1371-1484). */
int advhmm_log_probability_batch(advhmm_model* model,
                                 const uint8_t* seqs, const int64_t* seq_off, int32_t n_reads,
                                 uint32_t flags, double* logp);

/* ---- decoding, many loci in one launch ---------------------------------------------------
 * The batched form of the per-locus loop (genome_analyzer.py:280 x vntr_finder.py:727-767):
 * reads group_off[g] .. group_off[g+1]-1 are decoded against models[g] (the caller loops over
 * loci anyway, so reads arrive grouped).  All models must belong to `ctx`.
 * group_off[n_models+1] and seq_off[n_reads+1] are always HOST arrays (planning metadata).
 * With ADVHMM_DEVICE_BUFFERS every other buffer (seqs, logp, path_len, path_off, path,
 * path_total) is a DEVICE pointer on ctx's device: nothing is copied, the call only queues
 * work on the context's stream and returns; *path_total is a device int64; reads whose path
 * did not fit in path_cap get path_len = -2.  A code >= n_symbols cannot be reported by the return
 * value of an asynchronous call: such a read is decoded with the bad code read as symbol 0 and the
 * index of the first such read is kept on the device; advhmm_context_bad_symbol() fetches it (the host
 * path returns ADVHMM_ESYMBOL, the reference raises ValueError, hmm.pyx:72-79). */
#define ADVHMM_DEVICE_BUFFERS 0x100u
int advhmm_viterbi_multi(advhmm_context* ctx,
                         advhmm_model* const* models, int32_t n_models,
                         const int64_t* group_off,
                         const uint8_t* seqs, const int64_t* seq_off, int32_t n_reads,
                         uint32_t flags,
                         double* logp, int32_t* path_len, int64_t* path_off,
                         int32_t* path, int64_t path_cap, int64_t* path_total);

/* After an ADVHMM_DEVICE_BUFFERS call: synchronise the context's stream and report the first read of
 * that call holding a code >= n_symbols (*first_bad_read = -1: none). */
int advhmm_context_bad_symbol(advhmm_context* ctx, int32_t* first_bad_read);

/* ---- on-device path reducers ----------------------------------------------------------------
 * What adVNTR derives from a Viterbi path (hmm_utils.py:155-286), computed by the backtrack
 * kernel so that the host does not have to walk state names read by read.  The state classes
 * are what those functions parse out of the state NAMES; the host sets them once per model:
 *   bits 0-2 kind: 0 other, 1 match (name starts with 'M'), 2 insert ('I'), 3 delete ('D'),
 *                  4 unit_start*, 5 unit_end*
 *   bits 3-4 part: 1 name ends with 'suffix', 2 ends with 'prefix', 3 repeat unit, 0 n/a
 *   bits 5-6      the flank base (0..3) a suffix/prefix match state expects
 * repeats = get_number_of_repeats_in_vpath; n_match = get_number_of_matches_in_vpath;
 * repeat_bp = get_number_of_repeat_bp_matches_in_vpath; left_bp / right_bp =
 * get_left/right_flanking_region_size_in_vpath; left_hits / right_hits = the numerators of
 * get_flanking_regions_matching_rate.  repeats = -1 for an impossible read. */
typedef struct advhmm_read_summary {
    int32_t repeats, n_match, repeat_bp, left_bp, right_bp, left_hits, right_hits, unit_starts_ends;
} advhmm_read_summary;

int advhmm_model_set_state_classes(advhmm_model* model, const uint8_t* classes /* [n_states] */);

/* advhmm_viterbi_multi plus summaries[n_out] (host or device buffer like the others).  With
 * ADVHMM_WANT_SUMMARY and without ADVHMM_WANT_PATH no path is written (path may be NULL), only
 * path_len is filled. */
int advhmm_viterbi_multi_summary(advhmm_context* ctx,
                                 advhmm_model* const* models, int32_t n_models,
                                 const int64_t* group_off,
                                 const uint8_t* seqs, const int64_t* seq_off, int32_t n_reads,
                                 uint32_t flags,
                                 double* logp, int32_t* path_len, int64_t* path_off,
                                 int32_t* path, int64_t path_cap, int64_t* path_total,
                                 advhmm_read_summary* summaries);

/* ---- from the per-read results of many loci to their genotype calls (the step after the path) ----
 * Replaces, for whole batches of loci and on all host threads, what VNTRFinder does with the Viterbi
 * results of a locus's reads: recruit_read (vntr_finder.py:179-190), the better strand of every filtered
 * unmapped read (:235-254), the spanning test (:311-322), spanning + flanking repeat counts (:846-875)
 * and find_genotype_based_on_observed_repeats (:486-532).  Host code: the inputs are what
 * advhmm_viterbi_multi_summary delivered (host arrays).  Locus g owns reads group_off[g] ..
 * group_off[g+1]-1: first its n_mapped[g] mapped reads, then BOTH strands (forward, reverse complement)
 * of its n_unmapped[g] keyword-filtered unmapped reads.  min_score[g] = the locus's minimum Viterbi
 * score (scaled_score * read length, vntr_finder.py:166-177) or NaN when none is known (recruit_read's
 * fallback rule).  max_prob is the reference's float (same factors, same order, libm pow).
 * read_class (optional, [n_reads]): 0 not recruited, 1 spanning read, 2 flanking read. */
#define ADVHMM_CALL_ACCURACY_FILTER 0x1u   /* settings / --accuracy-filter: drop counts seen < 3 times, no flanking reads */
#define ADVHMM_CALL_HAPLOID         0x2u   /* --haploid */
typedef struct advhmm_locus_call {
    int32_t has_call;        /* 0: nothing observed (the reference returns None)                  */
    int32_t c1, c2;          /* the genotype's two copy numbers (c2 = 0: one allele observed)     */
    int32_t recruited;       /* reads that passed recruit_read                                     */
    int32_t spanning;        /* of those, spanning reads that entered the call                    */
    int32_t flanking;        /* flanking reads                                                     */
    double  max_prob;        /* posterior of the call (1e-20 when has_call = 0)                    */
} advhmm_locus_call;
int advhmm_genotypes_from_summaries(int64_t n_loci, const int64_t* group_off, const int32_t* n_mapped,
                                    const int32_t* n_unmapped, const double* min_score,
                                    const double* logp, const advhmm_read_summary* summaries, const int32_t* path_len,
                                    const int64_t* seq_off, uint32_t flags, int32_t min_repeat_bp, int32_t n_threads,
                                    advhmm_locus_call* calls, uint8_t* read_class);
/* find_genotype_based_on_observed_repeats (vntr_finder.py:486-532) for count lists: list i =
 * observed[obs_off[i] .. obs_off[i+1]) (PacBio spanning reads, :566-585: with ADVHMM_CALL_ACCURACY_FILTER
 * counts seen < 3 times are dropped first). */
int advhmm_genotypes_from_counts(int64_t n_lists, const int32_t* observed, const int64_t* obs_off, uint32_t flags,
                                 advhmm_locus_call* calls);

/* --frameshift mode (genome_analyzer.py:260, vntr_finder.py:776-780): find_frameshift_from_selected_reads up
 * to its binomial test (vntr_finder.py:265-300) for many loci on all host threads, from the FULL state paths
 * advhmm_viterbi_multi_summary returned with ADVHMM_WANT_PATH | ADVHMM_WANT_SUMMARY (host arrays).  Per locus:
 * every read of group_off[g] .. group_off[g+1]-1 that passes recruit_read is walked; insert / delete states of
 * repeat units whose emitted length is off by 1 or 2 bases from pattern_len[g] are counted per (state label,
 * inserted base); the candidate is the most frequent one (the last one found among equals, as the reference's
 * stable sort leaves it).  States are described by their class byte (as for advhmm_model_set_state_classes)
 * and, for the I / D states of the repeat units, the number in the state's name ("I7_2" -> 7): state s of
 * locus g is state_class / state_label[state_off[g] + s].  seqs = the reads as they were decoded (codes). */
typedef struct advhmm_frameshift_call {
    int32_t kind;            /* 0: no candidate, 2: insert state, 3: delete state (class-byte kinds)         */
    int32_t column;          /* the number in the candidate state's name                                      */
    int32_t base;            /* insert states: code of the base inserted at the first visit; otherwise -1     */
    int32_t count;           /* occurrences over the recruited reads                                          */
    int32_t selected;        /* reads that passed recruit_read                                                */
    int32_t reserved;
    int64_t repeat_bp;       /* repeating base pairs in the recruited reads (the coverage of the test)       */
} advhmm_frameshift_call;
int advhmm_frameshift_candidates(int64_t n_loci, const int64_t* group_off, const int32_t* pattern_len, const double* min_score,
                                 const int64_t* state_off, const uint8_t* state_class, const int32_t* state_label,
                                 const double* logp, const advhmm_read_summary* summaries, const int32_t* path_len,
                                 const int64_t* path_off, const int32_t* path, const uint8_t* seqs, const int64_t* seq_off,
                                 int32_t n_threads, advhmm_frameshift_call* calls);

/* ---- keyword pre-filter (the step before the hot path) ----------------------------------------
 * Replaces the Aho-Corasick scan of the `adVNTR-Filtering` binary (filtering/main.cc:229-300,
 * fed by genome_analyzer.py:173-197): count, for every read and locus, the occurrences of the
 * locus's keywords in the read.  Keyword i = keywords[keyword_off[i] .. keyword_off[i+1]), any
 * length from 1 to 4096, at most 16 distinct lengths per filter (adVNTR writes 15-mers for short
 * reads and two 80-mers per locus for long reads, vntr_finder.py:140-154); keywords and reads are
 * ASCII, anything but upper-case ACGT is the reference's fifth symbol (main.cc:43-54).
 * keyword_locus[i] is any caller-chosen int32 id of the locus that owns keyword i (the same keyword
 * may be listed for several loci; listing it twice for one locus counts twice, as two words of the
 * reference's machine would).
 * scan(): reads back to back, read r = seqs[seq_off[r] .. seq_off[r+1]).  Returns the triples
 * (read index, locus id, occurrences) with occurrences >= min_matches, in unspecified order;
 * *n_hits = number of triples (or needed, with ADVHMM_ECAPACITY).  The per-locus cap / ordering /
 * output format of main.cc:283-332 is host logic (advntr_b200/keyword_filter.py).
 * With ADVHMM_DEVICE_BUFFERS seqs, hit_* and n_hits are device pointers (seq_off stays a host
 * array unless ADVHMM_DEVICE_OFFSETS is set as well: then it is a device array too and nothing but
 * eight bytes crosses the bus); seqs must be 16-byte aligned and readable up to the next multiple
 * of 16 bytes. */
#define ADVHMM_DEVICE_OFFSETS 0x200u
typedef struct advhmm_kfilter advhmm_kfilter;
int  advhmm_kfilter_create(advhmm_context* ctx, int64_t n_keywords, const char* keywords,
                           const int64_t* keyword_off, const int32_t* keyword_locus, advhmm_kfilter** out);
void advhmm_kfilter_destroy(advhmm_kfilter* kf);
int  advhmm_kfilter_scan(advhmm_kfilter* kf, const char* seqs, const int64_t* seq_off, int32_t n_reads,
                         int32_t min_matches, uint32_t flags,
                         int32_t* hit_read, int32_t* hit_locus, int32_t* hit_count,
                         int64_t hit_cap, int64_t* n_hits);

/* ---- misc -------------------------------------------------------------------------------- */
const char* advhmm_last_error(void);
int         advhmm_abi_version(void);
/* Encode ASCII DNA to codes A,C,G,T -> 0..3 (both cases).  Returns -1 on success or the index
 * of the first byte that is not ACGT (the wrapper raises the reference's ValueError for it,
 * hmm.pyx:72-79). */
int64_t     advhmm_encode_acgt(const char* ascii, int64_t n, uint8_t* codes);

#ifdef __cplusplus
}
#endif
#endif /* ADVHMM_H_ */
