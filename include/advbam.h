/* advbam.h -- C ABI of the read-ingest side of the B200 adVNTR engine (libadvbam.so).
 *
 * What it replaces in the reference: the pysam calls that feed the
 * Viterbi hot path,
 *
 *   pysam.AlignmentFile(path, 'rb')                      vntr_finder.py:709, :453
 *   samfile.references                                   sam_utils.py:32-39
 *   samfile.head(5)                                      vntr_finder.py:716
 *   samfile.fetch(chromosome, vntr_start, vntr_end)      vntr_finder.py:727, :457
 *   the per-read tests of select_illumina_reads          vntr_finder.py:728-737 + utils.py:20-38
 *   the CIGAR walk of check_if_pacbio_mapped_read_spans_vntr   vntr_finder.py:373-420
 *   samtools view -f4 | samtools bam2fq | fastq->fasta   sam_utils.py:9-23
 *
 * evaluated on whole batches of records in native code: BGZF blocks are inflated (zlib) from the
 * memory-mapped file -- in parallel for whole-file scans --, records are parsed into columns, the
 * read-level decisions are taken on the columns, and the surviving reads are written as the 0..3 codes
 * advhmm_viterbi_multi takes.  File formats follow hts-specs SAMv1 (section 4 BAM, 4.1 BGZF, 5.2 BAI).
 * Host-only: no CUDA in this library; SAM text and CRAM are not read.
 *
 * Conventions: every function returns 0 or a negative code and leaves a message for
 * advbam_last_error() (per thread).  The caller owns the buffers it passes; the library owns handles.
 */
#ifndef ADVBAM_H
#define ADVBAM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADVBAM_OK 0
#define ADVBAM_E_ARG (-1)
#define ADVBAM_E_IO (-2)
#define ADVBAM_E_FORMAT (-3)
#define ADVBAM_E_INDEX (-4)

typedef struct advbam_file advbam_file;
typedef struct advbam_reads advbam_reads;

const char* advbam_last_error(void);

/* pysam.AlignmentFile(bam_path, 'rb'[, index_filename=bai_path]).  bai_path NULL: "<bam>.bai", then
 * "<bam without .bam>.bai"; a missing index is only an error when advbam_fetch is called. */
int advbam_open(const char* bam_path, const char* bai_path, advbam_file** out);
void advbam_close(advbam_file* f);

/* samfile.references / samfile.lengths / samfile.get_tid(name) (-1 when absent) */
int32_t advbam_n_references(const advbam_file* f);
const char* advbam_reference_name(const advbam_file* f, int32_t tid);
int64_t advbam_reference_length(const advbam_file* f, int32_t tid);
int32_t advbam_reference_id(const advbam_file* f, const char* name);

/* samfile.head(n): the first n records of the file. */
int advbam_head(advbam_file* f, int32_t n, advbam_reads** out);

/* samfile.fetch(reference, beg, end): records of `tid` with pos < end and bam_endpos > beg, in file
 * order (0-based half-open region; needs the index). */
int advbam_fetch(advbam_file* f, int32_t tid, int64_t beg, int64_t end, advbam_reads** out);

/* samtools view -f require -F exclude over the whole file (parallel inflate, n_threads <= 0: all cores). */
int advbam_scan(advbam_file* f, uint32_t require_flags, uint32_t exclude_flags, int32_t n_threads,
                advbam_reads** out);

/* In place: what `samtools bam2fq` prints for these records (sam_utils.py:20): reverse-strand records
 * are reverse-complemented (qualities reversed), names get "/1" or "/2" from the READ1/READ2 flags.
 * Secondary and supplementary records must have been excluded by the scan (-F 0x900). */
int advbam_reads_to_fastq_orientation(advbam_reads* r);

void advbam_reads_free(advbam_reads* r);

/* Columns of a batch; pointers stay valid until advbam_reads_free.  Sequences are ASCII as pysam's
 * read.seq ("=ACMGRSVTWYHKDBN"), qualities are phred values (read.query_qualities); has_qual[i] == 0
 * when the record stores none (0xff).  ref_end is read.reference_end, -1 for None (unmapped or no
 * CIGAR).  cigar holds op | len << 4 words (the CG:B,I tag replaces a long-CIGAR placeholder). */
typedef struct advbam_view {
    int64_t n;
    const uint16_t* flag;
    const uint8_t* mapq;
    const int32_t* tid;
    const int32_t* pos;
    const int32_t* ref_end;
    const uint8_t* has_qual;
    const int64_t* seq_off;  /* n + 1 */
    const char* seq;
    const uint8_t* qual;     /* same offsets as seq */
    const int64_t* name_off; /* n + 1 */
    const char* names;
    const int64_t* cigar_off; /* n + 1 */
    const uint32_t* cigar;
} advbam_view;
int advbam_reads_view(const advbam_reads* r, advbam_view* out);

/* The per-read tests of select_illumina_reads (vntr_finder.py:728-737) and is_low_quality_read
 * (utils.py:20-38) for every record of a fetch; decision[i] is one of the codes below and *vntr_bp is
 * vntr_bp_in_mapped_reads (:751-753). */
#define ADVBAM_DECODE 0           /* goes to Viterbi */
#define ADVBAM_SKIP_FLAGS 1       /* unmapped or duplicate (:728) */
#define ADVBAM_SKIP_SHORT 2       /* len(seq) < min_read_length (:731) */
#define ADVBAM_SKIP_REGION 3      /* fails the position test (:735) */
#define ADVBAM_SKIP_N 4           /* contains N (:736) */
#define ADVBAM_SKIP_LOW_QUALITY 5 /* is_low_quality_read (:739); the reference decodes it, then drops it */
#define ADVBAM_BAD_SYMBOL 6       /* a base outside ACGTN: hmm.viterbi raises ValueError in the reference */
#define ADVBAM_NO_QUALITIES 7     /* query_qualities is None: TypeError in the reference (utils.py:24) */
typedef struct advbam_illumina_params {
    int64_t vntr_start, vntr_end;
    int32_t read_length;        /* median of head(5), vntr_finder.py:714-718 */
    int32_t min_read_length;    /* int(read_length * 0.9) or settings.MIN_READ_LENGTH */
    int32_t mapq_cutoff;        /* settings.MAPQ_CUTOFF = 0 */
    int32_t quality_cutoff;     /* settings.QUALITY_SCORE_CUTOFF = 20 */
    double low_quality_fraction; /* settings.LOW_QUALITY_BP_TO_DISCARD_READ = 0.10 */
} advbam_illumina_params;
int advbam_select_illumina(const advbam_reads* r, const advbam_illumina_params* p, uint8_t* decision,
                           int64_t* vntr_bp);

/* Codes 0..3 (A C G T) of the records with decision[i] == ADVBAM_DECODE, concatenated, with their
 * offsets (n_selected + 1) and record indices: the layout advhmm_viterbi_multi takes.  decision NULL:
 * all records.  Sizes: call once with codes == NULL to get *n_selected and *n_codes. */
int advbam_gather_codes(const advbam_reads* r, const uint8_t* decision, uint8_t* codes, int64_t* off,
                        int64_t* index, int64_t* n_selected, int64_t* n_codes);

/* check_if_pacbio_mapped_read_spans_vntr (vntr_finder.py:373-420) for every record: seg_start[i] < 0
 * when the read is rejected, else the reference's read.seq[seg_start : seg_end] (already clamped to
 * the read) with left_bp / right_bp flanking bases inside it.  Records without aligned positions
 * (:458) are rejected. */
int advbam_spanning_segments(const advbam_reads* r, int64_t vntr_start, int64_t vntr_end,
                             int32_t hmm_flank, int32_t min_flank_bp, int64_t* seg_start, int64_t* seg_end,
                             int32_t* left_bp, int32_t* right_bp);

#ifdef __cplusplus
}
#endif
#endif
