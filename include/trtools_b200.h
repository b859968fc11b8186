/*
 * trtools_b200 — C-ABI of the B200-native TRTools hot path.
 *
 * The reference (gymrek-lab/TRTools v6.1.0) is pure Python and has no FFI; its seams for this
 * path are Python protocols.  This header is what a ctypes binding on the
 * reference side would bind (INTEGRATION.md shows the stub): one shared object,
 * libtrtools_b200.so, built with nvcc for sm_100a.  Every entry point cites the reference
 * function(s) it replaces as  file:line  relative to the reference tree.
 *
 * Conventions
 *   - every function returns TRT_OK (0) or a negative TRT_E* code; the message is available from
 *     trt_last_error(ctx) (or trt_last_error(NULL) for failures of trt_init itself);
 *   - no exceptions cross the ABI; no torch types; plain pointers and sizes only;
 *   - the caller owns all host buffers (the library never retains them past the call unless
 *     they came from trt_host_alloc); the context owns device buffers and streams;
 *   - a context is bound to one CUDA device and is not thread-safe: one context per GPU,
 *     driven by one host thread/process;
 *   - there is NO CPU fallback: without a CUDA device trt_init fails with TRT_ENODEV.
 */
#ifndef TRTOOLS_B200_H
#define TRTOOLS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRT_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------------------------ */
#define TRT_OK          0
#define TRT_ENODEV     -1   /* no usable CUDA device                                             */
#define TRT_ECUDA      -2   /* CUDA runtime error (message has the cudaError string)             */
#define TRT_EINVAL     -3   /* bad argument                                                      */
#define TRT_ESTATE     -4   /* call out of order (e.g. stats before trt_harmonize)               */
#define TRT_ENOMEM     -5
#define TRT_ERECORD    -6   /* malformed record -> reference raises ValueError/TypeError         */
#define TRT_ENCCL      -7

/* ---- VCF caller types: trtools/utils/tr_harmonizer.py:23-38 (VcfTypes) ---------------------- */
#define TRT_VCF_GANGSTR 0
#define TRT_VCF_ADVNTR  1
#define TRT_VCF_HIPSTR  2
#define TRT_VCF_EH      3
#define TRT_VCF_POPSTR  4
#define TRT_VCF_LONGTR  5

/* ---- numeric FORMAT fields a block can carry (cyvcf2 Variant.format(key) arrays) ---------- */
#define TRT_FMT_DP           0   /* int32 [L][S]                                                  */
#define TRT_FMT_DSTUTTER     1   /* int32 [L][S]                                                  */
#define TRT_FMT_DFLANKINDEL  2   /* int32 [L][S]                                                  */
#define TRT_FMT_Q            3   /* float32 [L][S]                                                */
#define TRT_FMT_QEXP         4   /* float32 [L][S][3]  (GangSTR)                                  */
#define TRT_FMT_AUX0         5   /* AUX0..AUX7: any other Integer/Float [L][S] field a call filter */
#define TRT_FMT_NAUX         8   /*   reads (SR, FR, ML, LC, ADFL ..., or host-evaluated filters)  */
#define TRT_FMT_NFIELDS      13

typedef struct trt_ctx trt_ctx;

typedef struct {
    char     name[128];
    int32_t  cc_major, cc_minor;
    int32_t  sm_count;
    int64_t  total_mem_bytes;
    int64_t  free_mem_bytes;
    int32_t  l2_bytes;
    int32_t  abi_version;
} trt_devinfo;

/* ---- context ----------------------------------------------------------------------------- */
int         trt_device_count(void);
int         trt_init(int device_ordinal, trt_ctx** out);
void        trt_destroy(trt_ctx* ctx);
const char* trt_last_error(const trt_ctx* ctx);
int         trt_device_info(trt_ctx* ctx, trt_devinfo* out);
int         trt_synchronize(trt_ctx* ctx);
/* pinned host staging memory for the ingest ring (cudaHostAlloc)                              */
void*       trt_host_alloc(trt_ctx* ctx, size_t bytes);
int         trt_host_free(trt_ctx* ctx, void* p);
/* number of kernels this context has launched since creation (bench.py's gpu_launches)        */
int64_t     trt_launch_count(const trt_ctx* ctx);
/* device milliseconds of the most recent call's kernels, measured with CUDA events on the
 * context's stream (0 if the call launched nothing)                                           */
double      trt_last_kernel_ms(const trt_ctx* ctx);
/* device milliseconds of the dominant sample-axis kernel(s) (GT scan / call filter / OLS moments)
 * inside the most recent call — the numerator of bench.py's roofline figure                       */
double      trt_last_scan_ms(const trt_ctx* ctx);
/* user-level CUDA-event stopwatch on the context's stream (bench.py times K steps with it)       */
int         trt_stopwatch_start(trt_ctx* ctx);
int         trt_stopwatch_stop(trt_ctx* ctx, double* ms_out);

/* ---- block ingest ------------------------------------------------------------------------
 * A block is L consecutive VCF records ("loci") over the same S samples, in exactly the arrays
 * cyvcf2 hands the reference per record, stacked over loci.
 * Replaces the per-record pulls  vcfrecord.genotype.array()  (tr_harmonizer.py:829-862) and
 * vcfrecord.format(key)  (tr_harmonizer.py:561-588).                                           */
int trt_block_begin(trt_ctx* ctx, int64_t n_loci, int64_t n_samples, int ploidy, int vcftype);

/* GT in cyvcf2 layout: int16 [L][S][P+1]; allele index, -1 = '.', -2 = ploidy pad, last col phased */
int trt_block_set_gt(trt_ctx* ctx, const int16_t* gt_host);
/* same, but the array already lives in device memory (row pitch in bytes, multiple of 16; the
 * buffer must stay valid until the next trt_block_begin).  Used for device-generated blocks.     */
int trt_block_set_gt_device(trt_ctx* ctx, const int16_t* gt_dev, size_t row_pitch_bytes);

/* Compact transfer form of a DIPLOID GT block whose loci have at most 253 alleles: two bytes per call — allele index
 * 0..252, 254 = ploidy pad (-2), 255 = no-call (-1) — and, optionally, one phase bit per call (phase_bits_host
 * [L][ceil(S/8)], bit s & 7 of byte s >> 3; NULL = every call unphased) instead of cyvcf2's three int16.  The copy is a
 * third of the size (PCIe bounds the host-buffer path); an expansion kernel rebuilds the native rows in HBM, so every
 * other entry point is unchanged.  trt_vcf_block_parse_packed produces this form straight from the VCF text.           */
int trt_block_set_gt_packed(trt_ctx* ctx, const uint8_t* gt2_host /*[L][S][2]*/, const uint8_t* phase_bits_host);
/* the block's genotypes (a locus range) back in the packed form                                                        */
int trt_block_get_gt_packed(trt_ctx* ctx, int64_t locus0, int64_t n, uint8_t* gt2_out_host, uint8_t* phase_bits_out_host);
/* Nibble transfer form for DIPLOID blocks whose loci have at most 14 alleles (nearly every TR locus): ONE byte per call,
 * first haplotype in the low nibble, second in the high nibble; 0..13 allele index, 14 = ploidy pad (-2), 15 = no-call
 * (-1); phase bits as above.  A sixth of cyvcf2's bytes across PCIe.  trt_vcf_block_parse_nibble produces it from text. */
int trt_block_set_gt_nibble(trt_ctx* ctx, const uint8_t* g4_host /*[L][S]*/, const uint8_t* phase_bits_host);
int trt_block_get_gt_nibble(trt_ctx* ctx, int64_t locus0, int64_t n, uint8_t* g4_out_host, uint8_t* phase_bits_out_host);

int trt_block_set_format_i32(trt_ctx* ctx, int field_id, const int32_t* v_host /*[L][S]*/);
int trt_block_set_format_f32(trt_ctx* ctx, int field_id, const float* v_host /*[L][S][ncol]*/, int ncol);
int trt_block_set_format_device(trt_ctx* ctx, int field_id, const void* v_dev, int ncol, int is_float);

/* Allele table of the block.  seqs: concatenated REF/ALT strings exactly as in the VCF (any
 * case); allele_off[nA+1] byte offsets; locus_off[L+1] allele index ranges (first allele of a
 * locus = REF).  pos = VCF POS; start/end = INFO START/END (HipSTR/LongTR; pass pos and
 * pos+len(REF)-1 for callers without flanks); period = INFO PERIOD or len(RU/Motif).
 * given_len (may be NULL): per-allele repeat-unit length for callers that report lengths
 * (EH <STRn>, popSTR <n>; NaN = derive from the sequence).
 * motifs (may be NULL): concatenated INFO RU / Motif strings, period[l] bytes per locus, for the
 * callers that state the motif (GangSTR, adVNTR, EH, popSTR); HipSTR/LongTR motifs are inferred.
 * Replaces the string work of  _HarmonizeHipSTRRecord  tr_harmonizer.py:336-408,
 * _HarmonizeGangSTRRecord :303-333, _HarmonizeAdVNTRRecord :411-436, _HarmonizePopSTRRecord
 * :473-512, _HarmonizeEHRecord :515-550 and  TRRecord.__init__ :693-773.                        */
int trt_block_set_alleles(trt_ctx* ctx, const char* seqs, const int64_t* allele_off, const int32_t* locus_off,
                          const int32_t* pos, const int32_t* start, const int32_t* end, const int32_t* period,
                          const double* given_len, const char* motifs);

/* ---- harmonize ---------------------------------------------------------------------------
 * Warp-per-locus kernel: trims flanks (python slice semantics), upper-cases, derives
 * repeat-unit lengths (float64), length/sequence equivalence classes and their sort order,
 * infers the HipSTR motif (utils.InferRepeatSequence utils.py:465-508 + GetCanonicalOneStrand
 * :396-427) and the homopolymer run of the full REF (utils.GetHomopolymerRun :340-360).          */
int trt_harmonize(trt_ctx* ctx);

typedef struct {
    /* per allele (nA entries, locus-major) */
    double*  allele_len;   /* repeat units: [ref_len, *alt_lens]  tr_harmonizer.py:740,757-759   */
    int32_t* trim_off;     /* trimmed allele = seqs[allele_off[a]+trim_off[a] ...+trim_len[a])   */
    int32_t* trim_len;
    int32_t* len_class;    /* first allele index (within locus) with the same length             */
    int32_t* seq_class;    /* first allele index (within locus) with the same trimmed sequence   */
    int32_t* len_order;    /* allele indices of the locus sorted by (length, index)              */
    int32_t* seq_order;    /* allele indices of the locus sorted by (trimmed sequence, index)    */
    /* per locus (L entries) */
    int32_t* hrun;         /* GetHomopolymerRun(full REF)                                        */
    int32_t* flags;        /* TRT_HF_* bits                                                      */
    char*    motif;        /* [sum(period)] inferred motif bytes, locus l at motif_off[l]        */
    int64_t* motif_off;    /* [L+1] (exclusive prefix sum of period)                             */
} trt_harmonize_out;
#define TRT_HF_HAS_FULL     1   /* start/end offsets non-zero -> full_alleles kept (:360-369)      */
#define TRT_HF_MOTIF_N      2   /* period > len(seq): motif is 'N'*period (utils.py:488-489)       */
#define TRT_HF_MOTIF_NONACGT 4  /* GetCanonicalOneStrand would raise KeyError                       */
#define TRT_HF_LEN_DUPS     8   /* some alleles share a length                                     */
#define TRT_HF_SEQ_DUPS    16   /* some alleles share a trimmed sequence                           */
#define TRT_HF_BAD_PERIOD  32   /* period <= 0                                                     */
int trt_get_harmonized(trt_ctx* ctx, trt_harmonize_out* out /* any member may be NULL */);

/* Packed length-genotype tensor: int16 [L][S][P]; entry = rank of the haplotype's length among
 * the locus' sorted distinct lengths (table via trt_get_len_table), -1 no-call, -2 ploidy pad.
 * Device-side equivalent of TRRecord.GetLengthGenotypes (tr_harmonizer.py:1210-1245) — the
 * float64 [S][P+1] array is only materialised at the Python API edge.                            */
int trt_pack_length_genotypes(trt_ctx* ctx);
int trt_get_packed_gt(trt_ctx* ctx, int16_t* out_host /*[L][S][P]*/);

/* ---- per-locus statistics (statSTR) ---------------------------------------------------------
 * One pass over the native GT array per sample group.  Replaces, per locus and per group,
 *   GetAlleleCounts tr_harmonizer.py:1420-1499, GetGenotypeCounts :1326-1418,
 *   GetAlleleFreqs :1501-1540, GetMaxAllele :1542-1575, GetCalledSamples :864-897,
 *   utils.GetHeterozygosity/GetEntropy/GetMean/GetMode/GetVariance/
 *   GetHardyWeinbergBinomialTest utils.py:142-338 and the statSTR wrappers statSTR.py:104-426.  */
typedef struct {
    /* [G][nA]: allele counts keyed by allele INDEX (partial calls contribute their called hap) */
    int32_t* ac;
    /* [G][L] each; NULL = not wanted */
    int64_t* n_called;           /* samples with no '.' haplotype  (= sum of genotype counts)    */
    int64_t* n_called_nonstrict; /* samples with at least one called haplotype                   */
    int64_t* n_hom;              /* fully-called samples whose two smallest alleles are equal
                                    under the selected equivalence (length or sequence)          */
    int64_t* n_padded;           /* fully-called samples carrying a -2 ploidy pad                */
    double*  thresh;             /* GetMaxAllele                                                 */
    double*  het;
    double*  entropy;
    double*  mean;
    double*  mode;
    double*  var;
    double*  hwep;
    int32_t* nalleles;           /* classes with freq >= nalleles_thresh                         */
} trt_locus_stats_out;

/* group_masks: [G][S] bytes (0/1) or NULL with G = 1 for "all samples" (statSTR.py:520-542).
 * use_length: collapse alleles by length (1) or by trimmed sequence (0) for
 * het/entropy/hwep/nalleles — mean/mode/var/thresh always use lengths (statSTR.py:347,375,402). */
int trt_locus_stats(trt_ctx* ctx, int use_length, const uint8_t* group_masks, int n_groups,
                    double nalleles_thresh, trt_locus_stats_out* out);

/* TRRecord.GetGenotypeCounts tr_harmonizer.py:1326-1418 for ONE locus (API edge, optional sample
 * mask [S] bytes): counts of index genotypes with haplotypes sorted, as a dense base-(A+2) table of
 * (A+2)^P entries where digit = allele + 2 (-2 pad -> 0, -1 no-call -> 1).                         */
int trt_genotype_counts(trt_ctx* ctx, int64_t locus, const uint8_t* mask, int64_t* table, int64_t table_len);

/* ---- dumpSTR ---------------------------------------------------------------------------------
 * Call-level filter operators (filters.py:327-484, 573-674) evaluated in the order given,
 * followed by ApplyCallFilters' bookkeeping (dumpSTR.py:613-774).                                 */
#define TRT_CF_MIN       0   /* value <  threshold filtered   (CallFilterMinValue)                */
#define TRT_CF_MAX       1   /* value >  threshold filtered   (CallFilterMaxValue)                */
#define TRT_CF_RATIO_GT  2   /* field/DP > threshold filtered (HipSTRCallFlankIndels/Stutter)     */
#define TRT_CF_QEXP_HET  3   /* QEXP[1] < thr on called samples                                   */
#define TRT_CF_QEXP_HOM  4   /* QEXP[2] < thr                                                     */
#define TRT_CF_QEXP_TOT  5   /* QEXP[1]+QEXP[2] < thr                                             */
#define TRT_CF_HOST_VALUE 6  /* float32 field holds the operator's own output (NaN = keep): the
                                string-parsing filters of filters.py:486-567,676-757,835-867 are
                                evaluated by the caller and merged here in filter order           */
typedef struct {
    int32_t kind;        /* TRT_CF_*                                                              */
    int32_t field_id;    /* TRT_FMT_* the operator reads (numerator for RATIO_GT)                 */
    double  threshold;
} trt_call_filter_spec;

#define TRT_MAX_CALL_FILTERS 16
typedef struct {
    /* per call, [L][S]: bit f set = filter f fired (value non-NaN); bit 31 = no-call before
     * filtering.  May be NULL.                                                                   */
    uint32_t* call_mask;
    /* [n_specs][L][S] float64 triggering values (NaN = not filtered) — only what the host needs
     * to print FORMAT:FILTER ('%g');  may be NULL                                                */
    double*   trigger_values;
    /* masked genotypes int16 [L][S][P+1] (filtered calls -> all -1, phase 0); may be NULL        */
    int16_t*  gt_masked;
    /* per-sample accumulators, ADDED to the caller's arrays (they persist across blocks):
     * filter_counts [n_specs][S]; numcalls [S]; totaldp [S] (NaN-poisoned like dumpSTR.py:710)    */
    int64_t*  filter_counts;
    int64_t*  numcalls;
    double*   totaldp;
    int32_t*  negative_dp_locus; /* out: first locus with a PASS call of negative DP, or -1       */
} trt_call_filter_out;
int trt_call_filters(trt_ctx* ctx, const trt_call_filter_spec* specs, int n_specs, int dp_field_id /* TRT_FMT_DP or -1 */,
                     trt_call_filter_out* out);
/* after trt_call_filters: subsequent trt_locus_stats / trt_locus_filters / trt_assoc_* read the
 * MASKED genotypes (the rebuilt TRRecord of dumpSTR.py:748-774) until the next trt_block_begin.  */

/* Locus-level filters (filters.py:35-217) + ApplyLocusFilters (dumpSTR.py:917-973) + INFO
 * recompute (dumpSTR.py:1307-1336).                                                               */
#define TRT_LF_CALLRATE 0
#define TRT_LF_HWE      1
#define TRT_LF_HETLOW   2
#define TRT_LF_HETHIGH  3
#define TRT_LF_HRUN     4
typedef struct { int32_t kind; double threshold; } trt_locus_filter_spec;
typedef struct {
    uint32_t* flags;      /* [L] bit i = locus filter i fired; bit 31 = NO_CALLS_REMAINING        */
    int64_t*  n_called;   /* [L]                                                                  */
    double*   het;        /* [L] INFO HET  (-1 when nothing is called)                            */
    double*   hwep;       /* [L] INFO HWEP (-1 when nothing is called)                            */
    int32_t*  ac;         /* [nA] index allele counts (INFO AC / REFAC)                           */
    int32_t*  hrun;       /* [L]                                                                  */
} trt_locus_filter_out;
int trt_locus_filters(trt_ctx* ctx, const trt_locus_filter_spec* specs, int n_specs, int use_length,
                      trt_locus_filter_out* out);

/* ---- associaTR ---------------------------------------------------------------------------------
 * Design = standardised covariates with column 0 reserved for the genotype and column 1 the
 * intercept (associaTR.py:190-194), restricted to the samples kept by the sample filter.
 * sample_index[n_design]: VCF sample index of each design row (ascending).                        */
int trt_assoc_set_design(trt_ctx* ctx, const double* covars /*[n_design][K] row-major, col 0 ignored*/,
                         const double* outcome /*[n_design]*/, const int32_t* sample_index, int64_t n_design, int K);
#define TRT_AF_OK              0
#define TRT_AF_NO_CALLED       1   /* 'No called samples'          lafg.py:228-229                */
#define TRT_AF_ONE_ALLELE      2   /* 'Only one called allele'     lafg.py:230-231                */
#define TRT_AF_NON_MAJOR       3   /* 'non-major allele count<c'   lafg.py:232-236                */
#define TRT_AF_NCOVARS         4   /* 'n covars >= n samples'      associaTR.py:257-258           */
typedef struct {
    int32_t* filter_code;   /* [L] TRT_AF_*                                                       */
    int64_t* n_tested;      /* [L] called samples among the design rows                           */
    double*  p;             /* [L] two-sided t-test p of the genotype coefficient                 */
    double*  coef;          /* [L] coefficient on the standardised scale / std(g)  (x pheno_std by caller) */
    double*  se;            /* [L]                                                                */
    double*  r2;            /* [L]                                                                */
    double*  std_g;         /* [L] population std of the summed length genotype                   */
    int32_t* ac_len;        /* [nA] allele counts among tested samples keyed by allele index      */
} trt_assoc_out;
/* load_and_filter_genotypes.load_trs lafg.py:157-259 (non-dosage branch) + per-locus regression
 * associaTR.py:246-291 (statsmodels OLS on called rows: params, bse, pvalues, rsquared).          */
int trt_assoc_ols(trt_ctx* ctx, double non_major_cutoff, trt_assoc_out* out);

/* ---- dosages ----------------------------------------------------------------
 * Beagle allele probabilities of the block: FORMAT AP1 / AP2 exactly as cyvcf2 returns them per record — float32
 * [S][A_l - 1] — stacked over loci (locus l starts S * (alternate alleles of the loci before it) floats in).
 * has_ap (may be NULL = all): [L] bytes, 0 = the record carries no AP1 / AP2 (its rows are skipped: zero alts' worth).   */
int trt_block_set_ap(trt_ctx* ctx, const float* ap1_host, const float* ap2_host, const uint8_t* has_ap_host);
/* TRRecord.GetDosages tr_harmonizer.py:1098-1208 for every locus of the block -> float32 [L][S] (the tensor annotaTR
 * writes, annotaTR.py:673-703).  error_out [L]: record-level validation result; the dosages of an invalid record are
 * not meaningful (the reference raises ValueError, or warns and returns NaN when strict=False).                          */
#define TRT_DOSAGE_BESTGUESS       0
#define TRT_DOSAGE_BEAGLEAP        1
#define TRT_DOSAGE_BESTGUESS_NORM  2
#define TRT_DOSAGE_BEAGLEAP_NORM   3
#define TRT_DE_OK            0
#define TRT_DE_NO_AP         1   /* 'Requested Beagle dosages ... but AP1/AP2 fields not found'      :1132-1140 */
#define TRT_DE_AP_SUM        2   /* 'AP1 or AP2 field summing to more than 1 detected'               :1162-1168 */
#define TRT_DE_AP_NEGATIVE   3   /* 'Negative AP1 or AP2 fields detected'                            :1169-1175 */
#define TRT_DE_NORM_RANGE    4   /* 'Error normalizing dosages: value >=2.1 or <=-0.1 detected'      :1199-1205 */
int trt_dosages(trt_ctx* ctx, int dosage_type, float* dosage_out /*[L][S]*/, int32_t* error_out /*[L]*/);

/* associaTR --beagle-dosages: load_trs' dosage branch lafg.py:175-214 + the regression on the summed dosage
 * associaTR.py:266-291.  Per allele a of the block the caller passes the length class it belongs to after rounding to
 * two decimals (cls[a] = index within the locus of the first allele with the same rounded length), that rounded length
 * (python round) and the numpy-rounded length used for the best-guess comparison (np.around).                          */
typedef struct {
    int64_t* n_tested;      /* [L] called samples among the design rows                                                */
    double*  p;             /* [L] as trt_assoc_out (regression of the outcome on the summed dosage)                     */
    double*  coef;
    double*  se;
    double*  r2;
    double*  std_g;
    int32_t* ncovars_code;  /* [L] TRT_AF_NCOVARS when the design has at least as many columns as tested samples, else 0 */
    double*  class_stats;   /* [nA][4] at class representatives: sum d, sum d^2, #(best guess in class), sum d[best guess in class] */
    double*  length_stats;  /* [L][5]: sum x, sum x^2, sum y, sum y^2, sum xy over the 2 n haplotype entries (x best-guess
                               length, y expected length) for r2_length_dosages_vs_best_guess_lengths                   */
} trt_assoc_dosage_out;
int trt_assoc_dosage_ols(trt_ctx* ctx, const int32_t* cls /*[nA]*/, const double* len_round /*[nA]*/,
                         const double* len_around /*[nA]*/, trt_assoc_dosage_out* out);

/* ---- reductions of the qcSTR / compareSTR consumers --------------------------------------------
 * qcSTR's record loop trtools/qcSTR/qcSTR.py:523-570 over the block: calls per sample (a call = not every haplotype
 * '.', :532-535), calls per locus (chrom_calls), and the quality-field sums behind the quality plots (:536-556; a
 * no-call or missing quality reads 0, or is skipped with ignore_no_call).  sample_mask (may be NULL): [S] bytes.
 * The per-sample arrays are ADDED to (they persist across blocks).                                                     */
typedef struct {
    int64_t* sample_calls;     /* [S]                                                                                    */
    int64_t* locus_calls;      /* [L]                                                                                    */
    double*  sample_quality;   /* [S] sum of the quality field over the loci (NULL when quality_field < 0)               */
    double*  locus_quality;    /* [L] mean quality over the selected samples (NaN when nothing is averaged)              */
} trt_qc_out;
int trt_qc_reduce(trt_ctx* ctx, const uint8_t* sample_mask_host,
                  const int32_t* rec_ploidy_host /* [L] GT columns of each record's own cyvcf2 array, or NULL = block ploidy:
                                                    a '.' of a record narrower than the block is [-1, -2] in the block but
                                                    [-1] to the reference's np.all(idx_gts == -1) */,
                  int quality_field /* TRT_FMT_* or -1 */, int ignore_no_call, trt_qc_out* out);

/* compareSTR.UpdateComparisonResults trtools/compareSTR/compareSTR.py:508-643 for the block (call set 1, resident) against
 * a second call set of the SAME loci: gt2 in cyvcf2 layout [L][S2][P+1], the shared samples as index pairs, and per allele
 * of set 2 its repeat-unit length and the set-1 sequence class it equals (trt_get_harmonized seq_class of that allele) or
 * a negative id of its own.  reflen[l] = len(record1.ref_allele) / period.  Per locus: samples called in both sets,
 * sequence / length concordant calls (haplotypes compared as unordered pairs when the calls are unphased or
 * ignore_phasing is set), the sums of d1, d2, d1^2, d1 d2, d2^2 with d = sum over haplotypes of (length - reflen), and a
 * status (1: a sample's ploidy differs between the sets, 2: mixed phasedness — the reference raises ValueError).         */
typedef struct {
    const int16_t* gt2;
    int64_t        S2;
    const int32_t* idx1;        /* [n_shared] sample index in set 1 ...                                                  */
    const int32_t* idx2;        /* ... and the same sample's index in set 2                                               */
    int64_t        n_shared;
    const int32_t* locus_off2;  /* [L+1] allele ranges of set 2                                                           */
    const int32_t* seq_id2;     /* [nA2]                                                                                  */
    const double*  len2;        /* [nA2]                                                                                  */
    const double*  reflen;      /* [L]                                                                                    */
    int32_t        ignore_phasing;
} trt_compare_in;
typedef struct {
    int64_t* numcalls;          /* [L]                                                                                    */
    int64_t* conc_seq;          /* [L]                                                                                    */
    int64_t* conc_len;          /* [L]                                                                                    */
    double*  len_sums;          /* [L][5]                                                                                 */
    int32_t* status;            /* [L]                                                                                    */
    int64_t* sample_numcalls;   /* [n_shared] ADDED to                                                                    */
    int64_t* sample_conc_seq;   /* [n_shared] ADDED to                                                                    */
    int64_t* sample_conc_len;   /* [n_shared] ADDED to                                                                    */
} trt_compare_out;
int trt_compare(trt_ctx* ctx, const trt_compare_in* in, trt_compare_out* out);

/* ---- synthetic blocks (bench / parity at sizes that do not fit through PCIe) ------------------
 * Device twin of trtools_b200/synth.py::fill_calls — bit-identical arrays.                        */
int trt_synth_fill(trt_ctx* ctx, uint64_t seed, int64_t locus_offset, int64_t n_loci, int64_t n_samples,
                   const uint32_t* cum_freq_host /*[n_loci][16]*/, uint32_t miss_thresh, uint32_t half_thresh,
                   int with_format /* bitmask: 1 DP, 2 DSTUTTER, 4 DFLANKINDEL, 8 Q (1 alone = all four) */);
/* after trt_synth_fill the block's GT (and FORMAT) arrays are the generated ones; these copy
 * a locus range back for parity checks                                                          */
int trt_block_get_gt(trt_ctx* ctx, int64_t locus0, int64_t n, int16_t* out_host /*[n][S][P+1]*/);
int trt_block_get_format(trt_ctx* ctx, int field_id, int64_t locus0, int64_t n, void* out_host);

/* ---- native block VCF ingest (host side; SURVEY.md §8f "next" row 1) ---------------------------
 * Replaces the per-record cyvcf2/htslib pulls the reference makes in its record loop:
 * `for record in vcf` + vcfrecord.genotype.array() (tr_harmonizer.py:829-862, 1761-1779) and
 * vcfrecord.format(key) (tr_harmonizer.py:561-588; dumpSTR/filters.py:365,446).  BGZF members are
 * inflated in parallel; a block of records is held as text and GT + the requested numeric
 * FORMAT keys of all its records are parsed in one multi-threaded pass into the stacked arrays
 * trt_block_set_gt / trt_block_set_format_* take (pass trt_host_alloc memory for pinned staging).
 * Pure host code: no CUDA call, no trt_ctx.                                                      */
typedef struct trt_vcf trt_vcf;
typedef struct trt_vcf_block trt_vcf_block;
/* plain text, gzip or BGZF; n_threads <= 0 = all hardware threads.  Reads the header.           */
int         trt_vcf_open(const char* path, int n_threads, trt_vcf** out);
void        trt_vcf_close(trt_vcf* v);
/* message of the last failure on v (v == NULL: of the calling thread's last failed trt_vcf_open)  */
const char* trt_vcf_last_error(const trt_vcf* v);
/* every leading '#' line, verbatim (cyvcf2 VCF.raw_header)                                        */
int         trt_vcf_header(trt_vcf* v, const char** text, int64_t* len);
/* sample columns named by the #CHROM line                                                         */
int64_t     trt_vcf_n_samples(const trt_vcf* v);
/* restrict parsing to these sample columns (strictly increasing, 0-based; cyvcf2 VCF(samples=))   */
int         trt_vcf_set_samples(trt_vcf* v, const int64_t* cols, int64_t n);
/* BGZF input only: continue reading at a virtual file offset from a tabix / CSI index (member at
 * byte coffset of the file, uoffset bytes into its inflated data) — cyvcf2 VCF(region) queries    */
int         trt_vcf_seek(trt_vcf* v, int64_t coffset, int32_t uoffset);
/* next run of up to max_loci records / about max_bytes of text (<= 0: no byte cap; at least one
 * record is returned).  *n_loci == 0 and *out == NULL at end of file.  The block owns its text
 * and outlives the reader's later reads; free it with trt_vcf_block_free.                         */
int         trt_vcf_read_block(trt_vcf* v, int64_t max_loci, int64_t max_bytes, trt_vcf_block** out, int64_t* n_loci);
void        trt_vcf_block_free(trt_vcf_block* b);
/* record text: record i is text[line_off[i] .. line_off[i+1]) (ends with '\n'); its first nine
 * columns (CHROM..FORMAT) are the first fixed_len[i] bytes; fixed_len[i] < 0 = fewer than eight
 * columns (the reference's reader fails on such a record)                                         */
int         trt_vcf_block_text(const trt_vcf_block* b, const char** text, const int64_t** line_off,
                               const int64_t** fixed_len);
/* One pass over the block's sample columns.
 *   gt_out (may be NULL): int16 [n][S][ploidy+1] in cyvcf2 layout (trt_block_set_gt's input)
 *   keys[k] / key_is_float[k] / key_out[k]: numeric scalar FORMAT keys -> int32 or float32 [n][S]
 *     (INT32_MIN / NaN for '.'); at most 32 keys per pass
 *   present[n][n_keys]: 0 = the record's FORMAT lacks the key, 1 = parsed, 2 = the record has a
 *     value this parser does not take (vector-valued, non-numeric): re-parse that record's key
 *   rec_ploidy[n]: most alleles in one GT of the record (> ploidy: call again with that ploidy)
 *   rec_status[n]: 0 ok, 1 malformed (< 8 columns), 2 = not handled natively (odd GT token, no GT
 *     key, ragged sample columns): re-parse the record from its text                              */
int         trt_vcf_block_parse(const trt_vcf_block* b, int ploidy, int16_t* gt_out, int n_keys,
                                const char* const* keys, const int32_t* key_is_float, void* const* key_out,
                                uint8_t* present, int32_t* rec_ploidy, uint8_t* rec_status);

/* The same pass with GT in the packed transfer form of trt_block_set_gt_packed: gt2_out uint8 [n][S][2] (allele 0..252,
 * 254 = ploidy pad, 255 = no-call), phase_out (may be NULL) [n][ceil(S/8)] one bit per call.  rec_status 3 = the record
 * does not fit (an allele index above 252, or a call with more than two haplotypes): parse the block in the plain form.  */
int         trt_vcf_block_parse_packed(const trt_vcf_block* b, uint8_t* gt2_out, uint8_t* phase_out, int n_keys,
                                       const char* const* keys, const int32_t* key_is_float, void* const* key_out,
                                       uint8_t* present, int32_t* rec_ploidy, uint8_t* rec_status);
/* the same into the nibble form (trt_block_set_gt_nibble): g4_out [n][S]; rec_status 3 = the record has an allele index
 * above 13 or more than two haplotypes (parse the block with trt_vcf_block_parse_packed instead)                        */
int         trt_vcf_block_parse_nibble(const trt_vcf_block* b, uint8_t* g4_out, uint8_t* phase_out, int n_keys,
                                       const char* const* keys, const int32_t* key_is_float, void* const* key_out,
                                       uint8_t* present, int32_t* rec_ploidy, uint8_t* rec_status);

/* raw text of the field at position field_index of the record's FORMAT column, for every kept
 * sample, as fixed-width NUL-padded byte strings [S][width] (String FORMAT keys: vcfrecord.format
 * on HipSTR GB / ALLREADS, GangSTR RC / REPCI — dumpSTR/filters.py:573-757; and the writer's
 * pass-through of fields nothing decoded).  *max_len = longest token; out is filled only when
 * out != NULL and width >= *max_len (call once to size, once to fill).                           */
int         trt_vcf_block_field(const trt_vcf_block* b, int64_t rec, int field_index, int32_t width, char* out,
                                int32_t* max_len);
/* Sample columns of one record as text (dumpSTR's record emission, trtools/dumpSTR/dumpSTR.py:1338
 * vcf writer.write_record): per sample the fields joined by ':', samples joined by TAB.
 *   kind[f]: 0 fixed-width byte strings [S] of ncol[f] bytes (NUL padded), 1 GT int16 [S][ncol]
 *   (cyvcf2 layout), 2 int32 [S][ncol] (INT32_MIN prints '.', INT32_MIN+1 = vector end, skipped),
 *   3 float32 / 4 float64 [S][ncol] ("%g", NaN prints '.')
 * Returns the bytes written, or -(bytes needed) if cap is too small.                             */
int64_t     trt_vcf_join_samples(int64_t n_samples, int n_fields, const int32_t* kind, const void* const* data,
                                 const int32_t* ncol, char* out, int64_t cap);

/* ---- multi-GPU: loci shard by contiguous ranges, one context per rank --------------------------
 * NCCL is used only to gather fixed-width per-locus result rows and to sum per-sample counters.  */
int trt_dist_unique_id(void* out_128_bytes);
int trt_dist_init(trt_ctx* ctx, int rank, int world, const void* unique_id_128_bytes);
int trt_dist_allgather_f64(trt_ctx* ctx, const double* send_host, int64_t count, double* recv_host /*[world*count]*/);
int trt_dist_allreduce_sum_i64(trt_ctx* ctx, int64_t* inout_host, int64_t count);
int trt_dist_allreduce_sum_f64(trt_ctx* ctx, double* inout_host, int64_t count);
int trt_dist_allreduce_max_f64(trt_ctx* ctx, double* inout_host, int64_t count);
int trt_dist_barrier(trt_ctx* ctx);
/* Gather of per-locus result rows on rank dst, DEVICE buffers on both sides (no host bounce): every rank sends
 * `nbytes` bytes starting `offset_bytes` into one of its device-resident result regions (the outputs of its most
 * recent trt_locus_stats / trt_assoc_ols / trt_locus_filters call, which may be made with NULL host pointers);
 * the context stream only copies the rows into a staging buffer; the exchange (grouped ncclSend / ncclRecv, ragged
 * sizes: nbytes_per_rank[world], rank order on dst) and dst's copy of the gathered table to host_out (pinned memory
 * from trt_host_alloc for full speed) run on a side stream, so the next block's kernels overlap both
 * (TRT_DIST_MAIN_STREAM=1 in the environment keeps the exchange on the context stream).  async != 0: returns without
 * blocking the host; trt_dist_wait completes it.
 * Region layouts (n = G*L rows of the call):
 *   TRT_REGION_STATS          12 arrays of n 8-byte values: thresh, het, entropy, mean, mode, var, hwep (f64),
 *                             nalleles (i32, first 4n bytes of its 8n slot), n_hom, n_called, n_called_nonstrict,
 *                             n_padded (i64)
 *   TRT_REGION_ALLELE_COUNTS  int32 [G][nA] allele counts by allele index
 *   TRT_REGION_ASSOC          p, coef, se, r2, std_g (f64 [L] each), n_tested (i64 [L]), filter_code (i32 [L], padded
 *                             to a multiple of 4 entries), ac_len (i32 [nA])
 *   TRT_REGION_LOCUS_FILTERS  flags (u32 [L], padded to 16 bytes), het, hwep (f64 [L]), n_called (i64 [L])           */
#define TRT_REGION_STATS          0
#define TRT_REGION_ALLELE_COUNTS  1
#define TRT_REGION_ASSOC          2
#define TRT_REGION_LOCUS_FILTERS  3
int trt_dist_gather_region(trt_ctx* ctx, int region, int64_t offset_bytes, int64_t nbytes, const int64_t* nbytes_per_rank,
                           int dst, void* host_out /* dst only; may be NULL */, int async);
/* same with host buffers on both sides (the CLIs' formatted output rows: ragged byte strings), blocking            */
int trt_dist_gather_host(trt_ctx* ctx, const void* send_host, int64_t nbytes, const int64_t* nbytes_per_rank, int dst,
                         void* recv_host);
int trt_dist_wait(trt_ctx* ctx);
int trt_dist_finalize(trt_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* TRTOOLS_B200_H */
