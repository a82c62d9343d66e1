/*
 * seqlib_b200.h -- C ABI of the B200-native seed-and-extend engine.
 *
 * This is the drop-in boundary: plain pointers and
 * sizes, no C++/torch types.  Every entry point names the reference
 * interface it replaces (paths relative to the SeqLib tree).  The library
 * behind it is CUDA only: there is no CPU fallback, every call that needs
 * the device fails with B200_ERR_CUDA when no sm_100 device is usable.
 *
 * Conventions
 *   - return value: 0 on success, a negative B200_ERR_* code otherwise;
 *     b200_last_error() returns a thread-local message.
 *   - every buffer handed back through an out-parameter of
 *     b200_mem_align_batch() is owned by the b200_results_t handle and is
 *     released by b200_results_free().
 */
#ifndef SEQLIB_B200_H
#define SEQLIB_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK            0
#define B200_ERR_ARG      -1
#define B200_ERR_CUDA     -2
#define B200_ERR_NOMEM    -3
#define B200_ERR_IO       -4
#define B200_ERR_LIMIT    -5   /* input outside the supported envelope */

const char *b200_last_error(void);

/* ------------------------------------------------------------------ */
/* Options: field-for-field the layout of bwa's mem_opt_t             */
/* (bwa/bwamem.h:52-84) so a binding can pass its own struct through. */
/* ------------------------------------------------------------------ */
typedef struct b200_mem_opt {
    int a, b;
    int o_del, e_del;
    int o_ins, e_ins;
    int pen_unpaired;
    int pen_clip5, pen_clip3;
    int w;
    int zdrop;
    uint64_t max_mem_intv;
    int T;
    int flag;
    int min_seed_len;
    int min_chain_weight;
    int max_chain_extend;
    float split_factor;
    int split_width;
    int max_occ;
    int max_chain_gap;
    int n_threads;
    int chunk_size;
    float mask_level;
    float drop_ratio;
    float XA_drop_ratio;
    float mask_level_redun;
    float mapQ_coef_len;
    int mapQ_coef_fac;
    int max_ins;
    int max_matesw;
    int max_XA_hits, max_XA_hits_alt;
    int8_t mat[25];
} b200_mem_opt_t;

/* replaces mem_opt_init() (bwa/bwamem.c:74-110) + bwa_fill_scmat()
 * (bwa/bwa.c:136-145); fills *opt in place. */
void b200_mem_opt_init(b200_mem_opt_t *opt);
void b200_fill_scmat(int a, int b, int8_t mat[25]);

/* ------------------------------------------------------------------ */
/* Index                                                              */
/* ------------------------------------------------------------------ */
typedef struct b200_index b200_index_t;

/* One reference contig as the host sees it (bntann1_t, bwa/bntseq.h:41-48). */
typedef struct b200_contig {
    int64_t offset;
    int32_t len;
    int32_t n_ambs;
    uint32_t gi;
    int32_t is_alt;
    const char *name;
    const char *anno;
} b200_contig_t;

/* Host view of an index in bwa's own layout; all pointers owned by the
 * index handle.  bwt is the Occ-interleaved array of bwt_bwtupdate_core()
 * (bwa/bwtindex.c:149-171), sa is sampled every sa_intv ranks with
 * sa[0] = -1 (bwa/bwt.c:62-84), pac is the forward strand, 4 bases/byte. */
typedef struct b200_index_view {
    uint64_t primary, L2[5], seq_len, bwt_size;
    const uint32_t *bwt;
    int sa_intv;
    uint64_t n_sa;
    const uint64_t *sa;
    int64_t l_pac;
    const uint8_t *pac;
    int32_t n_seqs;
    const b200_contig_t *contigs;
} b200_index_view_t;

/* replaces BWAIndex::ConstructIndex (src/BWAIndex.cpp:83-180): names/seqs
 * are n NUL-terminated strings.  Ambiguous bases become lrand48()&3, drawn
 * in the reference's order (two passes, src/BWAIndex.cpp:107-113,217).
 * The suffix sort, BWT, Occ interleave and SA sampling run on the GPU.
 * flags: bit0 = keep a host copy in bwa layout (needed for WriteIndex). */
int b200_index_construct(int n, const char *const *names, const char *const *seqs,
                         int flags, b200_index_t **out);

/* Builds an index over a 2-bit forward pac that is already on the host
 * (4 bases/byte, bwa order).  Same engine as b200_index_construct without
 * the ASCII pass; used for the 3 Gb synthetic reference.  `pac` holds
 * ceil(l_pac / 4) bytes; nothing beyond them is read. */
int b200_index_construct_pac(int64_t l_pac, const uint8_t *pac, int n_seqs,
                             const b200_contig_t *contigs, int flags, b200_index_t **out);

/* replaces BWAIndex::LoadIndex -> bwa_idx_load (bwa/bwa.c:289-316). */
int b200_index_load(const char *prefix, b200_index_t **out);
/* replaces BWAIndex::WriteIndex (src/BWAIndex.cpp:382-406): .bwt .sa .pac .ann .amb */
int b200_index_write(const b200_index_t *idx, const char *prefix);

/* replaces bwa_idx_destroy (bwa/bwa.c:323-335). */
void b200_index_destroy(b200_index_t *idx);

/* Host view (valid when the host copy was kept / loaded). */
int b200_index_view(const b200_index_t *idx, b200_index_view_t *view);

/* Device image: one contiguous blob (like bwa_idx2mem, bwa/bwa.c:362-401) so
 * a single ncclBroadcast ships it.  export copies device->device into a
 * caller-provided device buffer of b200_index_blob_bytes(); attach builds
 * an index handle over such a buffer (the buffer must outlive the handle). */
int64_t b200_index_blob_bytes(const b200_index_t *idx);
int b200_index_export_blob(const b200_index_t *idx, void *dev_dst);
int b200_index_attach_blob(void *dev_blob, int64_t nbytes, b200_index_t **out);
/* contig table of an attached blob is read back from the blob itself. */

int b200_index_n_seqs(const b200_index_t *idx);
const char *b200_index_seq_name(const b200_index_t *idx, int rid);
int64_t b200_index_seq_len(const b200_index_t *idx, int rid);
int64_t b200_index_l_pac(const b200_index_t *idx);

/* ------------------------------------------------------------------ */
/* Alignment                                                          */
/* ------------------------------------------------------------------ */

/* One alignment region + its finished alignment: the union of
 * mem_alnreg_t (bwa/bwamem.h:86-105) and mem_aln_t (bwa/bwamem.h:115-126)
 * as produced by mem_align1() followed by mem_reg2aln(). */
typedef struct b200_hit {
    int64_t rb, re;
    int64_t pos;
    uint64_t hash;
    int32_t qb, qe;
    int32_t rid;
    int32_t score, truesc, sub, alt_sc, csub, sub_n, w, seedcov;
    int32_t secondary, secondary_all, seedlen0, n_comp, is_alt;
    float frac_rep;
    int32_t flag;       /* mem_aln_t.flag (0x100 for secondary) */
    int32_t is_rev;
    int32_t mapq;
    int32_t NM;
    int32_t aln_sub;    /* mem_aln_t.sub = max(sub, csub) */
    int32_t n_cigar;
    int32_t md_len;     /* strlen(MD) */
    int64_t cigar_off;  /* index of the first CIGAR word in the cigar pool */
    int64_t md_off;     /* byte offset of the MD string in the md pool */
} b200_hit_t;

typedef struct b200_results b200_results_t;

typedef struct b200_results_view {
    int64_t n_reads;
    const int64_t *hit_off;   /* n_reads+1 entries; hits of read i are [hit_off[i], hit_off[i+1]) */
    const b200_hit_t *hits;
    const uint32_t *cigar;    /* BAM encoding len<<4|op, op 3 = clip (rewritten to S/H by the caller) */
    const char *md;           /* NUL-terminated MD strings */
    int64_t n_hits, n_cigar, n_md;
} b200_results_view_t;

/* Batch form of BWAAligner::alignSequence's compute
 * (src/BWAAligner.cpp:104-128): per read mem_align1()
 * (bwa/bwamem_extra.c:103-115) then mem_reg2aln() (bwa/bwamem.c:1119-1189)
 * for every region, in region order.
 *   seqs/seq_off : concatenated ASCII reads; read i = seqs[seq_off[i], seq_off[i+1])
 *   hash_ids     : the per-read tie-break id, i.e. the caller's lrand48()
 *                  draw (bwa/bwamem_extra.c:112); NULL = draw lrand48()
 *                  here, one per read in order.
 * Host buffers in, host buffers out; H2D/D2H copies are inside the call. */
int b200_mem_align_batch(const b200_index_t *idx, const b200_mem_opt_t *opt,
                         int64_t n_reads, const char *seqs, const int64_t *seq_off,
                         const int64_t *hash_ids, b200_results_t **out);
int b200_results_view(const b200_results_t *res, b200_results_view_t *view);
void b200_results_free(b200_results_t *res);

/* Page-locked host memory from the library's pool (cudaHostAlloc, recycled across calls): reads handed to b200_mem_align_batch
 * in such a buffer travel by DMA straight from it instead of through the driver's staging copies (a third of the call's time for
 * pageable buffers).  The drop-in BWAAligner::alignSequences flattens its std::string reads into one (src/BWAAligner.cpp has no
 * analogue: the reference aligns in place on the host). */
int b200_host_alloc(size_t bytes, void **out);
void b200_host_free(void *p);

/* Device-resident form used by bench.py's "value" leg: the reads are
 * uploaded once, each call runs all kernels and leaves results on the
 * device.  n_launches (optional) receives the number of kernel launches. */
typedef struct b200_batch b200_batch_t;
int b200_batch_create(const b200_index_t *idx, const b200_mem_opt_t *opt,
                      int64_t n_reads, const char *seqs, const int64_t *seq_off,
                      const int64_t *hash_ids, b200_batch_t **out);
int b200_batch_run(b200_batch_t *b, int *n_launches);
int b200_batch_fetch(b200_batch_t *b, b200_results_t **out);
void b200_batch_destroy(b200_batch_t *b);

/* ------------------------------------------------------------------ */
/* Multi-GPU: one process per GPU.  The index image is   */
/* replicated by ONE NCCL broadcast (the analogue of shipping the single */
/* block of bwa_idx2mem, bwa/bwa.c:362-401); a read batch is sharded by  */
/* contiguous read index and scattered over NVLink; after that every     */
/* rank calls b200_mem_align_batch on its shard -- no collective in the  */
/* data path.  NCCL is bound at run time (dlopen).                       */
/* ------------------------------------------------------------------ */
typedef struct b200_comm b200_comm_t;
/* reads [*beg, *end) of rank: contiguous, sizes differ by at most one */
void b200_shard_bounds(int64_t n_total, int world, int rank, int64_t *beg, int64_t *end);
int b200_comm_unique_id(char id[128]);      /* rank 0 makes it, the caller ships it to the other ranks (any channel) */
int b200_comm_init(const char id[128], int rank, int world, b200_comm_t **out);   /* on the calling thread's current device */
void b200_comm_destroy(b200_comm_t *c);
/* root passes its index and gets it back; the other ranks get a replica that owns its device memory */
int b200_index_bcast(b200_comm_t *c, const b200_index_t *root_idx, int root, b200_index_t **out);
float b200_comm_last_bcast_ms(const b200_comm_t *c);   /* device time of the image broadcast */
/* fixed-length reads held by root (host, n_total * read_len bytes) -> this rank's shard (host, (end - beg) * read_len bytes) */
int b200_reads_scatter(b200_comm_t *c, int root, int64_t n_total, int read_len, const char *seqs, char *shard);

/* Per-stage device timings (ms, CUDA events on the launching stream) and
 * work counters of the last b200_batch_run / b200_mem_align_batch. */
typedef struct b200_stage_stats {
    float ms_seed, ms_chain, ms_extend, ms_finalize, ms_total;
    uint64_t occ_blocks;      /* Occ blocks fetched (32 B each)            */
    uint64_t sa_reads;        /* suffix-array entries fetched (8 B each)   */
    uint64_t ref_bytes;       /* packed reference bytes fetched            */
    uint64_t sw_cells;        /* DP cells inside the adaptive band         */
    uint64_t n_ext, n_global; /* ksw_extend2 / ksw_global2 calls           */
    uint64_t n_overflow;      /* reads re-run with spill buffers           */
    int n_launches;
    uint64_t tab_lookups_lo;  /* seeding, text path: suffix-array / text sector requests (one or two 32-B sectors each) */
    uint64_t tab_lookups_hi;  /* seeding: prefix-chain table entries fetched (32 B each, HBM gathers) */
    uint64_t ext_fallback;    /* reads whose extensions were re-run by the row-synchronous kernel */
    uint64_t n_failed;        /* reads beyond every working-set limit: reported without hits (b200_last_error says so) */
} b200_stage_stats_t;
int b200_last_stats(b200_stage_stats_t *out);

/* Stage dumps for parity tests (each mirrors one reference intermediate). */
typedef struct b200_intv { uint64_t x0, x1, x2, info; } b200_intv_t; /* bwtintv_t, bwa/bwt.h:62-64 */
int b200_debug_collect_intv(const b200_index_t *idx, const b200_mem_opt_t *opt,
                            int64_t n_reads, const char *seqs, const int64_t *seq_off,
                            int64_t **intv_off, b200_intv_t **intv);   /* free() both */

/* ------------------------------------------------------------------ */
/* ksw_extend2 batch (config 3 microbench + unit parity)              */
/* replaces ksw_extend2 (bwa/ksw.c:416-515)                           */
/* ------------------------------------------------------------------ */
typedef struct b200_ext_job {
    int32_t qlen, tlen;
    int64_t q_off, t_off;    /* offsets into the query / target byte pools (nt4 codes 0..4) */
    int32_t w, end_bonus, zdrop, h0;
} b200_ext_job_t;
typedef struct b200_ext_out {
    int32_t score, qle, tle, gtle, gscore, max_off;
} b200_ext_out_t;
int b200_ksw_extend2_batch(int64_t n, const b200_ext_job_t *jobs,
                           const uint8_t *qpool, int64_t qpool_len,
                           const uint8_t *tpool, int64_t tpool_len,
                           const int8_t mat[25], int o_del, int e_del, int o_ins, int e_ins,
                           b200_ext_out_t *out, uint64_t *cells, float *kernel_ms);

/* ------------------------------------------------------------------ */
/* fermi-lite half: k-mer counting, error correction, unique filter   */
/*                                         */
/* ------------------------------------------------------------------ */

/* == fseq1_t (fermi-lite/fml.h:8-11): NUL-terminated, malloc'd strings. */
typedef struct b200_fseq1 {
    int32_t l_seq;
    char *seq, *qual;
} b200_fseq1_t;

/* == magopt_t (fermi-lite/fml.h:17-20) */
typedef struct b200_magopt {
    int flag, min_ovlp, min_elen, min_ensr, min_insr, max_bdist, max_bdiff, max_bvtx, min_merge_len, trim_len, trim_depth;
    float min_dratio1, max_bcov, max_bfrac;
} b200_magopt_t;

/* == fml_opt_t (fermi-lite/fml.h:22-29) */
typedef struct b200_fml_opt {
    int n_threads;
    int ec_k;
    int min_cnt, max_cnt;
    int min_asm_ovlp;
    int min_merge_len;
    b200_magopt_t mag_opt;
} b200_fml_opt_t;

/* replaces fml_opt_init (fermi-lite/misc.c:31-41) + mag_init_opt (fermi-lite/mag.c:539-557). */
void b200_fml_opt_init(b200_fml_opt_t *opt);
/* replaces fml_opt_adjust (fermi-lite/misc.c:43-54); only the read lengths are looked at. */
void b200_fml_opt_adjust(b200_fml_opt_t *opt, int n_seqs, const b200_fseq1_t *seqs);
void b200_fml_opt_adjust_lens(b200_fml_opt_t *opt, int64_t n_seqs, int64_t tot_len);

/* replaces fml_correct (fermi-lite/bfc.c:568-571): count ec_k-mers of all reads, then correct every read
 * IN PLACE (changed bases lower case, qualities recoded, bfc_ec1 at fermi-lite/bfc.c:401-466).
 * *kcov receives the value fml_correct returns.  ec_k <= 0 leaves the reads untouched and reports 255
 * (what the reference's undefined shifts amount to on x86-64). */
int b200_fml_correct(const b200_fml_opt_t *opt, int n, b200_fseq1_t *seqs, float *kcov);
/* replaces fml_fltuniq (fermi-lite/bfc.c:573-576): trims every read to its longest run of min_asm_ovlp-mers
 * present in the count table; dropped reads are free()d and get l_seq = 0, seq = qual = NULL like the reference. */
int b200_fml_fltuniq(const b200_fml_opt_t *opt, int n, b200_fseq1_t *seqs, float *kcov);

/* The same two calls on flat pools (what the wrappers above marshal into): read i occupies
 * [off[i], off[i+1]) of seqs (and of quals unless quals == NULL).  The pools are rewritten in place;
 * len_out[i] receives the new length (flt_uniq: the kept run now starts at off[i]; 0 = dropped). */
int b200_fml_correct_flat(const b200_fml_opt_t *opt, int flt_uniq, int64_t n, char *seqs, char *quals,
                          const int64_t *off, int32_t *len_out, float *kcov);

/* k-mer count table: replaces fml_count (fermi-lite/bfc.c:86-99) / bfc_ch_t (fermi-lite/htab.c). */
typedef struct b200_kmer_table b200_kmer_table_t;
int b200_fml_count(int64_t n, const char *seqs, const char *quals, const int64_t *off,
                   int k, int q, int l_pre, b200_kmer_table_t **out);
/* replaces bfc_ch_hist (fermi-lite/htab.c:104-127): returns the mode through *mode (-1 if none). */
int b200_kmer_table_hist(const b200_kmer_table_t *tab, uint64_t cnt[256], uint64_t high[64], int *mode);
/* replaces bfc_ch_count (fermi-lite/htab.c:95-102): distinct keys. */
int64_t b200_kmer_table_size(const b200_kmer_table_t *tab);
/* replaces bfc_ch_kmer_occ (fermi-lite/htab.c:85-93) for a batch of k-mers given as ASCII (n * k bytes);
 * occ[i] = -1 if absent, else high << 8 | total. */
int b200_kmer_table_lookup(const b200_kmer_table_t *tab, int64_t n, const char *kmers, int32_t *occ);
/* replaces kmer_correct (fermi-lite/bfc.c:556-566) as BFC::ErrorCorrect drives it (src/BFC.cpp:221-262):
 * corrects (or, with flt_uniq, trims) reads against an existing table with explicit min_cov / mode. */
int b200_kmer_correct_flat(const b200_kmer_table_t *tab, int min_cov, int mode, int flt_uniq, int64_t n,
                           char *seqs, char *quals, const int64_t *off, int32_t *len_out);
void b200_kmer_table_destroy(b200_kmer_table_t *tab);

/* ------------------------------------------------------------------ */
/* fermi-lite half: FMD-index, unitig graph, cleaning, unitigs        */
/*                                         */
/* ------------------------------------------------------------------ */

/* == fml_ovlp_t / fml_utg_t (fermi-lite/fml.h:34-46) */
typedef struct b200_utg_ovlp {
    uint32_t len:31, from:1;
    uint32_t id:31, to:1;
} b200_utg_ovlp_t;
typedef struct b200_utg {
    int32_t len;
    int32_t nsr;
    char *seq;
    char *cov;
    int n_ovlp[2];
    b200_utg_ovlp_t *ovlp;
} b200_utg_t;

typedef struct b200_utgs b200_utgs_t;

/* replaces fml_assemble (fermi-lite/misc.c:280-302) on flat pools: fml_opt_adjust, fml_correct (ec_k >= 0),
 * fml_fltuniq, fml_seq2fmi, fml_fmi2mag, fml_mag_clean with min_ensr/min_insr derived from kcov, fml_mag2utg.
 * The caller's pools are NOT modified (the reference leaves its reads in an unspecified state).  *out may be
 * an empty set (the reference returns NULL when every read was filtered out). */
int b200_fml_assemble_flat(const b200_fml_opt_t *opt, int64_t n, const char *seqs, const char *quals,
                           const int64_t *off, b200_utgs_t **out);
/* Many independent assemblies in one call -- what a caller of FermiAssembler does window after window (one genomic window
 * per assembler, SeqLib/FermiAssembler.h:25-123): window w owns reads [win_off[w], win_off[w+1]) of the flat pools and gets
 * exactly the unitigs b200_fml_assemble_flat returns for those reads alone (out[w], each released with b200_utgs_free).
 * A window-sized assembly is latency bound on a B200 (one string per thread for a few thousand strings), so the windows
 * run concurrently: n_threads host threads (0 = 4, at most 32), each with its own stream, claim windows from a shared
 * counter.  Returns the first error any window produced (the other windows still complete).  b200_fml_last_stats()
 * afterwards holds the stage times summed over the windows. */
int b200_fml_assemble_windows(const b200_fml_opt_t *opt, int64_t n_windows, const int64_t *win_off,
                              const char *seqs, const char *quals, const int64_t *off, int n_threads, b200_utgs_t **out);

/* the assembly half alone, as FermiAssembler::DirectAssemble drives it (src/FermiAssembler.cpp:24-39):
 * fml_seq2fmi + fml_fmi2mag + fml_mag_clean(opt as given) + fml_mag2utg on reads taken as they are. */
int b200_fml_seqs2utg_flat(const b200_fml_opt_t *opt, int64_t n, const char *seqs, const int64_t *off,
                           b200_utgs_t **out);
/* fseq1_t form of fml_assemble; *utg is a malloc'd fml_utg_t-compatible array released by b200_fml_utg_destroy
 * (== fml_utg_destroy, fermi-lite/misc.c:267-276). */
int b200_fml_assemble(const b200_fml_opt_t *opt, int n_seqs, const b200_fseq1_t *seqs, int *n_utg, b200_utg_t **utg);
void b200_fml_utg_destroy(int n_utg, b200_utg_t *utg);

int b200_utgs_view(const b200_utgs_t *u, int *n_utg, const b200_utg_t **utg);   /* pointers owned by the handle */
void b200_utgs_free(b200_utgs_t *u);

/* Stage dumps for parity tests.
 * b200_fmd_*: the FMD-index (replaces fml_seq2fmi -> rld_t, fermi-lite/misc.c:65-128): the BWT as one symbol per
 * byte (0..5 = $ACGTN), cnt[7] / mcnt[7] like rld_t, and rld_rank1a (fermi-lite/rld0.c:402-421) for a batch. */
typedef struct b200_fmd b200_fmd_t;
int b200_fmd_build(int64_t n, const char *seqs, const int64_t *off, b200_fmd_t **out);
int64_t b200_fmd_len(const b200_fmd_t *f);
int b200_fmd_info(const b200_fmd_t *f, uint64_t cnt[7], uint64_t mcnt[7]);
int b200_fmd_bwt(const b200_fmd_t *f, uint8_t *bwt);
int b200_fmd_rank1a(const b200_fmd_t *f, int64_t n_q, const uint64_t *q, uint64_t *ranks /* 6 per query */, int32_t *sym);
void b200_fmd_destroy(b200_fmd_t *f);
/* the unitig graph in mag_g_print's text format (fermi-lite/mag.c:151-194): stage 0 = out of fml_fmi2mag
 * (fermi-lite/unitig.c:406-455), stage 1 = after fml_mag_clean with opt as given; *text is malloc'd. */
int b200_fml_mag_text(const b200_fml_opt_t *opt, int stage, int64_t n, const char *seqs, const int64_t *off,
                      char **text, int64_t *text_len, float *rdist);

/* Device timings (ms) and work counters of the last fermi call on this thread. */
typedef struct b200_fml_stats {
    float ms_count, ms_table, ms_ec, ms_flt, ms_total;
    float ms_fmd, ms_nodes, ms_walk_host, ms_clean_host;   /* index build, overlap records (device); walk, cleaning (host) */
    uint64_t fmd_symbols, n_strings, n_vertices, n_utg;
    uint64_t n_kmers, n_distinct, table_bytes;
    uint64_t n_lookups;      /* count-table probes issued by the correction kernel (16 B each) */
    uint64_t n_spill;        /* reads re-run with the large scratch                            */
    uint64_t ec_codes[8];    /* reads per ecstat_t.ec_code (fermi-lite/bfc.h:23-35), [6] = none */
    int n_launches;
} b200_fml_stats_t;
int b200_fml_last_stats(b200_fml_stats_t *out);

/* ------------------------------------------------------------------ */
/* SAM text of single-end reads                  */
/* replaces mem_reg2sam (bwa/bwamem.c:1034-1086) = mem_gen_alt        */
/* (bwa/bwamem_extra.c:125-173) + mem_aln2sam (bwa/bwamem.c:851-976)  */
/* ------------------------------------------------------------------ */
/* One SAM record per reported alignment of every read of `view` (b200_results_view), rnames[rid] = contig names, reads in order, exactly the text bwa's mem_reg2sam puts
 * in bseq1_t.sam for mem_align1's regions (extra_flag 0, no mate): supplementary records hard-clipped unless
 * MEM_F_SOFTCLIP, MAPQ capped by the primary's, SA / XA (or XB) / pa tags, the unaligned record when nothing reaches opt->T.
 * seqs/names (and optionally quals/comments) in the flat layout of b200_mem_align_batch / b200_fastq_batch_t; a read
 * whose quality or comment range is empty prints '*' / nothing.  *sam is malloc'd and NUL-terminated: free() it.
 * Host-side formatting on all host threads; B200_ERR_LIMIT for MEM_F_REF_HDR (no contig annotations in the index image). */
int b200_results_to_sam(const b200_results_view_t *view, const b200_mem_opt_t *opt, const char *const *rnames, int n_rnames,
                        const char *seqs, const int64_t *seq_off, const char *quals, const int64_t *qual_off,
                        const char *names, const int64_t *name_off, const char *comments, const int64_t *comment_off,
                        char **sam, int64_t *sam_len);

/* ------------------------------------------------------------------ */
/* FASTA/FASTQ ingest                            */
/* replaces FastqReader::Open / GetNextSequence                        */
/* (src/FastqReader.cpp:8-59) = kseq_read over gzread                  */
/* (bwa/kseq.h:176-226): same records, same name/comment split, same   */
/* multi-line, blank-line, CR and truncation behaviour.                */
/* ------------------------------------------------------------------ */
typedef struct b200_fastq b200_fastq_t;
/* One batch of records in the flat layout b200_mem_align_batch() takes:
 * field f of record i = f[f_off[i], f_off[i+1]).  The buffers belong to the reader
 * and stay valid until its next b200_fastq_next_batch / b200_fastq_close. */
typedef struct b200_fastq_batch {
    int64_t n;                                   /* records in this batch                                   */
    const char *seq;     const int64_t *seq_off;
    const char *qual;    const int64_t *qual_off;    /* empty for FASTA records                          */
    const char *name;    const int64_t *name_off;
    const char *comment; const int64_t *comment_off;
    int32_t status;      /* 0: more may follow; 1: end of input; -2: kseq's "truncated quality" (-2) ended the stream */
    int32_t parsed_on_device;                    /* 1 if the batch came from the GPU line parser            */
    const uint8_t *has;  /* per record: bit 0 = kseq's comment string exists, bit 1 = its quality string exists, after this
                          * record -- FastqReader::GetNextSequence only assigns Com / Qual then (src/FastqReader.cpp:49-56) */
} b200_fastq_batch_t;
/* path "-" = stdin, like the reference; gzip or plain.  B200_ERR_IO if the file cannot be opened. */
int b200_fastq_open(const char *path, b200_fastq_t **out);
/* an in-memory (already decompressed) FASTA/FASTQ text; the text is not copied and must outlive the reader */
int b200_fastq_open_mem(const char *text, int64_t len, b200_fastq_t **out);
/* up to max_records records (FastqReader::GetNextSequence x max_records) */
int b200_fastq_next_batch(b200_fastq_t *r, int64_t max_records, b200_fastq_batch_t *out);
void b200_fastq_close(b200_fastq_t *r);
/* FastqReader::GetNextSequence assigns Com / Qual only once kseq has allocated those strings (src/FastqReader.cpp:49-56):
 * bit 0 = a comment has been read, bit 1 = a '+' line has been seen, as of the last record returned. */
int b200_fastq_buffers_seen(const b200_fastq_t *r);
/* Strict four-line FASTQ text parsed on the device: newline positions by a count / scan / write pass, one thread per record for
 * the '@' / '+' / length checks and the name/comment split, bases gathered into the contiguous layout above.  Gives exactly
 * what b200_fastq_next_batch gives on such text; returns B200_ERR_ARG (and no batch) if the text is not strict four-line
 * FASTQ (multi-line records, FASTA, blank lines, truncated last record) -- the caller then uses the stream parser.
 * `text` is a host buffer (pinned or not); copies are inside the call.  Buffers in *out are owned by the reader. */
int b200_fastq_parse_device(b200_fastq_t *r, const char *text, int64_t len, b200_fastq_batch_t *out);

/* Device selection for multi-GPU processes (one process per GPU). */
int b200_set_device(int ordinal);
int b200_device_count(void);

#ifdef __cplusplus
}
#endif
#endif
