/*
 * vg_oracle.h -- CPU ORACLE for the `vargeno geno` hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded, literal restatement of stages E-F of the reference
 * (/root/reference/src/qv.cc:760-1626).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library, and only as
 * the checker.  Nothing under vargeno_b200/ links, imports or calls it.
 *
 * Parity status: PINNED.  tests/golden/ holds outputs of the compiled reference itself
 * (oracle/_ref/vargeno_instr: per-read vote trace, pileup dump, %.17g confidences, VCF) on
 * adversarial synthetic inputs; tests/test_oracle_golden.py checks this restatement against
 * them, and the caller against the known-answer vector of the reference's own
 * test/expected_output (GQ 846).
 */
#ifndef VG_ORACLE_H
#define VG_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vgo_index vgo_index;

/* per-read summary (one per FASTQ record, in file order) */
typedef struct {
	uint32_t flags;      /* VGO_F_* */
	uint32_t target;     /* winning read position X (target_index, qv.cc:1376); 0 if no best */
	uint16_t freq;       /* best->freq (uint8 in the reference) */
	uint16_t n_ref;      /* recorded ref hit contexts of the final pass */
	uint16_t n_snp;      /* recorded snp hit contexts of the final pass */
	uint16_t passes;     /* orientation passes run: 0 (skipped), 1 or 2 */
	uint64_t ctx_hash;   /* order-independent digest of all recorded contexts of the final pass */
} vgo_read_result;

enum {
	VGO_F_SKIPPED   = 1u << 0,  /* N in the first 32K bases: read contributes nothing (qv.cc:815-828) */
	VGO_F_REVCOMPL  = 1u << 1,  /* the final pass was the reverse-complement pass */
	VGO_F_PROCESS   = 1u << 2,  /* process_read (qv.cc:1375) */
	VGO_F_AMBIGUOUS = 1u << 3,  /* index_table.ambiguous after the final pass */
	VGO_F_HASBEST   = 1u << 4   /* index_table.best != NULL after the final pass */
};

typedef struct {
	uint64_t reads, skipped_n, passes, placed;
	uint64_t exact_lookups;     /* query_ref_dict + query_snp_dict calls on the read's own k-mers */
	uint64_t nbr_query_lookups; /* the same calls on substituted k-mers */
	uint64_t nbr_scan_reads;    /* strided entry reads of iterate_ref_dict / iterate_snp_dict */
	uint64_t bf_probes;
	uint64_t lowq_kmers;
	uint64_t events;            /* recorded hit contexts, all passes */
	uint64_t pileup_incr;       /* counter increments attempted (before saturation) */
	uint64_t big_kmers;         /* low-quality k-mers whose ref block was >= BLOCK_SIZE_THRESHOLD */
} vgo_stats;

/* one pileup site as the reference's dense table holds it */
typedef struct {
	uint32_t pos;               /* 1-based position in the concatenation of all contigs */
	uint8_t ref, alt, ref_cnt, alt_cnt, ref_freq, alt_freq, pad0, pad1;
} vgo_site;

/* Index images are the on-disk record layouts (13 B / 40 B / 16 B / 78 B, little endian). */
vgo_index *vgo_index_create(const uint8_t *ref_rec, uint64_t n, const uint32_t *ref_aux, uint64_t aux_n,
                            const uint8_t *snp_rec, uint64_t m, const uint8_t *snp_aux, uint64_t aux_m,
                            const uint64_t *ref_bf, uint64_t ref_bf_bits, uint64_t ref_bf_words,
                            const uint64_t *snp_bf, uint64_t snp_bf_bits, uint64_t snp_bf_words);
void vgo_index_free(vgo_index *ix);

/* exact dictionary query of one k-mer: returns 1 if present; which = 0 ref, 1 snp */
int vgo_lookup(const vgo_index *ix, int which, uint64_t kmer, uint32_t *pos, uint8_t *flag, uint8_t *snp_info);
/* ref HI32 block (jumpgate pair) and snp HI24 block of a k-mer */
void vgo_blocks(const vgo_index *ix, uint64_t kmer, uint32_t *ref_lo, uint32_t *ref_size, uint32_t *snp_lo, uint32_t *snp_size);
int vgo_bf_check(const vgo_index *ix, int which, uint64_t value);

/* Stage E over FASTQ text (complete records).  results may be NULL; trace_path may be NULL, else a trace in
 * the format of oracle/instr.sed is written.  Returns the number of records, or <0: -1 truncated record,
 * -2 base outside ACGTNacgtn (reference aborts), -3 more than 2000 contexts (reference overflows),
 * -4 line longer than 1022 characters, -5 quality line shorter than the k-mer count. */
int64_t vgo_process_fastq(vgo_index *ix, const char *text, uint64_t nbytes, vgo_read_result *results,
                          uint64_t results_cap, const char *trace_path);

void vgo_reset_pileup(vgo_index *ix);
void vgo_get_stats(const vgo_index *ix, vgo_stats *out);
/* sites with ref != 0 || alt != 0 in position order; returns count (call with out == NULL to size) */
uint64_t vgo_get_sites(const vgo_index *ix, vgo_site *out, uint64_t cap);
/* add counters of another shard (saturating at 63) -- multi-GPU reduction check */
void vgo_add_counts(vgo_index *ix, const vgo_site *sites, uint64_t n);

/* choose_best_genotype, qv.cc:1789-1848: returns GTYPE (0 none, 1 ref, 2 alt, 3 het) */
int vgo_call(int ref_cnt, int alt_cnt, uint8_t ref_freq, uint8_t alt_freq, double *confidence);
/* (int)(-1*10*log(conf)), qv.cc:1681 */
int vgo_gq(double confidence);
/* the caller's tables, for checking the product's host-built tables: g[64*64*3], poisson[127] */
void vgo_tables(double *g, double *poisson);

uint64_t vgo_ctx_digest(int list_id, uint32_t position, uint32_t kmer_pos, uint64_t kmer, uint32_t modified_pos);

#ifdef __cplusplus
}
#endif
#endif
