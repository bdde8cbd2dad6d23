/*
 * vg_oracle.c -- CPU ORACLE (test infrastructure, see vg_oracle.h).
 *
 * Literal single-threaded restatement of the reference's read loop and caller.  Every block
 * cites the reference lines it follows (paths relative to /root/reference).  It deliberately
 * keeps the reference's evaluation ORDER (running best/ambiguous vote state, uint8 frequency
 * wrap, hit-context list order) so that the CUDA path's order-independent formulation is
 * checked against the real thing, not against itself.  Layout is NOT copied: block bounds come
 * from binary searches over the sorted 64-bit k-mers instead of the 16 GiB jumpgate
 * (src/qv.cc:531-584); by construction jumpgate[h] == lower_bound(h << 32).
 *
 * Reference behaviours kept on purpose: F6 (SNP Bloom filter content is
 * whatever the file says), F8 (k-mer i gated on qual[i]), F9 (non-overlapping 32-mers, quality
 * not reversed), F10 (saturation at 63), F13 (strided block scan: entry lo+S*t examined, entry
 * lo+t reported; S = 9 ref / 11 snp).  One documented divergence: a strided read past the end of
 * the array (undefined behaviour in the reference) is "no match" here.
 */
#include "vg_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAX_COV 63                     /* src/vartype.h:27 */
#define POS_AMBIGUOUS 0xFFFFFFFFu      /* src/vartype.h:33 */
#define AUX_COLS 10                    /* src/vartype.h:93 */
#define BLOCK_SIZE_THRESHOLD 100       /* src/vartype.h:103 */
#define QUALITY_SCORE '8'              /* src/vartype.h:17 */
#define MAX_HITS 2000                  /* src/qv.cc:709 */
#define NO_MODIFICATION 10086u         /* src/qv.cc:710 */
#define REF_STRIDE 9                   /* sizeof(struct kmer_entry), src/vartype.h:64-73, qv.cc:356-359 */
#define SNP_STRIDE 11                  /* sizeof(struct snp_kmer_entry), src/vartype.h:75-80, qv.cc:445-448 */

#define HI32(k) ((uint32_t)((k) >> 32))
#define LO32(k) ((uint32_t)((k) & 0xFFFFFFFFu))
#define HI24(k) ((uint32_t)((k) >> 40))
#define LO40(k) ((k) & 0xFFFFFFFFFFull)
#define SNP_INFO_POS(s) (((s) & 0xF8) >> 3)
#define SNP_INFO_REF(s) ((s) & 0x07)

typedef struct { uint64_t kmer; uint32_t position, kmer_pos, modified_pos; } ctx_t;

typedef struct { uint8_t ref, alt, ref_freq, alt_freq; uint8_t ref_cnt, alt_cnt; } pile_t;

struct vgo_index {
	uint64_t n, aux_n, m, aux_m;
	uint64_t *ref_kmer; uint32_t *ref_pos; uint8_t *ref_flag; uint32_t *ref_aux;
	uint64_t *snp_kmer; uint32_t *snp_pos; uint8_t *snp_info; uint8_t *snp_flag;
	uint32_t *snp_aux_pos; uint8_t *snp_aux_info;
	const uint64_t *ref_bf; uint64_t ref_bf_bits, ref_bf_words;
	const uint64_t *snp_bf; uint64_t snp_bf_bits, snp_bf_words;
	pile_t *pile; uint64_t pile_size;
	vgo_stats st;
};

static uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static uint32_t rd32(const uint8_t *p) { return (uint32_t)rd16(p) | ((uint32_t)rd16(p + 2) << 16); }
static uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

static unsigned kmer_get_base(uint64_t kmer, unsigned base) { return (unsigned)((kmer >> (2 * base)) & 3u); } /* util.c:129 */

vgo_index *vgo_index_create(const uint8_t *ref_rec, uint64_t n, const uint32_t *ref_aux, uint64_t aux_n,
                            const uint8_t *snp_rec, uint64_t m, const uint8_t *snp_aux, uint64_t aux_m,
                            const uint64_t *ref_bf, uint64_t ref_bf_bits, uint64_t ref_bf_words,
                            const uint64_t *snp_bf, uint64_t snp_bf_bits, uint64_t snp_bf_words)
{
	vgo_index *ix = (vgo_index *)calloc(1, sizeof(*ix));
	ix->n = n; ix->aux_n = aux_n; ix->m = m; ix->aux_m = aux_m;
	ix->ref_kmer = (uint64_t *)malloc((n + 1) * 8);
	ix->ref_pos = (uint32_t *)malloc((n + 1) * 4);
	ix->ref_flag = (uint8_t *)malloc(n + 1);
	ix->ref_aux = (uint32_t *)malloc((aux_n + 1) * AUX_COLS * 4);
	ix->snp_kmer = (uint64_t *)malloc((m + 1) * 8);
	ix->snp_pos = (uint32_t *)malloc((m + 1) * 4);
	ix->snp_info = (uint8_t *)malloc(m + 1);
	ix->snp_flag = (uint8_t *)malloc(m + 1);
	ix->snp_aux_pos = (uint32_t *)malloc((aux_m + 1) * AUX_COLS * 4);
	ix->snp_aux_info = (uint8_t *)malloc((aux_m + 1) * AUX_COLS);
	ix->ref_bf = ref_bf; ix->ref_bf_bits = ref_bf_bits; ix->ref_bf_words = ref_bf_words;
	ix->snp_bf = snp_bf; ix->snp_bf_bits = snp_bf_bits; ix->snp_bf_words = snp_bf_words;

	/* qv.cc:541-555 -- max_pos runs over every record's pos field, aux indices and POS_AMBIGUOUS included;
	 * the dense table is only ever addressed at genome positions, so size it by what can be addressed. */
	uint64_t max_pos = 0;
	for (uint64_t i = 0; i < n; i++) {
		const uint8_t *r = ref_rec + 13 * i;
		ix->ref_kmer[i] = rd64(r); ix->ref_pos[i] = rd32(r + 8); ix->ref_flag[i] = r[12];
		if (r[12] == 0 && ix->ref_pos[i] != POS_AMBIGUOUS && ix->ref_pos[i] > max_pos) max_pos = ix->ref_pos[i];
	}
	for (uint64_t i = 0; i < aux_n * AUX_COLS; i++) {
		ix->ref_aux[i] = ref_aux[i];
		if (ref_aux[i] > max_pos) max_pos = ref_aux[i];
	}
	for (uint64_t i = 0; i < m; i++) {
		const uint8_t *r = snp_rec + 16 * i;
		ix->snp_kmer[i] = rd64(r); ix->snp_pos[i] = rd32(r + 8); ix->snp_info[i] = r[12]; ix->snp_flag[i] = r[13];
		if (r[13] == 0 && ix->snp_pos[i] != POS_AMBIGUOUS && ix->snp_pos[i] > max_pos) max_pos = ix->snp_pos[i];
	}
	for (uint64_t i = 0; i < aux_m; i++) {
		const uint8_t *r = snp_aux + 78 * i + 8;
		for (int c = 0; c < AUX_COLS; c++) {
			ix->snp_aux_pos[i * AUX_COLS + c] = rd32(r + 7 * c);
			ix->snp_aux_info[i * AUX_COLS + c] = r[7 * c + 4];
			if (ix->snp_aux_pos[i * AUX_COLS + c] > max_pos) max_pos = ix->snp_aux_pos[i * AUX_COLS + c];
		}
	}
	ix->pile_size = max_pos + 32 + 1;                       /* qv.cc:602 */
	ix->pile = (pile_t *)calloc(ix->pile_size, sizeof(pile_t));

	/* static SNP-site fields, file order, last writer wins -- qv.cc:637-659 */
	for (uint64_t i = 0; i < m; i++) {
		const uint8_t *r = snp_rec + 16 * i;
		const unsigned snp = r[12];
		const unsigned iref = SNP_INFO_REF(snp);
		if ((iref & 4) == 0 && ix->snp_pos[i] != POS_AMBIGUOUS && r[13] == 0) {
			const uint64_t sp = (uint64_t)ix->snp_pos[i] + SNP_INFO_POS(snp);
			if (sp >= ix->pile_size) continue;              /* reference reallocs (buggy, qv.cc:649-654): out of contract */
			ix->pile[sp].ref = (uint8_t)iref;
			ix->pile[sp].alt = (uint8_t)kmer_get_base(ix->snp_kmer[i], SNP_INFO_POS(snp));
			ix->pile[sp].ref_freq = r[14];
			ix->pile[sp].alt_freq = r[15];
		}
	}
	return ix;
}

void vgo_index_free(vgo_index *ix)
{
	if (!ix) return;
	free(ix->ref_kmer); free(ix->ref_pos); free(ix->ref_flag); free(ix->ref_aux);
	free(ix->snp_kmer); free(ix->snp_pos); free(ix->snp_info); free(ix->snp_flag);
	free(ix->snp_aux_pos); free(ix->snp_aux_info); free(ix->pile); free(ix);
}

static uint64_t lower_bound64(const uint64_t *a, uint64_t n, uint64_t key)
{
	uint64_t lo = 0, hi = n;
	while (lo < hi) { uint64_t mid = lo + (hi - lo) / 2; if (a[mid] < key) lo = mid + 1; else hi = mid; }
	return lo;
}

/* ref block [lo, hi) of prefix HI32 -- the jumpgate pair of qv.cc:219-233 */
static void ref_block(const vgo_index *ix, uint64_t key, uint64_t *lo, uint64_t *hi)
{
	const uint32_t h = HI32(key);
	*lo = lower_bound64(ix->ref_kmer, ix->n, (uint64_t)h << 32);
	*hi = (h == 0xFFFFFFFFu) ? ix->n : lower_bound64(ix->ref_kmer, ix->n, ((uint64_t)h + 1) << 32);
}
static void snp_block(const vgo_index *ix, uint64_t key, uint64_t *lo, uint64_t *hi)
{
	const uint32_t h = HI24(key);
	*lo = lower_bound64(ix->snp_kmer, ix->m, (uint64_t)h << 40);
	*hi = (h == 0xFFFFFFu) ? ix->m : lower_bound64(ix->snp_kmer, ix->m, ((uint64_t)h + 1) << 40);
}

/* query_ref_dict qv.cc:206-240 / query_snp_dict qv.cc:385-411: entry rank or -1 (dict k-mers are unique) */
static int64_t query_ref(const vgo_index *ix, uint64_t key)
{
	uint64_t i = lower_bound64(ix->ref_kmer, ix->n, key);
	return (i < ix->n && ix->ref_kmer[i] == key) ? (int64_t)i : -1;
}
static int64_t query_snp(const vgo_index *ix, uint64_t key)
{
	uint64_t i = lower_bound64(ix->snp_kmer, ix->m, key);
	return (i < ix->m && ix->snp_kmer[i] == key) ? (int64_t)i : -1;
}

int vgo_lookup(const vgo_index *ix, int which, uint64_t kmer, uint32_t *pos, uint8_t *flag, uint8_t *snp_info)
{
	int64_t i = which ? query_snp(ix, kmer) : query_ref(ix, kmer);
	if (i < 0) return 0;
	if (which) { *pos = ix->snp_pos[i]; *flag = ix->snp_flag[i]; *snp_info = ix->snp_info[i]; }
	else { *pos = ix->ref_pos[i]; *flag = ix->ref_flag[i]; *snp_info = 0; }
	return 1;
}

void vgo_blocks(const vgo_index *ix, uint64_t kmer, uint32_t *ref_lo, uint32_t *ref_size, uint32_t *snp_lo, uint32_t *snp_size)
{
	uint64_t lo, hi;
	ref_block(ix, kmer, &lo, &hi); *ref_lo = (uint32_t)lo; *ref_size = (uint32_t)(hi - lo);
	snp_block(ix, kmer, &lo, &hi); *snp_lo = (uint32_t)lo; *snp_size = (uint32_t)(hi - lo);
}

/* src/generate_bf.h:126-142 */
static uint32_t hash32(uint32_t x)
{
	x = ((x >> 16) ^ x) * 0x45d9f3bu;
	x = ((x >> 16) ^ x) * 0x45d9f3bu;
	x = (x >> 16) ^ x;
	return x;
}
static uint64_t hash40(uint64_t x)
{
	x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
	x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
	x = x ^ (x >> 31);
	return x;
}
static int bf_bit(const uint64_t *w, uint64_t words, uint64_t bit)
{
	const uint64_t wi = bit >> 6;
	if (wi >= words) return 0;   /* callers may pass only the addressable prefix of the ref filter */
	return (int)((w[wi] >> (bit & 63)) & 1u);
}
/* BloomFilter::check_value, src/generate_bf.h:112-121; value_range 32 for ref, 40 for snp (qv.cc:2140-2141) */
int vgo_bf_check(const vgo_index *ix, int which, uint64_t v)
{
	if (which == 0) return bf_bit(ix->ref_bf, ix->ref_bf_words, (uint64_t)hash32((uint32_t)v) % ix->ref_bf_bits);
	return bf_bit(ix->snp_bf, ix->snp_bf_words, hash40(v) % ix->snp_bf_bits);
}

/* one_hamming_distance_32/64, qv.cc:267-312 with the masks of qv.cc:2146-2158 and diff_base_dictionary
 * qv.cc:2160-2173: true iff x = a^b is non-zero and confined to one 2-bit slot; *d = slot index */
static int one_base_apart(uint64_t a, uint64_t b, int *d)
{
	const uint64_t x = a ^ b;
	if (x == 0) return 0;
	if ((x & (x - 1)) == 0) { *d = __builtin_ctzll(x) / 2; return 1; }
	const uint64_t y = x & 0xAAAAAAAAAAAAAAAAull;
	if ((y & (y - 1)) != 0) return 0;
	const uint64_t z = x & 0x5555555555555555ull;
	if ((z & (z - 1)) != 0) return 0;
	if (y == (z << 1)) { *d = __builtin_ctzll(x) / 2; return 1; }
	return 0;
}

/* ---- position vote: IndexTable + index_2_kmer_pos_set, qv.cc:57-178 ---- */
typedef struct { uint32_t index; uint8_t freq; int nk; int cap; uint32_t *kpos; } vote_entry;
typedef struct { vote_entry *e; int n, cap; int best; int ambiguous; } vote_t;

static void vote_clear(vote_t *v)
{
	for (int i = 0; i < v->n; i++) free(v->e[i].kpos);
	v->n = 0; v->best = -1; v->ambiguous = 0;
}
static int vote_find(const vote_t *v, uint32_t index)
{
	for (int i = 0; i < v->n; i++) if (v->e[i].index == index) return i;
	return -1;
}
/* improved_index_table_add, qv.cc:132-178 */
static void vote_add(vote_t *v, uint32_t index, uint32_t kmer_pos, int is_neighbor)
{
	int t = vote_find(v, index);
	if (is_neighbor && t < 0) return;                       /* :134-139 */
	if (t < 0) {                                            /* :154-161 (entry and map key are created together) */
		if (v->n == v->cap) { v->cap = v->cap ? 2 * v->cap : 64; v->e = (vote_entry *)realloc(v->e, (size_t)v->cap * sizeof(vote_entry)); }
		t = v->n++;
		v->e[t].index = index; v->e[t].freq = 1; v->e[t].nk = 0; v->e[t].cap = 0; v->e[t].kpos = NULL;
	} else {
		++v->e[t].freq;                                     /* :148 uint8_t: wraps at 256 */
	}
	vote_entry *e = &v->e[t];
	int seen = 0;
	for (int i = 0; i < e->nk; i++) if (e->kpos[i] == kmer_pos) { seen = 1; break; }
	if (!seen) {                                            /* :163 */
		if (e->nk == e->cap) { e->cap = e->cap ? 2 * e->cap : 4; e->kpos = (uint32_t *)realloc(e->kpos, (size_t)e->cap * 4); }
		e->kpos[e->nk++] = kmer_pos;
	}
	if (e->nk <= 1) return;                                 /* :165 */
	if (v->best < 0) { v->best = t; v->ambiguous = 0; }     /* :167-177 */
	else if (t == v->best) v->ambiguous = 0;
	else if (e->freq == v->e[v->best].freq) v->ambiguous = 1;
	else if (e->freq > v->e[v->best].freq) { v->best = t; v->ambiguous = 0; }
}

/* ---- per-read state ---- */
typedef struct {
	vgo_index *ix;
	ctx_t ref_ctx[MAX_HITS], snp_ctx[MAX_HITS];
	size_t n_ref, n_snp;
	vote_t vote;
	int overflow;
} rs_t;

static void emit_ref(rs_t *s, uint64_t kmer, uint32_t pos, uint32_t offset, uint32_t mod, int is_neighbor)
{
	if (s->n_ref >= MAX_HITS) { s->overflow = 1; return; }
	s->ref_ctx[s->n_ref++] = (ctx_t){ kmer, pos - offset, pos, mod };
	s->ix->st.events++;
	vote_add(&s->vote, pos - offset, pos, is_neighbor);
}
static void emit_snp(rs_t *s, uint64_t kmer, uint32_t pos, uint32_t offset, uint32_t mod, int is_neighbor)
{
	if (s->n_snp >= MAX_HITS) { s->overflow = 1; return; }
	s->snp_ctx[s->n_snp++] = (ctx_t){ kmer, pos - offset, pos, mod };
	s->ix->st.events++;
	vote_add(&s->vote, pos - offset, pos, is_neighbor);
}

static int site_is_zero(const vgo_index *ix, uint64_t p)
{
	return p < ix->pile_size ? (ix->pile[p].ref == 0 && ix->pile[p].alt == 0) : 1;
}

/* exact hit of the read's own k-mer: qv.cc:850-890 (ref), :897-937 (snp) */
static void exact_ref(rs_t *s, int64_t h, uint64_t kmer, uint32_t offset)
{
	const vgo_index *ix = s->ix;
	if (h < 0 || ix->ref_pos[h] == POS_AMBIGUOUS) return;
	if (ix->ref_flag[h] == 0) emit_ref(s, kmer, ix->ref_pos[h], offset, NO_MODIFICATION, 0);
	else {
		const uint32_t *pl = &ix->ref_aux[(uint64_t)ix->ref_pos[h] * AUX_COLS];
		for (int c = 0; c < AUX_COLS; c++) { if (pl[c] == 0) break; emit_ref(s, kmer, pl[c], offset, NO_MODIFICATION, 0); }
	}
}
static void exact_snp(rs_t *s, int64_t h, uint64_t kmer, uint32_t offset)
{
	const vgo_index *ix = s->ix;
	if (h < 0 || ix->snp_pos[h] == POS_AMBIGUOUS) return;
	if (ix->snp_flag[h] == 0) emit_snp(s, kmer, ix->snp_pos[h], offset, NO_MODIFICATION, 0);
	else {
		const uint32_t *pl = &ix->snp_aux_pos[(uint64_t)ix->snp_pos[h] * AUX_COLS];
		for (int c = 0; c < AUX_COLS; c++) { if (pl[c] == 0) break; emit_snp(s, kmer, pl[c], offset, NO_MODIFICATION, 0); }
	}
}
/* neighbour hit entry h with modified base d: qv.cc:985-1046 and twins :1131-1172, :1227-1291 */
static void nbr_ref(rs_t *s, int64_t h, uint64_t nb, uint32_t d, uint32_t offset)
{
	const vgo_index *ix = s->ix;
	if (h < 0 || ix->ref_pos[h] == POS_AMBIGUOUS) return;
	if (ix->ref_flag[h] == 0) {
		if (site_is_zero(ix, (uint64_t)ix->ref_pos[h] + d)) emit_ref(s, nb, ix->ref_pos[h], offset, d, 1);
	} else if (ix->ref_flag[h] == 1) {
		const uint32_t *pl = &ix->ref_aux[(uint64_t)ix->ref_pos[h] * AUX_COLS];
		for (int c = 0; c < AUX_COLS; c++) {
			if (pl[c] == 0) break;
			if (site_is_zero(ix, (uint64_t)pl[c] + d)) emit_ref(s, nb, pl[c], offset, d, 1);
		}
	}
}
/* qv.cc:1053-1101 and twins :1178-1207, :1311-1356 */
static void nbr_snp(rs_t *s, int64_t h, uint64_t nb, uint32_t d, uint32_t offset)
{
	const vgo_index *ix = s->ix;
	if (h < 0 || ix->snp_pos[h] == POS_AMBIGUOUS) return;
	if (ix->snp_flag[h] == 0 && SNP_INFO_POS(ix->snp_info[h]) != d) emit_snp(s, nb, ix->snp_pos[h], offset, d, 1);
	else if (ix->snp_flag[h] == 1) {
		const uint32_t *pl = &ix->snp_aux_pos[(uint64_t)ix->snp_pos[h] * AUX_COLS];
		const uint8_t *sl = &ix->snp_aux_info[(uint64_t)ix->snp_pos[h] * AUX_COLS];
		for (int c = 0; c < AUX_COLS; c++) {
			if (pl[c] == 0) break;
			if (SNP_INFO_POS(sl[c]) != d) emit_snp(s, nb, pl[c], offset, d, 1);
		}
	}
}

/* pileup update for one recorded context: qv.cc:1386-1440 / :1447-1501 */
static void pile_update(vgo_index *ix, const ctx_t *c)
{
	for (unsigned b = 0; b < 32; b++) {
		if (b == c->modified_pos) continue;
		const uint64_t p = (uint64_t)c->kmer_pos + b;
		if (p >= ix->pile_size) continue;
		pile_t *e = &ix->pile[p];
		if (e->ref != e->alt) {
			const unsigned base = kmer_get_base(c->kmer, b);
			if (base == e->ref) { ix->st.pileup_incr++; if (e->ref_cnt != MAX_COV) ++e->ref_cnt; }
			else if (base == e->alt) { ix->st.pileup_incr++; if (e->alt_cnt != MAX_COV) ++e->alt_cnt; }
		}
	}
}

static uint64_t mix64(uint64_t x)
{
	x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
	x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
	return x ^ (x >> 31);
}
uint64_t vgo_ctx_digest(int list_id, uint32_t position, uint32_t kmer_pos, uint64_t kmer, uint32_t modified_pos)
{
	uint64_t h = mix64(kmer + 0x9E3779B97F4A7C15ull);
	h = mix64(h ^ (((uint64_t)position << 32) | kmer_pos));
	h = mix64(h ^ (((uint64_t)modified_pos << 8) | (uint64_t)list_id));
	return h;
}

/* One orientation pass over the encoded k-mers: qv.cc:830-1367 */
static void run_pass(rs_t *s, const uint64_t *kmers, size_t kmer_count, const char *qual)
{
	vgo_index *ix = s->ix;
	s->n_ref = s->n_snp = 0;
	for (size_t ki = 0; ki < kmer_count; ki++) {
		const uint64_t kmer = kmers[ki];
		const char qual_char = qual[ki];                    /* qv.cc:836 -- the ki-th quality CHARACTER (F8) */
		const uint32_t offset = (uint32_t)(32 * ki);

		const int64_t ref_hit = query_ref(ix, kmer);        /* :840-841 */
		const int64_t snp_hit = query_snp(ix, kmer);
		ix->st.exact_lookups += 2;
		uint64_t rlo, rhi;
		ref_block(ix, kmer, &rlo, &rhi);
		const uint64_t block_size = rhi - rlo;              /* check_block_size :242-264, :843 */

		exact_ref(s, ref_hit, kmer, offset);
		exact_snp(s, snp_hit, kmer, offset);

		if ((signed char)qual_char - QUALITY_SCORE >= 0) continue;   /* :943 (char is signed on x86-64) */
		ix->st.lowq_kmers++;

		uint32_t ref_search_bound = 64, snp_search_bound = 64;       /* :946-956 */
		ix->st.bf_probes += 2;
		if (vgo_bf_check(ix, 0, LO32(kmer)) == 0) ref_search_bound = 32;
		if (vgo_bf_check(ix, 1, LO40(kmer)) == 0) snp_search_bound = 40;
		const int big = block_size >= BLOCK_SIZE_THRESHOLD;
		if (big) ix->st.big_kmers++;

		if (big) {                                          /* :962-1109 lower half by exact queries */
			for (unsigned i = 0; i < 32; i += 2) {
				const unsigned d = i / 2;
				const uint64_t mask = 3ull << i;
				const uint64_t base = (kmer & mask) >> i;
				for (uint64_t j = 0; j < 4; j++) {
					if (j == base) continue;
					const uint64_t nb = (kmer & ~mask) | (j << i);
					ix->st.nbr_query_lookups += 2;
					const int64_t rh = query_ref(ix, nb);
					const int64_t sh = query_snp(ix, nb);
					nbr_ref(s, rh, nb, d, offset);
					nbr_snp(s, sh, nb, d, offset);
				}
			}
		} else {                                            /* :1110-1209 lower half by the strided block scan (F13) */
			/* iterate_ref_dict :316-376 -- collect first, then process (:1124-1173) */
			int64_t hit[BLOCK_SIZE_THRESHOLD]; uint64_t nbv[BLOCK_SIZE_THRESHOLD]; int dd[BLOCK_SIZE_THRESHOLD]; int nh = 0;
			for (uint64_t t = 0; t < block_size; t++) {
				const uint64_t e = rlo + REF_STRIDE * t;
				ix->st.nbr_scan_reads++;
				if (e >= ix->n) continue;                   /* UB in the reference; "no match" here */
				const uint32_t entry_lo = LO32(ix->ref_kmer[e]);
				int d;
				if (one_base_apart(LO32(kmer), entry_lo, &d)) {
					hit[nh] = (int64_t)(rlo + t);
					nbv[nh] = ((uint64_t)HI32(kmer) << 32) | entry_lo;
					dd[nh] = d; nh++;
				}
			}
			for (int h = 0; h < nh; h++) nbr_ref(s, hit[h], nbv[h], (uint32_t)dd[h], offset);

			/* iterate_snp_dict :413-464, then :1178-1208 */
			uint64_t slo, shi;
			snp_block(ix, kmer, &slo, &shi);
			nh = 0;
			for (uint64_t t = 0; t < shi - slo; t++) {
				const uint64_t e = slo + SNP_STRIDE * t;
				ix->st.nbr_scan_reads++;
				if (e >= ix->m) continue;
				const uint64_t entry_lo = LO40(ix->snp_kmer[e]);
				int d;
				if (one_base_apart(LO40(kmer), entry_lo, &d)) {
					if (nh >= BLOCK_SIZE_THRESHOLD) { s->overflow = 1; break; }   /* reference overflows its arrays */
					hit[nh] = (int64_t)(slo + t);
					nbv[nh] = ((uint64_t)HI24(kmer) << 40) | entry_lo;
					dd[nh] = d; nh++;
				}
			}
			for (int h = 0; h < nh; h++) nbr_snp(s, hit[h], nbv[h], (uint32_t)dd[h], offset);
		}

		for (unsigned i = 32; i < 64; i += 2) {             /* :1213-1365 upper half by exact queries */
			const unsigned d = i / 2;
			const uint64_t mask = 3ull << i;
			const uint64_t base = (kmer & mask) >> i;
			for (uint64_t j = 0; j < 4; j++) {
				if (j == base) continue;
				const uint64_t nb = (kmer & ~mask) | (j << i);
				if (i < ref_search_bound) {                 /* :1225 */
					ix->st.nbr_query_lookups++;
					nbr_ref(s, query_ref(ix, nb), nb, d, offset);
				}
				if (big || i >= 40) {                       /* :1305 */
					if (i >= snp_search_bound) continue;    /* :1307 */
					ix->st.nbr_query_lookups++;
					nbr_snp(s, query_snp(ix, nb), nb, d, offset);
				}
			}
		}
	}
}

int64_t vgo_process_fastq(vgo_index *ix, const char *text, uint64_t nbytes, vgo_read_result *results,
                          uint64_t results_cap, const char *trace_path)
{
	FILE *tr = trace_path ? fopen(trace_path, "w") : NULL;
	rs_t *s = (rs_t *)calloc(1, sizeof(rs_t));
	s->ix = ix; s->vote.best = -1;
	int64_t rc = 0;
	uint64_t p = 0, ord = 0;
	char rbuf[1024];
	uint64_t kmers[32];

	while (p < nbytes) {
		/* four fgets, qv.cc:760-763 */
		const char *line[4]; uint64_t len[4]; int has_nl[4];
		int ok = 1;
		for (int l = 0; l < 4; l++) {
			if (p >= nbytes) { ok = 0; break; }
			const char *e = (const char *)memchr(text + p, '\n', nbytes - p);
			line[l] = text + p;
			if (e) { len[l] = (uint64_t)(e - (text + p)); has_nl[l] = 1; p += len[l] + 1; }
			else { len[l] = nbytes - p; has_nl[l] = 0; p = nbytes; }
			if (len[l] > 1022) { rc = -4; goto done; }
		}
		if (!ok) { rc = -1; goto done; }
		vgo_read_result res; memset(&res, 0, sizeof(res));
		ix->st.reads++;

		/* qv.cc:778-779: strlen(read) - 1 */
		const uint64_t read_len_true = has_nl[1] ? len[1] : (len[1] ? len[1] - 1 : 0);
		const uint64_t rlen = (read_len_true / 32) * 32;
		const size_t kmer_count = (size_t)(rlen / 32);
		if (len[3] < kmer_count) { rc = -5; goto done; }
		memcpy(rbuf, line[1], rlen);
		const char *qual = line[3];

		int revcompl = 0, process_read = 0, skipped = 0;
		uint32_t target_index = 0;
		for (;;) {
			if (revcompl) {                                 /* qv.cc:787-806 over the first len bases only */
				char tmp[1024];
				for (uint64_t i = 0; i < rlen; i++) {
					char r = 0;
					switch (rbuf[i]) {
					case 'a': case 'A': r = 'T'; break;
					case 'c': case 'C': r = 'G'; break;
					case 'g': case 'G': r = 'C'; break;
					case 't': case 'T': r = 'A'; break;
					default: skipped = 1; break;
					}
					if (skipped) break;
					tmp[rlen - i - 1] = r;
				}
				if (skipped) break;
				memcpy(rbuf, tmp, rlen);
			}
			/* encode, qv.cc:810-828 with encode_kmer util.c:89-111 (scans base 31 down to 0) */
			for (size_t k = 0; k < kmer_count && !skipped; k++) {
				uint64_t enc = 0;
				for (int b = 31; b >= 0; b--) {
					enc <<= 2;
					switch (rbuf[32 * k + b]) {
					case 'A': case 'a': break;
					case 'C': case 'c': enc |= 1; break;
					case 'G': case 'g': enc |= 2; break;
					case 'T': case 't': enc |= 3; break;
					case 'N': case 'n': skipped = 1; break;
					default: rc = -2; goto done;
					}
					if (skipped) break;
				}
				kmers[k] = enc;
			}
			if (skipped) break;

			ix->st.passes++;
			res.passes++;
			run_pass(s, kmers, kmer_count, qual);
			if (s->overflow) { rc = -3; goto done; }

			/* qv.cc:1375-1376 */
			const vote_t *v = &s->vote;
			process_read = (v->best >= 0 && v->e[v->best].freq > 1 && !v->ambiguous);
			target_index = v->best >= 0 ? v->e[v->best].index : 0;
			for (size_t i = 0; i < s->n_ref; i++)
				if (process_read && s->ref_ctx[i].position == target_index) pile_update(ix, &s->ref_ctx[i]);
			for (size_t i = 0; i < s->n_snp; i++)
				if (process_read && s->snp_ctx[i].position == target_index) pile_update(ix, &s->snp_ctx[i]);
			if (!process_read && !revcompl) {               /* :1504-1510 */
				revcompl = 1;
				vote_clear(&s->vote);
				continue;
			}
			break;
		}

		if (skipped) {
			res.flags = VGO_F_SKIPPED;
			ix->st.skipped_n++;
		} else {
			const vote_t *v = &s->vote;
			res.flags = (revcompl ? VGO_F_REVCOMPL : 0) | (process_read ? VGO_F_PROCESS : 0) |
			            (v->ambiguous ? VGO_F_AMBIGUOUS : 0) | (v->best >= 0 ? VGO_F_HASBEST : 0);
			res.target = target_index;
			res.freq = v->best >= 0 ? v->e[v->best].freq : 0;
			res.n_ref = (uint16_t)s->n_ref; res.n_snp = (uint16_t)s->n_snp;
			uint64_t dg = 0;
			for (size_t i = 0; i < s->n_ref; i++)
				dg += vgo_ctx_digest(0, s->ref_ctx[i].position, s->ref_ctx[i].kmer_pos, s->ref_ctx[i].kmer, s->ref_ctx[i].modified_pos);
			for (size_t i = 0; i < s->n_snp; i++)
				dg += vgo_ctx_digest(1, s->snp_ctx[i].position, s->snp_ctx[i].kmer_pos, s->snp_ctx[i].kmer, s->snp_ctx[i].modified_pos);
			res.ctx_hash = dg;
			if (process_read) ix->st.placed++;
			if (tr) {                                       /* same format as oracle/instr.sed */
				fprintf(tr, "R %lu %d %d %u %d %d %zu %zu\n", (unsigned long)ord, revcompl, process_read, target_index,
				        (int)res.freq, (int)v->ambiguous, s->n_ref, s->n_snp);
				for (size_t i = 0; i < s->n_ref; i++)
					fprintf(tr, "r %u %u %lu %u\n", s->ref_ctx[i].position, s->ref_ctx[i].kmer_pos,
					        (unsigned long)s->ref_ctx[i].kmer, s->ref_ctx[i].modified_pos);
				for (size_t i = 0; i < s->n_snp; i++)
					fprintf(tr, "s %u %u %lu %u\n", s->snp_ctx[i].position, s->snp_ctx[i].kmer_pos,
					        (unsigned long)s->snp_ctx[i].kmer, s->snp_ctx[i].modified_pos);
			}
		}
		vote_clear(&s->vote);
		if (results && ord < results_cap) results[ord] = res;
		ord++;
	}
	rc = (int64_t)ord;
done:
	vote_clear(&s->vote);
	free(s->vote.e);
	free(s);
	if (tr) fclose(tr);
	return rc;
}

void vgo_reset_pileup(vgo_index *ix)
{
	for (uint64_t i = 0; i < ix->pile_size; i++) ix->pile[i].ref_cnt = ix->pile[i].alt_cnt = 0;
	memset(&ix->st, 0, sizeof(ix->st));
}

void vgo_get_stats(const vgo_index *ix, vgo_stats *out) { *out = ix->st; }

uint64_t vgo_get_sites(const vgo_index *ix, vgo_site *out, uint64_t cap)
{
	uint64_t n = 0;
	for (uint64_t i = 0; i < ix->pile_size; i++) {
		const pile_t *e = &ix->pile[i];
		if (e->ref == 0 && e->alt == 0) continue;
		if (out && n < cap) out[n] = (vgo_site){ (uint32_t)i, e->ref, e->alt, e->ref_cnt, e->alt_cnt, e->ref_freq, e->alt_freq, 0, 0 };
		n++;
	}
	return n;
}

void vgo_add_counts(vgo_index *ix, const vgo_site *sites, uint64_t n)
{
	for (uint64_t i = 0; i < n; i++) {
		pile_t *e = &ix->pile[sites[i].pos];
		unsigned r = e->ref_cnt + sites[i].ref_cnt, a = e->alt_cnt + sites[i].alt_cnt;
		e->ref_cnt = (uint8_t)(r > MAX_COV ? MAX_COV : r);
		e->alt_cnt = (uint8_t)(a > MAX_COV ? MAX_COV : a);
	}
}

/* ---- caller: choose_best_genotype, qv.cc:1789-1848 ---- */
#define ERR_RATE 0.01
#define AVG_COV 7.1
static double g_cache[MAX_COV + 1][MAX_COV + 1][3];
static double g_poisson[2 * MAX_COV + 1];
static int g_init = 0;

static void init_tables(void)
{
	if (g_init) return;
	for (int r = 0; r <= MAX_COV; r++)
		for (int a = 0; a <= MAX_COV; a++) {
			g_cache[r][a][0] = pow(1.0 - ERR_RATE, r) * pow(ERR_RATE, a);
			g_cache[r][a][1] = pow(0.5, r + a);
			g_cache[r][a][2] = pow(ERR_RATE, r) * pow(1.0 - ERR_RATE, a);
		}
	const double M = exp(-AVG_COV);
	for (int i = 0; i <= 2 * MAX_COV; i++) g_poisson[i] = (M * pow(AVG_COV, i)) / exp(lgamma(i + 1.0));
	g_init = 1;
}

void vgo_tables(double *g, double *poisson)
{
	init_tables();
	memcpy(g, g_cache, sizeof(g_cache));
	memcpy(poisson, g_poisson, sizeof(g_poisson));
}

int vgo_call(int ref_cnt, int alt_cnt, uint8_t ref_freq_enc, uint8_t alt_freq_enc, double *confidence)
{
	init_tables();
	*confidence = 0.0;
	if ((ref_cnt == 0 && alt_cnt == 0) || (ref_cnt == MAX_COV && alt_cnt == MAX_COV)) return 0;
	const double g0 = g_cache[ref_cnt][alt_cnt][0];
	const double g1 = g_cache[ref_cnt][alt_cnt][1];
	const double g2 = g_cache[ref_cnt][alt_cnt][2];
	const double p = ref_freq_enc / 255.0;
	const double q = alt_freq_enc / 255.0;
	const double p2 = p * p;
	const double q2 = q * q;
	const double p_g0 = p2 * g0;
	const double p_g1 = (1.0 - p2 - q2) * g1;
	const double p_g2 = q2 * g2;
	const double total = p_g0 + p_g1 + p_g2;
	const int n = ref_cnt + alt_cnt;
	if (p_g0 > p_g1 && p_g0 > p_g2) { *confidence = ((double)(p_g0 / total)) * g_poisson[n]; return 1; }
	else if (p_g1 > p_g0 && p_g1 > p_g2) { *confidence = ((double)(p_g1 / total)) * g_poisson[n]; return 3; }
	else { *confidence = ((double)(p_g2 / total)) * g_poisson[n]; return 2; }
}

int vgo_gq(double confidence) { int q = -1 * 10 * log(confidence); return q; }
