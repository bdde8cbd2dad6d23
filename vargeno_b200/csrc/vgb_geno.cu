// vgb_geno.cu -- the fused per-read kernel (K2 + K3 + K4): one warp owns one read from packed k-mers to pileup.
//
// Replaces the body of the reference's FASTQ loop, src/qv.cc:778-1510 (SURVEY.md 3.3 is the normative
// restatement this follows):
//   encode      src/util.c:89-111 encode_kmer, reverse complement src/qv.cc:787-806
//   exact       query_ref_dict / query_snp_dict src/qv.cc:206-240,385-411; event emission :850-937
//   neighbours  Bloom gates :946-956, big-block queries :962-1109, strided small-block scan :316-376,413-464,
//               :1110-1209 (F13 kept: entry lo+S*t is examined, entry lo+t is reported), upper half :1213-1365
//   vote        improved_index_table_add :132-178 in its order-independent form
//   pileup      :1382-1502 as integer atomics on compact per-site counters (saturation deferred)
//
// Mapping: persistent grid, 8 warps per CTA, reads claimed dynamically in batches.  A warp's 32 lanes are
// 32 independent probe chains during the exact and neighbour phases (each probe = jumpgate pair -> short
// search -> entry, all aligned loads that stay inside one 32 B sector), and 32 pileup positions during the
// update.  Hit contexts live in shared memory (spill to a per-warp global area only beyond 64 per read).
#include <cstring>

#include "vgb_internal.h"

namespace vgb {

constexpr int GW = 8;                 // warps per CTA
constexpr int EV_SMEM = 64;           // hit contexts per warp kept in shared memory
constexpr int EV_CAP = 4096;          // >= 2 * MAX_HITS
constexpr int READ_BATCH = 4;

struct __align__(8) Event {
	uint64_t kmer;
	uint32_t X;       // read position this context votes for (kmer_context.position)
	uint32_t kpos;    // kmer_context.kmer_pos
	uint32_t meta;    // bits 0-7 modified base (0xFF none) | 8-15 k-mer index | 16 list (0 ref, 1 snp) | 17 votes
	uint32_t pad;
};

struct WarpSmem {
	Event ev[EV_SMEM];
	uint32_t rlo[32], rB[32], slo[32], sB[32];
	uint32_t ev_count;
	uint32_t pad[3];
};

struct GenoArgs {
	DevIndex ix;
	const char *text;
	const uint32_t *line_start;
	uint32_t *meta;               // [1] n_reads [2] work counter [3] error bits [11] first line of the chunk's own records
	DevStats *stats;
	vgb_read_result *trace;       // nullptr unless VGB_CFG_TRACE
	Event *spill;                 // [grid warps][EV_CAP - EV_SMEM]
	const uint32_t *list;         // warp kernel: nullptr = every read of the chunk, else the deferred reads (count in meta[in_cnt])
	uint32_t *defer;              // group kernels: where reads go that need the next kernel (count in meta[defer_cnt]; bit 31: start at the retry pass)
	const uint32_t *klist;        // group kernels behind the 4-lane one: the reads they were handed (count in meta[in_cnt]); nullptr = whole chunk
	uint32_t *kdefer;             // 4-lane kernel: reads with 5..8 k-mers, for the 8-lane kernel (count in meta[9])
	uint32_t in_cnt, defer_cnt;   // meta slots: in_cnt = length of list / klist (in_cnt + 1: its work counter), defer_cnt = length of defer
};

struct LaneStats {
	uint32_t exact = 0, nbrq = 0, scan = 0, bf = 0, lowq = 0, events = 0, incr = 0, big = 0;
};

__device__ __forceinline__ uint64_t spread32(uint32_t x)
{
	uint64_t v = x;
	v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
	v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
	v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
	v = (v | (v << 2)) & 0x3333333333333333ull;
	v = (v | (v << 1)) & 0x5555555555555555ull;
	return v;
}

// reverse complement of a packed 32-mer (base b at bits 2b): reverse the base order, complement = 3 - code
__device__ __forceinline__ uint64_t revcomp64(uint64_t k)
{
	uint64_t r = __brevll(k);                                               // reverses bits: base order reversed, bit pairs swapped
	r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
	return ~r;
}

__device__ __forceinline__ Event *ev_at(WarpSmem *ws, Event *spill, uint32_t i)
{
	return i < EV_SMEM ? &ws->ev[i] : &spill[i - EV_SMEM];
}

__device__ __forceinline__ void emit(WarpSmem *ws, Event *spill, uint64_t kmer, uint32_t pos, uint32_t offset, uint32_t mod,
                                     uint32_t kidx, uint32_t list, LaneStats &st)
{
	const uint32_t i = atomicAdd(&ws->ev_count, 1u);                        // order of contexts is irrelevant
	st.events++;
	if (i >= EV_CAP) return;
	Event e;
	e.kmer = kmer; e.X = pos - offset; e.kpos = pos; e.meta = mod | (kidx << 8) | (list << 16); e.pad = 0;
	*ev_at(ws, spill, i) = e;
}

// exact hit of the read's own k-mer -> contexts (src/qv.cc:850-890 ref, :897-937 snp)
__device__ __forceinline__ void exact_ref_events(const DevIndex &ix, WarpSmem *ws, Event *spill, uint64_t kmer, uint32_t posx,
                                                 uint32_t offset, uint32_t kidx, LaneStats &st)
{
	if (posx == POS_AMBIGUOUS) return;
	if (posx < ix.amb_lo) { emit(ws, spill, kmer, posx, offset, NO_MOD, kidx, 0, st); return; }
	const uint32_t *row = ix.ref_aux + (uint64_t)(0xFFFFFFFEu - posx) * AUX_COLS;
	for (int c = 0; c < AUX_COLS; c++) {
		const uint32_t p = __ldg(row + c);
		if (p == 0) break;
		emit(ws, spill, kmer, p, offset, NO_MOD, kidx, 0, st);
	}
}
__device__ __forceinline__ void exact_snp_events(const DevIndex &ix, WarpSmem *ws, Event *spill, uint64_t kmer, const SnpEntry &e,
                                                 uint32_t offset, uint32_t kidx, LaneStats &st)
{
	if (e.pos == POS_AMBIGUOUS) return;
	if (snp_flag_of(e) == 0) { emit(ws, spill, kmer, e.pos, offset, NO_MOD, kidx, 1, st); return; }
	const uint32_t *row = ix.snp_aux_pos + (uint64_t)e.pos * AUX_COLS;
	for (int c = 0; c < AUX_COLS; c++) {
		const uint32_t p = __ldg(row + c);
		if (p == 0) break;
		emit(ws, spill, kmer, p, offset, NO_MOD, kidx, 1, st);
	}
}
// neighbour hit with modified base d (src/qv.cc:985-1046 and twins :1131-1172, :1227-1291)
__device__ __forceinline__ void nbr_ref_events(const DevIndex &ix, WarpSmem *ws, Event *spill, uint64_t nb, uint32_t posx, uint32_t d,
                                               uint32_t offset, uint32_t kidx, LaneStats &st)
{
	if (posx == POS_AMBIGUOUS) return;
	if (posx < ix.amb_lo) {
		if (!pile_nonzero(ix, (uint64_t)posx + d)) emit(ws, spill, nb, posx, offset, d, kidx, 0, st);
		return;
	}
	const uint32_t *row = ix.ref_aux + (uint64_t)(0xFFFFFFFEu - posx) * AUX_COLS;
	for (int c = 0; c < AUX_COLS; c++) {
		const uint32_t p = __ldg(row + c);
		if (p == 0) break;
		if (!pile_nonzero(ix, (uint64_t)p + d)) emit(ws, spill, nb, p, offset, d, kidx, 0, st);
	}
}
// src/qv.cc:1053-1101 and twins :1178-1207, :1311-1356
__device__ __forceinline__ void nbr_snp_events(const DevIndex &ix, WarpSmem *ws, Event *spill, uint64_t nb, const SnpEntry &e, uint32_t d,
                                               uint32_t offset, uint32_t kidx, LaneStats &st)
{
	if (e.pos == POS_AMBIGUOUS) return;
	const uint32_t flag = snp_flag_of(e);
	if (flag == 0) {
		if ((snp_info_of(e) >> 3) != d) emit(ws, spill, nb, e.pos, offset, d, kidx, 1, st);
		return;
	}
	const uint32_t *row = ix.snp_aux_pos + (uint64_t)e.pos * AUX_COLS;
	const uint8_t *inf = ix.snp_aux_info + (uint64_t)e.pos * AUX_COLS;
	for (int c = 0; c < AUX_COLS; c++) {
		const uint32_t p = __ldg(row + c);
		if (p == 0) break;
		if (((uint32_t)__ldg(inf + c) >> 3) != d) emit(ws, spill, nb, p, offset, d, kidx, 1, st);
	}
}

// the t-th substitution of base slot d: the three bases other than the current one, ascending (src/qv.cc:970-973)
__device__ __forceinline__ uint64_t substitute(uint64_t kmer, uint32_t d, uint32_t which)
{
	const uint32_t sh = 2 * d;
	const uint64_t base = (kmer >> sh) & 3ull;
	const uint64_t j = which + (which >= base ? 1 : 0);
	return (kmer & ~(3ull << sh)) | (j << sh);
}

template <int MINB>
__global__ void __launch_bounds__(GW * 32, MINB) k_geno(const GenoArgs a)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	WarpSmem *ws = reinterpret_cast<WarpSmem *>(smem_raw) + (threadIdx.x >> 5);
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t gwarp = blockIdx.x * GW + (threadIdx.x >> 5);
	Event *spill = a.spill + (uint64_t)gwarp * (EV_CAP - EV_SMEM);
	const DevIndex &ix = a.ix;
	// list mode: only the reads the 8-lane kernel deferred (more than 8 k-mers, or more hit contexts than its shared memory holds)
	const uint32_t n_reads = a.list ? a.meta[a.in_cnt] : a.meta[1];
	uint32_t *work = a.meta + (a.list ? a.in_cnt + 1 : 2);
	LaneStats st;
	uint32_t w_reads = 0, w_skipped = 0, w_passes = 0, w_placed = 0, w_bad = 0, w_overflow = 0, w_wrap = 0;

	for (;;) {
		uint32_t r0 = 0;
		if (lane == 0) r0 = atomicAdd(work, (uint32_t)READ_BATCH);
		r0 = __shfl_sync(0xffffffffu, r0, 0);
		if (r0 >= n_reads) break;
		const uint32_t r1 = min(r0 + READ_BATCH, n_reads);
		for (uint32_t ri = r0; ri < r1; ri++) {
			// list entries: bit 31 = the 8-lane kernel already ran (and accounted for) the forward pass, start at the retry
			const uint32_t rr = a.list ? __ldg(a.list + ri) : ri;
			const uint32_t r = rr & 0x7FFFFFFFu;
			// ---- record framing: lines 4r .. 4r+3 (src/qv.cc:760-779) ----
			uint32_t lsv = lane < 5 ? __ldg(a.line_start + a.meta[11] + 4ull * r + lane) : 0;   // meta[11]: first line of the chunk's own records (BGZF window)
			const uint32_t id_s = __shfl_sync(0xffffffffu, lsv, 0);
			const uint32_t seq_s = __shfl_sync(0xffffffffu, lsv, 1);
			const uint32_t sep_s = __shfl_sync(0xffffffffu, lsv, 2);
			const uint32_t qual_s = __shfl_sync(0xffffffffu, lsv, 3);
			const uint32_t next_s = __shfl_sync(0xffffffffu, lsv, 4);
			const uint32_t L = sep_s - 1 - seq_s;                 // strlen(read) - 1
			const uint32_t qlen = next_s - 1 - qual_s;
			const uint32_t K = L >> 5;
			w_reads++;
			vgb_read_result res;
			res.flags = 0; res.target = 0; res.freq = 0; res.n_ref = 0; res.n_snp = 0; res.passes = 0; res.ctx_hash = 0;
			bool bad = (seq_s - 1 - id_s > 1022) || (L > 1022) || (qual_s - 1 - sep_s > 1022) || (qlen > 1022) || (qlen < K);
			bool skipped = false;

			// ---- 2-bit packing: lane j ends up holding k-mer j (src/util.c:89-111) ----
			uint64_t kmer = 0;
			if (!bad) {
				for (uint32_t j = 0; j < K; j++) {
					const unsigned char ch = (unsigned char)__ldg(a.text + seq_s + 32u * j + lane);
					uint32_t code = 0, isn = 0, isx = 0;
					switch (ch) {
					case 'A': case 'a': code = 0; break;
					case 'C': case 'c': code = 1; break;
					case 'G': case 'g': code = 2; break;
					case 'T': case 't': code = 3; break;
					case 'N': case 'n': isn = 1; break;
					default: isx = 1; break;                      // the reference aborts here (src/util.c:103)
					}
					const uint32_t b0 = __ballot_sync(0xffffffffu, code & 1);
					const uint32_t b1 = __ballot_sync(0xffffffffu, code & 2);
					const uint32_t bn = __ballot_sync(0xffffffffu, isn);
					const uint32_t bx = __ballot_sync(0xffffffffu, isx);
					if (bn | bx) {
						// encode_kmer scans base 31 down to 0: the highest offending base decides N-skip vs abort
						const uint32_t top = 31 - __clz(bn | bx);
						if ((bx >> top) & 1u) bad = true; else skipped = true;
						break;
					}
					if (lane == j) kmer = spread32(b0) | (spread32(b1) << 1);
				}
			}
			if (bad) {
				w_bad++;
				if (lane == 0) atomicOr(&a.meta[5], 2u);
				if (a.trace && lane == 0) { res.flags = VGB_RF_SKIPPED; a.trace[r] = res; }
				continue;
			}
			if (skipped) {
				w_skipped++;
				if (a.trace && lane == 0) { res.flags = VGB_RF_SKIPPED; a.trace[r] = res; }
				continue;
			}
			// quality gate of k-mer i is the i-th quality CHARACTER (src/qv.cc:836,943; F8); char is signed
			bool lowq = false;
			if (lane < K) lowq = ((int)(signed char)__ldg(a.text + qual_s + lane) - QUALITY_SCORE) < 0;
			const uint32_t lowmask = __ballot_sync(0xffffffffu, lowq);
			const uint64_t kmer_fwd = kmer;

			bool process = false, has_best = false, ambiguous = false;
			uint32_t target = 0, best_freq = 0, E = 0;
			uint32_t pass = a.list ? rr >> 31 : 0u;
			for (;; pass++) {
				if (pass == 1) {                                   // src/qv.cc:787-806: reverse complement of the first 32K bases
					const uint64_t o = __shfl_sync(0xffffffffu, kmer_fwd, (K - 1 - lane) & 31);
					kmer = lane < K ? revcomp64(o) : 0;
				}
				w_passes++;
				if (lane == 0) ws->ev_count = 0;
				__syncwarp();

				// ---- exact queries: lane q handles k-mer q/2 against dictionary q%2 ----
				for (uint32_t qb = 0; qb < 2 * K; qb += 32) {
					const uint32_t q = qb + lane;
					const uint32_t kidx = q >> 1;
					const uint64_t km = __shfl_sync(0xffffffffu, kmer, kidx & 31);
					if (q < 2 * K) {
						st.exact++;
						const uint32_t offset = 32u * kidx;
						if ((q & 1) == 0) {
							uint32_t lo, hi, posx = 0;
							ref_block(ix, km, lo, hi);
							ws->rlo[kidx] = lo; ws->rB[kidx] = hi - lo;
							if (lo < hi && ref_find_in_block(ix, (uint32_t)km, lo, hi, posx) >= 0)
								exact_ref_events(ix, ws, spill, km, posx, offset, kidx, st);
						} else {
							uint32_t lo, hi;
							SnpEntry e;
							snp_block(ix, km, lo, hi);
							ws->slo[kidx] = lo; ws->sB[kidx] = hi - lo;
							if (lo < hi && snp_find_in_block(ix, km & 0xFFFFFFFFFFull, lo, hi, e) >= 0)
								exact_snp_events(ix, ws, spill, km, e, offset, kidx, st);
						}
					}
				}
				__syncwarp();

				// ---- Hamming-1 neighbours of the low-quality k-mers ----
				uint32_t lm = lowmask;
				while (lm) {
					const uint32_t i = __ffs(lm) - 1;
					lm &= lm - 1;
					const uint64_t km = __shfl_sync(0xffffffffu, kmer, i);
					const uint32_t offset = 32u * i;
					const uint32_t rlo = ws->rlo[i], rB = ws->rB[i], slo = ws->slo[i], sB = ws->sB[i];
					uint32_t gate = 0;
					if (lane == 0) gate = bf_ref(ix, (uint32_t)km) ? 1u : 0u;                 // src/qv.cc:955
					if (lane == 1) gate = bf_snp(ix, km & 0xFFFFFFFFFFull) ? 1u : 0u;         // src/qv.cc:956
					const bool rb = __shfl_sync(0xffffffffu, gate, 0) != 0;
					const bool sb = __shfl_sync(0xffffffffu, gate, 1) != 0;
					const bool big = rB >= BLOCK_SIZE_THRESHOLD;                               // src/qv.cc:843,962
					if (lane == 0) { st.lowq++; st.bf += 2; if (big) st.big++; }
					// task segments (all bounds warp-uniform)
					// upper half, ref, d = 16..31 (:1225): the 48 substituted k-mers share LO32, so every dictionary k-mer
					// that can answer one of those 48 queries sits in the LO32 bucket; one bucket walk replaces 48 probes
					uint32_t bs = 0, be = 0;
					if (rb) {
						uint32_t v0 = 0, v1 = 0;
						if (lane == 0) ref_lo_bucket(ix, (uint32_t)km, v0, v1);
						bs = __shfl_sync(0xffffffffu, v0, 0);
						be = __shfl_sync(0xffffffffu, v1, 0);
						if (lane == 0) st.nbrq += 48;             // the 48 reference queries this walk stands for
					}
					const uint32_t n0 = be - bs;
					const uint32_t n1 = sb ? 36u : 0u;            // upper half, snp:  d = 20..31 if sb  (:1305-1307)
					const uint32_t n2 = big ? 12u : 0u;           // upper half, snp:  d = 16..19 if big (:1305-1307)
					const uint32_t n3 = big ? 48u : rB;           // lower half, ref: queries (:975) or strided scan (:358-373)
					const uint32_t n4 = big ? 48u : sB;           // lower half, snp: queries (:977) or strided scan (:447-462)
					const uint32_t e0 = n0, e1 = e0 + n1, e2 = e1 + n2, e3 = e2 + n3, e4 = e3 + n4;
					for (uint32_t t = lane; t < e4; t += 32) {
						if (t < e0) {                                                        // LO32 bucket entry
							const uint2 en = __ldg(reinterpret_cast<const uint2 *>(ix.ref_by_lo + bs + t));
							const int sl = one_base_slot((uint64_t)(en.x ^ (uint32_t)(km >> 32)));
							if (sl >= 0) {
								const uint64_t nb = ((uint64_t)en.x << 32) | (uint32_t)km;
								nbr_ref_events(ix, ws, spill, nb, en.y, 16u + (uint32_t)sl, offset, i, st);
							}
						} else if (t >= e2 && t < e3 && big) {                               // ref query (big mode, lower half)
							const uint32_t u = t - e2;
							const uint32_t d = u / 3;
							const uint64_t nb = substitute(km, d, u % 3);
							uint32_t posx;
							st.nbrq++;
							if (ref_query(ix, nb, posx) >= 0) nbr_ref_events(ix, ws, spill, nb, posx, d, offset, i, st);
						} else if (t < e2 || (t >= e3 && big)) {                            // snp query
							uint32_t u, d;
							if (t < e1) { u = t - e0; d = 20u + u / 3; }
							else if (t < e2) { u = t - e1; d = 16u + u / 3; }
							else { u = t - e3; d = u / 3; }
							const uint64_t nb = substitute(km, d, u % 3);
							SnpEntry e;
							st.nbrq++;
							if (snp_query(ix, nb, e) >= 0) nbr_snp_events(ix, ws, spill, nb, e, d, offset, i, st);
						} else if (t < e3) {                                                 // ref strided scan step (F13)
							const uint32_t s = t - e2;
							const uint64_t ex = (uint64_t)rlo + (uint64_t)REF_STRIDE * s;
							st.scan++;
							if (ex < ix.n_ref) {
								const uint32_t entry_lo = __ldg(&ix.ref[ex].lo);
								const int d = one_base_slot((uint64_t)((uint32_t)km ^ entry_lo));
								if (d >= 0) {
									const uint32_t posx = __ldg(&ix.ref[rlo + s].posx);
									const uint64_t nb = (km & 0xFFFFFFFF00000000ull) | entry_lo;
									nbr_ref_events(ix, ws, spill, nb, posx, (uint32_t)d, offset, i, st);
								}
							}
						} else {                                                             // snp strided scan step (F13)
							const uint32_t s = t - e3;
							const uint64_t ex = (uint64_t)slo + (uint64_t)SNP_STRIDE * s;
							st.scan++;
							if (ex < ix.n_snp) {
								uint64_t entry_lo = 0;
								const int d = snp_scan_step(ix, slo, s, km, entry_lo);
								if (d >= 0) {
									const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(ix.snp + slo + s));
									SnpEntry e;
									e.key = ((uint64_t)raw.y << 32) | raw.x; e.pos = raw.z; e.extra = raw.w;
									const uint64_t nb = (km & 0xFFFFFF0000000000ull) | entry_lo;
									nbr_snp_events(ix, ws, spill, nb, e, (uint32_t)d, offset, i, st);
								}
							}
						}
					}
				}
				__syncwarp();

				// ---- vote (src/qv.cc:132-178, order-independent form) ----
				E = ws->ev_count;
				const bool overflow = E > EV_CAP;
				if (overflow) E = EV_CAP;
				// (1) which contexts vote: exact ones always; a neighbour of k-mer i iff an exact context with the same X
				//     exists from a k-mer j <= i (the map key must exist, :134-139, and exact contexts of k-mer i precede its neighbours)
				for (uint32_t e = lane; e < E; e += 32) {
					Event *p = ev_at(ws, spill, e);
					const uint32_t m = p->meta, X = p->X;
					bool v = (m & 0xFF) == NO_MOD;
					if (!v) {
						const uint32_t ki = (m >> 8) & 0xFF;
						for (uint32_t f = 0; f < E && !v; f++) {
							const Event *q = ev_at(ws, spill, f);
							v = ((q->meta & 0xFF) == NO_MOD) && q->X == X && ((q->meta >> 8) & 0xFF) <= ki;
						}
					}
					p->meta = (m & ~(1u << 17)) | (v ? (1u << 17) : 0u);
				}
				__syncwarp();
				// (2) per voting context: frequency of its X and whether two distinct k-mer positions support it (:163-165)
				uint32_t bf_ = 0, bxmin = 0xFFFFFFFFu, bxmax = 0;
				uint32_t nref = 0, nsnp = 0;
				uint64_t dg = 0;
				for (uint32_t e = lane; e < E; e += 32) {
					const Event *p = ev_at(ws, spill, e);
					const uint32_t m = p->meta, X = p->X, kp = p->kpos;
					if ((m >> 16) & 1) nsnp++; else nref++;
					if (a.trace) {
						const uint32_t mod = (m & 0xFF) == NO_MOD ? 10086u : (m & 0xFF);
						dg += ctx_digest((m >> 16) & 1, X, kp, p->kmer, mod);
					}
					if (!((m >> 17) & 1)) continue;
					uint32_t f = 0;
					bool distinct = false;
					for (uint32_t g = 0; g < E; g++) {
						const Event *q = ev_at(ws, spill, g);
						if (((q->meta >> 17) & 1) && q->X == X) { f++; distinct |= (q->kpos != kp); }
					}
					if (!distinct) continue;
					if (f > bf_) { bf_ = f; bxmin = X; bxmax = X; }
					else if (f == bf_) { bxmin = min(bxmin, X); bxmax = max(bxmax, X); }
				}
				const uint32_t maxf = __reduce_max_sync(0xffffffffu, bf_);
				const uint32_t xmin = __reduce_min_sync(0xffffffffu, bf_ == maxf ? bxmin : 0xFFFFFFFFu);
				const uint32_t xmax = __reduce_max_sync(0xffffffffu, bf_ == maxf ? bxmax : 0u);
				has_best = maxf > 0;
				ambiguous = has_best && xmin != xmax;
				process = has_best && !ambiguous;                  // freq > 1 is implied by two distinct k-mer positions (:1375)
				target = xmin;
				best_freq = maxf;
				if (maxf > 255) w_wrap++;                          // uint8 freq would have wrapped in the reference
				if (overflow) w_overflow++;
				nref = __reduce_add_sync(0xffffffffu, nref);
				nsnp = __reduce_add_sync(0xffffffffu, nsnp);
				if (nref > MAX_HITS || nsnp > MAX_HITS) w_overflow++;
				if (a.trace) {
#pragma unroll
					for (int o = 16; o; o >>= 1) dg += __shfl_xor_sync(0xffffffffu, dg, o);
					res.n_ref = (uint16_t)nref; res.n_snp = (uint16_t)nsnp; res.ctx_hash = dg;
				}

				// ---- pileup update: every recorded context at the winning position (src/qv.cc:1382-1502) ----
				if (process) {
					for (uint32_t e = 0; e < E; e++) {
						const Event *p = ev_at(ws, spill, e);
						if (p->X != target) continue;
						const uint32_t mod = p->meta & 0xFF;
						const uint64_t pos = (uint64_t)p->kpos + lane;
						if (lane == mod || pos >= ix.pile_len) continue;
						const uint4 blk = __ldg(reinterpret_cast<const uint4 *>(ix.pile + (pos >> 6)));
						const uint64_t bits = ((uint64_t)blk.y << 32) | blk.x;
						const uint32_t off = pos & 63;
						if (!((bits >> off) & 1ull)) continue;
						const uint32_t sid = blk.z + __popcll(bits & ((1ull << off) - 1));
						const uint32_t code = __ldg(ix.site_code + sid);
						const uint32_t rbase = code & 3, abase = code >> 2;
						if (rbase == abase) continue;                  // p->ref != p->alt (:1404)
						const uint32_t base = (uint32_t)(p->kmer >> (2 * lane)) & 3u;
						if (base == rbase) { atomicAdd(ix.cnt + 2ull * sid, 1u); st.incr++; }
						else if (base == abase) { atomicAdd(ix.cnt + 2ull * sid + 1, 1u); st.incr++; }
					}
				}
				__syncwarp();
				if (!process && pass == 0) continue;               // retry once on the reverse complement (:1504-1510)
				break;
			}
			if (process) w_placed++;
			if (a.trace && lane == 0) {
				res.flags = (pass ? VGB_RF_REVCOMPL : 0) | (process ? VGB_RF_PROCESS : 0) | (ambiguous ? VGB_RF_AMBIGUOUS : 0) |
				            (has_best ? VGB_RF_HASBEST : 0);
				res.target = target; res.freq = (uint16_t)(best_freq & 0xFF); res.passes = (uint16_t)(pass + 1);
				a.trace[r] = res;
			}
		}
	}

	// ---- statistics: one set of atomics per warp that did anything (an empty deferred list costs nothing) ----
	if (w_reads == 0) return;
	unsigned long long v[8] = { st.exact, st.nbrq, st.scan, st.bf, st.lowq, st.events, st.incr, st.big };
#pragma unroll
	for (int k = 0; k < 8; k++) {
#pragma unroll
		for (int o = 16; o; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
	}
	if (lane == 0) {
		DevStats *s = a.stats;
		atomicAdd(&s->reads, (unsigned long long)w_reads); atomicAdd(&s->skipped_n, (unsigned long long)w_skipped);
		atomicAdd(&s->passes, (unsigned long long)w_passes); atomicAdd(&s->placed, (unsigned long long)w_placed);
		atomicAdd(&s->exact_lookups, v[0]); atomicAdd(&s->nbr_query_lookups, v[1]); atomicAdd(&s->nbr_scan_reads, v[2]);
		atomicAdd(&s->bf_probes, v[3]); atomicAdd(&s->lowq_kmers, v[4]); atomicAdd(&s->events, v[5]);
		atomicAdd(&s->pileup_incr, v[6]); atomicAdd(&s->big_kmers, v[7]);
		if (w_bad) atomicAdd(&s->bad_records, (unsigned long long)w_bad);
		if (w_overflow) atomicAdd(&s->overflow_reads, (unsigned long long)w_overflow);
		if (w_wrap) atomicAdd(&s->freq_wrap_reads, (unsigned long long)w_wrap);
	}
}

#include "vgb_geno8.inl"

typedef void (*geno_kernel_t)(const GenoArgs);
static size_t grp_smem_bytes(int G, bool wide = false)
{
	const size_t R = 32 / G;
	return (wide ? sizeof(OctSmem<EV_WIDE>) : sizeof(OctSmem<EV_GROUP>)) * GW * R + GW * 16 * sizeof(uint32_t) + (G == 4 ? sizeof(Pend<4>) : sizeof(Pend<8>)) * GW * 2 * R;
}

template <int G>
static void pick_group_kernels(int minb, bool shortscan, geno_kernel_t &k, geno_kernel_t &kt)
{
	if (minb == 4 && shortscan) { k = k_geno8<4, false, G, EV_GROUP, true>; kt = k_geno8<4, true, G, EV_GROUP, true>; return; }
	if (minb >= 6) { k = k_geno8<6, false, G, EV_GROUP>; kt = k_geno8<6, true, G, EV_GROUP>; }
	else if (minb == 5) { k = k_geno8<5, false, G, EV_GROUP>; kt = k_geno8<5, true, G, EV_GROUP>; }
	else if (minb == 3) { k = k_geno8<3, false, G, EV_GROUP>; kt = k_geno8<3, true, G, EV_GROUP>; }
	else { k = k_geno8<4, false, G, EV_GROUP>; kt = k_geno8<4, true, G, EV_GROUP>; }
}

int geno_prepare(vgb_ctx *c)
{
	// VGB_GENO_KERNEL=warp: one warp per read for everything (the first version, kept as the path for deferred reads);
	//                =oct:  no 4-lane kernel in front (reads of 129..287 bases dominate)
	// VGB_GENO_MINB / VGB_GENO8_MINB / VGB_GENO4_MINB: register budget variants (CTAs per SM the compiler must make room for)
	int minb = 4, minb8 = 4, minb4 = 4;   // measured on B200 (profiles/r01_summary.md): 62 registers, no spills, 32 warps per SM beats 40 registers / 48 warps
	if (const char *e = getenv("VGB_GENO_MINB")) minb = atoi(e);
	if (const char *e = getenv("VGB_GENO8_MINB")) minb8 = atoi(e);
	if (const char *e = getenv("VGB_GENO4_MINB")) minb4 = atoi(e);
	// VGB_NO_TAIL_OVERLAP: hand-over kernels stay on the kernel stream (measurement switch); per-read results (trace) imply it
	c->tail_overlap = !(c->cfg.flags & VGB_CFG_TRACE) && !getenv("VGB_NO_TAIL_OVERLAP");
	const char *kk = getenv("VGB_GENO_KERNEL");
	const bool warp_only = kk && !strcmp(kk, "warp");
	c->use_quad = !(kk && !strcmp(kk, "oct"));
	geno_kernel_t k = minb >= 8 ? k_geno<8> : (minb >= 6 ? k_geno<6> : (minb == 5 ? k_geno<5> : k_geno<4>));
	geno_kernel_t grp[3][2];
	// mean HI24 block length of the SNP dictionary decides the scan instantiation (see k_geno8): below four entries a scan ends
	// inside its first load; VGB_SHORTSCAN=0/1 overrides (measurement switch)
	bool shortscan = (c->ix.n_snp >> 24) < 4;
	if (const char *e = getenv("VGB_SHORTSCAN")) shortscan = atoi(e) != 0;
	pick_group_kernels<4>(minb4, shortscan, grp[0][0], grp[0][1]);
	pick_group_kernels<8>(minb8, shortscan, grp[1][0], grp[1][1]);
	grp[2][0] = k_geno8<4, false, 8, EV_WIDE>; grp[2][1] = k_geno8<4, true, 8, EV_WIDE>;
	if (getenv("VGB_NO_WIDE")) grp[2][0] = grp[2][1] = nullptr;      // tuning: hand-overs go straight to the warp kernel, as before
	if (warp_only) for (int i = 0; i < 3; i++) grp[i][0] = grp[i][1] = nullptr;
	c->warp_kernel = (void *)k;
	for (int i = 0; i < 3; i++) for (int t = 0; t < 2; t++) c->grp_kernel[i][t] = (void *)grp[i][t];
	int occ = 0;
	const size_t smem = sizeof(WarpSmem) * GW;
	VGB_CUDA(c, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	VGB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, GW * 32, smem));
	if (occ < 1) occ = 1;
	c->geno_grid = (uint32_t)(c->sm_count * occ);
	for (int gi = 0; gi < 3 && !warp_only; gi++) {
		if (!grp[gi][0]) continue;
		const size_t sm = grp_smem_bytes(gi ? 8 : 4, gi == 2);   // hit contexts + one row of counters per warp + the warp's parked reads
		for (int t = 0; t < 2; t++) {
			VGB_CUDA(c, cudaFuncSetAttribute(grp[gi][t], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
			// VGB_CARVEOUT: preferred shared-memory share of the unified L1 / shared array in percent (tuning knob; default: the driver's choice)
			if (const char *e = getenv("VGB_CARVEOUT")) VGB_CUDA(c, cudaFuncSetAttribute(grp[gi][t], cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)));
		}
		VGB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, grp[gi][0], GW * 32, sm));
		if (occ < 1) occ = 1;
		c->grp_grid[gi] = (uint32_t)(c->sm_count * occ);
	}
	if (!c->d_spill) {
		Event *sp = nullptr;
		int rc = dev_alloc(c, &sp, (uint64_t)c->geno_grid * GW * (EV_CAP - EV_SMEM));
		if (rc) return rc;
		c->d_spill = sp;
	}
	return VGB_OK;
}

int geno_launch(vgb_ctx *c, Chunk &ck, uint64_t nbytes, uint64_t first_read_id)
{
	(void)nbytes; (void)first_read_id;
	GenoArgs a;
	a.ix = c->ix;
	a.text = ck.d_text;
	a.line_start = ck.d_line_start;
	a.meta = ck.d_meta;
	a.stats = c->d_stats;
	a.trace = c->d_trace ? c->d_trace + c->trace_n : nullptr;
	a.spill = (Event *)c->d_spill;
	a.list = nullptr;
	a.defer = ck.d_defer; a.defer_cnt = 6;
	a.klist = nullptr; a.in_cnt = 0;
	a.kdefer = ck.d_defer2;
	const int tr = a.trace ? 1 : 0;   // per-read results (VGB_CFG_TRACE: tests) are a separate instantiation, so the production kernels carry none of it
	const geno_kernel_t warp_kernel = (geno_kernel_t)c->warp_kernel;
	const geno_kernel_t grp4 = (geno_kernel_t)c->grp_kernel[0][tr], grp8 = (geno_kernel_t)c->grp_kernel[1][tr], grpw = (geno_kernel_t)c->grp_kernel[2][tr];
	if (grp8) {
		// 4 lanes per read (up to 4 k-mers: 128..159 bases), then 8 lanes per read for what it handed over (5..8 k-mers), then 8
		// lanes per read with the wide context list for the reads of both with too many hit contexts, then one warp per read for
		// the rest (more than 8 k-mers, more contexts than the wide list holds)
		if (c->use_quad) {
			grp4<<<c->grp_grid[0], GW * 32, grp_smem_bytes(4), c->stream>>>(a);
			a.klist = ck.d_defer2; a.in_cnt = 9;
			c->launches++;
		}
		a.kdefer = nullptr;
		grp8<<<c->grp_grid[1], GW * 32, grp_smem_bytes(8), c->stream>>>(a);
		c->launches++;
		// hand-over kernels: behind the main kernels of this chunk, and (tail_overlap) under the framing and main kernel of the next
		cudaStream_t ts = c->stream;
		ck.tail = false;
		if (c->tail_overlap) {
			VGB_CUDA(c, cudaEventRecord(ck.g1, c->stream));
			VGB_CUDA(c, cudaStreamWaitEvent(c->tail_stream, ck.g1, 0));
			VGB_CUDA(c, cudaEventRecord(ck.h0, c->tail_stream));
			ts = c->tail_stream;
			ck.tail = true;
		}
		if (grpw) {
			a.klist = ck.d_defer; a.in_cnt = 6;
			a.defer = ck.d_defer3; a.defer_cnt = 13;
			grpw<<<c->grp_grid[2], GW * 32, grp_smem_bytes(8, true), ts>>>(a);
			c->launches++;
			a.list = ck.d_defer3; a.in_cnt = 13;
		} else {
			a.list = ck.d_defer; a.in_cnt = 6;
		}
		warp_kernel<<<c->geno_grid, GW * 32, sizeof(WarpSmem) * GW, ts>>>(a);
		c->launches++;
	} else {
		ck.tail = false;
		warp_kernel<<<c->geno_grid, GW * 32, sizeof(WarpSmem) * GW, c->stream>>>(a);
		c->launches++;
	}
	VGB_CUDA(c, cudaGetLastError());
	return VGB_OK;
}

}  // namespace vgb
