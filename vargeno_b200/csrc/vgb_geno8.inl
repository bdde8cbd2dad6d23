// vgb_geno8.inl -- the main per-read kernel: EIGHT lanes per read, four reads per warp.  Included by vgb_geno.cu.
//
// Same semantics as k_geno (one warp per read), different mapping.  A 150 bp read has 4 k-mers = 8 exact probes, so a
// full warp per read leaves 24 lanes idle in the phases that matter and puts only one read's dependent probe chain in
// flight per warp (profiles/r01_summary.md: ~1400 warp instructions and ~10 dependent DRAM accesses per read,
// long-scoreboard stalls everywhere).  Here lane j of an octet owns k-mer j of its read for packing, exact probes,
// Bloom gates and bucket bounds; neighbour tasks, the vote and the pileup window are spread over the octet's 8 lanes.
// All four octets of a warp move through the phases together, so every warp-wide shuffle / ballot is executed
// convergently; loops whose trip count differs per octet contain no warp-synchronous operation.
//
// Reads with more than 8 k-mers (>= 288 bases) or more than OCT_EV hit contexts in a pass are deferred, untouched, to
// k_geno in list mode (a.defer / meta[6]).

constexpr int OCT_EV = 24;            // hit contexts per read kept in shared memory

// per-read statistics are kept in the octet's shared memory (committed only when the read is finished here, not when it
// is deferred): at 32 registers per thread every counter held in a register would be a spill
enum { S_EXACT, S_NBRQ, S_SCAN, S_BF, S_LOWQ, S_EVENTS, S_INCR, S_BIG };

struct OctSmem {
	Event ev[OCT_EV];
	uint32_t st[8];
	uint32_t ev_count;
	// vote result of the read's final pass, written by the octet's lane 0 (kept here, not in registers, until bookkeeping)
	uint32_t res_flags, res_target, res_freq, res_nref, res_nsnp, res_passes;
	uint32_t pad;
	uint64_t res_dg;
};
typedef OctSmem ReadStats;     // the emit / event helpers count through the same pointer

__device__ __forceinline__ void emit8(OctSmem *os, uint64_t kmer, uint32_t pos, uint32_t offset, uint32_t mod, uint32_t kidx,
                                      uint32_t list, ReadStats *rs)
{
	const uint32_t i = atomicAdd(&os->ev_count, 1u);
	if (i >= OCT_EV) return;                                  // overflow: the read is deferred after this pass
	Event e;
	e.kmer = kmer; e.X = pos - offset; e.kpos = pos; e.meta = mod | (kidx << 8) | (list << 16); e.pad = 0;
	os->ev[i] = e;
}

__device__ __forceinline__ void exact_ref_events8(const DevIndex &ix, OctSmem *os, uint64_t kmer, uint32_t posx, uint32_t offset,
                                                  uint32_t kidx, ReadStats *rs)
{
	if (posx == POS_AMBIGUOUS) return;
	if (posx < ix.amb_lo) { emit8(os, kmer, posx, offset, NO_MOD, kidx, 0, rs); return; }
	const uint32_t *row = ix.ref_aux + (uint64_t)(0xFFFFFFFEu - posx) * AUX_COLS;
	for (int c = 0; c < AUX_COLS; c++) {
		const uint32_t p = __ldg(row + c);
		if (p == 0) break;
		emit8(os, kmer, p, offset, NO_MOD, kidx, 0, rs);
	}
}
__device__ __forceinline__ void exact_snp_events8(const DevIndex &ix, OctSmem *os, uint64_t kmer, const SnpEntry &e, uint32_t offset,
                                                  uint32_t kidx, ReadStats *rs)
{
	if (e.pos == POS_AMBIGUOUS) return;
	if (snp_flag_of(e) == 0) { emit8(os, kmer, e.pos, offset, NO_MOD, kidx, 1, rs); return; }
	const uint32_t *row = ix.snp_aux_pos + (uint64_t)e.pos * AUX_COLS;
	for (int c = 0; c < AUX_COLS; c++) {
		const uint32_t p = __ldg(row + c);
		if (p == 0) break;
		emit8(os, kmer, p, offset, NO_MOD, kidx, 1, rs);
	}
}
__device__ __forceinline__ void nbr_ref_events8(const DevIndex &ix, OctSmem *os, uint64_t nb, uint32_t posx, uint32_t d, uint32_t offset,
                                                uint32_t kidx, ReadStats *rs)
{
	if (posx == POS_AMBIGUOUS) return;
	if (posx < ix.amb_lo) {
		if (!pile_nonzero(ix, (uint64_t)posx + d)) emit8(os, nb, posx, offset, d, kidx, 0, rs);
		return;
	}
	const uint32_t *row = ix.ref_aux + (uint64_t)(0xFFFFFFFEu - posx) * AUX_COLS;
	for (int c = 0; c < AUX_COLS; c++) {
		const uint32_t p = __ldg(row + c);
		if (p == 0) break;
		if (!pile_nonzero(ix, (uint64_t)p + d)) emit8(os, nb, p, offset, d, kidx, 0, rs);
	}
}
__device__ __forceinline__ void nbr_snp_events8(const DevIndex &ix, OctSmem *os, uint64_t nb, const SnpEntry &e, uint32_t d, uint32_t offset,
                                                uint32_t kidx, ReadStats *rs)
{
	if (e.pos == POS_AMBIGUOUS) return;
	if (snp_flag_of(e) == 0) {
		if ((snp_info_of(e) >> 3) != d) emit8(os, nb, e.pos, offset, d, kidx, 1, rs);
		return;
	}
	const uint32_t *row = ix.snp_aux_pos + (uint64_t)e.pos * AUX_COLS;
	const uint8_t *inf = ix.snp_aux_info + (uint64_t)e.pos * AUX_COLS;
	for (int c = 0; c < AUX_COLS; c++) {
		const uint32_t p = __ldg(row + c);
		if (p == 0) break;
		if (((uint32_t)__ldg(inf + c) >> 3) != d) emit8(os, nb, p, offset, d, kidx, 1, rs);
	}
}

// 32 characters starting at p -> packed 32-mer; nmask / xmask: bit b set if base b is N/n / outside ACGTNacgtn.
// Word-at-a-time (SWAR) version of encode_kmer (src/util.c:89-111): nine aligned 32-bit loads, funnel shifts, byte compares.
__device__ __forceinline__ uint64_t pack32(const char *p, uint32_t &nmask, uint32_t &xmask)
{
	const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
	const uint32_t sh = (uint32_t)(addr & 3) * 8;
	const uint32_t *wp = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
	uint32_t w[9];
#pragma unroll
	for (int i = 0; i < 9; i++) w[i] = __ldg(wp + i);
	uint64_t km = 0;
	uint32_t nm = 0, xm = 0;
#pragma unroll
	for (int i = 0; i < 8; i++) {
		const uint32_t c4 = __funnelshift_r(w[i], w[i + 1], sh);      // characters 4i .. 4i+3, first character in the low byte
		const uint32_t u = c4 & 0xDFDFDFDFu;                           // fold case
		const uint32_t va = __vcmpeq4(u, 0x41414141u), vc = __vcmpeq4(u, 0x43434343u), vg = __vcmpeq4(u, 0x47474747u),
		               vt = __vcmpeq4(u, 0x54545454u), vn = __vcmpeq4(u, 0x4E4E4E4Eu);
		const uint32_t code = ((vc | vt) & 0x01010101u) | ((vg | vt) & 0x02020202u);   // A0 C1 G2 T3 per byte
		const uint32_t pk = (code | (code >> 6) | (code >> 12) | (code >> 18)) & 0xFFu;
		km |= (uint64_t)pk << (8 * i);
		const uint32_t nb = vn & 0x01010101u;
		const uint32_t xb = ~(va | vc | vg | vt | vn) & 0x01010101u;
		nm |= ((nb | (nb >> 7) | (nb >> 14) | (nb >> 21)) & 0xFu) << (4 * i);
		xm |= ((xb | (xb >> 7) | (xb >> 14) | (xb >> 21)) & 0xFu) << (4 * i);
	}
	nmask = nm; xmask = xm;
	return km;
}

template <int MINB>
__global__ void __launch_bounds__(GW * 32, MINB) k_geno8(const GenoArgs a)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t ol = lane & 7;                 // lane inside the octet = k-mer index this lane owns
	const uint32_t ob = lane & 24;                // first lane of the octet
	OctSmem *os = reinterpret_cast<OctSmem *>(smem_raw) + (threadIdx.x >> 3);
	const DevIndex &ix = a.ix;
	const uint32_t n_reads = a.meta[1];
	const uint32_t FULL = 0xffffffffu;
#define OSHFL(v, src) __shfl_sync(FULL, (v), ob | (src))
#define OBALLOT(p) ((__ballot_sync(FULL, (p)) >> ob) & 0xFFu)

	// statistics live in shared memory (one row of 16 counters per warp), not in registers: the kernel runs at 32
	// registers per thread and every counter kept in a register is a spill somewhere else
	uint32_t *acc = reinterpret_cast<uint32_t *>(smem_raw + sizeof(OctSmem) * GW * 4) + (threadIdx.x >> 5) * 16;
	enum { A_EXACT, A_NBRQ, A_SCAN, A_BF, A_LOWQ, A_EVENTS, A_INCR, A_BIG, A_READS, A_SKIPPED, A_PASSES, A_PLACED, A_BAD, A_WRAP };
	if (lane < 16) acc[lane] = 0;
	__syncwarp();

	for (;;) {
		uint32_t r0 = 0;
		if (lane == 0) r0 = atomicAdd(&a.meta[2], 4u);
		r0 = __shfl_sync(FULL, r0, 0);
		if (r0 >= n_reads) break;
		const uint32_t r = r0 + (lane >> 3);
		const bool have = r < n_reads;

		// ---- record framing (src/qv.cc:760-779) ----
		const uint32_t lsv = (have && ol < 5) ? __ldg(a.line_start + 4ull * r + ol) : 0;
		const uint32_t id_s = OSHFL(lsv, 0), seq_s = OSHFL(lsv, 1), sep_s = OSHFL(lsv, 2), qual_s = OSHFL(lsv, 3), next_s = OSHFL(lsv, 4);
		const uint32_t L = sep_s - 1 - seq_s;
		const uint32_t qlen = next_s - 1 - qual_s;
		const uint32_t K = L >> 5;
		const bool bad_frame = have && ((seq_s - 1 - id_s > 1022) || (L > 1022) || (qual_s - 1 - sep_s > 1022) || (qlen > 1022) || (qlen < K));
		bool defer = have && !bad_frame && K > 8;
		bool active = have && !bad_frame && !defer;

		// ---- 2-bit packing: lane j packs k-mer j (src/util.c:89-111) ----
		uint64_t kmer_fwd = 0;
		uint32_t nm = 0, xm = 0;
		if (active && ol < K) kmer_fwd = pack32(a.text + seq_s + 32u * ol, nm, xm);
		// the first k-mer with an N or a foreign character decides; inside it encode_kmer meets the HIGHEST base first
		const uint32_t offm = OBALLOT((nm | xm) != 0);
		bool bad = bad_frame, skipped = false;
		{
			// shuffles are warp-wide: every octet executes them, whether or not it has an offending k-mer
			const uint32_t j = offm ? (uint32_t)__ffs(offm) - 1 : 0u;
			const uint32_t nmj = OSHFL(nm, j), xmj = OSHFL(xm, j);
			if (offm && active) {
				const uint32_t top = 31 - __clz(nmj | xmj);
				if ((xmj >> top) & 1u) bad = true; else skipped = true;
			}
		}
		if (bad || skipped) active = false;
		// quality gate of k-mer i = i-th quality CHARACTER, signed compare (src/qv.cc:836,943; F8)
		bool lowq = false;
		if (active && ol < K) lowq = ((int)(signed char)__ldg(a.text + qual_s + ol) - QUALITY_SCORE) < 0;

		if (a.debug_stage == 1) continue;
		os->st[ol] = 0;
		ReadStats *rs = os;
		bool done = !active;
		if (ol == 0) { os->res_flags = 0; os->res_target = 0; os->res_freq = 0; os->res_nref = 0; os->res_nsnp = 0; os->res_passes = 0; os->res_dg = 0; }
		uint32_t E = 0;

		for (uint32_t pass = 0; pass < 2; pass++) {
			const bool run = active && !done;
			if (!__any_sync(FULL, run)) break;
			uint64_t kmer = kmer_fwd;
			if (pass == 1) {                                   // src/qv.cc:787-806 on the packed form
				const uint64_t o = OSHFL(kmer_fwd, (K - 1 - ol) & 7);
				kmer = ol < K ? revcomp64(o) : 0;
			}
			if (run && ol == 0) os->ev_count = 0;
			if (run && ol == 0) os->res_passes = pass + 1;
			__syncwarp();

			// ---- level 1: everything that depends only on the k-mer is put in flight together ----
			const bool mine = run && ol < K;
			uint32_t rlo = 0, rhi = 0, slo = 0, shi = 0, bfr_w = 0, bfs_w = 0, bs = 0, be = 0;
			uint32_t f_lo = 0, f_hi = 0;                       // SNP block of the top 30 bits: exact membership only
			uint64_t bfr_bit = 0, bfs_bit = 0;
			const bool gates = mine && lowq;
			if (mine) {
				ref_block(ix, kmer, rlo, rhi);
				snp_block30(ix, kmer, f_lo, f_hi);
				atomicAdd(&os->st[S_EXACT], 2u);
			}
			if (gates) {
				snp_block(ix, kmer, slo, shi);                 // HI24 block: the strided scan walks it by rank (F13)
				bfr_bit = hash32((uint32_t)kmer);
				if (ix.ref_bf_bits <= 0xFFFFFFFFull) bfr_bit %= ix.ref_bf_bits;
				bfs_bit = hash40(kmer & 0xFFFFFFFFFFull) % ix.snp_bf_bits;
				if ((bfr_bit >> 5) < ix.ref_bf_nw32) bfr_w = __ldg(ix.ref_bf + (bfr_bit >> 5));
				if ((bfs_bit >> 5) < ix.snp_bf_nw32) bfs_w = __ldg(ix.snp_bf + (bfs_bit >> 5));
				ref_lo_bucket(ix, (uint32_t)kmer, bs, be);     // speculative: used only if the ref Bloom gate is open
				atomicAdd(&os->st[S_BF], 2u);
				atomicAdd(&os->st[S_LOWQ], 1u);
			}
			// ---- level 2: exact entries (src/qv.cc:840-937) ----
			if (mine) {
				uint32_t posx = 0;
				SnpEntry e;
				if (rlo < rhi && ref_find_in_block(ix, (uint32_t)kmer, rlo, rhi, posx) >= 0) exact_ref_events8(ix, os, kmer, posx, 32u * ol, ol, rs);
				if (f_lo < f_hi && snp_find_in_block(ix, kmer & 0xFFFFFFFFFFull, f_lo, f_hi, e) >= 0) exact_snp_events8(ix, os, kmer, e, 32u * ol, ol, rs);
			}
			const bool rb = gates && ((bfr_w >> (bfr_bit & 31)) & 1u);
			const bool sb = gates && ((bfs_w >> (bfs_bit & 31)) & 1u);
			if (!rb) { bs = 0; be = 0; }
			const uint32_t rB = rhi - rlo, sB = shi - slo;
			const bool big = rB >= BLOCK_SIZE_THRESHOLD;       // src/qv.cc:843,962
			if (gates) { if (rb) atomicAdd(&os->st[S_NBRQ], 48u); if (big) atomicAdd(&os->st[S_BIG], 1u); }

			if (a.debug_stage == 2) { done = true; continue; }
			// ---- Hamming-1 neighbours: the octet works through its low-quality k-mers one at a time ----
			uint32_t lm = OBALLOT(gates);
			const uint32_t rounds = __reduce_max_sync(FULL, (uint32_t)__popc(lm));
			for (uint32_t it = 0; it < rounds; it++) {
				const uint32_t i = lm ? (uint32_t)__ffs(lm) - 1 : 0;
				const bool on = lm != 0;
				lm &= lm - 1;
				const uint64_t km = OSHFL(kmer, i);
				const uint32_t k_rlo = OSHFL(rlo, i), k_rB = OSHFL(rB, i), k_slo = OSHFL(slo, i), k_sB = OSHFL(sB, i);
				const uint32_t k_bs = OSHFL(bs, i), k_be = OSHFL(be, i);
				const bool k_sb = OSHFL((uint32_t)sb, i) != 0, k_big = OSHFL((uint32_t)big, i) != 0;
				if (!on) continue;
				const uint32_t offset = 32u * i;
				const uint32_t n0 = k_be - k_bs;                  // upper half, ref: LO32 bucket walk for the 48 queries of :1225
				const uint32_t n1 = k_sb ? 36u : 0u;              // upper half, snp, d = 20..31 (:1305-1307)
				const uint32_t n2 = k_big ? 12u : 0u;             // upper half, snp, d = 16..19 in big mode
				const uint32_t n3 = k_big ? 48u : k_rB;           // lower half, ref: queries (:975) or strided scan (:358-373)
				const uint32_t n4 = k_big ? 48u : k_sB;           // lower half, snp: queries (:977) or strided scan (:447-462)
				const uint32_t e0 = n0, e1 = e0 + n1, e2 = e1 + n2, e3 = e2 + n3, e4 = e3 + n4;
				for (uint32_t t = ol; t < e4; t += 8) {
					if (t < e0) {
						const uint2 en = __ldg(reinterpret_cast<const uint2 *>(ix.ref_by_lo + k_bs + t));
						const int sl = one_base_slot((uint64_t)(en.x ^ (uint32_t)(km >> 32)));
						if (sl >= 0) nbr_ref_events8(ix, os, ((uint64_t)en.x << 32) | (uint32_t)km, en.y, 16u + (uint32_t)sl, offset, i, rs);
					} else if (t >= e2 && t < e3 && k_big) {
						const uint32_t u = t - e2, d = u / 3;
						const uint64_t nb = substitute(km, d, u % 3);
						uint32_t posx;
						atomicAdd(&os->st[S_NBRQ], 1u);
						if (ref_query(ix, nb, posx) >= 0) nbr_ref_events8(ix, os, nb, posx, d, offset, i, rs);
					} else if (t < e2 || (t >= e3 && k_big)) {
						uint32_t u, d;
						if (t < e1) { u = t - e0; d = 20u + u / 3; }
						else if (t < e2) { u = t - e1; d = 16u + u / 3; }
						else { u = t - e3; d = u / 3; }
						const uint64_t nb = substitute(km, d, u % 3);
						SnpEntry e;
						atomicAdd(&os->st[S_NBRQ], 1u);
						if (snp_query(ix, nb, e) >= 0) nbr_snp_events8(ix, os, nb, e, d, offset, i, rs);
					} else if (t < e3) {                              // ref strided scan step (F13)
						const uint32_t s = t - e2;
						const uint64_t ex = (uint64_t)k_rlo + (uint64_t)REF_STRIDE * s;
						atomicAdd(&os->st[S_SCAN], 1u);
						if (ex < ix.n_ref) {
							const uint32_t entry_lo = __ldg(&ix.ref[ex].lo);
							const int d = one_base_slot((uint64_t)((uint32_t)km ^ entry_lo));
							if (d >= 0) nbr_ref_events8(ix, os, (km & 0xFFFFFFFF00000000ull) | entry_lo, __ldg(&ix.ref[k_rlo + s].posx), (uint32_t)d, offset, i, rs);
						}
					} else {                                          // snp strided scan step (F13)
						const uint32_t s = t - e3;
						const uint64_t ex = (uint64_t)k_slo + (uint64_t)SNP_STRIDE * s;
						atomicAdd(&os->st[S_SCAN], 1u);
						if (ex < ix.n_snp) {
							const uint64_t entry_lo = snp_scan_lo40(ix, k_slo, s);
							const int d = one_base_slot((km & 0xFFFFFFFFFFull) ^ entry_lo);
							if (d >= 0) {
								const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(ix.snp + k_slo + s));
								SnpEntry e;
								e.key = ((uint64_t)raw.y << 32) | raw.x; e.pos = raw.z; e.extra = raw.w;
								nbr_snp_events8(ix, os, (km & 0xFFFFFF0000000000ull) | entry_lo, e, (uint32_t)d, offset, i, rs);
							}
						}
					}
				}
			}
			__syncwarp();
			if (a.debug_stage == 3) { done = true; continue; }

			// ---- vote (src/qv.cc:132-178, order-independent form) ----
			E = run ? os->ev_count : 0;
			if (run && ol == 0) os->st[S_EVENTS] += E;
			if (E > OCT_EV) { defer = true; done = true; E = 0; }   // too many contexts for shared memory: redo in k_geno
			const bool vrun = run && !defer;
			for (uint32_t e = ol; e < E; e += 8) {
				Event *p = &os->ev[e];
				const uint32_t m = p->meta, X = p->X;
				bool v = (m & 0xFF) == NO_MOD;
				if (!v) {
					const uint32_t ki = (m >> 8) & 0xFF;
					for (uint32_t f = 0; f < E && !v; f++) {
						const Event *q = &os->ev[f];
						v = ((q->meta & 0xFF) == NO_MOD) && q->X == X && ((q->meta >> 8) & 0xFF) <= ki;
					}
				}
				p->meta = (m & ~(1u << 17)) | (v ? (1u << 17) : 0u);
			}
			__syncwarp();
			uint32_t bf_ = 0, bxmin = 0xFFFFFFFFu, bxmax = 0;
			uint32_t t_nref = 0, t_nsnp = 0;
			uint64_t t_dg = 0;
			for (uint32_t e = ol; e < E; e += 8) {
				const Event *p = &os->ev[e];
				const uint32_t m = p->meta, X = p->X, kp = p->kpos;
				if ((m >> 16) & 1) t_nsnp++; else t_nref++;
				if (a.trace) t_dg += ctx_digest((m >> 16) & 1, X, kp, p->kmer, (m & 0xFF) == NO_MOD ? 10086u : (m & 0xFF));
				if (!((m >> 17) & 1)) continue;
				uint32_t f = 0;
				bool distinct = false;
				for (uint32_t g = 0; g < E; g++) {
					const Event *q = &os->ev[g];
					if (((q->meta >> 17) & 1) && q->X == X) { f++; distinct |= (q->kpos != kp); }
				}
				if (!distinct) continue;
				if (f > bf_) { bf_ = f; bxmin = X; bxmax = X; }
				else if (f == bf_) { bxmin = min(bxmin, X); bxmax = max(bxmax, X); }
			}
			// octet reductions (xor 1, 2, 4 stay inside the octet)
			uint32_t maxf = bf_;
#pragma unroll
			for (int o = 1; o < 8; o <<= 1) maxf = max(maxf, __shfl_xor_sync(FULL, maxf, o));
			uint32_t xmin = bf_ == maxf ? bxmin : 0xFFFFFFFFu, xmax = bf_ == maxf ? bxmax : 0u;
#pragma unroll
			for (int o = 1; o < 8; o <<= 1) {
				xmin = min(xmin, __shfl_xor_sync(FULL, xmin, o));
				xmax = max(xmax, __shfl_xor_sync(FULL, xmax, o));
				t_nref += __shfl_xor_sync(FULL, t_nref, o);
				t_nsnp += __shfl_xor_sync(FULL, t_nsnp, o);
				t_dg += __shfl_xor_sync(FULL, t_dg, o);
			}
			const bool has_best = maxf > 0;
			const bool ambiguous = has_best && xmin != xmax;
			const bool process = has_best && !ambiguous;       // freq > 1 is implied by two distinct k-mer positions (:1375)
			const uint32_t target = xmin;
			if (vrun && ol == 0) {
				os->res_nref = t_nref; os->res_nsnp = t_nsnp; os->res_dg = t_dg; os->res_target = target; os->res_freq = maxf;
				os->res_flags = (pass ? VGB_RF_REVCOMPL : 0) | (process ? VGB_RF_PROCESS : 0) | (ambiguous ? VGB_RF_AMBIGUOUS : 0) |
				                (has_best ? VGB_RF_HASBEST : 0);
			}

			if (a.debug_stage == 4) { done = true; continue; }
			// ---- pileup update: every recorded context at the winning position (src/qv.cc:1382-1502) ----
			if (vrun && process) {
				const uint64_t n_blk = ix.pile_len >> 6;
				for (uint32_t e = 0; e < E; e++) {
					const Event *p = &os->ev[e];
					if (p->X != target) continue;
					const uint32_t mod = p->meta & 0xFF;
					const uint64_t kmer_e = p->kmer;
					const uint64_t kpos = p->kpos;
					// the 32-position window [kpos, kpos+32) lies in one or two 64-position blocks; every lane of the octet
					// loads the same 16-byte block records (one transaction), then looks only at its own four positions
					const uint64_t bA = kpos >> 6, bB = (kpos + 31) >> 6;
					uint4 ka = make_uint4(0, 0, 0, 0), kb = make_uint4(0, 0, 0, 0);
					if (bA < n_blk) ka = __ldg(reinterpret_cast<const uint4 *>(ix.pile + bA));
					if (bB != bA && bB < n_blk) kb = __ldg(reinterpret_cast<const uint4 *>(ix.pile + bB));
					const uint64_t bitsA = ((uint64_t)ka.y << 32) | ka.x, bitsB = ((uint64_t)kb.y << 32) | kb.x;
					const uint32_t sh = (uint32_t)(kpos & 63);
					uint32_t win = (uint32_t)(bitsA >> sh);
					if (sh > 32) win |= (uint32_t)(bitsB << (64 - sh));
					uint32_t mineb = win & (0x01010101u << ol);        // my positions: ol, ol+8, ol+16, ol+24
					if (mod < 32) mineb &= ~(1u << mod);               // the substituted base does not count (:1391)
					while (mineb) {
						const uint32_t b = __ffs(mineb) - 1;
						mineb &= mineb - 1;
						const uint32_t off = sh + b;                    // bit index relative to block A
						const uint32_t sid = off < 64 ? ka.z + __popcll(bitsA & ((1ull << off) - 1))
						                              : kb.z + __popcll(bitsB & ((1ull << (off - 64)) - 1));
						const uint32_t code = __ldg(ix.site_code + sid);
						const uint32_t rbase = code & 3, abase = code >> 2;
						if (rbase == abase) continue;                  // p->ref != p->alt (:1404)
						const uint32_t base = (uint32_t)(kmer_e >> (2 * b)) & 3u;
						if (base == rbase) { atomicAdd(ix.cnt + 2ull * sid, 1u); atomicAdd(&os->st[S_INCR], 1u); }
						else if (base == abase) { atomicAdd(ix.cnt + 2ull * sid + 1, 1u); atomicAdd(&os->st[S_INCR], 1u); }
					}
				}
			}
			__syncwarp();
			if (vrun && (process || pass == 1)) done = true;      // otherwise: retry once on the reverse complement (:1504-1510)
		}

		// ---- per-read bookkeeping (lane 0 of the octet speaks for the read) ----
		if (have && !defer) {
			const uint32_t v = os->st[ol];                     // lane ol commits counter ol (S_* and A_* share the first 8 slots)
			if (v) atomicAdd(&acc[ol], v);
		}
		if (have && ol == 0) {
			if (defer) {
				a.defer[atomicAdd(&a.meta[6], 1u)] = r;
			} else {
				atomicAdd(&acc[A_READS], 1u);
				const uint32_t passes = os->res_passes, flags = os->res_flags, best_freq = os->res_freq;
				if (passes) atomicAdd(&acc[A_PASSES], passes);
				if (bad) { atomicAdd(&acc[A_BAD], 1u); atomicOr(&a.meta[3], 2u); }
				else if (skipped) atomicAdd(&acc[A_SKIPPED], 1u);
				else {
					if (flags & VGB_RF_PROCESS) atomicAdd(&acc[A_PLACED], 1u);
					if (best_freq > 255) atomicAdd(&acc[A_WRAP], 1u);
				}
				if (a.trace) {
					vgb_read_result res;
					res.flags = (bad || skipped) ? VGB_RF_SKIPPED : flags;
					res.target = os->res_target; res.freq = (uint16_t)(best_freq & 0xFF);
					res.n_ref = (uint16_t)os->res_nref; res.n_snp = (uint16_t)os->res_nsnp; res.passes = (uint16_t)passes; res.ctx_hash = os->res_dg;
					if (bad || skipped) { res.target = 0; res.freq = 0; res.n_ref = 0; res.n_snp = 0; res.passes = 0; res.ctx_hash = 0; }
					a.trace[r] = res;
				}
			}
		}
	}
#undef OSHFL
#undef OBALLOT

	__syncwarp();
	if (lane == 0) {
		DevStats *s = a.stats;
		atomicAdd(&s->reads, (unsigned long long)acc[A_READS]); atomicAdd(&s->skipped_n, (unsigned long long)acc[A_SKIPPED]);
		atomicAdd(&s->passes, (unsigned long long)acc[A_PASSES]); atomicAdd(&s->placed, (unsigned long long)acc[A_PLACED]);
		atomicAdd(&s->exact_lookups, (unsigned long long)acc[A_EXACT]); atomicAdd(&s->nbr_query_lookups, (unsigned long long)acc[A_NBRQ]);
		atomicAdd(&s->nbr_scan_reads, (unsigned long long)acc[A_SCAN]); atomicAdd(&s->bf_probes, (unsigned long long)acc[A_BF]);
		atomicAdd(&s->lowq_kmers, (unsigned long long)acc[A_LOWQ]); atomicAdd(&s->events, (unsigned long long)acc[A_EVENTS]);
		atomicAdd(&s->pileup_incr, (unsigned long long)acc[A_INCR]); atomicAdd(&s->big_kmers, (unsigned long long)acc[A_BIG]);
		if (acc[A_BAD]) atomicAdd(&s->bad_records, (unsigned long long)acc[A_BAD]);
		if (acc[A_WRAP]) atomicAdd(&s->freq_wrap_reads, (unsigned long long)acc[A_WRAP]);
	}
}
