// vgb_geno8.inl -- the main per-read kernel: G = 4 or 8 lanes per read, 32 / G reads per warp, ONE PASS per round.
// Included by vgb_geno.cu.
//
// Same semantics as k_geno (one warp per read), different mapping.  A 150 bp read has 4 k-mers = 8 exact probes, so a
// full warp per read leaves 24 lanes idle in the phases that matter and puts only one read's dependent probe chain in
// flight per warp (profiles/r01_summary.md).  Here lane j of a group owns k-mer j of its read for packing, exact probes,
// Bloom gates and bucket bounds; neighbour tasks, the vote and the pileup contexts are spread over the group's lanes.
// All groups of a warp move through the phases together, so every warp-wide shuffle / ballot is executed convergently;
// loops whose trip count differs per group contain no warp-synchronous operation.
//
// Four kernels run per chunk: G = 4 over every read (up to 4 k-mers = up to 159 bases: eight reads and eight probe
// chains per warp), G = 8 over the reads that one handed over (5..8 k-mers), G = 8 with EV_WIDE hit contexts per read over the
// reads either of them gave up on because a pass recorded more than EV_GROUP contexts (reads from repeat families: four
// reads per warp instead of one, 0.16 -> ~0.05 ms per 2 M GRCh38-shaped reads), then k_geno (one warp per read) over what is
// left: longer reads and reads with more than EV_WIDE contexts in a pass.  Lists: a.klist / a.in_cnt in, a.defer /
// a.defer_cnt out; bit 31 of a list entry = the forward pass is already done and accounted for.
//
// Rounds: a warp round runs ONE pass (src/qv.cc:778-1510 is "forward pass, then one retry on the reverse complement",
// :1504-1510) for 32 / G reads.  Reads whose forward pass places nothing are parked in a small per-warp queue (packed
// k-mers + quality gates) and the warp runs a retry round as soon as a full set is waiting -- so the second pass, which
// half of all reads need (reverse-strand reads), is executed with every group busy instead of half of them.

#ifndef VGB_OCT_EV
#define VGB_OCT_EV 24
#endif
constexpr int EV_GROUP = VGB_OCT_EV;  // hit contexts per read kept in shared memory (16 measured the same on S1 and at GRCh38 size)
constexpr int EV_WIDE = 64;           // ... in the instantiation behind them (4 k-mers x 10 aux columns + neighbours fit)

// per-round counters of one read, in the group's shared memory (committed when the round ends, unless the read is
// deferred in this round); the first eight A_* slots of the per-warp accumulator have the same meaning
enum { S_EXACT, S_NBRQ, S_SCAN, S_BF, S_LOWQ, S_EVENTS, S_INCR, S_BIG };

// hit context, 16 bytes: kmer_context.position is X + 32 * (k-mer index), because every context of k-mer i is recorded
// with offset 32 i (src/qv.cc:850-937, 985-1101)
struct __align__(8) GEvent {
	uint64_t kmer;
	uint32_t X;       // read position this context votes for
	uint32_t meta;    // bits 0-7 modified base (0xFF none) | 8-15 k-mer index | 16 list (0 ref, 1 snp) | 17 votes
};
template <int EVN>
struct OctSmem {
	GEvent ev[EVN];
	uint32_t st[8];
	uint32_t ev_count;
	uint32_t pad;
};
template <int G>
struct Pend {
	uint64_t kmer[G];                 // forward-strand packed k-mers
	uint32_t r, K, lowq, pad;         // read index in the chunk, k-mer count, quality gates (bit i = k-mer i)
};

template <int EVN>
__device__ __forceinline__ void emit8(OctSmem<EVN> *os, uint64_t kmer, uint32_t pos, uint32_t offset, uint32_t mod, uint32_t kidx, uint32_t list)
{
	const uint32_t i = atomicAdd(&os->ev_count, 1u);
	if (i >= (uint32_t)EVN) return;                           // overflow: the read is deferred after this pass
	GEvent e;
	e.kmer = kmer; e.X = pos - offset; e.meta = mod | (kidx << 8) | (list << 16);
	os->ev[i] = e;
}

// A dictionary hit -> hit contexts (src/qv.cc:850-937 exact, :985-1101 and twins for neighbours with modified base d).
//   list 0: v = posx of the RefEntry;  list 1: v = pos field of the SnpEntry, fi = snp_info | ambig_flag << 8.
// d == NO_MOD: exact hit, nothing is vetoed.  Otherwise a reference neighbour is dropped when its modified base sits on
// a SNP site (:990-991) and a SNP neighbour when the modified base IS the SNP base (:1055).
// Entries that stand for 2..10 positions go through their aux row; the loop is kept rolled (rare, and this body is
// instantiated at three call sites: code size is what the instruction cache sees).
template <int EVN>
__device__ __forceinline__ void hit8(const DevIndex &ix, OctSmem<EVN> *os, uint32_t list, uint64_t kmer, uint32_t v, uint32_t fi, uint32_t d,
                                     uint32_t offset, uint32_t kidx)
{
	if (v == POS_AMBIGUOUS) return;
	const bool single = list ? (fi >> 8) == 0 : v < ix.amb_lo;
	if (single) {
		bool veto = false;
		if (d != NO_MOD) veto = list ? ((fi & 0xFFu) >> 3) == d : pile_nonzero(ix, (uint64_t)v + d);
		if (!veto) emit8(os, kmer, v, offset, d, kidx, list);
		return;
	}
	const uint64_t row = (uint64_t)(list ? v : 0xFFFFFFFEu - v) * AUX_COLS;
	const uint32_t *rowp = (list ? ix.snp_aux_pos : ix.ref_aux) + row;
#pragma unroll 1
	for (int c = 0; c < AUX_COLS; c++) {
		const uint32_t p = __ldg(rowp + c);
		if (p == 0) break;
		bool veto = false;
		if (d != NO_MOD) veto = list ? ((uint32_t)ldr(ix.snp_aux_info + row + c) >> 3) == d : pile_nonzero(ix, (uint64_t)p + d);
		if (!veto) emit8(os, kmer, p, offset, d, kidx, list);
	}
}

// 32 characters starting at p -> packed 32-mer (encode_kmer, src/util.c:89-111).  `off`: 0 = all of ACGTacgt; otherwise
// what encode_kmer meets first (it walks from the HIGHEST base down): 1 = an N/n (the read is skipped), 2 = a character
// outside ACGTNacgtn (the reference aborts there; counted as a bad record here).
// Word-at-a-time: nine aligned 32-bit loads, funnel shifts, then per 4 characters
//   t = (c >> 1) & 3  ->  A 0, C 1, T 2, G 3;   code = t ^ (t >> 1)  ->  A 0, C 1, G 2, T 3
// and the test "the case-folded character is the letter its code stands for" ('A' + {0, 2, 6, 19}) in one compare per
// word; the last word that fails it holds the highest offending character.  (Recomputing per-base N / foreign masks with
// byte compares for offending k-mers cost ~480 instructions per warp round on S1, where a fifth of the reads are N runs.)
__device__ __forceinline__ uint64_t pack32(const char *p, uint32_t &off)
{
	const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
	const uint32_t sh = (uint32_t)(addr & 3) * 8;
	const uint32_t *wp = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
	uint32_t w[9];
#pragma unroll
	for (int i = 0; i < 9; i++) w[i] = __ldg(wp + i);
	uint32_t klo = 0, khi = 0, bw = 0, uw = 0;
#pragma unroll
	for (int i = 0; i < 8; i++) {
		const uint32_t u = __funnelshift_r(w[i], w[i + 1], sh) & 0xDFDFDFDFu;   // characters 4i .. 4i+3, case folded
		const uint32_t t = (u >> 1) & 0x03030303u;
		const uint32_t code = t ^ ((t >> 1) & 0x01010101u);
		const uint32_t b0 = code & 0x01010101u, b1 = (code >> 1) & 0x01010101u;
		const uint32_t expect = 0x41414141u + 2u * b0 + 6u * b1 + 11u * (b0 & b1);
		const uint32_t b = u ^ expect;
		if (b) { bw = b; uw = u; }                                 // words ascend: the last one kept is the highest
		// bytes {c0, c1, c2, c3} (2 bits each) -> one byte c0 | c1 << 2 | c2 << 4 | c3 << 6: the four partial products land
		// on disjoint bit ranges, bits 24..31 of the 32-bit product are the packed byte
		const uint32_t pk = (code * 0x01041040u) >> 24;
		if (i < 4) klo |= pk << (8 * i); else khi |= pk << (8 * (i - 4));
	}
	off = 0;
	if (bw) {
		const uint32_t top = (31u - (uint32_t)__clz(bw)) & ~7u;   // bit offset of the highest offending character in its word
		off = ((uw >> top) & 0xFFu) == 0x4Eu ? 1u : 2u;           // 'N' after case folding
	}
	return ((uint64_t)khi << 32) | klo;
}

// SHORTSCAN: instantiation for SNP dictionaries whose HI24 blocks hold a few entries (a chr22-sized list: two), where the strided
// scan ends inside its first 16-byte load: the compares of the second load are then skipped by a branch.  With GRCh38-sized
// blocks (~23 entries) the branch costs more than it saves (-3 % / -5 % of the gain on S2 / S3), without it the short scans pay
// for eight compares they do not need (+3 % on S1): chosen per index at upload (geno_prepare).
template <int MINB, bool TRACE, int G, int EVN, bool SHORTSCAN = false>
__global__ void __launch_bounds__(GW * 32, MINB) k_geno8(const GenoArgs a)
{
	typedef OctSmem<EVN> OS;
	static_assert(G == 4 || G == 8, "lanes per read");
	constexpr uint32_t R = 32 / G;                // reads per warp round
	constexpr uint32_t GM = (1u << G) - 1u;
	constexpr int PEND_CAP = 2 * R;               // parked reads per warp: at most R - 1 waiting + R from one forward round
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t ol = lane & (G - 1);           // lane inside the group = k-mer index this lane owns
	const uint32_t ob = lane & ~(uint32_t)(G - 1);   // first lane of the group
	const uint32_t gi = lane / G;                 // group inside the warp
	OS *os = reinterpret_cast<OS *>(smem_raw) + (threadIdx.x / G);
	// one row of 16 counters per warp, then the warp's queue of parked reads
	uint32_t *acc = reinterpret_cast<uint32_t *>(smem_raw + sizeof(OS) * GW * R) + (threadIdx.x >> 5) * 16;
	Pend<G> *pend = reinterpret_cast<Pend<G> *>(smem_raw + sizeof(OS) * GW * R + GW * 16 * sizeof(uint32_t)) + (threadIdx.x >> 5) * PEND_CAP;
	const DevIndex &ix = a.ix;
	// list mode (the instantiations behind the 4-lane one): only the reads handed over by the kernels in front
	const uint32_t n_reads = a.klist ? a.meta[a.in_cnt] : a.meta[1];
	uint32_t *work = a.meta + (a.klist ? a.in_cnt + 1 : 2);
	const uint32_t *ls0 = a.line_start + a.meta[11];      // first line of the chunk's own records (non-zero only for BGZF windows)
	const uint32_t FULL = 0xffffffffu;
#define OSHFL(v, src) __shfl_sync(FULL, (v), ob | (src))
#define OBALLOT(p) ((__ballot_sync(FULL, (p)) >> ob) & GM)
	enum { A_EXACT, A_NBRQ, A_SCAN, A_BF, A_LOWQ, A_EVENTS, A_INCR, A_BIG, A_READS, A_SKIPPED, A_PASSES, A_PLACED, A_BAD, A_WRAP };
	if (n_reads == 0) return;                     // list mode with nothing handed over (the usual case for 150 bp reads)
	if (lane < 16) acc[lane] = 0;
	__syncwarp();

	uint32_t npend = 0;                           // warp-uniform
	bool fresh_left = true;                       // warp-uniform

	for (;;) {
		uint32_t pass, r = 0, K = 0;
		bool have = false, run = false, defer = false, wide = false, bad = false, skipped = false, lowq = false, dretry = false;
		uint64_t kmer = 0;

		if (npend >= R || (!fresh_left && npend)) {
			// ---- retry round: up to R parked reads, on the reverse complement (src/qv.cc:787-806 on the packed form) ----
			pass = 1;
			const uint32_t take = min(npend, R);
			have = gi < take;
			if (have) {
				const Pend<G> *p = &pend[npend - 1 - gi];
				r = p->r; K = p->K; lowq = (p->lowq >> ol) & 1u;
				if (ol < K) kmer = revcomp64(p->kmer[K - 1 - ol]);
				run = true;
			}
			npend -= take;
		} else {
			if (!fresh_left) break;
			uint32_t r0 = 0;
			if (lane == 0) r0 = atomicAdd(work, R);
			r0 = __shfl_sync(FULL, r0, 0);
			if (r0 >= n_reads) { fresh_left = false; continue; }
			pass = 0;
			have = r0 + gi < n_reads;
			if (have) r = a.klist ? __ldg(a.klist + r0 + gi) : r0 + gi;
			if (EVN == EV_WIDE) {                             // only this instantiation is handed reads in the middle of their two passes
				dretry = (r >> 31) != 0;                      // list entry: forward pass done where the read came from, retry only
				r &= 0x7FFFFFFFu;
			}

			// ---- record framing (src/qv.cc:760-779): line starts 4r .. 4r+4 ----
			uint32_t lsv = 0, lsn = 0;
			if (have && ol < 4) lsv = __ldg(ls0 + 4ull * r + ol);
			if (have && ol == 0) lsn = __ldg(ls0 + 4ull * r + 4);
			const uint32_t id_s = OSHFL(lsv, 0), seq_s = OSHFL(lsv, 1), sep_s = OSHFL(lsv, 2), qual_s = OSHFL(lsv, 3), next_s = OSHFL(lsn, 0);
			const uint32_t L = sep_s - 1 - seq_s;
			const uint32_t qlen = next_s - 1 - qual_s;
			K = L >> 5;
			const bool bad_frame = have && ((seq_s - 1 - id_s > 1022) || (L > 1022) || (qual_s - 1 - sep_s > 1022) || (qlen > 1022) || (qlen < K));
			wide = have && !bad_frame && K > (uint32_t)G;       // more k-mers than this instantiation has lanes
			bool active = have && !bad_frame && !wide;

			// ---- 2-bit packing: lane j packs k-mer j (src/util.c:89-111) ----
			uint32_t off = 0;
			if (active && ol < K) kmer = pack32(a.text + seq_s + 32u * ol, off);
			// the first k-mer with an N or a foreign character decides; inside it encode_kmer meets the HIGHEST base first
			const uint32_t offm = OBALLOT(off != 0);
			bad = bad_frame;
			if (__any_sync(FULL, offm != 0)) {
				// shuffles are warp-wide: every group executes them, whether or not it has an offending k-mer
				const uint32_t offj = OSHFL(off, offm ? (uint32_t)__ffs(offm) - 1 : 0u);
				if (offm && active) { if (offj == 2u) bad = true; else skipped = true; }
			}
			if (bad || skipped) active = false;
			// quality gate of k-mer i = i-th quality CHARACTER, signed compare (src/qv.cc:836,943; F8)
			if (active && ol < K) lowq = ((int)(signed char)__ldg(a.text + qual_s + ol) - QUALITY_SCORE) < 0;
			run = active;
			if (EVN == EV_WIDE) {
				run = active && !dretry;
				dretry = dretry && active;                    // packed and gated here, parked below, run in a retry round
			}
		}

		os->st[ol] = 0;
		if (G == 4) os->st[ol + 4] = 0;
		if (ol == 0) os->ev_count = 0;
		__syncwarp();

		// ---- level 1: everything that depends only on the k-mer is put in flight together ----
		const bool mine = run && ol < K;
		uint32_t rlo = 0, rhi = 0, slo = 0, shi = 0, bfs_w = 0, bs = 0, be = 0;
		bool rb = false;
		uint32_t f_lo = 0, f_hi = 0;                       // SNP block of the top 30 bits: exact membership only
		uint64_t bfs_bit = 0;
		const bool gates = mine && lowq;
		bool rmay = false, smay = false;                   // false: the directory's fingerprint says "not the entry of this block"
		if (mine) dir_lookup(ix, kmer, rlo, rhi, f_lo, f_hi, rmay, smay);
		if (gates) {
			snp_block(ix, kmer, slo, shi);                 // HI24 block: the strided scan walks it by rank (F13)
			bfs_bit = hash40(kmer & 0xFFFFFFFFFFull) % ix.snp_bf_bits;
			if ((bfs_bit >> 5) < ix.snp_bf_nw32) bfs_w = ldr(ix.snp_bf + (bfs_bit >> 5));
			ref_lo_gate_bucket(ix, (uint32_t)kmer, rb, bs, be);   // the reference Bloom gate and the LO32 bucket bounds: one record
		}
		const bool sb = gates && ((bfs_w >> (bfs_bit & 31)) & 1u);
		// SNP gate open: bounds of the "same LO40" group that answers the 36 upper-half SNP queries (vgb_common.cuh); issued
		// here so that it travels under the exact-entry searches
		uint32_t qs = 0, qe = 0;
		if (sb) snp_lo_bucket(ix, kmer & 0xFFFFFFFFFFull, qs, qe);
		// ---- level 2: exact entries (src/qv.cc:840-937) ----
		bool rfound = false;
		if (mine) {
			uint32_t posx = 0;
			SnpEntry e;
			if (rlo < rhi && rmay && ref_find_in_block(ix, (uint32_t)kmer, rlo, rhi, posx) >= 0) { rfound = true; hit8(ix, os, 0, kmer, posx, 0, NO_MOD, 32u * ol, ol); }
			if (f_lo < f_hi && smay && snp_find_in_block(ix, kmer & 0xFFFFFFFFFFull, f_lo, f_hi, e) >= 0)
				hit8(ix, os, 1, kmer, e.pos, (uint32_t)(e.key >> 40) & 0xFFFFu, NO_MOD, 32u * ol, ol);
		}
		if (!rb) { bs = 0; be = 0; }
		// the k-mer is in the reference dictionary and its LO32 bucket holds one entry: that entry is the k-mer itself, none of
		// the 48 upper-half neighbours exists -- no need to read the bucket (half of all low-quality k-mers of correct-strand passes)
		else if (rfound && be - bs == 1u) be = bs;
		const uint32_t rB = rhi - rlo, sB = shi - slo;
		const bool big = rB >= BLOCK_SIZE_THRESHOLD;       // src/qv.cc:843,962
		// lookup accounting: 2K exact queries per pass; per low-quality k-mer the neighbour probes its gates open
		if (gates) {
			atomicAdd(&os->st[S_LOWQ], 1u);
			atomicAdd(&os->st[S_NBRQ], (rb ? 48u : 0u) + (sb ? 36u : 0u) + (big ? 12u + 96u : 0u));
			if (big) atomicAdd(&os->st[S_BIG], 1u);
			else if (rB + sB) atomicAdd(&os->st[S_SCAN], rB + sB);
		}

		// ---- Hamming-1 neighbours: the group works through its low-quality k-mers one at a time ----
		uint32_t lm = OBALLOT(gates);
		const uint32_t rounds = __reduce_max_sync(FULL, (uint32_t)__popc(lm));
		for (uint32_t it = 0; it < rounds; it++) {
			const uint32_t i = lm ? (uint32_t)__ffs(lm) - 1 : 0;
			const bool on = lm != 0;
			lm &= lm - 1;
			const uint64_t km = OSHFL(kmer, i);
			const uint32_t k_rlo = OSHFL(rlo, i), k_rB = OSHFL(rB, i), k_slo = OSHFL(slo, i), k_sB = OSHFL(sB, i);
			const uint32_t k_bs = OSHFL(bs, i), k_be = OSHFL(be, i);
			const uint32_t k_qs = OSHFL(qs, i), k_qe = OSHFL(qe, i);
			const bool k_big = OSHFL((uint32_t)big, i) != 0;
			if (!on) continue;
			const uint32_t offset = 32u * i;
			const uint32_t n0 = k_be - k_bs;                  // upper half, ref: LO32 bucket walk for the 48 queries of :1225
			const uint32_t n1 = k_qe - k_qs;                  // upper half, snp, d = 20..31 (:1305-1307): LO40 group walk for the 36 queries
			const uint32_t n2 = k_big ? 12u : 0u;             // upper half, snp, d = 16..19 in big mode
			const uint32_t n3 = k_big ? 48u : k_rB;           // lower half, ref: queries (:975) or strided scan (:358-373)
			const uint32_t n4 = k_big ? 48u : 0u;             // lower half, snp: queries in big mode (:977); the strided scan runs below
			const uint32_t e0 = n0, e1 = e0 + n1, e2 = e1 + n2, e3 = e2 + n3, e4 = e3 + n4;
			for (uint32_t t = ol; t < e4; t += G) {
				// every kind of task ends in "a dictionary entry was found" -> one shared tail (hit8)
				bool hit = false;
				uint32_t list = 0, v = 0, fi = 0, d = 0;
				uint64_t nb = 0;
				if (t < e0) {                                     // LO32 bucket entry
					const uint2 en = ldr(reinterpret_cast<const uint2 *>(ix.ref_by_lo + k_bs + t));
					const int sl = one_base_slot((uint64_t)(en.x ^ (uint32_t)(km >> 32)));
					if (sl >= 0) { hit = true; nb = ((uint64_t)en.x << 32) | (uint32_t)km; v = en.y; d = 16u + (uint32_t)sl; }
				} else if (t < e1) {                              // LO40 group entry: same lower 20 bases, one base of the upper 12 differs?
					const uint4 en = ldr(ix.snp_by_lo + k_qs + (t - e0));
					const uint64_t ek = ((uint64_t)en.y << 32) | en.x;
					const uint64_t x = ek ^ km;
					if ((x & 0xFFFFFFFFFFull) == 0) {
						const int sl = one_base_slot(x >> 40);
						if (sl >= 0) { hit = true; list = 1; nb = ek; v = en.z; fi = en.w; d = 20u + (uint32_t)sl; }
					}
				} else if (t < e2 || t >= e3) {                   // snp query (big mode only)
					uint32_t u;
					if (t < e2) { u = t - e1; d = 16u + u / 3; }
					else { u = t - e3; d = u / 3; }
					nb = substitute(km, d, u % 3);
					SnpEntry e;
					if (snp_query(ix, nb, e) >= 0) { hit = true; list = 1; v = e.pos; fi = (uint32_t)(e.key >> 40) & 0xFFFFu; }
				} else if (k_big) {                               // ref query (big mode, lower half): t in [e2, e3)
					const uint32_t u = t - e2;
					d = u / 3;
					nb = substitute(km, d, u % 3);
					uint32_t posx;
					if (ref_query(ix, nb, posx) >= 0) { hit = true; v = posx; }
				} else {                                          // ref strided scan step (F13): t in [e2, e3)
					const uint32_t st = t - e2;
					const uint64_t ex = (uint64_t)k_rlo + (uint64_t)REF_STRIDE * st;
					if (ex < ix.n_ref) {
						const uint32_t entry_lo = ldr(&ix.ref[ex].lo);
						const int dd = one_base_slot((uint64_t)((uint32_t)km ^ entry_lo));
						if (dd >= 0) { hit = true; nb = (km & 0xFFFFFFFF00000000ull) | entry_lo; v = ldr(&ix.ref[k_rlo + st].posx); d = (uint32_t)dd; }
					}
				}
				if (hit) hit8(ix, os, list, nb, v, fi, d, offset, i);
			}
			// SNP strided scan (F13, :447-462), small mode: step s examines the entry of rank slo + 11 s = element cbase + s of the
			// residue-major filter column (the low 32 bits of its LO40: vgb_common.cuh).  16 bytes = four steps per load, two loads in
			// flight per lane: a group covers 32 steps per trip (a GRCh38-sized block has ~23), and the trip holds nothing but loads
			// and compares.  An entry that passes the filter is rare: the lane leaves the inner loop, reads the entry's full key from
			// `snp`, records the hit contexts if it really is a neighbour, and the scan resumes behind it.
			if (!k_big && k_sB) {
				// steps whose examined rank slo + 11 s lies past the end of the dictionary match nothing: cut them off
				const uint32_t n_scan = min(k_sB, (uint32_t)((ix.n_snp - k_slo + (SNP_STRIDE - 1)) / SNP_STRIDE));
				const uint64_t cbase = (uint64_t)(k_slo % SNP_STRIDE) * ix.snp_scan_stride + k_slo / SNP_STRIDE;
				const uint32_t mis = (uint32_t)cbase & 3u;    // position of step 0 inside its 16-byte quad
				const uint4 *colp = reinterpret_cast<const uint4 *>(ix.snp_scan + (cbase - mis));
				const uint32_t nquads = (n_scan + mis + 3u) >> 2;
				uint32_t sq = ol, sub = 0;                    // next quad of this lane, entry of the trip to resume at
				while (sq < nquads) {
					uint32_t found = 8, st_hit = 0;
					do {
						const uint4 A = ldr(colp + sq);
						uint4 B = make_uint4(0, 0, 0, 0);
						const bool hasB = sq + G < nquads;
						if (hasB) B = ldr(colp + sq + G);
						const uint32_t w[8] = { A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w };
#pragma unroll
						for (uint32_t e = 0; e < 8; e++) {
							if (SHORTSCAN && e == 4 && !hasB) break;
							const uint32_t st = 4u * (sq + (e >> 2) * G) + (e & 3u) - mis;   // scan step of this entry (below 0 wraps: fails the range test)
							if (found == 8 && e >= sub && st < n_scan && scan_candidate((uint32_t)km, w[e])) { found = e; st_hit = st; }
						}
						sub = found + 1;
						if (sub >= 8) { sq += 2 * G; sub = 0; }
					} while (found == 8 && sq < nquads);
					if (found < 8) {
						// the filter let this step through: the EXAMINED entry (rank slo + 11 step) decides, the REPORTED one is rank slo + step (F13)
						const uint64_t entry_lo = ldr(&ix.snp[(uint64_t)k_slo + (uint64_t)SNP_STRIDE * st_hit].key) & 0xFFFFFFFFFFull;
						const int dd = one_base_slot((km & 0xFFFFFFFFFFull) ^ entry_lo);
						if (dd >= 0) {
							const uint4 raw = ldr(reinterpret_cast<const uint4 *>(ix.snp + k_slo + st_hit));
							hit8(ix, os, 1, (km & 0xFFFFFF0000000000ull) | entry_lo, raw.z, (raw.y >> 8) & 0xFFFFu, (uint32_t)dd, offset, i);
						}
					}
				}
			}
		}
		__syncwarp();

		// ---- vote (src/qv.cc:132-178, order-independent form) ----
		// two contexts with the same X come from different k-mer positions iff their k-mer indices differ (kmer_pos = X + 32 i)
		uint32_t E = run ? os->ev_count : 0;
		if (run && ol == 0) os->st[S_EVENTS] = E;
		if (E > (uint32_t)EVN) { defer = true; E = 0; }       // too many contexts for shared memory: redo this pass in the next kernel
		const bool vrun = run && !defer;
		for (uint32_t e = ol; e < E; e += G) {
			GEvent *p = &os->ev[e];
			const uint32_t m = p->meta, X = p->X;
			bool vt = (m & 0xFF) == NO_MOD;
			if (!vt) {
				const uint32_t ki = (m >> 8) & 0xFF;
				for (uint32_t f = 0; f < E && !vt; f++) {
					const GEvent *q = &os->ev[f];
					vt = ((q->meta & 0xFF) == NO_MOD) && q->X == X && ((q->meta >> 8) & 0xFF) <= ki;
				}
			}
			p->meta = (m & ~(1u << 17)) | (vt ? (1u << 17) : 0u);
		}
		__syncwarp();
		uint32_t bf_ = 0, bxmin = 0xFFFFFFFFu, bxmax = 0;
		uint32_t t_nref = 0, t_nsnp = 0;
		uint64_t t_dg = 0;
		for (uint32_t e = ol; e < E; e += G) {
			const GEvent *p = &os->ev[e];
			const uint32_t m = p->meta, X = p->X, ki = (m >> 8) & 0xFF;
			if (TRACE) {
				if ((m >> 16) & 1) t_nsnp++; else t_nref++;
				t_dg += ctx_digest((m >> 16) & 1, X, X + 32u * ki, p->kmer, (m & 0xFF) == NO_MOD ? 10086u : (m & 0xFF));
			}
			if (!((m >> 17) & 1)) continue;
			uint32_t f = 0;
			bool distinct = false;
			for (uint32_t g = 0; g < E; g++) {
				const GEvent *q = &os->ev[g];
				if (((q->meta >> 17) & 1) && q->X == X) { f++; distinct |= (((q->meta >> 8) & 0xFF) != ki); }
			}
			if (!distinct) continue;
			if (f > bf_) { bf_ = f; bxmin = X; bxmax = X; }
			else if (f == bf_) { bxmin = min(bxmin, X); bxmax = max(bxmax, X); }
		}
		// group reductions (xor offsets below G stay inside the group)
		uint32_t maxf = bf_;
#pragma unroll
		for (int o = 1; o < G; o <<= 1) maxf = max(maxf, __shfl_xor_sync(FULL, maxf, o));
		uint32_t xmin = bf_ == maxf ? bxmin : 0xFFFFFFFFu, xmax = bf_ == maxf ? bxmax : 0u;
#pragma unroll
		for (int o = 1; o < G; o <<= 1) {
			xmin = min(xmin, __shfl_xor_sync(FULL, xmin, o));
			xmax = max(xmax, __shfl_xor_sync(FULL, xmax, o));
			if (TRACE) {
				t_nref += __shfl_xor_sync(FULL, t_nref, o);
				t_nsnp += __shfl_xor_sync(FULL, t_nsnp, o);
				t_dg += __shfl_xor_sync(FULL, t_dg, o);
			}
		}
		const bool has_best = maxf > 0;
		const bool ambiguous = has_best && xmin != xmax;
		const bool process = vrun && has_best && !ambiguous;   // freq > 1 is implied by two distinct k-mer positions (:1375)
		const uint32_t target = xmin;
		const bool retry = (vrun && !process && pass == 0) || (EVN == EV_WIDE && dretry);   // park it: one retry on the reverse complement (:1504-1510)

		// ---- park the reads that go to a retry round (warp-wide compaction over the group leaders) ----
		const uint32_t lowqm = OBALLOT(lowq);
		const uint32_t rmask = __ballot_sync(FULL, retry && ol == 0);
		if (retry) {
			Pend<G> *p = &pend[npend + __popc(rmask & ((1u << ob) - 1u))];
			p->kmer[ol] = kmer;
			if (ol == 0) { p->r = r; p->K = K; p->lowq = lowqm; }
		}
		npend += __popc(rmask);

		// ---- per-read bookkeeping (lane 0 of the group speaks for the read) ----
		if (have && ol == 0) {
			if (wide && G == 4) {
				a.kdefer[atomicAdd(&a.meta[9], 1u)] = r;          // 5..8 k-mers: the 8-lane instantiation takes it, from the start
			} else if (wide || defer) {
				a.defer[atomicAdd(&a.meta[a.defer_cnt], 1u)] = r | (pass << 31);
			} else {
				if (run) { atomicAdd(&acc[A_PASSES], 1u); os->st[S_EXACT] = 2u * K; os->st[S_BF] = 2u * os->st[S_LOWQ]; }
				if (!retry) {
					atomicAdd(&acc[A_READS], 1u);
					if (bad) { atomicAdd(&acc[A_BAD], 1u); atomicOr(&a.meta[5], 2u); }
					else if (skipped) atomicAdd(&acc[A_SKIPPED], 1u);
					else {
						if (process) atomicAdd(&acc[A_PLACED], 1u);
						if (maxf > 255) atomicAdd(&acc[A_WRAP], 1u);
					}
					if (TRACE) {
						vgb_read_result res;
						res.flags = (pass ? VGB_RF_REVCOMPL : 0) | (process ? VGB_RF_PROCESS : 0) | (ambiguous ? VGB_RF_AMBIGUOUS : 0) |
						            (has_best ? VGB_RF_HASBEST : 0);
						res.target = target; res.freq = (uint16_t)(maxf & 0xFF);
						res.n_ref = (uint16_t)t_nref; res.n_snp = (uint16_t)t_nsnp; res.passes = (uint16_t)(pass + 1); res.ctx_hash = t_dg;
						if (bad || skipped) { res.flags = VGB_RF_SKIPPED; res.target = 0; res.freq = 0; res.n_ref = 0; res.n_snp = 0; res.passes = 0; res.ctx_hash = 0; }
						a.trace[r] = res;
					}
				}
			}
		}

		// ---- pileup update: every recorded context at the winning position (src/qv.cc:1382-1502), one lane per context ----
		if (process) {
			const uint64_t n_blk = ix.pile_len >> 6;
			for (uint32_t e = ol; e < E; e += G) {
				const GEvent *p = &os->ev[e];
				if (p->X != target) continue;
				const uint32_t mod = p->meta & 0xFF;
				const uint64_t kmer_e = p->kmer;
				const uint64_t kpos = (uint64_t)(uint32_t)(target + 32u * ((p->meta >> 8) & 0xFF));
				// the 32-position window [kpos, kpos+32) lies in one or two 64-position blocks of the site bitmap
				const uint64_t bA = kpos >> 6, bB = (kpos + 31) >> 6;
				uint4 ka = make_uint4(0, 0, 0, 0), kb = make_uint4(0, 0, 0, 0);
				if (bA < n_blk) ka = ldr(reinterpret_cast<const uint4 *>(ix.pile + bA));
				if (bB != bA && bB < n_blk) kb = ldr(reinterpret_cast<const uint4 *>(ix.pile + bB));
				const uint64_t bitsA = ((uint64_t)ka.y << 32) | ka.x, bitsB = ((uint64_t)kb.y << 32) | kb.x;
				const uint32_t sh = (uint32_t)(kpos & 63);
				uint32_t win = (uint32_t)(bitsA >> sh);
				if (sh > 32) win |= (uint32_t)(bitsB << (64 - sh));
				if (mod < 32) win &= ~(1u << mod);                 // the substituted base does not count (:1391)
				while (win) {
					const uint32_t b = __ffs(win) - 1;
					win &= win - 1;
					const uint32_t off = sh + b;                    // bit index relative to block A
					const uint32_t sid = off < 64 ? ka.z + __popcll(bitsA & ((1ull << off) - 1))
					                              : kb.z + __popcll(bitsB & ((1ull << (off - 64)) - 1));
					const uint32_t code = ldr(ix.site_code + sid);
					const uint32_t rbase = code & 3, abase = code >> 2;
					if (rbase == abase) continue;                  // p->ref != p->alt (:1404)
					const uint32_t base = (uint32_t)(kmer_e >> (2 * b)) & 3u;
					if (base == rbase) { atomicAdd(ix.cnt + 2ull * sid, 1u); atomicAdd(&os->st[S_INCR], 1u); }
					else if (base == abase) { atomicAdd(ix.cnt + 2ull * sid + 1, 1u); atomicAdd(&os->st[S_INCR], 1u); }
				}
			}
		}
		__syncwarp();
		// commit the round's counters (S_* and A_* share the first 8 slots)
		if (have && !defer && !wide) {
			const uint32_t cv = os->st[ol];
			if (cv) atomicAdd(&acc[ol], cv);
			if (G == 4) { const uint32_t cw = os->st[ol + 4]; if (cw) atomicAdd(&acc[ol + 4], cw); }
		}
		__syncwarp();
	}
#undef OSHFL
#undef OBALLOT

	// ---- statistics: one set of atomics per CTA, and only for counters that moved ----
	__syncthreads();
	if (threadIdx.x < 14) {
		uint32_t *all = reinterpret_cast<uint32_t *>(smem_raw + sizeof(OS) * GW * R);
		unsigned long long v = 0;
#pragma unroll
		for (int w = 0; w < GW; w++) v += all[w * 16 + threadIdx.x];
		DevStats *s = a.stats;
		unsigned long long *dst = nullptr;
		switch (threadIdx.x) {
		case A_EXACT: dst = &s->exact_lookups; break;   case A_NBRQ: dst = &s->nbr_query_lookups; break;
		case A_SCAN: dst = &s->nbr_scan_reads; break;   case A_BF: dst = &s->bf_probes; break;
		case A_LOWQ: dst = &s->lowq_kmers; break;       case A_EVENTS: dst = &s->events; break;
		case A_INCR: dst = &s->pileup_incr; break;      case A_BIG: dst = &s->big_kmers; break;
		case A_READS: dst = &s->reads; break;           case A_SKIPPED: dst = &s->skipped_n; break;
		case A_PASSES: dst = &s->passes; break;         case A_PLACED: dst = &s->placed; break;
		case A_BAD: dst = &s->bad_records; break;       default: dst = &s->freq_wrap_reads; break;
		}
		if (v) atomicAdd(dst, v);
	}
}
