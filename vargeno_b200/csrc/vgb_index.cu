// vgb_index.cu -- index upload and GPU re-layout.
//
// Replaces the load phase of the reference: dictionary + jumpgate construction (src/qv.cc:519-590, 606-695),
// static pileup initialisation (src/qv.cc:602-603, 637-659) and the Bloom filter loads (src/qv.cc:2140-2144).
// The raw on-disk records go over PCIe as they are (pinned staging, double buffered) and are converted on the
// device; the 2^32-entry reference jumpgate is produced by a scatter of block starts followed by a suffix-min fill
// instead of the reference's serial 16 GiB host loop.
#include <algorithm>
#include <cstring>
#include <vector>

#include "vgb_internal.h"

namespace vgb {

// --------------------------------------------------------------------------------------------------------
// generic exclusive scan (uint32), 2048 items per block
// --------------------------------------------------------------------------------------------------------
constexpr int SCAN_T = 256, SCAN_I = 8, SCAN_TILE = SCAN_T * SCAN_I;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total, uint32_t *sm /* >= 33 */)
{
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t inc = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
	if (lane == 31) sm[w] = inc;
	__syncthreads();
	if (w == 0) {
		uint32_t x = lane < (int)(blockDim.x >> 5) ? sm[lane] : 0, xi = x;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += t; }
		sm[lane] = xi - x;
		if (lane == 31) sm[32] = xi;
	}
	__syncthreads();
	const uint32_t r = sm[w] + inc - v;
	if (total) *total = sm[32];
	__syncthreads();
	return r;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_tile_sums(const uint32_t *in, uint64_t n, uint32_t *tsum)
{
	__shared__ uint32_t sm[33];
	const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
	uint32_t s = 0;
#pragma unroll
	for (int i = 0; i < SCAN_I; i++) { const uint64_t j = base + (uint64_t)i * SCAN_T + threadIdx.x; if (j < n) s += in[j]; }
	uint32_t tot;
	block_exclusive_scan(s, &tot, sm);
	if (threadIdx.x == 0) tsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) k_scan_tiles_serial(uint32_t *tsum, uint64_t nt, uint32_t *total)
{
	__shared__ uint32_t sm[33];
	uint32_t carry = 0;
	for (uint64_t b = 0; b < nt; b += 1024) {
		const uint64_t j = b + threadIdx.x;
		const uint32_t v = j < nt ? tsum[j] : 0;
		uint32_t tot;
		const uint32_t ex = block_exclusive_scan(v, &tot, sm);
		if (j < nt) tsum[j] = carry + ex;
		carry += tot;
	}
	if (threadIdx.x == 0) { tsum[nt] = carry; if (total) *total = carry; }
}

__global__ void __launch_bounds__(SCAN_T) k_scan_apply(const uint32_t *in, uint32_t *out, uint64_t n, const uint32_t *tsum)
{
	__shared__ uint32_t sm[33];
	const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_I;
	uint32_t v[SCAN_I], s = 0;
#pragma unroll
	for (int i = 0; i < SCAN_I; i++) { v[i] = (base + i < n) ? in[base + i] : 0; s += v[i]; }
	uint32_t run = block_exclusive_scan(s, nullptr, sm) + tsum[blockIdx.x];
#pragma unroll
	for (int i = 0; i < SCAN_I; i++) { if (base + i < n) out[base + i] = run; run += v[i]; }
}

int exclusive_scan_u32(vgb_ctx *c, const uint32_t *d_in, uint32_t *d_out, uint64_t n, uint32_t *d_tmp, uint32_t *d_total)
{
	const uint64_t nt = (n + SCAN_TILE - 1) / SCAN_TILE;
	if (nt == 0) { if (d_total) VGB_CUDA(c, cudaMemsetAsync(d_total, 0, 4, c->stream)); return VGB_OK; }
	k_scan_tile_sums<<<(unsigned)nt, SCAN_T, 0, c->stream>>>(d_in, n, d_tmp);
	k_scan_tiles_serial<<<1, 1024, 0, c->stream>>>(d_tmp, nt, d_total);
	k_scan_apply<<<(unsigned)nt, SCAN_T, 0, c->stream>>>(d_in, d_out, n, d_tmp);
	c->launches += 3;
	VGB_CUDA(c, cudaGetLastError());
	return VGB_OK;
}

// --------------------------------------------------------------------------------------------------------
// jumpgate: scatter of block starts (by the parse kernels) + suffix-min fill
// --------------------------------------------------------------------------------------------------------
constexpr int JG_T = 256, JG_I = 16, JG_TILE = JG_T * JG_I;

__global__ void __launch_bounds__(JG_T) k_jg_tile_min(const uint32_t *jg, uint32_t *tmin)
{
	__shared__ uint32_t sm[JG_T / 32];
	const uint4 *p = reinterpret_cast<const uint4 *>(jg + (uint64_t)blockIdx.x * JG_TILE);
	uint32_t m = 0xFFFFFFFFu;
#pragma unroll
	for (int i = 0; i < JG_I / 4; i++) { const uint4 v = p[i * JG_T + threadIdx.x]; m = min(m, min(min(v.x, v.y), min(v.z, v.w))); }
#pragma unroll
	for (int o = 16; o; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
	if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
	__syncthreads();
	if (threadIdx.x == 0) { for (int i = 1; i < JG_T / 32; i++) m = min(m, sm[i]); tmin[blockIdx.x] = m; }
}

// in place: tmin[t] <- min over tiles strictly after t (sentinel for the last)
__global__ void __launch_bounds__(1024) k_jg_suffix_tiles(uint32_t *tmin, int64_t nt, uint32_t sentinel)
{
	__shared__ uint32_t sm[1024];
	uint32_t carry = sentinel;
	for (int64_t hi = nt; hi > 0; hi -= 1024) {
		const int64_t j = hi - 1 - threadIdx.x;          // thread 0 owns the last tile of the chunk
		const uint32_t v = j >= 0 ? tmin[j] : 0xFFFFFFFFu;
		sm[threadIdx.x] = v;
		__syncthreads();
		// inclusive prefix-min over thread index (= suffix-min over tile index)
		for (int o = 1; o < 1024; o <<= 1) {
			const uint32_t t = threadIdx.x >= o ? sm[threadIdx.x - o] : 0xFFFFFFFFu;
			__syncthreads();
			sm[threadIdx.x] = min(sm[threadIdx.x], t);
			__syncthreads();
		}
		const uint32_t excl = threadIdx.x ? sm[threadIdx.x - 1] : 0xFFFFFFFFu;
		const uint32_t chunk_min = sm[1023];
		if (j >= 0) tmin[j] = min(carry, excl);
		__syncthreads();
		carry = min(carry, chunk_min);
	}
}

__global__ void __launch_bounds__(JG_T) k_jg_fill(uint32_t *jg, const uint32_t *tcarry)
{
	__shared__ uint32_t sm[JG_T];
	uint32_t *base = jg + (uint64_t)blockIdx.x * JG_TILE + (uint64_t)threadIdx.x * JG_I;
	uint32_t v[JG_I];
#pragma unroll
	for (int i = 0; i < JG_I / 4; i++) { const uint4 q = reinterpret_cast<const uint4 *>(base)[i]; v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w; }
	uint32_t m = 0xFFFFFFFFu;
#pragma unroll
	for (int i = 0; i < JG_I; i++) m = min(m, v[i]);
	sm[threadIdx.x] = m;
	__syncthreads();
	for (int o = 1; o < JG_T; o <<= 1) {   // inclusive suffix-min over thread index
		const uint32_t t = threadIdx.x + o < JG_T ? sm[threadIdx.x + o] : 0xFFFFFFFFu;
		__syncthreads();
		sm[threadIdx.x] = min(sm[threadIdx.x], t);
		__syncthreads();
	}
	uint32_t run = min(tcarry[blockIdx.x], threadIdx.x + 1 < JG_T ? sm[threadIdx.x + 1] : 0xFFFFFFFFu);
#pragma unroll
	for (int i = JG_I - 1; i >= 0; i--) { run = min(run, v[i]); v[i] = run; }
#pragma unroll
	for (int i = 0; i < JG_I / 4; i++) reinterpret_cast<uint4 *>(base)[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

// jg has 2^bits + 1 entries, all 0xFFFFFFFF except the scattered block starts
static int fill_jumpgate(vgb_ctx *c, uint32_t *jg, int bits, uint32_t n_entries, uint32_t *d_tmp)
{
	const uint64_t n = 1ull << bits;
	const int64_t nt = (int64_t)(n / JG_TILE);
	k_jg_tile_min<<<(unsigned)nt, JG_T, 0, c->stream>>>(jg, d_tmp);
	k_jg_suffix_tiles<<<1, 1024, 0, c->stream>>>(d_tmp, nt, n_entries);
	k_jg_fill<<<(unsigned)nt, JG_T, 0, c->stream>>>(jg, d_tmp);
	c->launches += 3;
	VGB_CUDA(c, cudaMemcpyAsync(jg + n, &n_entries, 4, cudaMemcpyHostToDevice, c->stream));
	VGB_CUDA(c, cudaStreamSynchronize(c->stream));   // n_entries is a stack variable
	return VGB_OK;
}

// --------------------------------------------------------------------------------------------------------
// record conversion
// --------------------------------------------------------------------------------------------------------
struct ParseOut { unsigned long long errors; uint32_t max_pos; uint32_t first_error_kind; };

__device__ __forceinline__ uint32_t ld32u(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

// raw: `count` 13-byte records starting at global rank `first`; prev_kmer = k-mer of rank first-1 (ignored if first == 0)
__global__ void __launch_bounds__(256) k_parse_ref(const uint8_t *raw, uint64_t first, uint64_t count, uint64_t prev_kmer,
                                                    RefEntry *out, uint32_t *jg, uint32_t *jg_lo, uint32_t amb_lo, uint32_t n_aux, ParseOut *po)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t my_max = 0;
	if (i < count) {
	const uint8_t *r = raw + 13 * i;
	const uint64_t kmer = (uint64_t)ld32u(r) | ((uint64_t)ld32u(r + 4) << 32);
	const uint32_t pos = ld32u(r + 8);
	const uint32_t flag = r[12];
	const uint64_t g = first + i;
	uint64_t pk = prev_kmer;
	if (i > 0) pk = (uint64_t)ld32u(r - 13) | ((uint64_t)ld32u(r - 9) << 32);
	uint32_t err = 0;
	if (g > 0 && pk >= kmer) err = 1;                               // not sorted / duplicate k-mer
	uint32_t posx;
	if (pos == POS_AMBIGUOUS) posx = POS_AMBIGUOUS;
	else if (flag == 0) { posx = pos; if (pos >= amb_lo) err = 2; else my_max = pos; }
	else if (flag == 1) { if (pos >= n_aux) { err = 3; posx = POS_AMBIGUOUS; } else posx = 0xFFFFFFFEu - pos; }
	else { err = 4; posx = POS_AMBIGUOUS; }                         // reference asserts (src/qv.cc:887-889)
	out[g] = RefEntry{ (uint32_t)kmer, posx };
	const uint32_t hi = (uint32_t)(kmer >> 32);
	if (g == 0 || (uint32_t)(pk >> 32) != hi) jg[hi] = (uint32_t)g;
	atomicAdd(&jg_lo[(uint32_t)kmer], 1u);                          // bucket sizes of the LO32-keyed view
	if (err) { atomicAdd(&po->errors, 1ull); atomicCAS(&po->first_error_kind, 0u, err); }
	}
	my_max = __reduce_max_sync(0xffffffffu, my_max);
	if ((threadIdx.x & 31) == 0 && my_max) atomicMax(&po->max_pos, my_max);
}

// LO32-keyed view: one thread per HI32 prefix walks its block and scatters {hi, posx} into the bucket of each entry's LO32.
// cursor[] holds the bucket starts on entry and the bucket ENDS on exit.
__global__ void __launch_bounds__(256) k_scatter_by_lo(const RefEntry *ref, const uint32_t *jg, uint32_t *cursor, RefEntry *by_lo)
{
	const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t lo = jg[h], hi = jg[h + 1];
	for (uint32_t i = lo; i < hi; i++) {
		const RefEntry e = ref[i];
		const uint32_t slot = atomicAdd(&cursor[e.lo], 1u);
		by_lo[slot] = RefEntry{ (uint32_t)h, e.posx };
	}
}

// LO40-keyed view of the SNP dictionary (vgb_common.cuh): one thread per HI24 block walks its entries; cursor[] holds the
// group starts on entry and the group ENDS on exit
__global__ void __launch_bounds__(256) k_count_snp_lo(const SnpEntry *snp, uint64_t n, uint32_t *count)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) atomicAdd(&count[(uint32_t)((snp[i].key & 0xFFFFFFFFFFull) >> 10)], 1u);
}
__global__ void __launch_bounds__(256) k_scatter_snp_lo(const SnpEntry *snp, const uint32_t *jg24, uint32_t *cursor, uint4 *by_lo)
{
	const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;      // 2^24 blocks
	const uint32_t lo = jg24[h], hi = jg24[h + 1];
	for (uint32_t i = lo; i < hi; i++) {
		const SnpEntry e = snp[i];
		const uint64_t lo40 = e.key & 0xFFFFFFFFFFull;
		const uint64_t kmer = (h << 40) | lo40;
		const uint32_t slot = atomicAdd(&cursor[(uint32_t)(lo40 >> 10)], 1u);
		by_lo[slot] = make_uint4((uint32_t)kmer, (uint32_t)(kmer >> 32), e.pos, (uint32_t)(e.key >> 40) & 0xFFFFu);
	}
}

// LO32 records (vgb_common.cuh, DevIndex::lo12): ends[] = bucket ends of the LO32-keyed view, bf = the reference Bloom filter as the
// file holds it (only its first 2^32 bits are addressable by hash32)
constexpr uint64_t LO12_GROUPS = ((1ull << 32) + 11) / 12;
constexpr uint32_t LO12_OVF_CAP = 1u << 20;
__global__ void __launch_bounds__(256) k_lo12(const uint32_t *ends, const uint32_t *bf, uint64_t bf_bits, uint64_t bf_nw32, uint4 *out,
                                               uint32_t *ovf, uint32_t *n_ovf)
{
	const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= LO12_GROUPS) return;
	const uint64_t l0 = 12 * g;
	const uint32_t start = l0 ? ends[l0 - 1] : 0u;
	uint32_t e[12], gates = 0, last = start;
	for (int j = 0; j < 12; j++) {
		const uint64_t l = l0 + j;
		if (l >> 32) { e[j] = last; continue; }                       // past the last LO32 value: empty buckets
		e[j] = last = ends[l];
		uint64_t bit = hash32((uint32_t)l);
		if (bf_bits <= 0xFFFFFFFFull) bit %= bf_bits;                 // 9.6e9 bits in practice: the modulo is the identity
		const uint64_t w = bit >> 5;
		if (w < bf_nw32 && ((bf[w] >> (bit & 31)) & 1u)) gates |= 1u << j;
	}
	uint32_t w[8] = { start, 0, 0, 0, 0, 0, 0, gates };
	if (e[11] - start > 0xFFFFu) {
		w[7] |= 0x80000000u;
		const uint32_t slot = atomicAdd(n_ovf, 1u);
		if (slot < LO12_OVF_CAP) { uint32_t *o = ovf + 14ull * slot; o[0] = (uint32_t)g; o[1] = start; for (int j = 0; j < 12; j++) o[2 + j] = e[j]; }
	} else {
		for (int j = 0; j < 12; j++) w[1 + (j >> 1)] |= (e[j] - start) << (16 * (j & 1));
	}
	out[2 * g] = make_uint4(w[0], w[1], w[2], w[3]);
	out[2 * g + 1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// sort a short device list of `n` records of `words` u32 by their first word (host side; the lists are rare-case tables)
static int sort_records_by_first_word(vgb_ctx *c, uint32_t *d_list, uint32_t n, uint32_t words)
{
	if (n < 2) return VGB_OK;
	std::vector<uint32_t> h((size_t)n * words), o(h.size()), order(n);
	VGB_CUDA(c, copy_sync(c, h.data(), d_list, h.size() * 4, cudaMemcpyDeviceToHost));
	for (uint32_t i = 0; i < n; i++) order[i] = i;
	std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return h[(size_t)a * words] < h[(size_t)b * words]; });
	for (uint32_t i = 0; i < n; i++) memcpy(&o[(size_t)i * words], &h[(size_t)order[i] * words], words * 4);
	VGB_CUDA(c, copy_sync(c, d_list, o.data(), o.size() * 4, cudaMemcpyHostToDevice));
	return VGB_OK;
}

__global__ void __launch_bounds__(256) k_max_u32(const uint32_t *a, uint64_t n, uint32_t *out)
{
	uint32_t m = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) m = max(m, a[i]);
#pragma unroll
	for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
	if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// raw: 16-byte records (aligned)
__global__ void __launch_bounds__(256) k_parse_snp(const uint4 *raw, uint64_t first, uint64_t count, uint64_t prev_kmer,
                                                    SnpEntry *out, uint32_t *jg, uint32_t *jg30, uint32_t n_aux, uint32_t *last_writer,
                                                    uint64_t pile_len, ParseOut *po)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	const uint4 r = raw[i];
	const uint64_t kmer = (uint64_t)r.x | ((uint64_t)r.y << 32);
	const uint32_t pos = r.z;
	const uint32_t info = r.w & 0xFF, flag = (r.w >> 8) & 0xFF, rf = (r.w >> 16) & 0xFF, af = (r.w >> 24) & 0xFF;
	const uint64_t g = first + i;
	uint64_t pk = prev_kmer;
	if (i > 0) { const uint4 q = raw[i - 1]; pk = (uint64_t)q.x | ((uint64_t)q.y << 32); }
	uint32_t err = 0;
	if (g > 0 && pk >= kmer) err = 1;
	if (flag > 1) err = 4;
	if (flag == 1 && pos != POS_AMBIGUOUS && pos >= n_aux) err = 3;
	const uint32_t ipos = info >> 3;
	const uint32_t alt = (uint32_t)(kmer >> (2 * ipos)) & 3u;       // kmer_get_base(kmer, snp_info_pos), src/qv.cc:656
	SnpEntry e;
	e.key = (kmer & 0xFFFFFFFFFFull) | ((uint64_t)info << 40) | ((uint64_t)flag << 48);
	e.pos = (flag == 1 && err == 3) ? POS_AMBIGUOUS : pos;
	e.extra = alt | (rf << 8) | (af << 16);
	out[g] = e;
	const uint32_t hi = (uint32_t)(kmer >> 40);
	if (g == 0 || (uint32_t)(pk >> 40) != hi) jg[hi] = (uint32_t)g;
	if (g == 0 || (pk >> 34) != (kmer >> 34)) jg30[kmer >> 34] = (uint32_t)g;
	// static pileup writer (src/qv.cc:637-659): unambiguous, reference base in ACGT; file order, last wins
	if ((info & 4) == 0 && pos != POS_AMBIGUOUS && flag == 0) {
		const uint64_t sp = (uint64_t)pos + ipos;
		if (sp >= pile_len) err = 5; else atomicMax(&last_writer[sp], (uint32_t)(g + 1));
	}
	if (err) { atomicAdd(&po->errors, 1ull); atomicCAS(&po->first_error_kind, 0u, err); }
}

// combined directory (vgb_common.cuh), reference half of record p: base, cumulative counts of the four HI32 blocks, their fp4
constexpr uint32_t XOVF_CAP = 4u << 20;   // overflow records per list (low-complexity prefixes); beyond that the upload fails loudly
__global__ void __launch_bounds__(256) k_xdir_ref(const uint32_t *jg, const RefEntry *ref, uint4 *x, uint32_t *ovf, uint32_t *n_ovf)
{
	const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= (1ull << 30)) return;
	const uint4 a = *reinterpret_cast<const uint4 *>(jg + 4 * p);
	const uint32_t j[5] = { a.x, a.y, a.z, a.w, jg[4 * p + 4] };
	uint4 r = make_uint4(j[0], 0, 0, 0);
	if (j[4] - j[0] > 254u) {
		r.z = 0xFFFFFFFFu;
		const uint32_t slot = atomicAdd(n_ovf, 1u);
		if (slot < XOVF_CAP) { uint32_t *o = ovf + 6ull * slot; o[0] = (uint32_t)p; for (int k = 0; k < 5; k++) o[1 + k] = j[k]; }
	} else {
		for (int k = 0; k < 4; k++) {
			r.z |= (j[k + 1] - j[0]) << (8 * k);
			if (j[k + 1] - j[k] == 1u) r.w |= fp4_ref(ref[j[k]].lo) << (16 + 4 * k);
		}
	}
	x[p] = r;
}
// SNP half of record p: base, count, fp8 of a single entry
__global__ void __launch_bounds__(256) k_xdir_snp(const uint32_t *sjg30, const SnpEntry *snp, uint4 *x, uint32_t *ovf, uint32_t *n_ovf)
{
	const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= (1ull << 30)) return;
	const uint32_t lo = sjg30[p], hi = sjg30[p + 1];
	uint4 r = x[p];
	r.y = lo;
	if (hi - lo > 254u) {
		r.w |= 0xFFu;
		const uint32_t slot = atomicAdd(n_ovf, 1u);
		if (slot < XOVF_CAP) { uint32_t *o = ovf + 3ull * slot; o[0] = (uint32_t)p; o[1] = lo; o[2] = hi; }
	} else {
		r.w |= hi - lo;
		// the entry keeps LO40 of its k-mer; bits 34..39 are the low bits of p: the fingerprint covers the low 34 bits, which is
		// what fp8_snp() takes from a full k-mer
		if (hi - lo == 1u) r.w |= fp8_snp(snp[lo].key & 0xFFFFFFFFFFull) << 8;
	}
	x[p] = r;
}

__global__ void __launch_bounds__(256) k_snp_scan_layout(const SnpEntry *snp, uint64_t n, uint64_t stride, uint32_t *scan)
{
	const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r < n) scan[(r % SNP_STRIDE) * stride + r / SNP_STRIDE] = (uint32_t)snp[r].key;     // low 32 bits of LO40: the scan's filter column
}

__global__ void __launch_bounds__(256) k_snp_aux(const uint8_t *raw78, uint64_t n, uint32_t *pos_out, uint8_t *info_out)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (row, col)
	if (i >= n * AUX_COLS) return;
	const uint8_t *p = raw78 + 78 * (i / AUX_COLS) + 8 + 7 * (i % AUX_COLS);
	pos_out[i] = ld32u(p);
	info_out[i] = p[4];
}

// one warp per 32 positions: site bitmap halves (no atomics)
__global__ void __launch_bounds__(256) k_site_bits(const uint32_t *last_writer, const SnpEntry *snp, uint64_t pile_len,
                                                    uint32_t *bits32, uint32_t *popc64)
{
	const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t code = 0;
	if (p < pile_len) {
		const uint32_t lw = last_writer[p];
		if (lw) { const SnpEntry e = snp[lw - 1]; code = ((uint32_t)(e.key >> 40) & 3u) | ((e.extra & 3u) << 2); }
	}
	const uint32_t m = __ballot_sync(0xffffffffu, code != 0);
	// pile_len is a multiple of 64, so a warp is entirely inside or entirely outside
	if ((threadIdx.x & 31) == 0 && p < pile_len) { bits32[p >> 5] = m; atomicAdd(&popc64[p >> 6], __popc(m)); }
}

__global__ void __launch_bounds__(256) k_site_blocks(const uint32_t *bits32, const uint32_t *rank, uint64_t n_blk, PileBlk *pile)
{
	const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n_blk) return;
	PileBlk k;
	k.bits = (uint64_t)bits32[2 * b] | ((uint64_t)bits32[2 * b + 1] << 32);
	k.rank = rank[b];
	k.pad = 0;
	pile[b] = k;
}

__global__ void __launch_bounds__(256) k_site_fill(const uint32_t *last_writer, const SnpEntry *snp, const PileBlk *pile, uint64_t pile_len,
                                                    uint32_t *site_pos, uint8_t *site_code, uint8_t *site_rf, uint8_t *site_af)
{
	const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= pile_len) return;
	const PileBlk k = pile[p >> 6];
	const uint32_t off = p & 63;
	if (!((k.bits >> off) & 1ull)) return;
	const uint32_t sid = k.rank + __popcll(k.bits & ((1ull << off) - 1));
	const SnpEntry e = snp[last_writer[p] - 1];
	site_pos[sid] = (uint32_t)p;
	site_code[sid] = (uint8_t)(((uint32_t)(e.key >> 40) & 3u) | ((e.extra & 3u) << 2));
	site_rf[sid] = (uint8_t)(e.extra >> 8);
	site_af[sid] = (uint8_t)(e.extra >> 16);
}

// --------------------------------------------------------------------------------------------------------
// host driver
// --------------------------------------------------------------------------------------------------------
static const char *parse_err_text(uint32_t k)
{
	switch (k) {
	case 1: return "records are not strictly sorted by k-mer";
	case 2: return "position collides with the ambiguous-entry encoding (genome + aux rows must stay below 2^32 - 2)";
	case 3: return "aux row index out of range";
	case 4: return "ambig_flag is neither 0 nor 1 (the reference asserts, src/qv.cc:887-889)";
	case 5: return "SNP position beyond the reference dictionary's last position + 32 (the reference reallocs, src/qv.cc:649-654)";
	default: return "unknown";
	}
}

// Streams `total` bytes from pageable/pinned host memory to the device in pieces and runs `launch` on each piece.
template <typename F>
static int stream_records(vgb_ctx *c, const uint8_t *src, uint64_t n_rec, uint32_t rec_bytes, F launch)
{
	if (c->upload_from_device) {
		// records already in HBM (vgb_build_index_device): convert in place, one launch, no staging
		if (n_rec) { launch(const_cast<uint8_t *>(src), 0, n_rec); c->launches++; }
		if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess)
			return set_err(c, VGB_E_CUDA, "index conversion kernel failed");
		return VGB_OK;
	}
	const uint64_t piece_rec = (64ull << 20) / rec_bytes * 16;             // ~1 GiB pieces
	uint8_t *h_pin[2] = { nullptr, nullptr }, *d_raw[2] = { nullptr, nullptr };
	cudaEvent_t done[2];
	const uint64_t piece_bytes = std::min(n_rec, piece_rec) * rec_bytes + 64;
	int rc = VGB_OK;
	for (int k = 0; k < 2; k++) {
		if (cudaMallocHost((void **)&h_pin[k], piece_bytes) != cudaSuccess || cudaMalloc((void **)&d_raw[k], piece_bytes) != cudaSuccess)
			rc = set_err(c, VGB_E_CUDA, "staging allocation of %llu bytes failed", (unsigned long long)piece_bytes);
		cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming);
	}
	int k = 0;
	for (uint64_t first = 0; rc == VGB_OK && first < n_rec; first += piece_rec, k ^= 1) {
		const uint64_t cnt = std::min(piece_rec, n_rec - first);
		cudaEventSynchronize(done[k]);                                     // staging pair k free again
		memcpy(h_pin[k], src + first * rec_bytes, cnt * rec_bytes);
		if (cudaMemcpyAsync(d_raw[k], h_pin[k], cnt * rec_bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
			rc = set_err(c, VGB_E_CUDA, "H2D copy of index records failed");
			break;
		}
		launch(d_raw[k], first, cnt);
		c->launches++;
		cudaEventRecord(done[k], c->stream);
	}
	cudaStreamSynchronize(c->stream);
	if (rc == VGB_OK && cudaGetLastError() != cudaSuccess) rc = set_err(c, VGB_E_CUDA, "index conversion kernel failed");
	for (int q = 0; q < 2; q++) { cudaFreeHost(h_pin[q]); cudaFree(d_raw[q]); cudaEventDestroy(done[q]); }
	return rc;
}

static uint64_t rd_kmer(const uint8_t *rec) { uint64_t k; memcpy(&k, rec, 8); return k; }

// release a device allocation that dev_alloc registered with the context
static void free_owned(vgb_ctx *c, void *p)
{
	for (int i = 0; i < c->n_owned; i++)
		if (c->owned[i] == p) { c->owned[i] = c->owned[--c->n_owned]; break; }
	cudaFree(p);
}

int index_upload(vgb_ctx *c, const vgb_index_view *v)
{
	if (c->have_index) return set_err(c, VGB_E_ARG, "an index is already resident in this context");
	if (!v || (!v->ref_records && v->n_ref) || (!v->snp_records && v->n_snp)) return set_err(c, VGB_E_ARG, "null index view");
	if (v->n_ref >= 0xFFFFFFFFull || v->n_snp >= 0xFFFFFFFFull)
		return set_err(c, VGB_E_INDEX, "dictionary too large (limit 2^32 32-mers, src/qv.cc:523,610)");
	if (v->n_ref_aux >= 0x7FFFFFFFull || v->n_snp_aux >= 0x7FFFFFFFull) return set_err(c, VGB_E_INDEX, "aux table too large");
	if (!v->ref_bf_bits || !v->snp_bf_bits) return set_err(c, VGB_E_INDEX, "empty Bloom filter");
	DevIndex &ix = c->ix;
	int rc;

	ParseOut *d_po = nullptr; uint32_t *d_tmp = nullptr;
	if ((rc = dev_alloc(c, &d_po, 1, false))) return rc;
	if ((rc = dev_alloc(c, &d_tmp, (1ull << 32) / SCAN_TILE + 8, false))) return rc;
	VGB_CUDA(c, cudaMemsetAsync(d_po, 0, sizeof(ParseOut), c->stream));

	// ---- reference dictionary ----
	RefEntry *d_ref = nullptr, *d_by_lo = nullptr; uint32_t *d_jg = nullptr, *d_jg_lo = nullptr; uint32_t *d_aux = nullptr;
	if ((rc = dev_alloc(c, &d_ref, v->n_ref))) return rc;
	if ((rc = dev_alloc(c, &d_by_lo, v->n_ref))) return rc;
	if ((rc = dev_alloc(c, &d_jg, (1ull << 32) + 1))) return rc;                 // temporary: folded into xdir below
	if ((rc = dev_alloc(c, &d_jg_lo, (1ull << 32) + 1))) return rc;
	VGB_CUDA(c, cudaMemsetAsync(d_jg_lo, 0, ((1ull << 32) + 1) * 4, c->stream));
	if ((rc = dev_alloc(c, &d_aux, v->n_ref_aux * AUX_COLS))) return rc;
	VGB_CUDA(c, cudaMemsetAsync(d_jg, 0xFF, ((1ull << 32) + 1) * 4, c->stream));
	if (v->n_ref_aux) VGB_CUDA(c, cudaMemcpyAsync(d_aux, v->ref_aux, v->n_ref_aux * AUX_COLS * 4, cudaMemcpyDefault, c->stream));
	const uint32_t amb_lo = 0xFFFFFFFFu - (uint32_t)v->n_ref_aux;
	rc = stream_records(c, v->ref_records, v->n_ref, 13, [&](uint8_t *d_raw, uint64_t first, uint64_t cnt) {
		const uint64_t prev = first ? rd_kmer(v->ref_records + 13 * (first - 1)) : 0;
		k_parse_ref<<<(unsigned)((cnt + 255) / 256), 256, 0, c->stream>>>(d_raw, first, cnt, prev, d_ref, d_jg, d_jg_lo, amb_lo, (uint32_t)v->n_ref_aux, d_po);
	});
	if (rc) return rc;
	if (v->n_ref_aux) { k_max_u32<<<296, 256, 0, c->stream>>>(d_aux, v->n_ref_aux * AUX_COLS, &d_po->max_pos); c->launches++; }
	if ((rc = fill_jumpgate(c, d_jg, 32, (uint32_t)v->n_ref, d_tmp))) return rc;
	// bucket sizes -> bucket starts (in place), then scatter; the cursor array ends up holding bucket ends
	if ((rc = exclusive_scan_u32(c, d_jg_lo, d_jg_lo, 1ull << 32, d_tmp, nullptr))) return rc;
	k_scatter_by_lo<<<(unsigned)((1ull << 32) / 256), 256, 0, c->stream>>>(d_ref, d_jg, d_jg_lo, d_by_lo);
	{
		// LO32 records: bucket bounds + reference Bloom gate per LO32 value; the 16 GiB end array and the filter are released
		// ref filter: hash32() < 2^32 <= 9.6e9 bits, so only the first 2^26 words are addressable
		uint64_t rw = v->ref_bf_nwords;
		if (v->ref_bf_bits > 0xFFFFFFFFull) rw = std::min<uint64_t>(rw, 1ull << 26);
		rw = std::min<uint64_t>(rw, (v->ref_bf_bits + 63) / 64);
		uint32_t *d_rbf = nullptr, *d_lovf = nullptr, *d_nl = nullptr; uint4 *d_lo12 = nullptr;
		if ((rc = dev_alloc(c, &d_rbf, rw * 2, false)) || (rc = dev_alloc(c, &d_lo12, 2 * LO12_GROUPS)) || (rc = dev_alloc(c, &d_lovf, 14ull * LO12_OVF_CAP)) ||
		    (rc = dev_alloc(c, &d_nl, 1, false))) return rc;
		if (rw) VGB_CUDA(c, cudaMemcpyAsync(d_rbf, v->ref_bf_words, rw * 8, cudaMemcpyDefault, c->stream));
		VGB_CUDA(c, cudaMemsetAsync(d_nl, 0, 4, c->stream));
		k_lo12<<<(unsigned)((LO12_GROUPS + 255) / 256), 256, 0, c->stream>>>(d_jg_lo, d_rbf, v->ref_bf_bits, rw * 2, d_lo12, d_lovf, d_nl);
		c->launches++;
		uint32_t nl = 0;
		VGB_CUDA(c, copy_sync(c, &nl, d_nl, 4, cudaMemcpyDeviceToHost));
		cudaFree(d_rbf); cudaFree(d_nl);
		free_owned(c, d_jg_lo);
		if (nl > LO12_OVF_CAP) return set_err(c, VGB_E_INDEX, "%u LO32 groups overflow their 16-bit counts (limit %u): sequence too repetitive for this layout", nl, LO12_OVF_CAP);
		if ((rc = sort_records_by_first_word(c, d_lovf, nl, 14))) return rc;
		ix.lo12 = d_lo12; ix.lo12_ovf = d_lovf; ix.n_lo12_ovf = nl;
	}
	// first half of the combined directory (vgb_common.cuh); the 16 GiB jumpgate itself is released before the SNP side allocates
	uint4 *d_xdir = nullptr; uint32_t *d_ovf_ref = nullptr, *d_ovf_snp = nullptr, *d_novf = nullptr;
	if ((rc = dev_alloc(c, &d_xdir, 1ull << 30))) return rc;
	if ((rc = dev_alloc(c, &d_ovf_ref, 6ull * XOVF_CAP)) || (rc = dev_alloc(c, &d_ovf_snp, 3ull * XOVF_CAP)) || (rc = dev_alloc(c, &d_novf, 2, false))) return rc;
	VGB_CUDA(c, cudaMemsetAsync(d_novf, 0, 8, c->stream));
	k_xdir_ref<<<(unsigned)((1ull << 30) / 256), 256, 0, c->stream>>>(d_jg, d_ref, d_xdir, d_ovf_ref, d_novf);
	c->launches += 2;
	VGB_CUDA(c, cudaStreamSynchronize(c->stream));
	free_owned(c, d_jg);
	ParseOut po;
	VGB_CUDA(c, vgb::copy_sync(c, &po, d_po, sizeof(po), cudaMemcpyDeviceToHost));
	if (po.errors) return set_err(c, VGB_E_INDEX, "reference dictionary: %llu bad records (%s)", po.errors, parse_err_text(po.first_error_kind));
	if (po.max_pos >= amb_lo) return set_err(c, VGB_E_INDEX, "reference dictionary: %s", parse_err_text(2));
	ix.ref = d_ref; ix.n_ref = v->n_ref; ix.ref_aux = d_aux; ix.n_ref_aux = (uint32_t)v->n_ref_aux; ix.amb_lo = amb_lo;
	ix.ref_by_lo = d_by_lo;

	// ---- SNP dictionary + static pileup ----
	// SNP k-mer starts are reference k-mer starts, so sites lie below max_pos + 32 (src/qv.cc:596-603)
	const uint64_t pile_len = (((uint64_t)po.max_pos + 32 + 1) + 63) / 64 * 64;
	SnpEntry *d_snp = nullptr; uint32_t *d_sjg, *d_sjg30, *d_sap, *d_lw; uint8_t *d_sai;
	if ((rc = dev_alloc(c, &d_snp, v->n_snp))) return rc;
	if ((rc = dev_alloc(c, &d_sjg, (1ull << 24) + 1))) return rc;
	if ((rc = dev_alloc(c, &d_sjg30, (1ull << 30) + 1))) return rc;             // temporary: folded into xdir below
	VGB_CUDA(c, cudaMemsetAsync(d_sjg30, 0xFF, ((1ull << 30) + 1) * 4, c->stream));
	if ((rc = dev_alloc(c, &d_sap, v->n_snp_aux * AUX_COLS))) return rc;
	if ((rc = dev_alloc(c, &d_sai, v->n_snp_aux * AUX_COLS))) return rc;
	if ((rc = dev_alloc(c, &d_lw, pile_len, false))) return rc;
	VGB_CUDA(c, cudaMemsetAsync(d_sjg, 0xFF, ((1ull << 24) + 1) * 4, c->stream));
	VGB_CUDA(c, cudaMemsetAsync(d_lw, 0, pile_len * 4, c->stream));
	VGB_CUDA(c, cudaMemsetAsync(d_po, 0, sizeof(ParseOut), c->stream));
	rc = stream_records(c, v->snp_records, v->n_snp, 16, [&](uint8_t *d_raw, uint64_t first, uint64_t cnt) {
		const uint64_t prev = first ? rd_kmer(v->snp_records + 16 * (first - 1)) : 0;
		k_parse_snp<<<(unsigned)((cnt + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<const uint4 *>(d_raw), first, cnt, prev, d_snp, d_sjg,
		                                                                  d_sjg30, (uint32_t)v->n_snp_aux, d_lw, pile_len, d_po);
	});
	if (rc) return rc;
	if (v->n_snp_aux) {
		rc = stream_records(c, v->snp_aux, v->n_snp_aux, 78, [&](uint8_t *d_raw, uint64_t first, uint64_t cnt) {
			k_snp_aux<<<(unsigned)((cnt * AUX_COLS + 255) / 256), 256, 0, c->stream>>>(d_raw, cnt, d_sap + first * AUX_COLS, d_sai + first * AUX_COLS);
		});
		if (rc) return rc;
	}
	if ((rc = fill_jumpgate(c, d_sjg, 24, (uint32_t)v->n_snp, d_tmp))) return rc;
	if ((rc = fill_jumpgate(c, d_sjg30, 30, (uint32_t)v->n_snp, d_tmp))) return rc;
	// second half of the combined directory; the separate array is not needed any more
	k_xdir_snp<<<(unsigned)((1ull << 30) / 256), 256, 0, c->stream>>>(d_sjg30, d_snp, d_xdir, d_ovf_snp, d_novf + 1);
	c->launches++;
	VGB_CUDA(c, cudaStreamSynchronize(c->stream));
	free_owned(c, d_sjg30);
	ix.xdir = d_xdir;
	{
		// the two overflow lists were appended in no particular order: sort them by p on the host (they are short)
		uint32_t novf[2];
		VGB_CUDA(c, copy_sync(c, novf, d_novf, 8, cudaMemcpyDeviceToHost));
		cudaFree(d_novf);
		if (novf[0] > XOVF_CAP || novf[1] > XOVF_CAP)
			return set_err(c, VGB_E_INDEX, "%u / %u directory records overflow their 8-bit counts (limit %u): sequence too repetitive for this layout", novf[0], novf[1], XOVF_CAP);
		if ((rc = sort_records_by_first_word(c, d_ovf_ref, novf[0], 6)) || (rc = sort_records_by_first_word(c, d_ovf_snp, novf[1], 3))) return rc;
		ix.xovf_ref = d_ovf_ref; ix.n_xovf_ref = novf[0];
		ix.xovf_snp = d_ovf_snp; ix.n_xovf_snp = novf[1];
	}
	{
		// LO40-keyed view for the upper-half SNP neighbours: group sizes -> starts (in place) -> scatter (leaves the ends)
		uint4 *d_sbl = nullptr; uint32_t *d_sdl = nullptr;
		if ((rc = dev_alloc(c, &d_sbl, v->n_snp))) return rc;
		if ((rc = dev_alloc(c, &d_sdl, (1ull << 30) + 1))) return rc;
		VGB_CUDA(c, cudaMemsetAsync(d_sdl, 0, ((1ull << 30) + 1) * 4, c->stream));
		if (v->n_snp) k_count_snp_lo<<<(unsigned)((v->n_snp + 255) / 256), 256, 0, c->stream>>>(d_snp, v->n_snp, d_sdl);
		if ((rc = exclusive_scan_u32(c, d_sdl, d_sdl, 1ull << 30, d_tmp, nullptr))) return rc;
		k_scatter_snp_lo<<<(unsigned)((1ull << 24) / 256), 256, 0, c->stream>>>(d_snp, d_sjg, d_sdl, d_sbl);
		c->launches += 2;
		ix.snp_by_lo = d_sbl; ix.snp_dir_lo = d_sdl;
	}
	{
		// residue-major filter column for the strided scan (read four entries per 16-byte load: a little slack behind the last column)
		const uint64_t stride = (v->n_snp + SNP_STRIDE - 1) / SNP_STRIDE + 1;
		uint32_t *d_scan = nullptr;
		if ((rc = dev_alloc(c, &d_scan, stride * SNP_STRIDE + 16))) return rc;
		VGB_CUDA(c, cudaMemsetAsync(d_scan, 0, (stride * SNP_STRIDE + 16) * 4, c->stream));
		if (v->n_snp) k_snp_scan_layout<<<(unsigned)((v->n_snp + 255) / 256), 256, 0, c->stream>>>(d_snp, v->n_snp, stride, d_scan);
		c->launches++;
		ix.snp_scan = d_scan; ix.snp_scan_stride = stride;
	}
	VGB_CUDA(c, vgb::copy_sync(c, &po, d_po, sizeof(po), cudaMemcpyDeviceToHost));
	if (po.errors) return set_err(c, VGB_E_INDEX, "SNP dictionary: %llu bad records (%s)", po.errors, parse_err_text(po.first_error_kind));
	ix.snp = d_snp; ix.n_snp = v->n_snp; ix.snp_jg = d_sjg; ix.snp_aux_pos = d_sap; ix.snp_aux_info = d_sai; ix.n_snp_aux = (uint32_t)v->n_snp_aux;

	// site bitmap + rank directory + compact per-site arrays
	const uint64_t n_blk = pile_len / 64;
	uint32_t *d_bits32 = nullptr, *d_popc = nullptr, *d_rank = nullptr, *d_total = nullptr; PileBlk *d_pile = nullptr;
	if ((rc = dev_alloc(c, &d_bits32, pile_len / 32, false))) return rc;
	if ((rc = dev_alloc(c, &d_popc, n_blk, false))) return rc;
	if ((rc = dev_alloc(c, &d_rank, n_blk, false))) return rc;
	if ((rc = dev_alloc(c, &d_total, 1, false))) return rc;
	if ((rc = dev_alloc(c, &d_pile, n_blk))) return rc;
	VGB_CUDA(c, cudaMemsetAsync(d_popc, 0, n_blk * 4, c->stream));
	k_site_bits<<<(unsigned)(pile_len / 256 + 1), 256, 0, c->stream>>>(d_lw, d_snp, pile_len, d_bits32, d_popc);
	c->launches++;
	if ((rc = exclusive_scan_u32(c, d_popc, d_rank, n_blk, d_tmp, d_total))) return rc;
	k_site_blocks<<<(unsigned)((n_blk + 255) / 256), 256, 0, c->stream>>>(d_bits32, d_rank, n_blk, d_pile);
	c->launches++;
	uint32_t n_sites = 0;
	VGB_CUDA(c, cudaMemcpyAsync(&n_sites, d_total, 4, cudaMemcpyDeviceToHost, c->stream));
	VGB_CUDA(c, cudaStreamSynchronize(c->stream));
	uint8_t *d_code = nullptr; uint32_t *d_cnt;
	if ((rc = dev_alloc(c, &c->d_site_pos, n_sites))) return rc;
	if ((rc = dev_alloc(c, &d_code, n_sites))) return rc;
	if ((rc = dev_alloc(c, &c->d_site_rf, n_sites))) return rc;
	if ((rc = dev_alloc(c, &c->d_site_af, n_sites))) return rc;
	if ((rc = dev_alloc(c, &d_cnt, 2ull * n_sites))) return rc;
	k_site_fill<<<(unsigned)(pile_len / 256 + 1), 256, 0, c->stream>>>(d_lw, d_snp, d_pile, pile_len, c->d_site_pos, d_code, c->d_site_rf, c->d_site_af);
	c->launches++;
	VGB_CUDA(c, cudaMemsetAsync(d_cnt, 0, 2ull * n_sites * 4, c->stream));
	ix.pile = d_pile; ix.pile_len = pile_len; ix.site_code = d_code; ix.n_sites = n_sites; ix.cnt = d_cnt;

	// ---- SNP Bloom filter (the reference one lives on as the gate bits of the LO32 records) ----
	const uint64_t sw = std::min<uint64_t>(v->snp_bf_nwords, (v->snp_bf_bits + 63) / 64);
	uint32_t *d_sbf;
	if ((rc = dev_alloc(c, &d_sbf, sw * 2))) return rc;
	if (sw) VGB_CUDA(c, cudaMemcpyAsync(d_sbf, v->snp_bf_words, sw * 8, cudaMemcpyDefault, c->stream));
	ix.snp_bf = d_sbf; ix.snp_bf_bits = v->snp_bf_bits; ix.snp_bf_nw32 = sw * 2;

	VGB_CUDA(c, cudaStreamSynchronize(c->stream));
	cudaFree(d_po); cudaFree(d_tmp); cudaFree(d_lw); cudaFree(d_bits32); cudaFree(d_popc); cudaFree(d_rank); cudaFree(d_total);
	VGB_CUDA(c, cudaGetLastError());
	c->have_index = true;
	return geno_prepare(c);
}

}  // namespace vgb
