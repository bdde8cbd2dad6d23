// vgb_build.cu -- GPU index builder: genome + accepted SNP lines -> the reference's on-disk index records, in HBM.
//
// SURVEY.md 8(f)-1 ("next" row): a fast builder of the SAME index `vargeno index` writes (src/dictgen.c:12-154,277-301 for
// the reference dictionary, :750-781 and :156-275 for the SNP dictionary, src/generate_bf.cc:107-154 and :247-262 for the
// two Bloom filters), needed to create chr22- and GRCh38-shaped indexes on a fresh box in seconds.  The records it
// produces are byte-identical to the reference's files (tests/test_gpu_index_build.py compares with the sha256 of files
// the compiled reference wrote); they stay in device memory and go straight into vgb_index_upload_device, or are copied
// back and written to disk by the Python tooling.
//
// Not on the hot path.  The sort is CUB's radix sort (library code, stable: equal k-mers keep ascending positions, which
// is what glibc's merge-sort qsort gives the reference); everything else is hand-written.
#include <cub/device/device_radix_sort.cuh>

#include <cstring>
#include <vector>

#include "vgb_internal.h"

namespace vgb {

constexpr uint64_t REF_BF_BITS = 1200000000ull * 8;   // src/generate_bf.h:201
constexpr uint64_t SNP_BF_BITS = 140000000ull * 8;    // src/generate_bf.h:203
constexpr uint64_t REF_LITE_BF_BITS = 2300000000ull * 8;   // src/generate_bf.h:202

__device__ __forceinline__ uint32_t base2(uint8_t c)   // A0 C1 G2 T3, anything else 4 (fasta_parser.c maps it to 'N')
{
	switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; }
}

// One thread = 32 consecutive k-mer starts of one contig [cs, ce): rolling encode (src/dictgen.c:12-51).
// pass 0: count valid (N-free) k-mers per thread; pass 1: write them at the scanned offsets.
template <int PASS>
__global__ void __launch_bounds__(256) k_build_kmers(const uint8_t *g, uint64_t cs, uint64_t ce, uint32_t *counts, uint64_t thread_base,
                                                      const uint32_t *offsets, uint64_t *keys, uint32_t *pos)
{
	const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t s0 = cs + t * 32;
	if (s0 + 32 > ce) return;                                  // no k-mer starts here (counts[] is pre-zeroed)
	uint64_t km = 0, nmask = 0;
	for (int j = 0; j < 63; j++) {
		const uint64_t p = s0 + j;
		const uint32_t c = p < ce ? base2(g[p]) : 4u;
		if (j < 32) km |= (uint64_t)(c & 3) << (2 * j);
		if (c == 4) nmask |= 1ull << j;
	}
	uint32_t n = 0, w = PASS ? offsets[thread_base + t] : 0;
	for (int j = 0; j < 32; j++) {
		const uint64_t s = s0 + j;
		if (s + 32 > ce) break;
		if (j > 0) {
			const uint64_t p = s + 31;
			km = (km >> 2) | ((uint64_t)(base2(g[p]) & 3) << 62);
		}
		if (((nmask >> j) & 0xFFFFFFFFull) == 0) {
			if (PASS) { keys[w] = km; pos[w] = (uint32_t)(s + 1); w++; }   // 1-based position in the concatenation (:289)
			n++;
		}
	}
	if (PASS == 0) counts[thread_base + t] = n;
}

__global__ void __launch_bounds__(256) k_heads(const uint64_t *keys, uint64_t n, uint32_t *head)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) k_group_starts(const uint32_t *head, const uint32_t *gidx, uint64_t n, uint32_t *gstart)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && head[i]) gstart[gidx[i]] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) k_small_flags(const uint32_t *gstart, uint64_t n_groups, uint32_t *small)
{
	const uint64_t gq = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (gq >= n_groups) return;
	const uint32_t sz = gstart[gq + 1] - gstart[gq];
	small[gq] = (sz >= 2 && sz <= (uint32_t)AUX_COLS) ? 1u : 0u;
}

__device__ __forceinline__ void st32u(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
__device__ __forceinline__ void st64u(uint8_t *p, uint64_t v) { st32u(p, (uint32_t)v); st32u(p + 4, (uint32_t)(v >> 32)); }

// write_kmers, src/dictgen.c:63-154: one 13-byte record per distinct k-mer, aux rows for 2..10 occurrences
__global__ void __launch_bounds__(256) k_write_ref(const uint64_t *keys, const uint32_t *pos, const uint32_t *gstart, const uint32_t *small,
                                                    const uint32_t *auxidx, uint64_t n_groups, uint8_t *rec, uint32_t *aux)
{
	const uint64_t gq = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (gq >= n_groups) return;
	const uint32_t s = gstart[gq], sz = gstart[gq + 1] - s;
	uint8_t *r = rec + 13 * gq;
	st64u(r, keys[s]);
	if (sz == 1) { st32u(r + 8, pos[s]); r[12] = 0; }
	else if (small[gq]) {
		const uint32_t a = auxidx[gq];
		st32u(r + 8, a); r[12] = 1;
		for (uint32_t c = 0; c < (uint32_t)AUX_COLS; c++) aux[(uint64_t)a * AUX_COLS + c] = c < sz ? pos[s + c] : 0u;
	} else { st32u(r + 8, POS_AMBIGUOUS); r[12] = 1; }
}

struct SnpLine { uint32_t pos0; uint8_t code, rf, af, pad; };   // global 0-based position, ref | alt << 2

// 32 alt-allele k-mers per accepted SNP line (src/dictgen.c:750-781): k-mer i starts at pos0 - 31 + i, ALT at offset 31 - i
__global__ void __launch_bounds__(256) k_snp_kmers(const uint8_t *g, const SnpLine *lines, uint64_t n_lines, uint64_t *keys, uint32_t *idx)
{
	const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= n_lines * 32) return;
	const SnpLine L = lines[q >> 5];
	const uint32_t i = (uint32_t)(q & 31);
	const uint64_t s = (uint64_t)L.pos0 - 31 + i;
	uint64_t km = 0;
	for (int j = 0; j < 32; j++) km |= (uint64_t)(base2(g[s + j]) & 3) << (2 * j);
	const uint32_t off = 31 - i;
	km = (km & ~(3ull << (2 * off))) | ((uint64_t)(L.code >> 2) << (2 * off));
	keys[q] = km;
	idx[q] = (uint32_t)q;
}

// write_snp_kmers, src/dictgen.c:156-275
__global__ void __launch_bounds__(256) k_write_snp(const uint64_t *keys, const uint32_t *idx, const SnpLine *lines, const uint32_t *gstart,
                                                    const uint32_t *small, const uint32_t *auxidx, uint64_t n_groups, uint8_t *rec, uint8_t *aux)
{
	const uint64_t gq = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (gq >= n_groups) return;
	const uint32_t s = gstart[gq], sz = gstart[gq + 1] - s;
	uint8_t *r = rec + 16 * gq;
	st64u(r, keys[s]);
	auto field = [&](uint32_t q, uint32_t &pos, uint8_t &info, uint8_t &rf, uint8_t &af) {
		const SnpLine L = lines[q >> 5];
		const uint32_t i = q & 31;
		pos = L.pos0 - 31 + i + 1;
		info = (uint8_t)(((31 - i) << 3) | (L.code & 3));
		rf = L.rf; af = L.af;
	};
	if (sz == 1) {
		uint32_t p; uint8_t info, rf, af;
		field(idx[s], p, info, rf, af);
		st32u(r + 8, p); r[12] = info; r[13] = 0; r[14] = rf; r[15] = af;
	} else if (small[gq]) {
		const uint32_t a = auxidx[gq];
		st32u(r + 8, a); r[12] = 0; r[13] = 1; r[14] = 0; r[15] = 0;
		uint8_t *ar = aux + 78ull * a;
		st64u(ar, keys[s]);
		for (uint32_t c = 0; c < (uint32_t)AUX_COLS; c++) {
			uint32_t p = 0; uint8_t info = 0, rf = 0, af = 0;
			if (c < sz) field(idx[s + c], p, info, rf, af);
			uint8_t *col = ar + 8 + 7 * c;
			st32u(col, p); col[4] = info; col[5] = rf; col[6] = af;
		}
	} else { st32u(r + 8, POS_AMBIGUOUS); r[12] = 0; r[13] = 1; r[14] = 0; r[15] = 0; }
}

// SNP Bloom filter from the UCSC table (constructBfFromUcsc, src/generate_bf.cc:538-556): per accepted record LO40 of the 32-mer that
// ends just before the SNP -- inserted even when it holds an N, as the value 0 that encode_kmer returns (src/util.c:103) -- and then
// LO40 of the 32 alternative-allele k-mers, up to the first N behind the SNP
__global__ void __launch_bounds__(256) k_bf_snp_ucsc(const uint8_t *g, const uint32_t *pos0, const uint8_t *alt, uint64_t n, uint32_t *words32)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint64_t p = pos0[i];
	uint64_t km = 0;
	bool had_n = false;
	for (int j = 0; j < 32; j++) {
		const uint32_t c = base2(g[p - 32 + j]);
		if (c == 4) { had_n = true; break; }
		km |= (uint64_t)c << (2 * j);
	}
	if (had_n) km = 0;
	uint64_t bit = hash40(km & 0xFFFFFFFFFFull) % SNP_BF_BITS;
	atomicOr(&words32[bit >> 5], 1u << (bit & 31));
	if (had_n) return;
	for (int j = 0; j < 32; j++) {
		const uint32_t c = j ? base2(g[p + j]) : (uint32_t)alt[i];
		if (c == 4) return;
		km = (km >> 2) | ((uint64_t)c << 62);
		bit = hash40(km & 0xFFFFFFFFFFull) % SNP_BF_BITS;
		atomicOr(&words32[bit >> 5], 1u << (bit & 31));
	}
}

int build_snp_bf_ucsc(vgb_ctx *c, const uint8_t *d_genome, const uint32_t *pos0, const uint8_t *alt, uint64_t n, uint64_t **d_words, uint64_t *bits, uint64_t *nwords)
{
	const uint64_t nw = (SNP_BF_BITS + 63) / 64;
	uint32_t *w = nullptr, *d_pos = nullptr; uint8_t *d_alt = nullptr;
	int rc;
	if ((rc = dev_alloc(c, &w, nw * 2, false))) return rc;
	if ((rc = dev_alloc(c, &d_pos, n, false)) || (rc = dev_alloc(c, &d_alt, n, false))) { cudaFree(w); cudaFree(d_pos); return rc; }
	cudaError_t e = cudaMemsetAsync(w, 0, nw * 8, c->stream);
	if (e == cudaSuccess && n) e = copy_sync(c, d_pos, pos0, n * 4, cudaMemcpyHostToDevice);
	if (e == cudaSuccess && n) e = copy_sync(c, d_alt, alt, n, cudaMemcpyHostToDevice);
	if (e == cudaSuccess && n) { k_bf_snp_ucsc<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_genome, d_pos, d_alt, n, w); c->launches++; }
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	if (e == cudaSuccess) e = cudaGetLastError();
	cudaFree(d_pos); cudaFree(d_alt);
	if (e != cudaSuccess) { cudaFree(w); return set_err(c, VGB_E_CUDA, "UCSC SNP filter build failed: %s", cudaGetErrorString(e)); }
	*d_words = reinterpret_cast<uint64_t *>(w); *bits = SNP_BF_BITS; *nwords = nw;
	return VGB_OK;
}

// <prefix>.ref.bf.lite.bf (src/generate_bf.cc:102-105,145-163): LO40 of every N-free 32-mer of every contig, value_range 40.
// Written by the reference's `index`, read by nothing; produced here so that `index` leaves the same set of files behind.
// Thread t owns the 32 k-mer starts cs + 32 t .. cs + 32 t + 31 (same walk as k_build_kmers).
__global__ void __launch_bounds__(256) k_bf_lite(const uint8_t *g, uint64_t cs, uint64_t ce, uint32_t *words32)
{
	const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t s0 = cs + t * 32;
	if (s0 + 32 > ce) return;
	uint64_t km = 0, nmask = 0;
	for (int j = 0; j < 63; j++) {
		const uint64_t p = s0 + j;
		const uint32_t c = p < ce ? base2(g[p]) : 4u;
		if (j < 32) km |= (uint64_t)(c & 3) << (2 * j);
		if (c == 4) nmask |= 1ull << j;
	}
	for (int j = 0; j < 32; j++) {
		const uint64_t s = s0 + j;
		if (s + 32 > ce) break;
		if (j > 0) km = (km >> 2) | ((uint64_t)(base2(g[s + 31]) & 3) << 62);
		if (((nmask >> j) & 0xFFFFFFFFull) == 0) {
			const uint64_t bit = hash40(km & 0xFFFFFFFFFFull) % REF_LITE_BF_BITS;
			atomicOr(&words32[bit >> 5], 1u << (bit & 31));
		}
	}
}

int build_ref_lite_bf(vgb_ctx *c, const uint8_t *d_genome, const uint64_t *cstart, const uint64_t *clen, uint32_t n_contigs,
                      uint64_t **d_words, uint64_t *bits, uint64_t *nwords)
{
	const uint64_t nw = (REF_LITE_BF_BITS + 63) / 64;
	uint32_t *w = nullptr;
	int rc;
	if ((rc = dev_alloc(c, &w, nw * 2, false))) return rc;
	VGB_CUDA(c, cudaMemsetAsync(w, 0, nw * 8, c->stream));
	for (uint32_t k = 0; k < n_contigs; k++) {
		if (clen[k] < 32) continue;
		const uint64_t threads = (clen[k] - 31 + 31) / 32;
		k_bf_lite<<<(unsigned)((threads + 255) / 256), 256, 0, c->stream>>>(d_genome, cstart[k], cstart[k] + clen[k], w);
		c->launches++;
	}
	VGB_CUDA(c, cudaStreamSynchronize(c->stream));
	VGB_CUDA(c, cudaGetLastError());
	*d_words = reinterpret_cast<uint64_t *>(w); *bits = REF_LITE_BF_BITS; *nwords = nw;
	return VGB_OK;
}

// reference Bloom filter: LO32 of every dictionary k-mer (src/generate_bf.cc:146-147); hash32 % 9.6e9 is the identity
__global__ void __launch_bounds__(256) k_bf_ref(const uint64_t *keys, uint64_t n, uint32_t *words32)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t bit = hash32((uint32_t)keys[i]);
	atomicOr(&words32[bit >> 5], 1u << (bit & 31));
}

// SNP Bloom filter, literally (src/generate_bf.cc:238-262): one value per accepted line, LO40 of the 32-mer that
// ends just before the SNP; lines whose 32 bases before the SNP contain an N are skipped (:241-242)
__global__ void __launch_bounds__(256) k_bf_snp(const uint8_t *g, const uint32_t *pos0, uint64_t n, uint32_t *words32)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint64_t s = (uint64_t)pos0[i] - 32;
	uint64_t km = 0;
	for (int j = 0; j < 32; j++) {
		const uint32_t c = base2(g[s + j]);
		if (c == 4) return;
		km |= (uint64_t)c << (2 * j);
	}
	const uint64_t bit = hash40(km & 0xFFFFFFFFFFull) % SNP_BF_BITS;
	atomicOr(&words32[bit >> 5], 1u << (bit & 31));
}

// sorted (key, payload) -> groups of equal keys: gstart[n_groups + 1], small flags, aux indices.  Returns counts.
static int group_sorted(vgb_ctx *c, const uint64_t *keys, uint64_t n, uint32_t **gstart_out, uint32_t **small_out, uint32_t **auxidx_out,
                        uint64_t *n_groups_out, uint64_t *n_aux_out)
{
	uint32_t *head = nullptr, *gidx = nullptr, *tmp = nullptr, *tot = nullptr;
	int rc;
	if ((rc = dev_alloc(c, &head, n, false)) || (rc = dev_alloc(c, &gidx, n, false)) || (rc = dev_alloc(c, &tmp, n / 2048 + 16, false)) ||
	    (rc = dev_alloc(c, &tot, 2, false)))
		return rc;
	const unsigned grid = (unsigned)((n + 255) / 256);
	if (n) k_heads<<<grid, 256, 0, c->stream>>>(keys, n, head);
	if ((rc = exclusive_scan_u32(c, head, gidx, n, tmp, tot))) return rc;
	uint32_t ng = 0;
	VGB_CUDA(c, cudaMemcpyAsync(&ng, tot, 4, cudaMemcpyDeviceToHost, c->stream));
	VGB_CUDA(c, cudaStreamSynchronize(c->stream));
	uint32_t *gstart = nullptr, *small = nullptr, *auxidx = nullptr;
	if ((rc = dev_alloc(c, &gstart, (uint64_t)ng + 1, false)) || (rc = dev_alloc(c, &small, (uint64_t)ng + 1, false)) ||
	    (rc = dev_alloc(c, &auxidx, (uint64_t)ng + 1, false)))
		return rc;
	if (n) k_group_starts<<<grid, 256, 0, c->stream>>>(head, gidx, n, gstart);
	const uint32_t n32 = (uint32_t)n;
	VGB_CUDA(c, cudaMemcpyAsync(gstart + ng, &n32, 4, cudaMemcpyHostToDevice, c->stream));
	VGB_CUDA(c, cudaStreamSynchronize(c->stream));
	if (ng) k_small_flags<<<(unsigned)((ng + 255) / 256), 256, 0, c->stream>>>(gstart, ng, small);
	if ((rc = exclusive_scan_u32(c, small, auxidx, ng, tmp, tot + 1))) return rc;
	uint32_t na = 0;
	VGB_CUDA(c, cudaMemcpyAsync(&na, tot + 1, 4, cudaMemcpyDeviceToHost, c->stream));
	VGB_CUDA(c, cudaStreamSynchronize(c->stream));
	cudaFree(head); cudaFree(gidx); cudaFree(tmp); cudaFree(tot);
	c->launches += 3;
	*gstart_out = gstart; *small_out = small; *auxidx_out = auxidx; *n_groups_out = ng; *n_aux_out = na;
	return VGB_OK;
}

template <typename V>
static int sort_pairs(vgb_ctx *c, uint64_t **keys, V **vals, uint64_t n)
{
	if (n == 0) return VGB_OK;
	uint64_t *k2 = nullptr; V *v2 = nullptr;
	int rc;
	if ((rc = dev_alloc(c, &k2, n, false)) || (rc = dev_alloc(c, &v2, n, false))) return rc;
	cub::DoubleBuffer<uint64_t> kb(*keys, k2);
	cub::DoubleBuffer<V> vb(*vals, v2);
	size_t tb = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, tb, kb, vb, n, 0, 64, c->stream);
	void *tmp = nullptr;
	VGB_CUDA(c, cudaMalloc(&tmp, tb ? tb : 1));
	cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tb, kb, vb, n, 0, 64, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	cudaFree(tmp);
	if (e != cudaSuccess) return set_err(c, VGB_E_CUDA, "radix sort failed: %s", cudaGetErrorString(e));
	uint64_t *kcur = kb.Current(); V *vcur = vb.Current();
	cudaFree(kcur == *keys ? k2 : *keys);
	cudaFree(vcur == *vals ? v2 : *vals);
	*keys = kcur; *vals = vcur;
	return VGB_OK;
}

int build_index_device(vgb_ctx *c, const uint8_t *d_genome, uint64_t genome_len, const uint64_t *cstart, const uint64_t *clen, uint32_t n_contigs,
                       const uint32_t *snp_pos0, const uint8_t *snp_code, const uint8_t *snp_rf, const uint8_t *snp_af, uint64_t n_snp_lines,
                       const uint32_t *bf_pos0, uint64_t n_bf_lines, vgb_index_view *out)
{
	if (genome_len >= 0xFFFFFFF0ull) return set_err(c, VGB_E_INDEX, "genome too long for 32-bit positions");
	memset(out, 0, sizeof(*out));
	int rc;
	// ---- reference dictionary ----
	uint64_t n_threads = 0;
	for (uint32_t k = 0; k < n_contigs; k++) {
		if (clen[k] < 32) return set_err(c, VGB_E_INDEX, "contig shorter than 32 bases (the reference asserts, src/dictgen.c:17)");
		n_threads += (clen[k] + 31) / 32;
	}
	uint32_t *counts = nullptr, *offs = nullptr, *tmp = nullptr, *tot = nullptr;
	if ((rc = dev_alloc(c, &counts, n_threads, false)) || (rc = dev_alloc(c, &offs, n_threads, false)) ||
	    (rc = dev_alloc(c, &tmp, n_threads / 2048 + 16, false)) || (rc = dev_alloc(c, &tot, 1, false)))
		return rc;
	VGB_CUDA(c, cudaMemsetAsync(counts, 0, n_threads * 4, c->stream));
	uint64_t tb = 0;
	for (uint32_t k = 0; k < n_contigs; k++) {
		const uint64_t nt = (clen[k] + 31) / 32;
		k_build_kmers<0><<<(unsigned)((nt + 255) / 256), 256, 0, c->stream>>>(d_genome, cstart[k], cstart[k] + clen[k], counts, tb, nullptr, nullptr, nullptr);
		tb += nt;
	}
	if ((rc = exclusive_scan_u32(c, counts, offs, n_threads, tmp, tot))) return rc;
	uint32_t n_k = 0;
	VGB_CUDA(c, cudaMemcpyAsync(&n_k, tot, 4, cudaMemcpyDeviceToHost, c->stream));
	VGB_CUDA(c, cudaStreamSynchronize(c->stream));
	uint64_t *keys = nullptr; uint32_t *pos = nullptr;
	if ((rc = dev_alloc(c, &keys, n_k, false)) || (rc = dev_alloc(c, &pos, n_k, false))) return rc;
	tb = 0;
	for (uint32_t k = 0; k < n_contigs; k++) {
		const uint64_t nt = (clen[k] + 31) / 32;
		k_build_kmers<1><<<(unsigned)((nt + 255) / 256), 256, 0, c->stream>>>(d_genome, cstart[k], cstart[k] + clen[k], counts, tb, offs, keys, pos);
		tb += nt;
	}
	c->launches += 2 * n_contigs;
	VGB_CUDA(c, cudaStreamSynchronize(c->stream));
	cudaFree(counts); cudaFree(offs); cudaFree(tmp); cudaFree(tot);
	if ((rc = sort_pairs<uint32_t>(c, &keys, &pos, n_k))) return rc;
	// reference Bloom filter from the sorted keys
	uint32_t *rbf = nullptr;
	if ((rc = dev_alloc(c, &rbf, 1ull << 27, false))) return rc;
	VGB_CUDA(c, cudaMemsetAsync(rbf, 0, (1ull << 27) * 4, c->stream));
	if (n_k) k_bf_ref<<<(unsigned)(((uint64_t)n_k + 255) / 256), 256, 0, c->stream>>>(keys, n_k, rbf);
	uint32_t *gstart, *small, *auxidx;
	uint64_t ng, na;
	if ((rc = group_sorted(c, keys, n_k, &gstart, &small, &auxidx, &ng, &na))) return rc;
	uint8_t *ref_rec = nullptr; uint32_t *ref_aux = nullptr;
	if ((rc = dev_alloc(c, &ref_rec, 13 * ng, false)) || (rc = dev_alloc(c, &ref_aux, na * AUX_COLS, false))) return rc;
	if (ng) k_write_ref<<<(unsigned)((ng + 255) / 256), 256, 0, c->stream>>>(keys, pos, gstart, small, auxidx, ng, ref_rec, ref_aux);
	VGB_CUDA(c, cudaStreamSynchronize(c->stream));
	cudaFree(keys); cudaFree(pos); cudaFree(gstart); cudaFree(small); cudaFree(auxidx);
	out->ref_records = ref_rec; out->n_ref = ng; out->ref_aux = ref_aux; out->n_ref_aux = na;
	out->ref_bf_words = reinterpret_cast<const uint64_t *>(rbf); out->ref_bf_bits = REF_BF_BITS; out->ref_bf_nwords = 1ull << 26;

	// ---- SNP dictionary ----
	SnpLine *lines = nullptr;
	{
		std::vector<SnpLine> h(n_snp_lines);
		for (uint64_t i = 0; i < n_snp_lines; i++) { h[i].pos0 = snp_pos0[i]; h[i].code = snp_code[i]; h[i].rf = snp_rf[i]; h[i].af = snp_af[i]; h[i].pad = 0; }
		if ((rc = dev_alloc(c, &lines, n_snp_lines, false))) return rc;
		if (n_snp_lines) VGB_CUDA(c, vgb::copy_sync(c, lines, h.data(), n_snp_lines * sizeof(SnpLine), cudaMemcpyHostToDevice));
	}
	const uint64_t n_sk = n_snp_lines * 32;
	if (n_sk >= 0xFFFFFFF0ull) return set_err(c, VGB_E_INDEX, "too many SNP k-mers");
	uint64_t *skeys = nullptr; uint32_t *sidx = nullptr;
	if ((rc = dev_alloc(c, &skeys, n_sk, false)) || (rc = dev_alloc(c, &sidx, n_sk, false))) return rc;
	if (n_sk) k_snp_kmers<<<(unsigned)((n_sk + 255) / 256), 256, 0, c->stream>>>(d_genome, lines, n_snp_lines, skeys, sidx);
	if ((rc = sort_pairs<uint32_t>(c, &skeys, &sidx, n_sk))) return rc;
	if ((rc = group_sorted(c, skeys, n_sk, &gstart, &small, &auxidx, &ng, &na))) return rc;
	uint8_t *snp_rec = nullptr, *snp_aux = nullptr;
	if ((rc = dev_alloc(c, &snp_rec, 16 * ng, false)) || (rc = dev_alloc(c, &snp_aux, 78 * na, false))) return rc;
	if (ng) k_write_snp<<<(unsigned)((ng + 255) / 256), 256, 0, c->stream>>>(skeys, sidx, lines, gstart, small, auxidx, ng, snp_rec, snp_aux);
	VGB_CUDA(c, cudaStreamSynchronize(c->stream));
	cudaFree(skeys); cudaFree(sidx); cudaFree(gstart); cudaFree(small); cudaFree(auxidx); cudaFree(lines);
	out->snp_records = snp_rec; out->n_snp = ng; out->snp_aux = snp_aux; out->n_snp_aux = na;

	// ---- SNP Bloom filter ----
	const uint64_t sw32 = (SNP_BF_BITS + 63) / 64 * 2;
	uint32_t *sbf = nullptr, *d_bfpos = nullptr;
	if ((rc = dev_alloc(c, &sbf, sw32, false)) || (rc = dev_alloc(c, &d_bfpos, n_bf_lines, false))) return rc;
	VGB_CUDA(c, cudaMemsetAsync(sbf, 0, sw32 * 4, c->stream));
	if (n_bf_lines) {
		VGB_CUDA(c, cudaMemcpyAsync(d_bfpos, bf_pos0, n_bf_lines * 4, cudaMemcpyHostToDevice, c->stream));
		k_bf_snp<<<(unsigned)((n_bf_lines + 255) / 256), 256, 0, c->stream>>>(d_genome, d_bfpos, n_bf_lines, sbf);
	}
	VGB_CUDA(c, cudaStreamSynchronize(c->stream));
	cudaFree(d_bfpos);
	out->snp_bf_words = reinterpret_cast<const uint64_t *>(sbf); out->snp_bf_bits = SNP_BF_BITS; out->snp_bf_nwords = sw32 / 2;
	c->launches += 6;
	VGB_CUDA(c, cudaGetLastError());
	return VGB_OK;
}

__global__ void __launch_bounds__(256) k_synth_genome(uint8_t *out, uint64_t cs, uint64_t len, uint64_t ci, uint64_t seed)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= len) return;
	// rnd64(seed, stream = 1, a = contig, b = i) of tools/synth.py
	uint64_t x = seed + 0x9E3779B97F4A7C15ull * 2;
	x ^= ci * 0xBF58476D1CE4E5B9ull;
	x += i * 0x94D049BB133111EBull;
	x = mix64(x);
	const char B[4] = { 'A', 'C', 'G', 'T' };
	out[cs + i] = (uint8_t)B[(x >> 33) & 3];
}

int synth_genome(vgb_ctx *c, uint8_t *d_out, const uint64_t *cstart, const uint64_t *clen, uint32_t n_contigs, uint64_t seed)
{
	for (uint32_t k = 0; k < n_contigs; k++)
		if (clen[k]) k_synth_genome<<<(unsigned)((clen[k] + 255) / 256), 256, 0, c->stream>>>(d_out, cstart[k], clen[k], k, seed);
	c->launches += n_contigs;
	VGB_CUDA(c, cudaStreamSynchronize(c->stream));
	VGB_CUDA(c, cudaGetLastError());
	return VGB_OK;
}

void free_index_device(vgb_index_view *v)
{
	cudaFree((void *)v->ref_records); cudaFree((void *)v->ref_aux); cudaFree((void *)v->snp_records); cudaFree((void *)v->snp_aux);
	cudaFree((void *)v->ref_bf_words); cudaFree((void *)v->snp_bf_words);
	memset(v, 0, sizeof(*v));
}

}  // namespace vgb
