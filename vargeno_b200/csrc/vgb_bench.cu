// vgb_bench.cu -- batch dictionary probes (parity + microbenchmark entry points) and measurement helpers.
#include <algorithm>

#include "vgb_internal.h"

namespace vgb {

// one thread per k-mer: query_ref_dict + query_snp_dict + check_block_size + both Bloom probes
__global__ void __launch_bounds__(256) k_lookup(const DevIndex ix, const uint64_t *kmers, uint64_t n, vgb_hit *out)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint64_t km = kmers[i];
	vgb_hit h;
	memset(&h, 0, sizeof(h));
	uint32_t lo, hi, posx = 0;
	ref_block(ix, km, lo, hi);
	h.ref_block_lo = lo; h.ref_block_n = hi - lo;
	if (lo < hi && ref_find_in_block(ix, (uint32_t)km, lo, hi, posx) >= 0) {
		h.ref_found = 1;
		if (posx == POS_AMBIGUOUS) { h.ref_pos = POS_AMBIGUOUS; h.ref_flag = 1; }   // > 10 copies (the only case the files hold)
		else if (posx < ix.amb_lo) { h.ref_pos = posx; h.ref_flag = 0; }
		else { h.ref_pos = 0xFFFFFFFEu - posx; h.ref_flag = 1; }
	}
	SnpEntry e;
	snp_block(ix, km, lo, hi);
	h.snp_block_lo = lo; h.snp_block_n = hi - lo;
	if (lo < hi && snp_find_in_block(ix, km & 0xFFFFFFFFFFull, lo, hi, e) >= 0) {
		h.snp_found = 1; h.snp_pos = e.pos; h.snp_flag = (uint8_t)snp_flag_of(e); h.snp_info = (uint8_t)snp_info_of(e);
	}
	h.ref_bf = bf_ref(ix, (uint32_t)km);
	h.snp_bf = bf_snp(ix, km & 0xFFFFFFFFFFull);
	out[i] = h;
}

int lookup_kmers(vgb_ctx *c, const uint64_t *kmers, uint64_t n, vgb_hit *out)
{
	if (!c->have_index) return set_err(c, VGB_E_ARG, "no index");
	if (n == 0) return VGB_OK;
	uint64_t *d_k = nullptr; vgb_hit *d_h = nullptr;
	int rc;
	if ((rc = dev_alloc(c, &d_k, n, false))) return rc;
	if ((rc = dev_alloc(c, &d_h, n, false))) { cudaFree(d_k); return rc; }
	cudaError_t e = cudaMemcpyAsync(d_k, kmers, n * 8, cudaMemcpyHostToDevice, c->stream);
	k_lookup<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->ix, d_k, n, d_h);
	c->launches++;
	if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_h, n * sizeof(vgb_hit), cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	cudaFree(d_k); cudaFree(d_h);
	if (e != cudaSuccess) return set_err(c, VGB_E_CUDA, "lookup kernel failed: %s", cudaGetErrorString(e));
	return VGB_OK;
}

// ---- probe microbenchmark (BASELINE config 5) ----
__global__ void __launch_bounds__(256) k_make_probe_kmers(const DevIndex ix, uint64_t *kmers, uint64_t n, int mode, uint64_t seed)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint64_t r = mix64(seed + 0x9E3779B97F4A7C15ull * (i + 1));
	const bool sample = mode == 1 || (mode == 2 && (i & 1));
	if (!sample || ix.n_ref == 0) { kmers[i] = r; return; }
	// sampled dictionary entry: its HI32 is the jumpgate slot h with jg[h] <= idx < jg[h+1]
	const uint32_t idx = (uint32_t)(r % ix.n_ref);
	uint64_t lo = 0, hi = 1ull << 32;                     // invariant: jg[lo] <= idx, jg[hi] > idx
	while (hi - lo > 1) { const uint64_t mid = (lo + hi) >> 1; if (ref_jg_at(ix, mid) <= idx) lo = mid; else hi = mid; }
	kmers[i] = (lo << 32) | ix.ref[idx].lo;
}

// Persistent grid, U k-mers per thread and trip, the probes of a trip issued level by level: U directory records, then the
// first entry of every block that can hold the k-mer (its fingerprint matches), then -- only for longer blocks -- the search.
// (One k-mer per thread and one tiny CTA per 256 k-mers spent more time launching CTAs than probing: 17 G fetches/s.)
constexpr int PROBE_U = 4;
__global__ void __launch_bounds__(256) k_probe(const DevIndex ix, const uint64_t *kmers, uint64_t n, unsigned long long *found)
{
	const uint64_t T = (uint64_t)gridDim.x * blockDim.x;
	uint32_t f = 0;
	for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; base < n; base += T * PROBE_U) {
		uint64_t km[PROBE_U];
		uint4 r[PROBE_U];
		bool on[PROBE_U];
#pragma unroll
		for (int j = 0; j < PROBE_U; j++) { on[j] = base + j * T < n; km[j] = on[j] ? __ldg(kmers + base + j * T) : 0; }
#pragma unroll
		for (int j = 0; j < PROBE_U; j++) r[j] = ldr(ix.xdir + (uint32_t)(km[j] >> 34));
		uint32_t rlo[PROBE_U], rhi[PROBE_U], flo[PROBE_U], fhi[PROBE_U];
		bool rq[PROBE_U], sq[PROBE_U];
		uint2 re[PROBE_U];
		uint4 se[PROBE_U];
#pragma unroll
		for (int j = 0; j < PROBE_U; j++) {
			bool rm, sm;
			dir_decode(ix, r[j], km[j], rlo[j], rhi[j], flo[j], fhi[j], rm, sm);
			rq[j] = on[j] && rm && rlo[j] < rhi[j];
			sq[j] = on[j] && sm && flo[j] < fhi[j];
			re[j] = make_uint2(0, 0); se[j] = make_uint4(0, 0, 0, 0);
			if (rq[j]) re[j] = ldr(reinterpret_cast<const uint2 *>(ix.ref + rlo[j]));
			if (sq[j]) se[j] = ldr(reinterpret_cast<const uint4 *>(ix.snp + flo[j]));
		}
#pragma unroll
		for (int j = 0; j < PROBE_U; j++) {
			if (rq[j]) {
				uint32_t posx;
				if (re[j].x == (uint32_t)km[j]) f++;
				else if (rhi[j] - rlo[j] > 1 && ref_find_in_block(ix, (uint32_t)km[j], rlo[j] + 1, rhi[j], posx) >= 0) f++;
			}
			if (sq[j]) {
				SnpEntry e;
				const uint64_t key = (((uint64_t)se[j].y << 32) | se[j].x) & 0xFFFFFFFFFFull;
				if (key == (km[j] & 0xFFFFFFFFFFull)) f++;
				else if (fhi[j] - flo[j] > 1 && snp_find_in_block(ix, km[j] & 0xFFFFFFFFFFull, flo[j] + 1, fhi[j], e) >= 0) f++;
			}
		}
	}
	f = __reduce_add_sync(0xffffffffu, f);
	if ((threadIdx.x & 31) == 0 && f) atomicAdd(found, (unsigned long long)f);
}

int probe_bench(vgb_ctx *c, uint64_t n, int mode, uint64_t seed, int repeats, double *ms, uint64_t *found)
{
	if (!c->have_index) return set_err(c, VGB_E_ARG, "no index");
	uint64_t *d_k = nullptr; unsigned long long *d_f = nullptr;
	int rc;
	if ((rc = dev_alloc(c, &d_k, n, false))) return rc;
	if ((rc = dev_alloc(c, &d_f, 1, false))) { cudaFree(d_k); return rc; }
	const unsigned mk_grid = (unsigned)((n + 255) / 256);
	int occ = 1;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_probe, 256, 0);
	const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)c->sm_count * (occ > 0 ? occ : 1), (n + 255) / 256);
	k_make_probe_kmers<<<mk_grid, 256, 0, c->stream>>>(c->ix, d_k, n, mode, seed);
	cudaMemsetAsync(d_f, 0, 8, c->stream);
	k_probe<<<grid, 256, 0, c->stream>>>(c->ix, d_k, n, d_f);       // warm-up
	cudaMemsetAsync(d_f, 0, 8, c->stream);
	cudaEventRecord(c->ev[0], c->stream);
	for (int r = 0; r < repeats; r++) k_probe<<<grid, 256, 0, c->stream>>>(c->ix, d_k, n, d_f);
	cudaEventRecord(c->ev[1], c->stream);
	c->launches += 2 + repeats;
	cudaError_t e = cudaStreamSynchronize(c->stream);
	float t = 0;
	cudaEventElapsedTime(&t, c->ev[0], c->ev[1]);
	unsigned long long f = 0;
	if (e == cudaSuccess) e = cudaMemcpy(&f, d_f, 8, cudaMemcpyDeviceToHost);
	cudaFree(d_k); cudaFree(d_f);
	if (e != cudaSuccess) return set_err(c, VGB_E_CUDA, "probe benchmark failed: %s", cudaGetErrorString(e));
	*ms = (double)t / std::max(1, repeats);
	*found = f / std::max(1, repeats);
	return VGB_OK;
}

// ---- HBM random-sector roofline denominator ----
// every thread issues `per_thread` independent 16-byte loads, each from a different uniformly random 32-byte sector
__global__ void __launch_bounds__(256) k_random_sectors(const uint4 *buf, uint64_t n_sectors, uint32_t per_thread, uint64_t seed, unsigned long long *sink)
{
	const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t acc = 0;
	uint64_t s = mix64(seed ^ (tid * 0x9E3779B97F4A7C15ull));
	for (uint32_t i = 0; i < per_thread; i += 8) {
		uint4 v[8];
#pragma unroll
		for (int k = 0; k < 8; k++) {
			s = mix64(s + k + 1);                      // same generator as tools/probes/sector_probe.cu (an LCG here measured ~25 % low)
			v[k] = __ldg(buf + 2 * (s % n_sectors));
		}
#pragma unroll
		for (int k = 0; k < 8; k++) acc += v[k].x ^ v[k].w;
	}
	if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

int random_sector_bench(vgb_ctx *c, uint64_t bytes, uint64_t n_loads, int repeats, double *gbs)
{
	uint4 *d_buf = nullptr; unsigned long long *d_sink = nullptr;
	int rc;
	if ((rc = dev_alloc(c, &d_buf, bytes / 16, false))) return rc;
	if ((rc = dev_alloc(c, &d_sink, 1, false))) { cudaFree(d_buf); return rc; }
	cudaMemsetAsync(d_buf, 1, bytes, c->stream);
	const uint32_t per_thread = 64;
	const uint64_t threads = (n_loads + per_thread - 1) / per_thread;
	const unsigned grid = (unsigned)((threads + 255) / 256);
	k_random_sectors<<<grid, 256, 0, c->stream>>>(d_buf, bytes / 32, per_thread, 1, d_sink);
	cudaEventRecord(c->ev[0], c->stream);
	for (int r = 0; r < repeats; r++) k_random_sectors<<<grid, 256, 0, c->stream>>>(d_buf, bytes / 32, per_thread, 2 + r, d_sink);
	cudaEventRecord(c->ev[1], c->stream);
	c->launches += 1 + repeats;
	cudaError_t e = cudaStreamSynchronize(c->stream);
	float t = 0;
	cudaEventElapsedTime(&t, c->ev[0], c->ev[1]);
	cudaFree(d_buf); cudaFree(d_sink);
	if (e != cudaSuccess) return set_err(c, VGB_E_CUDA, "random sector benchmark failed: %s", cudaGetErrorString(e));
	const double loads = (double)grid * 256.0 * per_thread * std::max(1, repeats);
	*gbs = loads * 32.0 / ((double)t * 1e-3) / 1e9;
	return VGB_OK;
}

// ---- synthetic reads on the device: twin of vargeno_b200/tools/synth.py simulate_reads (sub_rate / lowq only) ----
__device__ __forceinline__ uint64_t rnd64(uint64_t seed, uint64_t stream, uint64_t a, uint64_t b)
{
	uint64_t x = seed + 0x9E3779B97F4A7C15ull * (stream + 1);
	x ^= a * 0xBF58476D1CE4E5B9ull;
	x += b * 0x94D049BB133111EBull;
	return mix64(x);
}

struct SynthArgs {
	const uint8_t *hap0, *hap1;
	uint64_t genome_len;
	const uint64_t *cstart, *clen;
	uint32_t n_contigs;
	uint64_t n_reads;
	uint32_t L;
	uint64_t seed, first_id;
	uint32_t id_width;
	uint64_t sub_thr, lowq_thr;
	uint32_t lowq_chars;
	char *out;
};

__device__ __forceinline__ uint32_t base_code(uint8_t c)
{
	switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; }
}

__global__ void __launch_bounds__(256) k_synth_reads(const SynthArgs a)
{
	const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t lane = threadIdx.x & 31;
	if (warp >= a.n_reads) return;
	const uint64_t rid = a.first_id + warp;
	const uint32_t L = a.L;
	const uint64_t rec = 2 + a.id_width + 1 + L + 3 + L + 1;
	char *o = a.out + warp * rec;
	const uint64_t r0 = rnd64(a.seed, 30, rid, 0);
	uint64_t s = r0 % (a.genome_len - L + 1);
	uint32_t ci = 0;
	for (uint32_t k = 1; k < a.n_contigs; k++) if (a.cstart[k] <= s) ci = k;     // searchsorted(starts, s, 'right') - 1
	s = min(s, a.cstart[ci] + a.clen[ci] - L);
	const uint64_t r1 = rnd64(a.seed, 31, rid, 0);
	const uint8_t *hap = ((r1 >> 7) & 1) ? a.hap1 : a.hap0;
	const bool rev = (r1 >> 9) & 1;
	const char BASES[4] = { 'A', 'C', 'G', 'T' };
	if (lane == 0) { o[0] = '@'; o[1] = 'r'; }
	if (lane < a.id_width) {
		uint64_t p = 1;
		for (uint32_t d = 0; d < lane; d++) p *= 10;
		o[2 + a.id_width - 1 - lane] = (char)('0' + (rid / p) % 10);
	}
	char *seq = o + 2 + a.id_width + 1;
	char *qual = seq + L + 3;
	if (lane == 0) { seq[-1] = '\n'; seq[L] = '\n'; seq[L + 1] = '+'; seq[L + 2] = '\n'; qual[L] = '\n'; }
	for (uint32_t j = lane; j < L; j += 32) {
		const uint32_t col = rev ? (L - 1 - j) : j;            // column in genome orientation
		uint8_t b = hap[s + col];
		const uint64_t rb = rnd64(a.seed, 32, rid, col);
		if ((rb & 0xFFFFFFFFull) < a.sub_thr && b != 'N') b = (uint8_t)BASES[(base_code(b) + 1 + (uint32_t)((rb >> 40) % 3)) & 3];
		if (rev) { const uint32_t cd = base_code(b); b = cd < 4 ? (uint8_t)BASES[3 - cd] : (uint8_t)'N'; }
		seq[j] = (char)b;
		const uint64_t rq = rnd64(a.seed, 34, rid, j);
		const char hi_q = (char)('8' + (uint32_t)((rq >> 8) % 19));
		const char lo_q = (char)('#' + (uint32_t)((rq >> 16) % 21));
		const bool low = ((rq >> 32) < a.lowq_thr) && j < a.lowq_chars;
		qual[j] = low ? lo_q : hi_q;
	}
}

int synth_reads(vgb_ctx *c, const uint8_t *hap0, const uint8_t *hap1, uint64_t genome_len, const uint64_t *cstart,
                const uint64_t *clen, uint32_t n_contigs, uint64_t n_reads, uint32_t read_len, uint64_t seed, uint64_t first_id,
                uint32_t id_width, double sub_rate, double lowq_prob, uint32_t lowq_chars, char *out, uint64_t out_cap)
{
	const uint64_t rec = 2 + id_width + 1 + read_len + 3 + read_len + 1;
	if (n_reads * rec > out_cap) return set_err(c, VGB_E_ARG, "output buffer too small: need %llu bytes", (unsigned long long)(n_reads * rec));
	if (id_width > 19 || read_len == 0 || genome_len < read_len) return set_err(c, VGB_E_ARG, "bad synth parameters");
	uint64_t *d_cs = nullptr, *d_cl = nullptr;
	int rc;
	if ((rc = dev_alloc(c, &d_cs, n_contigs, false))) return rc;
	if ((rc = dev_alloc(c, &d_cl, n_contigs, false))) { cudaFree(d_cs); return rc; }
	cudaMemcpyAsync(d_cs, cstart, n_contigs * 8, cudaMemcpyHostToDevice, c->stream);
	cudaMemcpyAsync(d_cl, clen, n_contigs * 8, cudaMemcpyHostToDevice, c->stream);
	SynthArgs a;
	a.hap0 = hap0; a.hap1 = hap1; a.genome_len = genome_len; a.cstart = d_cs; a.clen = d_cl; a.n_contigs = n_contigs;
	a.n_reads = n_reads; a.L = read_len; a.seed = seed; a.first_id = first_id; a.id_width = id_width;
	a.sub_thr = (uint64_t)(sub_rate * 4294967296.0);
	a.lowq_thr = (uint64_t)(lowq_prob * 4294967296.0);
	a.lowq_chars = lowq_chars; a.out = out;
	const uint64_t threads = n_reads * 32;
	k_synth_reads<<<(unsigned)((threads + 255) / 256), 256, 0, c->stream>>>(a);
	c->launches++;
	cudaError_t e = cudaStreamSynchronize(c->stream);
	cudaFree(d_cs); cudaFree(d_cl);
	if (e != cudaSuccess) return set_err(c, VGB_E_CUDA, "synthetic read kernel failed: %s", cudaGetErrorString(e));
	return VGB_OK;
}

}  // namespace vgb
