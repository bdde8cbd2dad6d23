// vgb_fastq.cu -- K1: FASTQ record framing on the device.
//
// Replaces the four fgets() per record of the reference (src/qv.cc:760-763): the raw text chunk is already in
// HBM; three small kernels find every line start (newline count per 4 KiB tile -> exclusive scan -> scatter),
// so the per-read kernel can address record r as lines 4r .. 4r+3 without any host parsing.
// Streaming, 128-bit loads, bounded by HBM bandwidth (2 reads of the text + 4 B written per line).
#include "vgb_internal.h"

namespace vgb {

constexpr int FQ_T = 256;
constexpr int FQ_TILE = FQ_T * 16;   // bytes per block

// 16-bit mask of '\n' positions among the 16 bytes at text[off .. off+16)
__device__ __forceinline__ uint32_t nl_mask16(const char *text, uint64_t off, uint64_t n, bool aligned)
{
	uint32_t m = 0;
	if (aligned && off + 16 <= n) {
		const uint4 v = __ldg(reinterpret_cast<const uint4 *>(text + off));
		const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const uint32_t eq = __vcmpeq4(w[k], 0x0A0A0A0Au);   // 0xFF in each matching byte
			m |= ((eq & 1u) | ((eq >> 7) & 2u) | ((eq >> 14) & 4u) | ((eq >> 21) & 8u)) << (4 * k);
		}
	} else {
		for (int k = 0; k < 16; k++)
			if (off + k < n && text[off + k] == '\n') m |= 1u << k;
	}
	return m;
}

__global__ void __launch_bounds__(FQ_T) k_fq_count(const char *text, uint64_t n, uint32_t *blk_counts)
{
	__shared__ uint32_t sm[FQ_T / 32];
	const bool aligned = (reinterpret_cast<uintptr_t>(text) & 15) == 0;
	const uint64_t off = (uint64_t)blockIdx.x * FQ_TILE + (uint64_t)threadIdx.x * 16;
	uint32_t cnt = off < n ? __popc(nl_mask16(text, off, n, aligned)) : 0;
#pragma unroll
	for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
	if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = cnt;
	__syncthreads();
	if (threadIdx.x == 0) { uint32_t t = 0; for (int i = 0; i < FQ_T / 32; i++) t += sm[i]; blk_counts[blockIdx.x] = t; }
}

// meta: [0] n_lines [1] n_reads [2] work counter [3] format error bits
__global__ void k_fq_finish(const char *text, uint64_t n, const uint32_t *total_nl, uint32_t *meta, uint32_t *line_start, uint64_t line_cap)
{
	uint32_t lines = *total_nl;
	uint32_t err = 0;
	const bool open_tail = n > 0 && text[n - 1] != '\n';   // last line without '\n': treated as terminated
	if ((uint64_t)lines + 2 > line_cap) { err |= 4; lines = 0; }
	else {
		line_start[0] = 0;
		if (open_tail) { line_start[lines + 1] = (uint32_t)(n + 1); lines += 1; }
	}
	if (lines % 4) err |= 1;                               // truncated record (the reference would reuse stale buffers)
	meta[0] = lines; meta[1] = lines / 4; meta[2] = 0; meta[3] = err; meta[6] = 0; meta[7] = 0;
	meta[5] |= err;                                        // sticky until vgb_reset_counts (the slot is reused by later chunks)
}

__global__ void __launch_bounds__(FQ_T) k_fq_scatter(const char *text, uint64_t n, const uint32_t *blk_excl, const uint32_t *meta, uint32_t *line_start)
{
	__shared__ uint32_t sm[FQ_T / 32 + 1];
	if (meta[3] & 4) return;
	const bool aligned = (reinterpret_cast<uintptr_t>(text) & 15) == 0;
	const uint64_t off = (uint64_t)blockIdx.x * FQ_TILE + (uint64_t)threadIdx.x * 16;
	const uint32_t m = off < n ? nl_mask16(text, off, n, aligned) : 0;
	const uint32_t cnt = __popc(m);
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t inc = cnt;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
	if (lane == 31) sm[w] = inc;
	__syncthreads();
	if (threadIdx.x == 0) { uint32_t run = 0; for (int i = 0; i < FQ_T / 32; i++) { const uint32_t t = sm[i]; sm[i] = run; run += t; } }
	__syncthreads();
	uint32_t idx = blk_excl[blockIdx.x] + sm[w] + inc - cnt;   // global index of this thread's first newline
	uint32_t mm = m;
	while (mm) {
		const int b = __ffs(mm) - 1;
		mm &= mm - 1;
		line_start[++idx] = (uint32_t)(off + b + 1);           // line idx starts right after newline idx-1
	}
}

int fastq_index_lines(vgb_ctx *c, Chunk &ck, uint64_t nbytes)
{
	const uint64_t nblk = (nbytes + FQ_TILE - 1) / FQ_TILE;
	const uint64_t line_cap = c->max_chunk_bytes / 2 + 16;
	uint32_t *tmp = ck.d_blk_counts + nblk + 8;                // scratch for the scan's tile sums
	k_fq_count<<<(unsigned)nblk, FQ_T, 0, c->stream>>>(ck.d_text, nbytes, ck.d_blk_counts);
	c->launches++;
	int rc = exclusive_scan_u32(c, ck.d_blk_counts, ck.d_blk_counts, nblk, tmp, ck.d_meta + 4);
	if (rc) return rc;
	k_fq_finish<<<1, 1, 0, c->stream>>>(ck.d_text, nbytes, ck.d_meta + 4, ck.d_meta, ck.d_line_start, line_cap);
	k_fq_scatter<<<(unsigned)nblk, FQ_T, 0, c->stream>>>(ck.d_text, nbytes, ck.d_blk_counts, ck.d_meta, ck.d_line_start);
	c->launches += 2;
	VGB_CUDA(c, cudaGetLastError());
	return VGB_OK;
}

}  // namespace vgb
