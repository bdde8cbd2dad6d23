// vgb_inflate.cuh -- raw DEFLATE (RFC 1951) decoder for BGZF blocks, one warp per block.
//
// Real FASTQ input is gzip; the reference reads plain text only (src/qv.cc:760-763 `fgets`) and is fed through `zcat`
// (experiment/experiment.md:20-27).  BGZF (the blocked gzip of bgzip / htslib) is a series of independent gzip members of at
// most 64 KiB each, so the members of a chunk inflate in parallel on the device and only the compressed bytes cross PCIe.
//
// Mapping: lane 0 of the warp walks the Huffman stream (it is inherently serial) and writes literals itself; a match
// (length, distance) is broadcast and copied by all 32 lanes.  The decoding tables live in the warp's shared memory:
// a 10-bit direct table for the literal/length alphabet, an 8-bit one for distances, canonical count / symbol arrays for the
// longer codes.  The same source compiles for the host (one "lane"), which is how tests/test_inflate_host.py checks it against
// zlib without a GPU.
#pragma once
#include <cstdint>

namespace vgb {

constexpr int INF_LIT_BITS = 10, INF_DST_BITS = 8;

struct InflateTables {
	uint16_t lit[1 << INF_LIT_BITS];   // (symbol << 4) | code length, 0 = longer than INF_LIT_BITS: canonical walk
	uint16_t dst[1 << INF_DST_BITS];
	uint16_t lcount[16], dcount[16];   // codes per length
	uint16_t lsym[288], dsym[32];      // symbols in canonical order
	uint16_t code[320];                // canonical code of every symbol (table construction)
	uint8_t lens[320];                 // code lengths: literal/length alphabet first, distance alphabet behind it
};

enum { INF_OK = 0, INF_E_INPUT = 1, INF_E_OUTPUT = 2, INF_E_CODE = 3, INF_E_DIST = 4, INF_E_BTYPE = 5, INF_E_STORED = 6 };

#ifdef __CUDACC__
#define INF_HD __device__ __forceinline__
#define INF_TABLE static __device__ const
#define INF_LANE() (threadIdx.x & 31u)
#define INF_LANES 32u
#define INF_SYNC() __syncwarp()
#define INF_BCAST(v) __shfl_sync(0xffffffffu, (v), 0)
// bytes another lane of this warp wrote to global memory a moment ago: ordered by the __syncwarp in front, read past L1
#define INF_LOAD_OUT(p) __ldcg(p)
#else
#define INF_HD inline
#define INF_TABLE static const
#define INF_LANE() 0u
#define INF_LANES 1u
#define INF_SYNC() ((void)0)
#define INF_BCAST(v) (v)
#define INF_LOAD_OUT(p) (*(p))
#endif

// length / distance bases and extra bits (RFC 1951 3.2.5), order of the code-length code lengths (3.2.7)
INF_TABLE uint16_t INF_LBASE[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
INF_TABLE uint8_t INF_LEXT[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
INF_TABLE uint16_t INF_DBASE[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
INF_TABLE uint8_t INF_DEXT[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
INF_TABLE uint8_t INF_ORDER[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };

struct InflateBits {                    // LSB-first bit reader over [in, in + n)
	const uint8_t *in;
	uint64_t n, ip;
	uint64_t bb;
	uint32_t bc;
	bool over;                          // ran past the end of the input
};

// 32 more bits whenever at most 32 are left: one unaligned word per refill (two aligned loads on the device) instead of a byte
// at a time -- the refill sits on the critical path of every symbol
INF_HD void inf_refill(InflateBits &b)
{
	if (b.bc > 32) return;
	uint32_t w = 0;
	if (b.ip + 4 <= b.n) {
#ifdef __CUDACC__
		const uintptr_t a = reinterpret_cast<uintptr_t>(b.in + b.ip);
		const uint32_t *wp = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
		w = __funnelshift_r(wp[0], wp[1], (uint32_t)(a & 3) * 8);      // the bytes in front of / behind the payload are header / trailer bytes of the same buffer
#else
		for (int i = 0; i < 4; i++) w |= (uint32_t)b.in[b.ip + i] << (8 * i);
#endif
	} else {
		for (int i = 0; i < 4; i++) if (b.ip + i < b.n) w |= (uint32_t)b.in[b.ip + i] << (8 * i);
		if (b.ip >= b.n + 8) b.over = true;                 // a few zero bytes behind the end are legal look-ahead, more is not
	}
	b.ip += 4;
	b.bb |= (uint64_t)w << b.bc;
	b.bc += 32;
}
INF_HD uint32_t inf_take(InflateBits &b, uint32_t k)        // k <= 32
{
	if (b.bc < k) inf_refill(b);
	const uint32_t v = (uint32_t)(b.bb & ((1ull << k) - 1ull));
	b.bb >>= k; b.bc -= k;
	return v;
}

// canonical walk (codes longer than the direct table): one bit at a time, as the format defines it
INF_HD int inf_slow(InflateBits &b, const uint16_t *count, const uint16_t *sym)
{
	int code = 0, first = 0, index = 0;
	for (int len = 1; len <= 15; len++) {
		code |= (int)inf_take(b, 1);
		const int c = count[len];
		if (code - c < first) return sym[index + (code - first)];
		index += c; first += c; first <<= 1; code <<= 1;
	}
	return -1;
}

INF_HD uint32_t inf_rev(uint32_t v, int bits)
{
	uint32_t r = 0;
	for (int i = 0; i < bits; i++) { r = (r << 1) | (v & 1u); v >>= 1; }
	return r;
}

// lens[0 .. n) -> count[], sym[], code[] (lane 0), then the direct table (all lanes).  Returns false for an over-subscribed set.
INF_HD bool inf_build(const uint8_t *lens, int n, uint16_t *count, uint16_t *sym, uint16_t *code, uint16_t *fast, int fast_bits)
{
	const uint32_t lane = INF_LANE();
	uint32_t ok = 1;
	if (lane == 0) {
		for (int i = 0; i < 16; i++) count[i] = 0;
		for (int i = 0; i < n; i++) count[lens[i]]++;
		count[0] = 0;
		int left = 1;
		for (int len = 1; len <= 15; len++) { left <<= 1; left -= count[len]; if (left < 0) { ok = 0; break; } }
		// RFC 1951 3.2.2: next_code[len] = (next_code[len - 1] + count[len - 1]) << 1; symbols of one length in symbol order
		uint16_t offs[16], next[16];
		uint32_t c = 0;
		for (int len = 1; len <= 15; len++) { c = (c + count[len - 1]) << 1; next[len] = (uint16_t)c; }
		offs[1] = 0;
		for (int len = 1; len < 15; len++) offs[len + 1] = offs[len] + count[len];
		for (int i = 0; i < n; i++) {
			const int l = lens[i];
			if (l) { sym[offs[l]++] = (uint16_t)i; code[i] = next[l]++; }
		}
		count[0] = 0;
	}
	INF_SYNC();
	ok = INF_BCAST(ok);
	if (!ok) return false;
	for (uint32_t i = lane; i < (1u << fast_bits); i += INF_LANES) fast[i] = 0;
	INF_SYNC();
	for (int i = (int)lane; i < n; i += (int)INF_LANES) {
		const int l = lens[i];
		if (l == 0 || l > fast_bits) continue;
		const uint32_t r = inf_rev(code[i], l);
		for (uint32_t k = r; k < (1u << fast_bits); k += 1u << l) fast[k] = (uint16_t)((i << 4) | l);
	}
	INF_SYNC();
	return true;
}

INF_HD int inf_decode_lit(InflateBits &b, const InflateTables &t)
{
	if (b.bc < 32) inf_refill(b);
	const uint16_t e = t.lit[b.bb & ((1u << INF_LIT_BITS) - 1)];
	if (e) { b.bb >>= (e & 15); b.bc -= (e & 15); return e >> 4; }
	return inf_slow(b, t.lcount, t.lsym);
}
INF_HD int inf_decode_dst(InflateBits &b, const InflateTables &t)
{
	if (b.bc < 32) inf_refill(b);
	const uint16_t e = t.dst[b.bb & ((1u << INF_DST_BITS) - 1)];
	if (e) { b.bb >>= (e & 15); b.bc -= (e & 15); return e >> 4; }
	return inf_slow(b, t.dcount, t.dsym);
}

// One DEFLATE stream [in, in + in_len) -> out[0 .. out_cap).  Executed by a whole warp (every lane calls it with the same
// arguments); returns the same (status, bytes written) in every lane.
INF_HD int inflate_block(const uint8_t *in, uint64_t in_len, uint8_t *out, uint32_t out_cap, InflateTables &t, uint32_t *out_len)
{
	const uint32_t lane = INF_LANE();
	InflateBits b;
	b.in = in; b.n = in_len; b.ip = 0; b.bb = 0; b.bc = 0; b.over = false;
	uint32_t pos = 0;
	int status = INF_OK;
	for (;;) {
		// ---- block header (lane 0 reads, everybody learns the type) ----
		uint32_t hdr = 0;
		if (lane == 0) hdr = inf_take(b, 3);
		hdr = INF_BCAST(hdr);
		const uint32_t last = hdr & 1u, type = hdr >> 1;
		if (type == 0) {
			// stored: skip to the byte boundary, LEN / NLEN, raw bytes (copied by all lanes)
			uint32_t len = 0, src = 0, bad = 0;
			if (lane == 0) {
				inf_take(b, b.bc & 7u);
				len = inf_take(b, 16);
				const uint32_t nlen = inf_take(b, 16);
				if ((len ^ 0xFFFFu) != nlen) bad = 1;
				// whole bytes still in the bit buffer belong to the raw data: hand them back
				b.ip -= b.bc >> 3; b.bb = 0; b.bc = 0;
				src = (uint32_t)b.ip;
				if (b.ip + len > b.n) bad = 1;
				b.ip += len;
			}
			len = INF_BCAST(len); src = INF_BCAST(src); bad = INF_BCAST(bad);
			if (bad) { status = INF_E_STORED; break; }
			if (pos + len > out_cap) { status = INF_E_OUTPUT; break; }
			for (uint32_t i = lane; i < len; i += INF_LANES) out[pos + i] = in[src + i];
			pos += len;
			INF_SYNC();
		} else if (type == 1 || type == 2) {
			// ---- code lengths ----
			uint32_t nlen = 288, ndist = 30, bad = 0;
			if (type == 1) {
				for (uint32_t i = lane; i < 288; i += INF_LANES) t.lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
				for (uint32_t i = lane; i < 30; i += INF_LANES) t.lens[288 + i] = 5;
				INF_SYNC();
			} else {
				if (lane == 0) {
					nlen = inf_take(b, 5) + 257; ndist = inf_take(b, 5) + 1;
					const uint32_t ncode = inf_take(b, 4) + 4;
					if (nlen > 286 || ndist > 30) bad = 1;
					uint8_t cl[19];
					for (int i = 0; i < 19; i++) cl[i] = 0;
					for (uint32_t i = 0; i < ncode; i++) cl[INF_ORDER[i]] = (uint8_t)inf_take(b, 3);
					// the code-length code: 19 symbols, decoded with the canonical walk (tables borrowed from the distance slots)
					uint16_t *cc = t.dcount, *cs = t.dsym;
					for (int i = 0; i < 16; i++) cc[i] = 0;
					for (int i = 0; i < 19; i++) cc[cl[i]]++;
					cc[0] = 0;
					uint16_t offs[16];
					offs[1] = 0;
					for (int len = 1; len < 15; len++) offs[len + 1] = offs[len] + cc[len];
					for (int i = 0; i < 19; i++) if (cl[i]) cs[offs[cl[i]]++] = (uint16_t)i;
					uint32_t idx = 0;
					while (!bad && idx < nlen + ndist) {
						const int sym = inf_slow(b, cc, cs);
						if (sym < 0) { bad = 1; break; }
						if (sym < 16) { t.lens[idx++] = (uint8_t)sym; continue; }
						uint32_t rep, val = 0;
						if (sym == 16) { if (idx == 0) { bad = 1; break; } val = t.lens[idx - 1]; rep = 3 + inf_take(b, 2); }
						else if (sym == 17) rep = 3 + inf_take(b, 3);
						else rep = 11 + inf_take(b, 7);
						if (idx + rep > nlen + ndist) { bad = 1; break; }
						while (rep--) t.lens[idx++] = (uint8_t)val;
					}
					if (!bad && t.lens[256] == 0) bad = 1;          // no end-of-block code
					// distance lengths behind the literal/length lengths, at the fixed slot the table builder expects
					if (!bad) {
						for (int i = (int)ndist - 1; i >= 0; i--) t.lens[288 + i] = t.lens[nlen + i];
						for (uint32_t i = nlen; i < 288; i++) t.lens[i] = 0;
						for (uint32_t i = ndist; i < 30; i++) t.lens[288 + i] = 0;
					}
				}
				INF_SYNC();
				bad = INF_BCAST(bad);
				if (bad) { status = INF_E_CODE; break; }
				nlen = 288; ndist = 30;
			}
			if (!inf_build(t.lens, (int)nlen, t.lcount, t.lsym, t.code, t.lit, INF_LIT_BITS) ||
			    !inf_build(t.lens + 288, (int)ndist, t.dcount, t.dsym, t.code + 288, t.dst, INF_DST_BITS)) { status = INF_E_CODE; break; }
			// ---- symbols: lane 0 decodes and writes literals until a match or the end of the block ----
			for (;;) {
				uint32_t mlen = 0, mdist = 0, err = 0, eob = 0;
				if (lane == 0) {
					for (;;) {
						const int sym = inf_decode_lit(b, t);
						if (sym < 0) { err = INF_E_CODE; break; }
						if (sym < 256) {
							if (pos >= out_cap) { err = INF_E_OUTPUT; break; }
							out[pos++] = (uint8_t)sym;
							continue;
						}
						if (sym == 256) { eob = 1; break; }
						if (sym > 285) { err = INF_E_CODE; break; }
						mlen = INF_LBASE[sym - 257] + inf_take(b, INF_LEXT[sym - 257]);
						const int ds = inf_decode_dst(b, t);
						if (ds < 0 || ds > 29) { err = INF_E_CODE; break; }
						mdist = INF_DBASE[ds] + inf_take(b, INF_DEXT[ds]);
						if (mdist > pos) err = INF_E_DIST;
						else if (pos + mlen > out_cap) err = INF_E_OUTPUT;
						break;
					}
					if (b.over) err = INF_E_INPUT;
				}
				INF_SYNC();
				pos = INF_BCAST(pos); mlen = INF_BCAST(mlen); mdist = INF_BCAST(mdist); err = INF_BCAST(err); eob = INF_BCAST(eob);
				if (err) { status = (int)err; break; }
				if (eob) break;
				// the match: every byte comes from the part of the output that is already complete (i mod distance)
				if (mdist >= mlen) { for (uint32_t i = lane; i < mlen; i += INF_LANES) out[pos + i] = INF_LOAD_OUT(out + pos - mdist + i); }
				else { for (uint32_t i = lane; i < mlen; i += INF_LANES) out[pos + i] = INF_LOAD_OUT(out + pos - mdist + (i % mdist)); }
				pos += mlen;
				INF_SYNC();
			}
			if (status != INF_OK) break;
		} else { status = INF_E_BTYPE; break; }
		if (last) break;
	}
	*out_len = pos;
	return status;
}

}  // namespace vgb
