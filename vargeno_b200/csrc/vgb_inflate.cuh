// vgb_inflate.cuh -- raw DEFLATE (RFC 1951) decoder for BGZF blocks, one warp per block.
//
// Real FASTQ input is gzip; the reference reads plain text only (src/qv.cc:760-763 `fgets`) and is fed through `zcat`
// (experiment/experiment.md:20-27).  BGZF (the blocked gzip of bgzip / htslib) is a series of independent gzip members of at
// most 64 KiB each, so the members of a chunk inflate in parallel on the device and only the compressed bytes cross PCIe.
//
// Mapping: lane 0 of the warp walks the Huffman stream (it is inherently serial).  What it produces goes into an 8 KiB ring in
// the warp's shared memory, never straight to HBM: literals are one shared-memory store, and a match -- every fifth byte of
// gzip'ed FASTQ starts one, almost all of them 3..12 bytes long and a few hundred bytes back -- is copied ring to ring by lane 0
// itself.  (The first version wrote literals to global memory and sent every match through an L2 round trip plus a warp
// broadcast: ~600 cycles per output byte.)  The warp is called in only for what lane 0 cannot do cheaply: flushing the ring
// to the text buffer with aligned 16-byte stores every 2 KiB, long matches, and matches that reach further back than the ring
// (read from the text that is already flushed).  The input is read one aligned word per 32 bits, loaded one refill ahead.
// Decoding tables live beside the ring: a 10-bit direct table for the literal/length alphabet, an 8-bit one for distances, a
// 7-bit one for the code-length code, canonical count / symbol arrays for the longer codes.  The same source compiles for the
// host (one "lane"), which is how tests/test_inflate_host.py checks it against zlib without a GPU.
#pragma once
#include <cstdint>
#include <cstring>

namespace vgb {

constexpr int INF_LIT_BITS = 10, INF_DST_BITS = 8, INF_CL_BITS = 7;
constexpr uint32_t INF_WIN = 8192;                    // ring size (power of two)
constexpr uint32_t INF_FLUSH = 2048;                  // lane 0 hands over when this much is waiting in the ring
constexpr uint32_t INF_REACH = INF_WIN - 320;         // a match this far back (or less) is still whole in the ring while it is copied
constexpr uint32_t INF_SOLO = 12;                     // matches up to this length are copied by lane 0 alone

struct InflateTables {
	uint16_t lit[1 << INF_LIT_BITS];   // (symbol << 4) | code length, 0 = longer than INF_LIT_BITS: canonical walk
	uint16_t dst[1 << INF_DST_BITS];   // doubles as the direct table of the code-length code while the lengths are read
	uint16_t lcount[16], dcount[16];   // codes per length
	uint16_t lsym[288], dsym[32];      // symbols in canonical order
	uint16_t code[320];                // canonical code of every symbol (table construction)
	uint8_t lens[320];                 // code lengths: literal/length alphabet first, distance alphabet behind it
};
struct alignas(16) InflateWarp {       // what one warp needs in shared memory
	uint8_t ring[INF_WIN];
	InflateTables t;
};

enum { INF_OK = 0, INF_E_INPUT = 1, INF_E_OUTPUT = 2, INF_E_CODE = 3, INF_E_DIST = 4, INF_E_BTYPE = 5, INF_E_STORED = 6 };

#ifdef __CUDACC__
#define INF_HD __device__ __forceinline__
#define INF_COLD __device__ __noinline__
#define INF_TABLE static __constant__
#define INF_LANE() (threadIdx.x & 31u)
#define INF_LANES 32u
#define INF_SYNC() __syncwarp()
#define INF_BCAST(v) __shfl_sync(0xffffffffu, (v), 0)
// bytes another lane of this warp wrote to global memory a moment ago: ordered by the __syncwarp in front, read past L1
#define INF_LOAD_OUT(p) __ldcg(p)
#define INF_COPY16(d, s) (*reinterpret_cast<uint4 *>(d) = *reinterpret_cast<const uint4 *>(s))
#else
#define INF_HD inline
#define INF_COLD inline
#define INF_TABLE static const
#define INF_LANE() 0u
#define INF_LANES 1u
#define INF_SYNC() ((void)0)
#define INF_BCAST(v) (v)
#define INF_LOAD_OUT(p) (*(p))
#define INF_COPY16(d, s) memcpy((d), (s), 16)
#endif

// length / distance bases and extra bits (RFC 1951 3.2.5), order of the code-length code lengths (3.2.7)
INF_TABLE uint16_t INF_LBASE[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
INF_TABLE uint8_t INF_LEXT[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
INF_TABLE uint16_t INF_DBASE[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
INF_TABLE uint8_t INF_DEXT[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
INF_TABLE uint8_t INF_ORDER[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };

struct InflateBits {                    // LSB-first bit reader over [in, in + n); used by lane 0 only
	const uint8_t *in;
	uint32_t n, ip;                     // ip: bytes of the stream that have been moved into bb (a BGZF member is < 64 KiB)
	uint64_t bb;
	uint32_t bc;
	bool over;                          // ran past the end of the input
#ifdef __CUDACC__
	// the stream is read as aligned words W[k]; 32 stream bits = funnel shift of two neighbours by the (constant) misalignment
	const uint32_t *wp;                 // address of `wnx`
	uint32_t wlo, wnx, sh;              // W[k], W[k + 1] (loaded one refill ahead of its use), 8 * (address & 3)
#endif
};

INF_HD void inf_start(InflateBits &b, uint32_t at)          // (re)start reading at byte `at` of the stream
{
	b.ip = at; b.bb = 0; b.bc = 0;
#ifdef __CUDACC__
	// words in front of / behind the payload belong to the same buffer (member header, trailer, the buffer's 64 spare bytes)
	const uintptr_t a = reinterpret_cast<uintptr_t>(b.in + at);
	b.wp = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
	b.sh = (uint32_t)(a & 3) * 8;
	b.wlo = b.wp[0];
	b.wp += 1;
	b.wnx = b.wp[0];
#endif
}

// 32 more bits (callers make sure that at most 32 are left)
INF_HD void inf_refill(InflateBits &b)
{
	uint32_t w = 0;
#ifdef __CUDACC__
	w = __funnelshift_r(b.wlo, b.wnx, b.sh);
	b.wlo = b.wnx;
	b.wp += 1;
	b.wnx = b.wp[0];                                        // needed at the next refill: its latency hides behind ~8 symbols
#else
	for (int i = 0; i < 4; i++) if (b.ip + i < b.n) w |= (uint32_t)b.in[b.ip + i] << (8 * i);
#endif
	if (b.ip >= b.n + 8) b.over = true;                     // a few bytes behind the end are legal look-ahead, more is not
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
INF_COLD int inf_slow_cold(InflateBits &b, const uint16_t *count, const uint16_t *sym)
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

// the out-of-line helpers work on a COPY of the reader: taking the address of the caller's own would pin it to local memory,
// and the per-symbol loop would store and reload its bit buffer around every symbol
INF_HD int inf_slow(InflateBits &b, const uint16_t *count, const uint16_t *sym)
{
	InflateBits c = b;
	const int r = inf_slow_cold(c, count, sym);
	b = c;
	return r;
}

INF_HD uint32_t inf_rev(uint32_t v, int bits)
{
#ifdef __CUDACC__
	return __brev(v) >> (32 - bits);
#else
	uint32_t r = 0;
	for (int i = 0; i < bits; i++) { r = (r << 1) | (v & 1u); v >>= 1; }
	return r;
#endif
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

// HLIT / HDIST / HCLEN and the code lengths of a dynamic block (RFC 1951 3.2.7), lane 0 only.  The 19-symbol code-length code
// (at most 7 bits) is decoded through a direct table borrowed from t.dst, which is built only afterwards.  Returns 0 or an error.
INF_COLD uint32_t inf_read_lengths_cold(InflateBits &b, InflateTables &t)
{
	const uint32_t nlen = inf_take(b, 5) + 257, ndist = inf_take(b, 5) + 1, ncode = inf_take(b, 4) + 4;
	if (nlen > 286 || ndist > 30) return 1;
	uint8_t cl[19];
	for (int i = 0; i < 19; i++) cl[i] = 0;
	for (uint32_t i = 0; i < ncode; i++) cl[INF_ORDER[i]] = (uint8_t)inf_take(b, 3);
	uint16_t cnt[8], next[8];
	for (int i = 0; i < 8; i++) cnt[i] = 0;
	for (int i = 0; i < 19; i++) cnt[cl[i]]++;
	cnt[0] = 0;
	int left = 1;
	for (int len = 1; len <= 7; len++) { left <<= 1; left -= cnt[len]; if (left < 0) return 1; }
	uint32_t c = 0;
	for (int len = 1; len <= 7; len++) { c = (c + cnt[len - 1]) << 1; next[len] = (uint16_t)c; }
	uint16_t *fast = t.dst;
	for (int i = 0; i < (1 << INF_CL_BITS); i++) fast[i] = 0;
	for (int i = 0; i < 19; i++) {
		const int l = cl[i];
		if (!l) continue;
		const uint32_t r = inf_rev(next[l]++, l);
		for (uint32_t k = r; k < (1u << INF_CL_BITS); k += 1u << l) fast[k] = (uint16_t)((i << 4) | l);
	}
	uint32_t idx = 0;
	while (idx < nlen + ndist) {
		if (b.bc < 32) inf_refill(b);
		const uint16_t e = fast[b.bb & ((1u << INF_CL_BITS) - 1)];
		if (!e) return 1;                                   // a bit pattern no code of an incomplete set stands for
		b.bb >>= (e & 15); b.bc -= (e & 15);
		const uint32_t sym = e >> 4;
		if (sym < 16) { t.lens[idx++] = (uint8_t)sym; continue; }
		uint32_t rep, val = 0;
		if (sym == 16) { if (idx == 0) return 1; val = t.lens[idx - 1]; rep = 3 + inf_take(b, 2); }
		else if (sym == 17) rep = 3 + inf_take(b, 3);
		else rep = 11 + inf_take(b, 7);
		if (idx + rep > nlen + ndist) return 1;
		while (rep--) t.lens[idx++] = (uint8_t)val;
	}
	if (t.lens[256] == 0) return 1;                         // no end-of-block code
	// distance lengths behind the literal/length lengths, at the fixed slot the table builder expects
	for (int i = (int)ndist - 1; i >= 0; i--) t.lens[288 + i] = t.lens[nlen + i];
	for (uint32_t i = nlen; i < 288; i++) t.lens[i] = 0;
	for (uint32_t i = ndist; i < 30; i++) t.lens[288 + i] = 0;
	return 0;
}

INF_HD uint32_t inf_read_lengths(InflateBits &b, InflateTables &t)
{
	InflateBits c = b;
	const uint32_t r = inf_read_lengths_cold(c, t);
	b = c;
	return r;
}

// Ring -> text buffer, bytes [from, to) of the output, by the whole warp.  Byte x of the output sits at ring[(x + a0) % INF_WIN]
// with a0 = the text buffer's address mod 16, so a 16-byte aligned piece of the text is a 16-byte aligned piece of the ring.
// Without `all` the copy stops at the last 16-byte boundary of the text; returns the new flushed mark (the same in every lane).
INF_HD uint32_t inf_flush(const uint8_t *ring, uint8_t *out, uint32_t a0, uint32_t from, uint32_t to, bool all)
{
	const uint32_t lane = INF_LANE();
	const uint32_t M = INF_WIN - 1;
	uint32_t f = from;
	const uint32_t head = (16u - ((f + a0) & 15u)) & 15u;
	const uint32_t nh = head < to - f ? head : to - f;
	for (uint32_t i = lane; i < nh; i += INF_LANES) out[f + i] = ring[(f + i + a0) & M];
	f += nh;
	const uint32_t nvec = (to - f) >> 4;
	for (uint32_t v = lane; v < nvec; v += INF_LANES) INF_COPY16(out + f + 16u * v, ring + ((f + 16u * v + a0) & M));
	f += 16u * nvec;
	if (all) {
		for (uint32_t i = lane; i < to - f; i += INF_LANES) out[f + i] = ring[(f + i + a0) & M];
		f = to;
	}
	INF_SYNC();
	return f;
}

// One DEFLATE stream [in, in + in_len) -> out[0 .. out_cap).  Executed by a whole warp (every lane calls it with the same
// arguments); returns the same (status, bytes written) in every lane.
INF_HD int inflate_block(const uint8_t *in, uint64_t in_len, uint8_t *out, uint32_t out_cap, InflateWarp &ws, uint32_t *out_len)
{
	enum { EX_EOB = 1, EX_MATCH = 2, EX_FLUSH = 3, EX_ERR = 4 };
	const uint32_t lane = INF_LANE();
	const uint32_t M = INF_WIN - 1;
	InflateTables &t = ws.t;
	uint8_t *ring = ws.ring;
	const uint32_t a0 = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15u);
	InflateBits b;
	b.in = in; b.n = (uint32_t)(in_len < 0xFFFFFF00ull ? in_len : 0xFFFFFF00ull); b.over = false;
	inf_start(b, 0);
	uint32_t pos = 0, flushed = 0;     // bytes produced / bytes of them that are in the text buffer; the same in every lane
	int status = INF_OK;
	for (;;) {
		// ---- block header (lane 0 reads, everybody learns the type) ----
		uint32_t hdr = 0;
		if (lane == 0) hdr = inf_take(b, 3);
		hdr = INF_BCAST(hdr);
		const uint32_t last = hdr & 1u, type = hdr >> 1;
		if (type == 0) {
			// stored: skip to the byte boundary, LEN / NLEN, raw bytes (copied by all lanes, straight into the text buffer; the
			// ring keeps the end of them for the matches of the blocks behind)
			uint32_t len = 0, src = 0, bad = 0;
			if (lane == 0) {
				inf_take(b, b.bc & 7u);
				len = inf_take(b, 16);
				const uint32_t nlen = inf_take(b, 16);
				if ((len ^ 0xFFFFu) != nlen) bad = 1;
				const uint32_t at = b.ip - (b.bc >> 3);           // whole bytes still in the bit buffer belong to the raw data
				src = at;
				if ((uint64_t)at + len > b.n) bad = 1;
				else inf_start(b, at + len);
			}
			len = INF_BCAST(len); src = INF_BCAST(src); bad = INF_BCAST(bad);
			if (bad) { status = INF_E_STORED; break; }
			if (pos + len > out_cap) { status = INF_E_OUTPUT; break; }
			flushed = inf_flush(ring, out, a0, flushed, pos, true);
			for (uint32_t i = lane; i < len; i += INF_LANES) {
				const uint8_t v = in[src + i];
				out[pos + i] = v;
				if (len - i <= INF_WIN) ring[(pos + i + a0) & M] = v;
			}
			pos += len;
			flushed = pos;
			INF_SYNC();
		} else if (type == 1 || type == 2) {
			// ---- code lengths ----
			uint32_t bad = 0;
			if (type == 1) {
				for (uint32_t i = lane; i < 288; i += INF_LANES) t.lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
				for (uint32_t i = lane; i < 30; i += INF_LANES) t.lens[288 + i] = 5;
				INF_SYNC();
			} else {
				if (lane == 0) bad = inf_read_lengths(b, t);
				INF_SYNC();
				bad = INF_BCAST(bad);
				if (bad) { status = INF_E_CODE; break; }
			}
			if (!inf_build(t.lens, 288, t.lcount, t.lsym, t.code, t.lit, INF_LIT_BITS) ||
			    !inf_build(t.lens + 288, 30, t.dcount, t.dsym, t.code + 288, t.dst, INF_DST_BITS)) { status = INF_E_CODE; break; }
			// ---- symbols: lane 0 decodes into the ring until something needs the warp ----
			for (;;) {
				uint32_t ex = 0, mlen = 0, mdist = 0, err = 0;
				if (lane == 0) {
					for (;;) {
						if (pos - flushed >= INF_FLUSH) { ex = EX_FLUSH; break; }
						// two literals per refill test: 32 bits are enough for a literal of any length and a second one from the
						// direct table (the pair is what gzip'ed FASTQ mostly consists of)
						if (b.bc < 32) inf_refill(b);
						int sym;
						const uint32_t e = t.lit[b.bb & ((1u << INF_LIT_BITS) - 1)];
						if (e && e < (256u << 4)) {
							if (pos >= out_cap) { ex = EX_ERR; err = INF_E_OUTPUT; break; }
							b.bb >>= (e & 15); b.bc -= (e & 15);
							ring[(pos + a0) & M] = (uint8_t)(e >> 4);
							pos++;
							const uint32_t e2 = t.lit[b.bb & ((1u << INF_LIT_BITS) - 1)];
							if (e2 && e2 < (256u << 4) && pos < out_cap) {
								b.bb >>= (e2 & 15); b.bc -= (e2 & 15);
								ring[(pos + a0) & M] = (uint8_t)(e2 >> 4);
								pos++;
							}
							continue;
						}
						if (e) { b.bb >>= (e & 15); b.bc -= (e & 15); sym = (int)(e >> 4); }
						else sym = inf_slow(b, t.lcount, t.lsym);
						if (sym < 256) {
							if (sym < 0) { ex = EX_ERR; err = INF_E_CODE; break; }
							if (pos >= out_cap) { ex = EX_ERR; err = INF_E_OUTPUT; break; }
							ring[(pos + a0) & M] = (uint8_t)sym;
							pos++;
							continue;
						}
						if (sym == 256) { ex = EX_EOB; break; }
						if (sym > 285) { ex = EX_ERR; err = INF_E_CODE; break; }
						mlen = INF_LBASE[sym - 257] + inf_take(b, INF_LEXT[sym - 257]);
						const int ds = inf_decode_dst(b, t);
						if (ds < 0 || ds > 29) { ex = EX_ERR; err = INF_E_CODE; break; }
						mdist = INF_DBASE[ds] + inf_take(b, INF_DEXT[ds]);
						if (mdist > pos) { ex = EX_ERR; err = INF_E_DIST; break; }
						if (pos + mlen > out_cap) { ex = EX_ERR; err = INF_E_OUTPUT; break; }
						if (mlen > INF_SOLO || mdist > INF_REACH) { ex = EX_MATCH; break; }
						// short match, whole in the ring: byte by byte, four at a time when the pieces cannot overlap
						uint32_t d = pos + a0, s = pos + a0 - mdist, left = mlen;
						if (mdist >= 4) {
							for (; left >= 4; left -= 4, d += 4, s += 4) {
								const uint8_t v0 = ring[s & M], v1 = ring[(s + 1) & M], v2 = ring[(s + 2) & M], v3 = ring[(s + 3) & M];
								ring[d & M] = v0; ring[(d + 1) & M] = v1; ring[(d + 2) & M] = v2; ring[(d + 3) & M] = v3;
							}
						}
						for (; left; left--, d++, s++) ring[d & M] = ring[s & M];
						pos += mlen;
					}
					if (b.over) { ex = EX_ERR; err = INF_E_INPUT; }
				}
				INF_SYNC();
				// what lane 0 found: ex (3 bits) | match length (9 bits) | distance (16 bits), and how far it got
				uint32_t w0 = ex | (err << 3) | (mlen << 6) | (mdist << 15);
				w0 = INF_BCAST(w0); pos = INF_BCAST(pos);
				ex = w0 & 7u; err = (w0 >> 3) & 7u; mlen = (w0 >> 6) & 511u; mdist = w0 >> 15;
				if (ex == EX_ERR) { status = (int)err; break; }
				if (ex == EX_EOB) break;
				if (ex == EX_MATCH) {
					// every byte comes from the part of the output that is already complete (i mod distance): out of the ring, or --
					// further back than the ring reaches -- out of the text buffer, where those bytes were flushed long ago
					if (mdist <= INF_REACH) {
						for (uint32_t i = lane; i < mlen; i += INF_LANES) {
							const uint32_t j = mdist >= mlen ? i : i % mdist;
							ring[(pos + i + a0) & M] = ring[(pos - mdist + j + a0) & M];
						}
					} else {
						for (uint32_t i = lane; i < mlen; i += INF_LANES) ring[(pos + i + a0) & M] = INF_LOAD_OUT(out + pos - mdist + i);
					}
					pos += mlen;
					INF_SYNC();
				}
				if (pos - flushed >= INF_FLUSH) flushed = inf_flush(ring, out, a0, flushed, pos, false);
			}
			if (status != INF_OK) break;
		} else { status = INF_E_BTYPE; break; }
		if (last) break;
	}
	if (status == INF_OK) flushed = inf_flush(ring, out, a0, flushed, pos, true);
	*out_len = pos;
	return status;
}

}  // namespace vgb
