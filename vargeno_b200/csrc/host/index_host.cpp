// index_host.cpp -- `vargeno-b200 index <ref.fa> <snps.vcf> <prefix>`: the reference's `vargeno index`
// (src/qv.cc:2316-2346 -> src/dictgen.c, src/generate_bf.cc) with the sorting / collapsing / Bloom-filter work done on
// the GPU (vgb_build_index_device, csrc/vgb_build.cu).  The host keeps what is text: the two FASTA readers, the two VCF
// walks with their filters, and the file formats.  The five files `geno` reads come out byte-identical to the
// reference's (tests/test_gpu_cli.py compares them with the sha256 of reference-built files in tests/golden/).
// <prefix>.ref.bf.lite.bf (2.3 GB, written by the reference, read by nothing: src/generate_bf.cc:102-105,145-163) is written
// too unless --no-lite is given.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "geno_host.h"

namespace vgh {

namespace {

constexpr uint8_t BASE_X = 7, BASE_N = 4;

struct CodeTable {                      // encode_base, src/util.c:53-87
	uint8_t c[256];
	CodeTable()
	{
		memset(c, BASE_X, sizeof(c));
		const char *acgt = "ACGT";
		for (int i = 0; i < 4; i++) { c[(uint8_t)acgt[i]] = (uint8_t)i; c[(uint8_t)acgt[i] + 32] = (uint8_t)i; }
		c[(uint8_t)'N'] = BASE_N; c[(uint8_t)'n'] = BASE_N;
	}
};
const CodeTable CODE;

inline bool c_isspace(uint8_t ch) { return ch == ' ' || (ch >= '\t' && ch <= '\r'); }
inline uint8_t up(uint8_t ch) { return (ch >= 'a' && ch <= 'z') ? (uint8_t)(ch - 32) : ch; }

// Both views of the FASTA the reference takes:
//   raw  (src/generate_bf.cc:18-76):  name = the whole header line, characters as they are, newlines dropped
//   norm (src/fasta_parser.c:35-131): name cut at the first white space or '|' (<= 64 chars), upper case, non-ACGT -> 'N'
struct Fasta {
	std::vector<std::string> raw_names, names;
	std::vector<uint64_t> starts, lens;
	std::vector<uint8_t> raw, norm;      // contigs concatenated; same geometry in both views
};

bool read_fasta(const std::string &path, Fasta &fa, std::string &err)
{
	MappedFile f;
	if (!f.open(path, err)) return false;
	const uint8_t *d = f.data;
	const uint64_t n = f.size;
	fa.raw.reserve(n);
	uint64_t i = 0;
	while (i < n && d[i] != '>') i++;                  // anything in front of the first header is ignored
	while (i < n) {
		i++;                                           // '>'
		std::string name;
		while (i < n && d[i] != '\n' && d[i] != '>') name.push_back((char)d[i++]);
		if (i < n && d[i] == '\n') i++;
		const uint64_t start = fa.raw.size();
		while (i < n && d[i] != '>') { if (d[i] != '\n') fa.raw.push_back(d[i]); i++; }
		fa.raw_names.push_back(name);
		size_t cut = 0;
		while (cut < name.size() && cut < 64 && !c_isspace((uint8_t)name[cut]) && name[cut] != '|') cut++;
		fa.names.push_back(name.substr(0, cut));
		fa.starts.push_back(start);
		fa.lens.push_back(fa.raw.size() - start);
	}
	if (fa.names.empty()) { err = "no sequence in " + path; return false; }
	if (fa.names.size() > 128) { err = "more than 128 contigs (the reference's chrlens[128], src/qv.cc:482)"; return false; }
	fa.norm.resize(fa.raw.size());
	for (uint64_t k = 0; k < fa.raw.size(); k++) {
		const uint8_t code = CODE.c[fa.raw[k]];
		if (code == BASE_X) { err = "FASTA has a character outside ACGTNacgtn (the reference's Bloom-filter pass aborts in encode_kmer, src/util.c:103)"; return false; }
		fa.norm[k] = code < 4 ? (uint8_t)"ACGT"[code] : (uint8_t)'N';
	}
	for (size_t c = 0; c < fa.lens.size(); c++)
		if (fa.lens[c] < 32) { err = "contig " + fa.names[c] + " is shorter than 32 bases (the reference asserts, src/dictgen.c:17)"; return false; }
	if (fa.raw.size() >= 0xFFFFFFFEull) { err = "genome too long for 32-bit positions"; return false; }
	return true;
}

std::vector<std::string> split_tabs(const char *b, const char *e)
{
	std::vector<std::string> out;
	const char *s = b;
	for (const char *p = b;; p++) {
		if (p == e || *p == '\t') { out.emplace_back(s, p); s = p + 1; if (p == e) break; }
	}
	return out;
}

uint8_t freq_enc(double f) { const float v = (float)f * 255.0f; return (uint8_t)((int)v & 0xFF); }   // src/dictgen.c:741-742

struct SnpLines {                       // what vgb_build_index_device takes, file order
	std::vector<uint32_t> pos0;         // 0-based in the concatenation
	std::vector<uint8_t> code, rf, af;  // ref | alt << 2, encoded CAF pair
	std::vector<uint32_t> bf_pos0;      // lines that reach the Bloom-filter insert loop
};

// dictionary side: src/dictgen.c:599-748, record by record (same restatement as tools/index_builder.parse_vcf_for_dict)
bool vcf_for_dict(const MappedFile &vcf, const Fasta &fa, SnpLines &out, std::string &err)
{
	const bool ref_has_chr = !fa.names[0].empty() && fa.names[0][0] == 'c';
	bool has_freq = true;
	long freq_index = -1;
	const char *p = (const char *)vcf.data, *end = p + vcf.size;
	while (p < end) {
		const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
		const char *le = nl ? nl + 1 : end;             // line including its '\n' (the last line may lack one)
		std::string line(p, le);
		p = le;
		if (line.empty() || line[0] == '#' || line[0] == '\n') continue;
		if (line.back() != '\n') line.push_back('\n');
		const std::vector<std::string> fld = split_tabs(line.data(), line.data() + line.size());
		if (fld.size() < 5) continue;
		std::string chrom;
		for (char ch : fld[0]) { if (c_isspace((uint8_t)ch)) break; chrom.push_back(ch); }
		if ((chrom.empty() || chrom[0] != 'c') && ref_has_chr) chrom = "chr" + chrom;
		if (chrom.size() > 49) chrom.resize(49);
		if (fld[3].empty()) continue;
		const uint8_t ref_b = up((uint8_t)fld[3][0]);
		const uint8_t rc = CODE.c[ref_b];
		if (rc == BASE_X) continue;
		if (fld[3].size() > 1 && !c_isspace((uint8_t)fld[3][1])) continue;       // REF longer than one base (:644)
		const std::string &alt_f = fld[4];
		if (alt_f.size() > 1 && !c_isspace((uint8_t)alt_f[1])) continue;         // ALT longer than one base (:649)
		long ci = -1;
		for (size_t k = 0; k < fa.names.size(); k++) if (fa.names[k] == chrom) { ci = (long)k; break; }
		if (ci < 0) { fprintf(stderr, "[Error] chromosome name %s in VCF file not found in reference\n", chrom.c_str()); continue; }
		const uint64_t index = (uint64_t)((long long)strtod(fld[1].c_str(), nullptr) - 1) & 0xFFFFFFFFull;
		const uint8_t *seq = fa.norm.data() + fa.starts[ci];
		const uint64_t len = fa.lens[ci];
		if (index >= len || up(seq[index]) != ref_b) {
			err = "Mismatch between reference sequence and SNP file at 0-based index " + std::to_string(index) + " in " + fa.names[ci] +
			      " (the reference exits, src/dictgen.c:666-672)";
			return false;
		}
		if (index < 32 || index + 32 > len) continue;
		const uint8_t a2 = alt_f.empty() ? 0 : up((uint8_t)alt_f[0]);
		if (rc > 3 || a2 == 0 || CODE.c[a2] > 3 || a2 >= 'a') continue;
		double f1 = 0.5, f2 = 0.5;
		std::vector<const char *> toks;
		std::string info;
		if (has_freq) {
			if (fld.size() > 7) for (char ch : fld[7]) { if (ch == ' ' || ch == '\t' || ch == '\n') break; info.push_back(ch); }
			size_t pos = 0;
			while (pos < info.size()) {
				toks.push_back(info.c_str() + pos);
				const size_t m = info.find_first_of(";=", pos);
				if (m == std::string::npos) break;
				pos = m + 1;
			}
			for (size_t k = 0; k < toks.size(); k++) if (!strncmp(toks[k], "CAF", 3)) freq_index = (long)k + 1;
			if (freq_index == -1) has_freq = false;
		}
		if (has_freq) {
			if (freq_index >= (long)toks.size()) { err = "VCF record without CAF after one with CAF: undefined in the reference (src/dictgen.c:733-737)"; return false; }
			const char *t = toks[freq_index];
			f1 = strtod(t, nullptr);
			const char *comma = strchr(t, ',');
			if (!comma) { err = "CAF without a comma: undefined in the reference (src/dictgen.c:736)"; return false; }
			f2 = strtod(comma + 1, nullptr);
		}
		if (a2 == ref_b) continue;
		bool skip = false;
		for (uint64_t k = index - 32; k < index; k++) if (CODE.c[seq[k]] == BASE_N) { skip = true; break; }
		for (uint64_t k = index + 1; k < index + 32 && !skip; k++) if (seq[k] == 'N') skip = true;
		if (skip) continue;
		out.pos0.push_back((uint32_t)(fa.starts[ci] + index));
		out.code.push_back((uint8_t)(rc | (CODE.c[a2] << 2)));
		out.rf.push_back(freq_enc(f1));
		out.af.push_back(freq_enc(f2));
	}
	return true;
}

// Bloom-filter side: src/generate_bf.cc:203-246 (same walk as tools/index_builder.bf_line_positions, including the stale
// sequence kept when a contig name is unknown); the N test of the preceding 32-mer is done by the GPU builder
bool vcf_for_bf(const MappedFile &vcf, const Fasta &fa, SnpLines &out, std::string &err)
{
	std::string pre = "XO";
	long ci = -1;
	const uint8_t *seq = nullptr;
	uint64_t len = 0;
	const char *p = (const char *)vcf.data, *end = p + vcf.size;
	while (p < end) {
		const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
		const char *le = nl ? nl : end;
		const char *b = p;
		p = nl ? nl + 1 : end;
		if (b == le || *b == '#') continue;
		const std::vector<std::string> col = split_tabs(b, le);
		if (col.size() < 5) continue;
		std::string chrom = col[0];
		if (chrom.empty() || chrom[0] != 'c') chrom = "chr" + chrom;
		const long long pos = strtoll(col[1].c_str(), nullptr, 10) - 1;
		const std::string &r = col[3], &a = col[4];
		if (r.size() > 1 || a.size() > 1) continue;
		if (chrom != pre) {
			for (size_t k = 0; k < fa.raw_names.size(); k++)
				if (fa.raw_names[k] == chrom) { seq = fa.raw.data() + fa.starts[k]; len = fa.lens[k]; ci = (long)k; break; }
			pre = chrom;
		}
		if (pos < 32 || (uint64_t)pos + 32 > len) continue;
		if (r.empty() || a.empty()) continue;
		if ((uint8_t)r[0] != seq[pos] || r == a || a == "N" || a == "n") continue;
		if (CODE.c[(uint8_t)a[0]] == BASE_X) { err = "ALT base '" + a + "' makes the reference abort in shift_kmer (src/util.c:122)"; return false; }
		out.bf_pos0.push_back((uint32_t)(fa.starts[ci] + (uint64_t)pos));
	}
	return true;
}

// ---------------------------------------------------------------------------------------------------------------
// UCSC snp-table input: the legacy SNP format of the reference (`vargeno ucscd` / `ucscbf`, src/qv.cc:1954-2008,
// 2225-2238; the same code sits commented out in its `index`, :2246-2314).  Columns: 1 chrom, 2 chromStart (0-based), 6 strand,
// 7 refNCBI, 8 refUCSC, 9 observed, 11 class, 21 alleleFreqCount, 22 alleles, 24 alleleFreqs.
// ---------------------------------------------------------------------------------------------------------------
struct UcscBfLines { std::vector<uint32_t> pos0; std::vector<uint8_t> alt; };   // lines that reach the insert loop of constructBfFromUcsc

inline uint8_t rev_base(uint8_t c)       // rev(), src/dictgen.c:322-345: complement of an upper- or lower-case base, anything else unchanged
{
	switch (c) { case 'A': case 'a': return 'T'; case 'C': case 'c': return 'G'; case 'G': case 'g': return 'C'; case 'T': case 't': return 'A'; default: return c; }
}
inline bool is_acgt(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

// dictionary side: make_snp_dict, src/dictgen.c:350-540.  snp_locs (optional): the bool-per-position table the reference
// keeps for `filt` (index = 1-based position in the concatenation; src/dictgen.c:459-466, written only #if GEN_FLT_DATA).
bool ucsc_for_dict(const MappedFile &txt, const Fasta &fa, SnpLines &out, std::vector<uint8_t> *snp_locs, std::string &err)
{
	if (snp_locs) snp_locs->assign(10, 0);
	long ci = -1;                                       // `chrom` of the reference's loop: survives records of unknown contigs
	const char *p = (const char *)txt.data, *end = p + txt.size;
	uint64_t lineno = 0;
	while (p < end) {
		const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
		const char *le = nl ? nl : end;
		const char *b = p;
		p = nl ? nl + 1 : end;
		lineno++;
		if (b == le || *b == '#') continue;
		if ((uint64_t)(le - b) >= 5999) { err = "UCSC table line " + std::to_string(lineno) + " is longer than the reference's 6000-byte buffer (src/dictgen.c:362)"; return false; }
		const std::vector<std::string> f = split_tabs(b, le);
		if (f.size() < 25) { err = "UCSC table line " + std::to_string(lineno) + " has fewer than 25 columns (undefined in the reference)"; return false; }
		std::string chrom;
		for (char ch : f[1]) { if (c_isspace((uint8_t)ch) || chrom.size() >= 49) break; chrom.push_back(ch); }
		if (f[7].empty() || f[8].empty()) continue;
		const uint8_t ref_b = up((uint8_t)f[7][0]);
		const uint8_t rc = CODE.c[ref_b];
		if (rc == BASE_X || strncmp(f[11].c_str(), "single", 6) != 0 || ref_b != up((uint8_t)f[8][0])) continue;
		if (f[7].size() != 1 || f[8].size() != 1) continue;                           // reference alleles one base long (:417)
		if (ci < 0 || fa.names[ci] != chrom) {
			ci = -1;
			for (size_t k = 0; k < fa.names.size(); k++) if (fa.names[k] == chrom) { ci = (long)k; break; }
			if (ci < 0) continue;
		}
		const uint64_t index = (uint32_t)atoi(f[2].c_str());
		const uint8_t *seq = fa.norm.data() + fa.starts[ci];
		const uint64_t len = fa.lens[ci];
		if (index >= len || up(seq[index]) != ref_b) {
			err = "Mismatch found between reference sequence and SNP file at 0-based index " + std::to_string(index) + " in " + fa.names[ci] +
			      " (the reference exits, src/dictgen.c:431-437)";
			return false;
		}
		if (index < 32 || index + 32 > len) continue;
		if (f[21].empty() || f[21][0] != '2') continue;                               // bi-allelic only (:444)
		const bool neg = !f[6].empty() && f[6][0] == '-';
		if (!neg && (f[6].empty() || f[6][0] != '+')) { err = "strand is neither + nor - at line " + std::to_string(lineno) + " (the reference asserts, src/dictgen.c:450)"; return false; }
		if (f[22].size() < 3) { err = "alleles column too short at line " + std::to_string(lineno); return false; }
		const uint8_t a1 = neg ? rev_base(up((uint8_t)f[22][0])) : up((uint8_t)f[22][0]);
		const uint8_t a2 = neg ? rev_base(up((uint8_t)f[22][2])) : up((uint8_t)f[22][2]);
		if (!is_acgt(a1) || !is_acgt(a2)) { err = "allele outside ACGT at line " + std::to_string(lineno) + " (the reference asserts, src/dictgen.c:455-456)"; return false; }
		if (a1 != ref_b && a2 != ref_b) continue;
		if (snp_locs) {
			const uint64_t loc = fa.starts[ci] + 1 + index;
			if (loc >= snp_locs->size()) snp_locs->resize(loc + 1, 0);
			(*snp_locs)[loc] = 1;
		}
		const char *fq = f[24].c_str();
		double f1 = (float)atof(fq);
		const char *comma = strchr(fq, ',');
		if (!comma) { err = "alleleFreqs without a comma at line " + std::to_string(lineno) + " (the reference runs off the buffer, src/dictgen.c:470)"; return false; }
		double f2 = (float)atof(comma + 1);
		if (a2 == ref_b) std::swap(f1, f2);
		for (char chv : f[9]) {                                                       // observed: the first usable alternative allele decides (:481-517)
			if (c_isspace((uint8_t)chv)) break;
			const uint8_t alt = neg ? rev_base(up((uint8_t)chv)) : up((uint8_t)chv);
			if (alt == ref_b || !is_acgt(alt)) continue;
			bool skip = false;
			for (uint64_t k = index - 32; k < index; k++) if (CODE.c[seq[k]] == BASE_N) { skip = true; break; }
			for (uint64_t k = index + 1; k < index + 32 && !skip; k++) if (seq[k] == 'N' || seq[k] == 'n') skip = true;
			if (!skip) {
				out.pos0.push_back((uint32_t)(fa.starts[ci] + index));
				out.code.push_back((uint8_t)(rc | (CODE.c[alt] << 2)));
				out.rf.push_back(freq_enc(f1));
				out.af.push_back(freq_enc(f2));
			}
			break;
		}
	}
	return true;
}

// Bloom-filter side: constructBfFromUcsc, src/generate_bf.cc:439-592 (raw FASTA view: whole header line as the name; the stale
// sequence is kept when a contig is unknown; "single" must match exactly here, a prefix is enough on the dictionary side)
bool ucsc_for_bf(const MappedFile &txt, const Fasta &fa, UcscBfLines &out, std::string &err)
{
	std::string pre = "XO";
	long ci = -1;
	const uint8_t *seq = nullptr;
	uint64_t len = 0;
	const char *p = (const char *)txt.data, *end = p + txt.size;
	uint64_t lineno = 0;
	while (p < end) {
		const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
		const char *le = nl ? nl : end;
		const char *b = p;
		p = nl ? nl + 1 : end;
		lineno++;
		if (b == le || *b == '#') continue;
		const std::vector<std::string> c = split_tabs(b, le);
		if (c.size() < 25) { err = "UCSC table line " + std::to_string(lineno) + " has fewer than 25 columns (undefined in the reference)"; return false; }
		if (c[7].empty() || c[8].empty()) continue;
		const uint8_t ref_b = up((uint8_t)c[7][0]);
		if (CODE.c[ref_b] == BASE_X || c[11] != "single" || ref_b != up((uint8_t)c[8][0])) continue;
		if (c[7].size() != 1 || c[8].size() != 1) continue;
		if (c[1] != pre) {
			bool found = false;
			for (size_t k = 0; k < fa.raw_names.size(); k++)
				if (fa.raw_names[k] == c[1]) { seq = fa.raw.data() + fa.starts[k]; len = fa.lens[k]; ci = (long)k; found = true; break; }
			if (!found) continue;
			pre = c[1];
		}
		const uint64_t index = (uint64_t)(uint32_t)strtol(c[2].c_str(), nullptr, 10);
		if (index >= len || up(seq[index]) != ref_b) { err = "Mismatch found between reference sequence and SNP file at 0-based index " + std::to_string(index) + " (the reference exits, src/generate_bf.cc:507-510)"; return false; }
		if (index < 32 || index + 32 > len) continue;
		if (c[21] != "2") continue;
		const bool neg = !c[6].empty() && c[6][0] == '-';
		if (!neg && (c[6].empty() || c[6][0] != '+')) { err = "strand is neither + nor - at line " + std::to_string(lineno); return false; }
		if (c[22].size() < 3) { err = "alleles column too short at line " + std::to_string(lineno); return false; }
		const uint8_t a1 = neg ? rev_base(up((uint8_t)c[22][0])) : up((uint8_t)c[22][0]);
		const uint8_t a2 = neg ? rev_base(up((uint8_t)c[22][2])) : up((uint8_t)c[22][2]);
		if (!is_acgt(a1) || !is_acgt(a2)) { err = "allele outside ACGT at line " + std::to_string(lineno); return false; }
		if (a1 != ref_b && a2 != ref_b) continue;
		bool any = false;
		for (char chv : c[9]) {
			if (c_isspace((uint8_t)chv)) break;
			const uint8_t alt = neg ? rev_base(up((uint8_t)chv)) : up((uint8_t)chv);
			if (alt == ref_b || !is_acgt(alt)) continue;
			out.pos0.push_back((uint32_t)(fa.starts[ci] + index));
			out.alt.push_back(CODE.c[alt]);
			any = true;
			break;
		}
		if (!any) { err = "observed column without a usable alternative allele at line " + std::to_string(lineno) + " (the reference runs off the string, src/generate_bf.cc:536)"; return false; }
	}
	return true;
}

bool write_all(int fd, const void *buf, uint64_t n)
{
	const uint8_t *p = (const uint8_t *)buf;
	while (n) {
		const ssize_t w = ::write(fd, p, (size_t)std::min<uint64_t>(n, 1ull << 30));
		if (w <= 0) return false;
		p += w; n -= (uint64_t)w;
	}
	return true;
}

// device records -> file, through a bounded host buffer
bool write_device(vgb_ctx *ctx, int fd, const void *dptr, uint64_t bytes, std::vector<uint8_t> &buf, std::string &err)
{
	const uint64_t step = buf.size();
	for (uint64_t off = 0; off < bytes; off += step) {
		const uint64_t n = std::min(step, bytes - off);
		if (vgb_memcpy_d2h(ctx, buf.data(), (const uint8_t *)dptr + off, n) != VGB_OK) { err = vgb_last_error(ctx); return false; }
		if (!write_all(fd, buf.data(), n)) { err = "write failed"; return false; }
	}
	return true;
}

// sdsl bit_vector container (src/generate_bf.h:83-89): u64 bit count, then ceil(bits/64) words; words beyond `nwords` are zero
bool write_bitvector(vgb_ctx *ctx, const std::string &path, const uint64_t *dwords, uint64_t bits, uint64_t nwords, std::vector<uint8_t> &buf, std::string &err)
{
	const int fd = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
	if (fd < 0) { err = "cannot create " + path; return false; }
	const uint64_t total_words = (bits + 63) / 64;
	nwords = std::min(nwords, total_words);
	bool ok = write_all(fd, &bits, 8) && write_device(ctx, fd, dwords, nwords * 8, buf, err);
	if (ok && ftruncate(fd, (off_t)(8 + total_words * 8)) != 0) { ok = false; err = "cannot size " + path; }
	::close(fd);
	if (!ok && err.empty()) err = "cannot write " + path;
	return ok;
}

}  // namespace

static bool ends_with(const std::string &s, const char *suf) { const size_t n = strlen(suf); return s.size() >= n && s.compare(s.size() - n, n, suf) == 0; }

int run_index(const std::string &fasta, const std::string &vcf_path, const std::string &prefix, int device, bool verbose, const std::string &dump_parse,
              bool write_lite, const std::string &snp_locs_path)
{
	const bool ucsc = ends_with(vcf_path, ".txt");      // the reference tells its two SNP formats apart by extension (src/qv.cc:2244-2246,2315)
	const auto t0 = std::chrono::steady_clock::now();
	auto secs = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
	std::string err;
	Fasta fa;
	if (!read_fasta(fasta, fa, err)) { fprintf(stderr, "vargeno-b200: %s\n", err.c_str()); return EXIT_FAILURE; }
	MappedFile vcf;
	if (!vcf.open(vcf_path, err)) { fprintf(stderr, "vargeno-b200: %s\n", err.c_str()); return EXIT_FAILURE; }
	SnpLines sl;
	UcscBfLines ub;
	std::vector<uint8_t> snp_locs;
	const bool ok_parse = ucsc ? (ucsc_for_dict(vcf, fa, sl, snp_locs_path.empty() ? nullptr : &snp_locs, err) && ucsc_for_bf(vcf, fa, ub, err))
	                           : (vcf_for_dict(vcf, fa, sl, err) && vcf_for_bf(vcf, fa, sl, err));
	if (!ok_parse) { fprintf(stderr, "vargeno-b200: %s\n", err.c_str()); return EXIT_FAILURE; }
	if (!snp_locs_path.empty() && !ucsc) { fprintf(stderr, "vargeno-b200: --snp-locs goes with the UCSC table input (src/qv.cc:1988-1996)\n"); return EXIT_FAILURE; }
	if (verbose) fprintf(stderr, "parsed %zu contigs, %zu bp, %zu dictionary SNP lines, %zu Bloom-filter SNP lines (%.2f s)\n", fa.names.size(),
	                     fa.norm.size(), sl.pos0.size(), ucsc ? ub.pos0.size() : sl.bf_pos0.size(), secs());
	if (!snp_locs_path.empty()) {
		// the table `filt` takes: u64 size, then one byte per 1-based concatenated position (src/qv.cc:1988-1996)
		const int fd = ::open(snp_locs_path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
		const uint64_t n = snp_locs.size();
		const bool ok = fd >= 0 && write_all(fd, &n, 8) && write_all(fd, snp_locs.data(), n);
		if (fd >= 0) ::close(fd);
		if (!ok) { fprintf(stderr, "vargeno-b200: cannot write %s\n", snp_locs_path.c_str()); return EXIT_FAILURE; }
	}

	if (!dump_parse.empty()) {
		// host-side half only (no GPU needed): what the two VCF walks hand to the device builder, as text
		FILE *f = fopen(dump_parse.c_str(), "w");
		if (!f) { fprintf(stderr, "vargeno-b200: cannot create %s\n", dump_parse.c_str()); return EXIT_FAILURE; }
		for (size_t c = 0; c < fa.names.size(); c++) fprintf(f, "C %s %llu %llu\n", fa.names[c].c_str(), (unsigned long long)fa.starts[c], (unsigned long long)fa.lens[c]);
		for (size_t k = 0; k < sl.pos0.size(); k++) fprintf(f, "D %u %u %u %u\n", sl.pos0[k], sl.code[k], sl.rf[k], sl.af[k]);
		for (size_t k = 0; k < sl.bf_pos0.size(); k++) fprintf(f, "B %u\n", sl.bf_pos0[k]);
		for (size_t k = 0; k < ub.pos0.size(); k++) fprintf(f, "U %u %u\n", ub.pos0[k], ub.alt[k]);
		if (!snp_locs.empty()) { uint64_t t = 0; for (uint8_t v : snp_locs) t += v; fprintf(f, "L %zu %llu\n", snp_locs.size(), (unsigned long long)t); }
		fclose(f);
		return EXIT_SUCCESS;
	}
	vgb_config cfg{};
	cfg.device = device; cfg.world_size = 1; cfg.rank = 0; cfg.max_chunk_bytes = 1 << 20;
	vgb_ctx *ctx = nullptr;
	if (vgb_ctx_create(&ctx, &cfg) != VGB_OK) { fprintf(stderr, "vargeno-b200: %s\n", vgb_last_error(nullptr)); return EXIT_FAILURE; }
	int status = EXIT_FAILURE;
	void *d_genome = vgb_device_alloc(ctx, fa.norm.size());
	vgb_index_view view{};
	bool built = false;
	do {
		if (!d_genome || vgb_memcpy_h2d(ctx, d_genome, fa.norm.data(), fa.norm.size()) != VGB_OK) { err = vgb_last_error(ctx); break; }
		if (vgb_build_index_device(ctx, (const uint8_t *)d_genome, fa.norm.size(), fa.starts.data(), fa.lens.data(), (uint32_t)fa.names.size(),
		                           sl.pos0.data(), sl.code.data(), sl.rf.data(), sl.af.data(), sl.pos0.size(), sl.bf_pos0.data(), sl.bf_pos0.size(),
		                           &view) != VGB_OK) { err = vgb_last_error(ctx); break; }
		built = true;
		if (verbose) fprintf(stderr, "built on the device: %llu reference 32-mers (%llu aux rows), %llu SNP 32-mers (%llu aux rows) (%.2f s)\n",
		                     (unsigned long long)view.n_ref, (unsigned long long)view.n_ref_aux, (unsigned long long)view.n_snp,
		                     (unsigned long long)view.n_snp_aux, secs());
		std::vector<uint8_t> buf(64u << 20);
		{   // .ref.dict: u64 n, u64 aux_n, n x 13 B, aux_n x 40 B (src/dictgen.c:63-154)
			const int fd = ::open((prefix + ".ref.dict").c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
			if (fd < 0) { err = "cannot create " + prefix + ".ref.dict"; break; }
			const uint64_t hdr[2] = { view.n_ref, view.n_ref_aux };
			const bool ok = write_all(fd, hdr, 16) && write_device(ctx, fd, view.ref_records, view.n_ref * 13, buf, err) &&
			                write_device(ctx, fd, view.ref_aux, view.n_ref_aux * 40, buf, err);
			::close(fd);
			if (!ok) { if (err.empty()) err = "cannot write " + prefix + ".ref.dict"; break; }
		}
		{   // .snp.dict: u64 m, u64 aux_m, m x 16 B, aux_m x 78 B (src/dictgen.c:156-275)
			const int fd = ::open((prefix + ".snp.dict").c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
			if (fd < 0) { err = "cannot create " + prefix + ".snp.dict"; break; }
			const uint64_t hdr[2] = { view.n_snp, view.n_snp_aux };
			const bool ok = write_all(fd, hdr, 16) && write_device(ctx, fd, view.snp_records, view.n_snp * 16, buf, err) &&
			                write_device(ctx, fd, view.snp_aux, view.n_snp_aux * 78, buf, err);
			::close(fd);
			if (!ok) { if (err.empty()) err = "cannot write " + prefix + ".snp.dict"; break; }
		}
		if (!write_bitvector(ctx, prefix + ".ref.bf", view.ref_bf_words, view.ref_bf_bits, view.ref_bf_nwords, buf, err)) break;
		if (!ucsc) {
			if (!write_bitvector(ctx, prefix + ".snp.bf", view.snp_bf_words, view.snp_bf_bits, view.snp_bf_nwords, buf, err)) break;
		} else {
			// UCSC input: the SNP filter holds 33 values per accepted record (src/generate_bf.cc:538-556), built by its own kernel
			uint64_t *d_bf = nullptr, bits = 0, nw = 0;
			if (vgb_build_snp_bf_ucsc_device(ctx, (const uint8_t *)d_genome, ub.pos0.data(), ub.alt.data(), ub.pos0.size(), &d_bf, &bits, &nw) != VGB_OK) {
				err = vgb_last_error(ctx);
				break;
			}
			const bool ok = write_bitvector(ctx, prefix + ".snp.bf", d_bf, bits, nw, buf, err);
			vgb_device_free(ctx, d_bf);
			if (!ok) break;
		}
		if (write_lite) {   // the sixth file of the reference's `index`
			uint64_t *d_lite = nullptr, bits = 0, nw = 0;
			if (vgb_build_ref_lite_bf_device(ctx, (const uint8_t *)d_genome, fa.starts.data(), fa.lens.data(), (uint32_t)fa.names.size(), &d_lite, &bits, &nw) != VGB_OK) {
				err = vgb_last_error(ctx);
				break;
			}
			const bool ok = write_bitvector(ctx, prefix + ".ref.bf.lite.bf", d_lite, bits, nw, buf, err);
			vgb_device_free(ctx, d_lite);
			if (!ok) break;
		}
		{   // .chrlens (src/qv.cc:2344-2346)
			FILE *f = fopen((prefix + ".chrlens").c_str(), "w");
			if (!f) { err = "cannot create " + prefix + ".chrlens"; break; }
			for (size_t c = 0; c < fa.names.size(); c++) fprintf(f, "%s %llu\n", fa.names[c].c_str(), (unsigned long long)fa.lens[c]);
			fclose(f);
		}
		status = EXIT_SUCCESS;
	} while (false);
	if (status != EXIT_SUCCESS) fprintf(stderr, "vargeno-b200: %s\n", err.c_str());
	if (built) vgb_free_index_device(ctx, &view);
	if (d_genome) vgb_device_free(ctx, d_genome);
	vgb_ctx_destroy(ctx);
	printf("Time: %.2f sec\n", secs());
	return status;
}

// `vargeno-b200 filt <ref.dict> <snp_locs> <out.dict>`: dict_filt, src/dict_filt.c:23-79 -- keeps the reference-dictionary entries
// that are ambiguous or lie within a read length (READ_LEN 101, src/vartype.h:12) of a SNP; aux rows are copied as they are.
int run_filt(const std::string &ref_dict, const std::string &snp_locs, const std::string &out_path)
{
	constexpr uint64_t READ_LEN = 101;
	std::string err;
	MappedFile rd, sl;
	if (!rd.open(ref_dict, err) || !sl.open(snp_locs, err)) { fprintf(stderr, "vargeno-b200: %s\n", err.c_str()); return EXIT_FAILURE; }
	if (sl.size < 8 || rd.size < 16) { fprintf(stderr, "vargeno-b200: truncated input\n"); return EXIT_FAILURE; }
	uint64_t n_loc, n, an;
	memcpy(&n_loc, sl.data, 8);
	memcpy(&n, rd.data, 8);
	memcpy(&an, rd.data + 8, 8);
	if (sl.size < 8 + n_loc || rd.size < 16 + 13 * n + 40 * an) { fprintf(stderr, "vargeno-b200: truncated input\n"); return EXIT_FAILURE; }
	if (n_loc < READ_LEN) { fprintf(stderr, "vargeno-b200: snp_locs table shorter than a read (the reference's window arithmetic wraps, src/dict_filt.c:15)\n"); return EXIT_FAILURE; }
	const uint8_t *loc = sl.data + 8;
	std::vector<uint32_t> pre(n_loc + 1, 0);            // pre[i] = number of SNP positions below i
	for (uint64_t i = 0; i < n_loc; i++) pre[i + 1] = pre[i] + (loc[i] ? 1u : 0u);
	FILE *out = fopen(out_path.c_str(), "wb");
	if (!out) { fprintf(stderr, "vargeno-b200: cannot create %s\n", out_path.c_str()); return EXIT_FAILURE; }
	uint64_t kept = 0;
	fwrite(&kept, 8, 1, out);
	fwrite(&an, 8, 1, out);
	std::vector<uint8_t> buf;
	buf.reserve(13u << 20);
	const uint8_t *rec = rd.data + 16;
	for (uint64_t i = 0; i < n; i++, rec += 13) {
		uint32_t pos;
		memcpy(&pos, rec + 8, 4);
		bool keep = pos == 0xFFFFFFFFu || rec[12] == 1;
		if (!keep && pos < n_loc) {
			const uint64_t lo = pos > READ_LEN - 32 ? pos - (READ_LEN - 32) : 0;
			const uint64_t hi = pos < n_loc - (READ_LEN - 1) ? pos + (READ_LEN - 1) : n_loc - 1;
			keep = pre[hi + 1] != pre[lo];
		}
		if (keep) {
			buf.insert(buf.end(), rec, rec + 13);
			kept++;
			if (buf.size() >= (13u << 20) - 13) { fwrite(buf.data(), 1, buf.size(), out); buf.clear(); }
		}
	}
	fwrite(buf.data(), 1, buf.size(), out);
	fwrite(rd.data + 16 + 13 * n, 1, 40 * an, out);
	bool ok = fflush(out) == 0 && fseek(out, 0, SEEK_SET) == 0 && fwrite(&kept, 8, 1, out) == 1;
	ok = (fclose(out) == 0) && ok;
	if (!ok) { fprintf(stderr, "vargeno-b200: writing %s failed\n", out_path.c_str()); return EXIT_FAILURE; }
	printf("New size: %llu\n", (unsigned long long)kept);
	printf("Removed:  %llu/%llu\n", (unsigned long long)(n - kept), (unsigned long long)n);
	return EXIT_SUCCESS;
}

}  // namespace vgh
