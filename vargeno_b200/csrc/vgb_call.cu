// vgb_call.cu -- K5: per-SNP genotype call over the compact site arrays.
//
// Replaces the calling scan src/qv.cc:1573-1626 and choose_best_genotype src/qv.cc:1789-1848.  One thread per
// site; the likelihood tables come from the host (vgb_tables.cpp); the nine IEEE operations are issued with
// explicit round-to-nearest intrinsics so nothing is contracted into an FMA (the reference's qv.o is built
// without -march=native, Makefile:32-33, hence without FMA).
#include "vgb_internal.h"

namespace vgb {

__global__ void __launch_bounds__(256) k_call(const uint32_t *cnt, const uint8_t *code, const uint8_t *rf, const uint8_t *af,
                                               uint64_t n, const double *tab, uint8_t *gtype, double *conf)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t c = code[i];
	uint8_t g = 0;
	double cf = 0.0;
	const int r = (int)min(cnt[2 * i], (uint32_t)MAX_COV);          // counters saturate at MAX_COV (src/qv.cc:1412-1413)
	const int a = (int)min(cnt[2 * i + 1], (uint32_t)MAX_COV);
	if ((c & 3) != (c >> 2) && !((r == 0 && a == 0) || (r == MAX_COV && a == MAX_COV))) {   // :1580, :1821
		const double *t = tab + (r * 64 + a) * 3;
		const double g0 = t[0], g1 = t[1], g2 = t[2];
		const double p = __ddiv_rn((double)rf[i], 255.0);
		const double q = __ddiv_rn((double)af[i], 255.0);
		const double p2 = __dmul_rn(p, p);
		const double q2 = __dmul_rn(q, q);
		const double p_g0 = __dmul_rn(p2, g0);
		const double p_g1 = __dmul_rn(__dsub_rn(__dsub_rn(1.0, p2), q2), g1);
		const double p_g2 = __dmul_rn(q2, g2);
		const double total = __dadd_rn(__dadd_rn(p_g0, p_g1), p_g2);
		const double po = tab[64 * 64 * 3 + r + a];
		if (p_g0 > p_g1 && p_g0 > p_g2) { g = 1; cf = __dmul_rn(__ddiv_rn(p_g0, total), po); }        // GTYPE_REF
		else if (p_g1 > p_g0 && p_g1 > p_g2) { g = 3; cf = __dmul_rn(__ddiv_rn(p_g1, total), po); }   // GTYPE_HET
		else { g = 2; cf = __dmul_rn(__ddiv_rn(p_g2, total), po); }                                    // GTYPE_ALT
	}
	gtype[i] = g;
	conf[i] = cf;
}

__global__ void __launch_bounds__(256) k_split_counts(const uint32_t *cnt, uint64_t n, uint32_t *ref_cnt, uint32_t *alt_cnt)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	ref_cnt[i] = min(cnt[2 * i], (uint32_t)MAX_COV);
	alt_cnt[i] = min(cnt[2 * i + 1], (uint32_t)MAX_COV);
}

int call_sites(vgb_ctx *c, uint8_t *gtype, double *conf, uint64_t n_sites)
{
	if (!c->have_index) return set_err(c, VGB_E_ARG, "no index");
	if (n_sites != c->ix.n_sites) return set_err(c, VGB_E_ARG, "n_sites is %llu, index has %llu", (unsigned long long)n_sites, (unsigned long long)c->ix.n_sites);
	if (n_sites == 0) return VGB_OK;
	int rc;
	// output staging is allocated once per context: cudaMalloc / cudaFree in the call path would serialise with every
	// other context of the process (and of the node) on the driver lock
	if (!c->d_call_gt && (rc = dev_alloc(c, &c->d_call_gt, n_sites))) return rc;
	if (!c->d_call_conf && (rc = dev_alloc(c, &c->d_call_conf, n_sites))) return rc;
	uint8_t *d_g = c->d_call_gt;
	double *d_c = c->d_call_conf;
	k_call<<<(unsigned)((n_sites + 255) / 256), 256, 0, c->stream>>>(c->ix.cnt, c->ix.site_code, c->d_site_rf, c->d_site_af, n_sites, c->d_tables, d_g, d_c);
	c->launches++;
	cudaError_t e = cudaMemcpyAsync(gtype, d_g, n_sites, cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(conf, d_c, n_sites * 8, cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	if (e != cudaSuccess) return set_err(c, VGB_E_CUDA, "caller kernel failed: %s", cudaGetErrorString(e));
	return VGB_OK;
}

int fetch_pileup(vgb_ctx *c, uint32_t *ref_cnt, uint32_t *alt_cnt, uint64_t n_sites)
{
	if (!c->have_index) return set_err(c, VGB_E_ARG, "no index");
	if (n_sites != c->ix.n_sites) return set_err(c, VGB_E_ARG, "n_sites is %llu, index has %llu", (unsigned long long)n_sites, (unsigned long long)c->ix.n_sites);
	if (n_sites == 0) return VGB_OK;
	int rc;
	if (!c->d_fetch_ref && (rc = dev_alloc(c, &c->d_fetch_ref, n_sites))) return rc;
	if (!c->d_fetch_alt && (rc = dev_alloc(c, &c->d_fetch_alt, n_sites))) return rc;
	uint32_t *d_r = c->d_fetch_ref, *d_a = c->d_fetch_alt;
	k_split_counts<<<(unsigned)((n_sites + 255) / 256), 256, 0, c->stream>>>(c->ix.cnt, n_sites, d_r, d_a);
	c->launches++;
	cudaError_t e = cudaMemcpyAsync(ref_cnt, d_r, n_sites * 4, cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(alt_cnt, d_a, n_sites * 4, cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	if (e != cudaSuccess) return set_err(c, VGB_E_CUDA, "pileup fetch failed: %s", cudaGetErrorString(e));
	return VGB_OK;
}

}  // namespace vgb
