// sector_probe.cu -- what does ONE random 32-byte sector load cost on this part?  Stand-alone measurement tool
// (not part of libvgb200.so):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sector_probe sector_probe.cu
//
// Every thread issues independent 16-byte loads from uniformly random 32-byte sectors of a large buffer, with one of
// several load flavours; the program prints useful GB/s (32 B per load) per flavour.  Run it under
//   ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum
// to see how many bytes each flavour really moves per load (fetch granularity of L1 -> L2 and L2 -> DRAM).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
	x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
	x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
	return x ^ (x >> 31);
}

template <int MODE>
__device__ __forceinline__ uint4 load16(const uint4 *p)
{
	uint4 v;
	if (MODE == 0) v = __ldg(p);                                                                 // ld.global.nc
	else if (MODE == 1) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	else if (MODE == 2) asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	else if (MODE == 3) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	else if (MODE == 4) asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	else if (MODE == 5) asm volatile("ld.global.nc.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	else if (MODE == 6) asm volatile("ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	else if (MODE == 7) asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	else {                                                                                       // 4-byte __ldg
		v.x = __ldg(reinterpret_cast<const uint32_t *>(p)); v.y = v.z = v.w = 0;
	}
	return v;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_probe(const uint4 *buf, uint64_t n_sectors, uint32_t per_thread, uint64_t seed, unsigned long long *sink)
{
	const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t acc = 0;
	uint64_t s = mix64(seed ^ (tid * 0x9E3779B97F4A7C15ull));
	for (uint32_t i = 0; i < per_thread; i += 8) {
		uint4 v[8];
#pragma unroll
		for (int k = 0; k < 8; k++) {
			s = mix64(s + k + 1);
			v[k] = load16<MODE>(buf + 2 * (s % n_sectors));
		}
#pragma unroll
		for (int k = 0; k < 8; k++) acc += v[k].x ^ v[k].w;
	}
	if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

// sweep variant: MLP independent DEPENDENT-ADDRESS chains per thread (the next address of a chain is computed from the loaded
// value, so exactly MLP loads per thread are in flight, whatever the compiler does), `ctas` CTAs of 256 threads per SM, and
// `ballast` bytes of dynamic shared memory per CTA (what a kernel's own shared memory takes away from the unified L1)
template <int MODE, int MLP>
__global__ void __launch_bounds__(256) k_probe_mlp(const uint4 *buf, uint64_t n_sectors, uint32_t per_thread, uint64_t seed, unsigned long long *sink)
{
	extern __shared__ unsigned char occupancy_ballast[];
	const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t s[MLP];
#pragma unroll
	for (int k = 0; k < MLP; k++) s[k] = mix64(seed ^ ((tid * MLP + k) * 0x9E3779B97F4A7C15ull));
	for (uint32_t i = 0; i < per_thread; i += MLP) {
		uint4 v[MLP];
#pragma unroll
		for (int k = 0; k < MLP; k++) v[k] = load16<MODE>(buf + 2 * (s[k] % n_sectors));
#pragma unroll
		for (int k = 0; k < MLP; k++) s[k] = mix64(s[k] + v[k].x + 1);
	}
	uint64_t acc = 0;
#pragma unroll
	for (int k = 0; k < MLP; k++) acc ^= s[k];
	if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

template <int MODE, int MLP>
static void sweep_one(const char *name, const uint4 *buf, uint64_t n_sectors, unsigned long long *sink, int ctas, size_t ballast)
{
	cudaFuncSetAttribute(k_probe_mlp<MODE, MLP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ballast);
	int occ = 0;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_probe_mlp<MODE, MLP>, 256, ballast);
	if (occ > ctas) occ = ctas;
	const unsigned grid = 148u * (unsigned)occ;                      // one wave: exactly `occ` CTAs per SM
	const uint32_t per_thread = 2048;
	const uint64_t loads = (uint64_t)grid * 256 * per_thread;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	k_probe_mlp<MODE, MLP><<<grid, 256, ballast>>>(buf, n_sectors, per_thread, 1, sink);
	cudaEventRecord(e0);
	for (int r = 0; r < 2; r++) k_probe_mlp<MODE, MLP><<<grid, 256, ballast>>>(buf, n_sectors, per_thread, 2 + r, sink);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	ms /= 2;
	printf("{\"sweep\": \"%s\", \"loads_in_flight_per_thread\": %d, \"ctas_per_sm\": %d, \"warps_per_sm\": %d, \"loads_in_flight_per_sm\": %d, \"smem_per_sm_kb\": %.0f, \"ms\": %.3f, \"g_loads_per_s\": %.2f, \"err\": \"%s\"}\n",
	       name, MLP, occ, occ * 8, occ * 256 * MLP, occ * (ballast + 1024) / 1024.0, ms, loads / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
	fflush(stdout);
}

template <int MODE>
static void sweep(const char *name, const uint4 *buf, uint64_t n_sectors, unsigned long long *sink)
{
	// (1) loads in flight per thread x resident CTAs per SM, no shared memory
	const int ctas[] = { 1, 2, 4, 8 };
	for (int c : ctas) {
		sweep_one<MODE, 1>(name, buf, n_sectors, sink, c, 0);
		sweep_one<MODE, 2>(name, buf, n_sectors, sink, c, 0);
		sweep_one<MODE, 4>(name, buf, n_sectors, sink, c, 0);
		sweep_one<MODE, 8>(name, buf, n_sectors, sink, c, 0);
		sweep_one<MODE, 16>(name, buf, n_sectors, sink, c, 0);
	}
	// (2) the shape of k_geno8 (4 CTAs of 256 threads per SM) with more and more shared memory per CTA
	const size_t ballast[] = { 0, 7 << 10, 15 << 10, 24 << 10, 32 << 10, 40 << 10, 48 << 10, 56 << 10 };
	for (size_t b : ballast) {
		sweep_one<MODE, 2>(name, buf, n_sectors, sink, 4, b);
		sweep_one<MODE, 4>(name, buf, n_sectors, sink, 4, b);
		sweep_one<MODE, 8>(name, buf, n_sectors, sink, 4, b);
	}
}

static unsigned g_grid = 148 * 64;
static uint32_t g_per_thread = 64;

template <int MODE>
static void run(const char *name, const uint4 *buf, uint64_t n_sectors, unsigned long long *sink)
{
	const uint32_t per_thread = g_per_thread;
	const unsigned grid = g_grid;
	const uint64_t loads = (uint64_t)grid * 256 * per_thread;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	k_probe<MODE><<<grid, 256>>>(buf, n_sectors, per_thread, 1, sink);
	cudaEventRecord(e0);
	for (int r = 0; r < 3; r++) k_probe<MODE><<<grid, 256>>>(buf, n_sectors, per_thread, 2 + r, sink);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	ms /= 3;
	printf("{\"flavour\": \"%s\", \"grid\": %u, \"per_thread\": %u, \"ms\": %.3f, \"loads_per_s\": %.4g, \"useful_gbs_at_32B\": %.1f, \"err\": \"%s\"}\n", name, grid, per_thread, ms, loads / (ms * 1e-3),
	       loads * 32.0 / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
	fflush(stdout);
}

int main(int argc, char **argv)
{
	const uint64_t gib = argc > 1 ? strtoull(argv[1], nullptr, 10) : 32;
	const int gran = argc > 2 ? atoi(argv[2]) : 0;
	if (argc > 4) g_grid = (unsigned)atoi(argv[4]);          // CTAs of 256 threads
	if (argc > 5) g_per_thread = (uint32_t)atoi(argv[5]);    // loads per thread (multiple of 8)
	if (gran) {
		cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran);
		size_t got = 0;
		cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
		printf("{\"set_l2_fetch_granularity\": %d, \"result\": \"%s\", \"now\": %zu}\n", gran, cudaGetErrorString(e), got);
	} else {
		size_t got = 0;
		cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
		printf("{\"l2_fetch_granularity_default\": %zu, \"buffer_gib\": %llu}\n", got, (unsigned long long)gib);
	}
	const uint64_t bytes = gib << 30;
	uint4 *buf;
	unsigned long long *sink;
	if (cudaMalloc(&buf, bytes) != cudaSuccess || cudaMalloc(&sink, 8) != cudaSuccess) { printf("alloc failed\n"); return 1; }
	cudaMemset(buf, 1, bytes);
	cudaMemset(sink, 0, 8);
	const uint64_t n_sectors = bytes / 32;
	if (argc > 3 && argv[3][0] == 's') {              // sweep: loads in flight per thread x CTAs per SM, for the two flavours that matter
		sweep<4>("ld.global.nc.L2::64B 16B", buf, n_sectors, sink);
		sweep<8>("ld.global.nc (__ldg) 4B", buf, n_sectors, sink);
		return 0;
	}
	run<0>("ld.global.nc (__ldg) 16B", buf, n_sectors, sink);
	if (argc > 3) {                                   // quick: only the two flavours that matter
		run<4>("ld.global.nc.L2::64B 16B", buf, n_sectors, sink);
		return 0;
	}
	run<1>("ld.global.cg 16B", buf, n_sectors, sink);
	run<2>("ld.global.cv 16B", buf, n_sectors, sink);
	run<3>("ld.global.nc.L1::no_allocate 16B", buf, n_sectors, sink);
	run<4>("ld.global.nc.L2::64B 16B", buf, n_sectors, sink);
	run<5>("ld.global.nc.L2::128B 16B", buf, n_sectors, sink);
	run<6>("ld.global.ca 16B", buf, n_sectors, sink);
	run<7>("ld.global.L1::no_allocate 16B", buf, n_sectors, sink);
	run<8>("ld.global.nc (__ldg) 4B", buf, n_sectors, sink);
	return 0;
}
