// ptbroofs — the three machine roofs the closest-hit (BVH traversal) kernel is measured against ("measure with an FMA
// microbench ... measure [L2] with a resident-set read microbench — don't trust the estimate").  Measurement tooling, not product code:
// bench.py runs these on the GPU it is about to time, before the timed region, and reports roofline.frac against what THEY measured.
//
//   ptbroofs_fp32     FP32 FMA throughput of the non-tensor pipes at the clock the GPU holds under that load (TFLOP/s)
//   ptbroofs_gather   random gather bandwidth: every lane reads `bytes_per_lane` (32 / 64 / 128) contiguous, aligned bytes at an independent
//                     pseudo-random offset inside a working set, with the traversal kernel's launch shape (128-thread blocks, a fixed number
//                     of resident blocks per SM) and its load instructions (ld.global.nc, 256-bit where the record allows):
//                       working set ~ the tree (10 MB): every SM's L1 misses, the 126 MB L2 serves it      -> the L2 gather roof
//                       working set 64 KB:              every SM's L1 holds it                              -> the L1 gather roof
//                     (GB/s of requested bytes).
// Built by pathtracerwithcuda_b200/build.py into tools/roofs/libptbroofs.so for sm_100a.
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>

namespace
{

__global__ void __launch_bounds__(256) k_fma(float* out, int iters, float a, float b)
{
	// 8 independent chains per thread: FFMA latency 4 cycles x 2 issue slots -> 8 in flight per warp keep the pipe full at any occupancy
	float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.0f, x2 = x0 + 2.0f, x3 = x0 + 3.0f, x4 = x0 + 4.0f, x5 = x0 + 5.0f, x6 = x0 + 6.0f, x7 = x0 + 7.0f;
	for (int i = 0; i < iters; i++)
	{
#pragma unroll
		for (int k = 0; k < 16; k++)
		{
			x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
			x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
		}
	}
	const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
	if (s == 12345.678f) out[0] = s;   // never true: keeps the chains alive
}

__device__ __forceinline__ void ld256(const void* p, float (&v)[8])
{
	asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}

// BYTES per lane per gather: 32 (one 256-bit load), 64 (two: a binary BVH node), 128 (four: a full line)
template <int BYTES>
__global__ void __launch_bounds__(128, 8) k_gather(const char* __restrict__ base, unsigned records, int iters, float* out)
{
	unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
	float acc = 0.0f;
	for (int i = 0; i < iters; i++)
	{
		// four independent gathers in flight per lane (the traversal kernel has one per lane; the roof is what the memory system can
		// deliver for this access shape, so latency is taken out of the picture)
		unsigned r[4];
#pragma unroll
		for (int k = 0; k < 4; k++)
		{
			s = s * 1664525u + 1013904223u;
			r[k] = (unsigned)(((unsigned long long)(s >> 4) * records) >> 28);   // uniform in [0, records)
		}
#pragma unroll
		for (int k = 0; k < 4; k++)
		{
			const char* p = base + (size_t)r[k] * BYTES;
#pragma unroll
			for (int q = 0; q < BYTES / 32; q++)
			{
				float v[8];
				ld256(p + q * 32, v);
				acc += v[0] + v[7];
			}
		}
	}
	if (acc == 12345.678f) out[0] = acc;
}

// L1 -> register-file write-back roof: every warp streams 256-bit loads (32 lanes x 32 B = 8 full lines per instruction) over a 32 KB buffer
// that stays in its SM's L1.  This is what a load INSTRUCTION costs the L1 data pipe whatever its addresses are — the resource ncu
// reports as l1tex__lsu_writeback_active / l1tex__data_pipe_lsu_wavefronts for the traversal kernels.
__global__ void __launch_bounds__(128, 8) k_l1_stream(const char* __restrict__ base, int iters, float* out)
{
	const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	float acc = 0.0f;
	for (int i = 0; i < iters; i++)
	{
#pragma unroll
		for (int k = 0; k < 8; k++)
		{
			// 32 KB = 32 chunks of 1 KB; every warp walks them from its own phase
			const unsigned chunk = (unsigned)(i * 8 + k + warp * 5 + blockIdx.x) & 31u;
			float v[8];
			ld256(base + chunk * 1024u + lane * 32u, v);
			acc += v[0] + v[7];
		}
	}
	if (acc == 12345.678f) out[0] = acc;
}

int run_gather(int bytes_per_lane, const char* base, unsigned records, int iters, int grid, cudaStream_t st)
{
	switch (bytes_per_lane)
	{
	case 32: k_gather<32><<<grid, 128, 0, st>>>(base, records, iters, (float*)base); break;
	case 64: k_gather<64><<<grid, 128, 0, st>>>(base, records, iters, (float*)base); break;
	case 128: k_gather<128><<<grid, 128, 0, st>>>(base, records, iters, (float*)base); break;
	default: return 1;
	}
	return 0;
}

} // namespace

extern "C"
{

// FP32 FMA roof.  out[0] = TFLOP/s (2 flop per FMA), out[1] = ms of the timed launch, out[2] = SM count
int ptbroofs_fp32(int device, double* out)
{
	if (cudaSetDevice(device) != cudaSuccess) return 1;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 1;
	float* d = nullptr;
	if (cudaMalloc(&d, 256) != cudaSuccess) return 1;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int grid = prop.multiProcessorCount * 8, iters = 4096;
	double best = 0.0, best_ms = 0.0;
	for (int rep = 0; rep < 5; rep++)
	{
		cudaEventRecord(e0);
		k_fma<<<grid, 256>>>(d, iters, 1.000001f, 1e-7f);
		cudaEventRecord(e1);
		if (cudaEventSynchronize(e1) != cudaSuccess) return 1;
		float ms = 0.0f;
		cudaEventElapsedTime(&ms, e0, e1);
		const double flops = 2.0 * 8 * 16 * (double)iters * 256.0 * grid;
		if (rep > 0 && flops / (ms * 1e-3) / 1e12 > best) { best = flops / (ms * 1e-3) / 1e12; best_ms = ms; }
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
	out[0] = best; out[1] = best_ms; out[2] = prop.multiProcessorCount;
	return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// Random-gather roof.  working_set_bytes: size of the region gathered from; bytes_per_lane: 32 | 64 | 128; blocks_per_sm: resident
// 128-thread blocks per SM (8 = the traversal kernels' launch bounds).  out[0] = GB/s requested, out[1] = ms, out[2] = bytes moved
int ptbroofs_gather(int device, long long working_set_bytes, int bytes_per_lane, int blocks_per_sm, double* out)
{
	if (cudaSetDevice(device) != cudaSuccess) return 1;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 1;
	if (working_set_bytes < 4096 || (bytes_per_lane != 32 && bytes_per_lane != 64 && bytes_per_lane != 128) || blocks_per_sm < 1 || blocks_per_sm > 16) return 1;
	char* d = nullptr;
	if (cudaMalloc(&d, (size_t)working_set_bytes) != cudaSuccess) return 1;
	cudaMemset(d, 0, (size_t)working_set_bytes);
	const unsigned records = (unsigned)(working_set_bytes / bytes_per_lane);
	const int grid = prop.multiProcessorCount * blocks_per_sm;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	// size the run for roughly a millisecond or more
	int iters = working_set_bytes <= (1 << 20) ? 2048 : 256;
	double best = 0.0, best_ms = 0.0, moved = 0.0;
	for (int rep = 0; rep < 5; rep++)
	{
		cudaEventRecord(e0);
		if (run_gather(bytes_per_lane, d, records, iters, grid, 0)) return 1;
		cudaEventRecord(e1);
		if (cudaEventSynchronize(e1) != cudaSuccess) return 1;
		float ms = 0.0f;
		cudaEventElapsedTime(&ms, e0, e1);
		const double bytes = (double)grid * 128.0 * iters * 4.0 * bytes_per_lane;
		if (rep > 0 && bytes / (ms * 1e-3) / 1e9 > best) { best = bytes / (ms * 1e-3) / 1e9; best_ms = ms; moved = bytes; }
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
	out[0] = best; out[1] = best_ms; out[2] = moved;
	return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// L1 write-back roof (see k_l1_stream).  out[0] = GB/s delivered to registers, out[1] = ms
int ptbroofs_l1_stream(int device, int blocks_per_sm, double* out)
{
	if (cudaSetDevice(device) != cudaSuccess) return 1;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 1;
	if (blocks_per_sm < 1 || blocks_per_sm > 16) return 1;
	char* d = nullptr;
	if (cudaMalloc(&d, 32 << 10) != cudaSuccess) return 1;
	cudaMemset(d, 0, 32 << 10);
	const int grid = prop.multiProcessorCount * blocks_per_sm, iters = 4096;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	double best = 0.0, best_ms = 0.0;
	for (int rep = 0; rep < 5; rep++)
	{
		cudaEventRecord(e0);
		k_l1_stream<<<grid, 128>>>(d, iters, (float*)d);
		cudaEventRecord(e1);
		if (cudaEventSynchronize(e1) != cudaSuccess) return 1;
		float ms = 0.0f;
		cudaEventElapsedTime(&ms, e0, e1);
		const double bytes = (double)grid * 128.0 * iters * 8.0 * 32.0;
		if (rep > 0 && bytes / (ms * 1e-3) / 1e9 > best) { best = bytes / (ms * 1e-3) / 1e9; best_ms = ms; }
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
	out[0] = best; out[1] = best_ms;
	return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

} // extern "C"
