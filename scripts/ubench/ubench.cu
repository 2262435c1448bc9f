// ubench.cu -- micro-benchmarks behind the design of the round-2 1CLJ force kernel (B200, sm_100a):
//   1. issue / pipe throughput of FFMA, FFMA2 (packed f32x2), FADD, HFMA2, IADD3 and mixes
//   2. shared-memory wavefronts of LDS.32 / .64 / .128 when the lanes of a warp read k distinct addresses (k lane groups)
//   3. cost of many small cp.async.bulk (TMA) copies against few large ones (same bytes)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu ; run on the GPU box, prints one line per test.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

constexpr int NCH = 8;      // independent dependency chains per thread

template <int MODE>
__global__ void __launch_bounds__(1024) k_pipe(float* out, int iters, long long* cyc) {
	float a[NCH], b = 1.0001f + iters * 1e-9f, c = 0.5f + iters * 1e-9f;
	unsigned long long A2[NCH]; unsigned ia[NCH]; unsigned h[NCH];
	for (int i = 0; i < NCH; i++) { a[i] = threadIdx.x * 1e-3f + i; float2 t = make_float2(a[i], a[i] + 1.f); A2[i] = *reinterpret_cast<unsigned long long*>(&t); ia[i] = threadIdx.x + i; h[i] = 0x3c003c00u + i; }
	float2 bb = make_float2(b, b + 1e-7f * iters), cc = make_float2(c, c + 1e-7f * iters);
	unsigned long long B2 = *reinterpret_cast<unsigned long long*>(&bb), C2 = *reinterpret_cast<unsigned long long*>(&cc);
	unsigned hb = 0x3c013c01u + (iters & 1), hc = 0x38003800u + (iters & 1);
	__syncthreads();
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
		#pragma unroll
		for (int i = 0; i < NCH; i++) {
			if (MODE == 0 || MODE == 5 || MODE == 7) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
			if (MODE == 1 || MODE == 6 || MODE == 7 || MODE == 9) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A2[i]) : "l"(B2), "l"(C2));
			if (MODE == 2) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
			if (MODE == 3 || MODE == 8) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h[i]) : "r"(hb), "r"(hc));
			if (MODE == 8) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
			if (MODE == 4 || MODE == 5 || MODE == 6 || MODE == 9) asm volatile("add.u32 %0, %0, %1;" : "+r"(ia[i]) : "r"(ia[(i + 1) % NCH]));
			if (MODE == 9) asm volatile("xor.b32 %0, %0, %1;" : "+r"(ia[i]) : "r"(ia[(i + 3) % NCH]));
		}
	}
	long long t1 = clock64();
	float s = 0; for (int i = 0; i < NCH; i++) { float2 t = *reinterpret_cast<float2*>(&A2[i]); s += a[i] + t.x + t.y + (float) ia[i] + (float) h[i]; }
	if (s == 12345.678f) out[0] = s;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// LDS: lanes of a warp form k contiguous groups; group g reads record (g * gstride + it) of W-byte records
template <int WB>
__global__ void __launch_bounds__(1024) k_lds(float* out, int iters, int k, int gstride_bytes, long long* cyc) {
	extern __shared__ __align__(16) unsigned char sm[];
	for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = (float) i;
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const int g = lane * k / 32;
	uint32_t base = (uint32_t) __cvta_generic_to_shared(sm) + (uint32_t) g * gstride_bytes + (threadIdx.x >> 5) * 64;
	float acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
		#pragma unroll
		for (int u = 0; u < 8; u++) {
			const uint32_t ad = base + (uint32_t) u * WB;
			if (WB == 4) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(ad)); acc0 += v; }
			if (WB == 8) { float v, w; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v), "=f"(w) : "r"(ad)); acc0 += v; acc1 += w; }
			if (WB == 16) { float v, w, x, y; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v), "=f"(w), "=f"(x), "=f"(y) : "r"(ad)); acc0 += v; acc1 += w; acc2 += x; acc3 += y; }
		}
		base += 8 * WB; if ((it & 15) == 15) base -= 16 * 8 * WB;
	}
	long long t1 = clock64();
	float s = acc0 + acc1 + acc2 + acc3;
	if (s == 12345.678f) out[0] = s;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// STS.U16 predicated: fraction of lanes active
__global__ void __launch_bounds__(1024) k_sts(float* out, int iters, int active_mod, long long* cyc) {
	extern __shared__ __align__(16) unsigned char sm[];
	uint32_t base = (uint32_t) __cvta_generic_to_shared(sm) + threadIdx.x * 2;
	const bool on = (threadIdx.x % active_mod) == 0;
	__syncthreads();
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
		#pragma unroll
		for (int u = 0; u < 8; u++) {
			if (on) asm volatile("st.shared.u16 [%0], %1;" :: "r"(base + u * 2048), "h"((unsigned short) it));
		}
	}
	long long t1 = clock64();
	__syncthreads();
	if (sm[threadIdx.x] == 77 && iters < 0) out[0] = 1.f;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// TMA: ncopy bulk copies of bytes each per iteration, issued by the first `issuers` threads; wait on the mbarrier
__global__ void __launch_bounds__(128) k_tma(const unsigned char* src, float* out, int iters, int ncopy, int bytes, int issuers, long long* cyc) {
	extern __shared__ __align__(16) unsigned char sm[];
	__shared__ __align__(8) unsigned long long bar;
	const uint32_t bar_s = (uint32_t) __cvta_generic_to_shared(&bar), sm_s = (uint32_t) __cvta_generic_to_shared(sm);
	if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_s));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncthreads();
	uint32_t parity = 0;
	const unsigned char* s0 = src + (size_t) blockIdx.x * 65536 * 4;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
		if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_s), "r"(ncopy * bytes) : "memory");
		__syncthreads();
		for (int c = threadIdx.x; c < ncopy; c += issuers) {
			if (threadIdx.x < issuers)
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				             :: "r"(sm_s + (uint32_t) c * bytes), "l"(s0 + ((size_t) ((c * 7 + it * 13) % 997) * 256)), "r"(bytes), "r"(bar_s) : "memory");
		}
		uint32_t ok;
		do {
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar_s), "r"(parity) : "memory");
		} while (!ok);
		parity ^= 1;
		__syncthreads();
	}
	long long t1 = clock64();
	if (sm[threadIdx.x] == 77 && iters < 0) out[0] = 1.f;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static long long maxcyc(long long* d, int n) { static long long h[4096]; cudaMemcpy(h, d, n * sizeof(long long), cudaMemcpyDeviceToHost); long long m = 0; for (int i = 0; i < n; i++) if (h[i] > m) m = h[i]; return m; }

int main() {
	int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
	const int sms = p.multiProcessorCount;
	printf("device %s, %d SMs\n", p.name, sms);
	float* out; long long* cyc; CK(cudaMalloc(&out, 1024)); CK(cudaMalloc(&cyc, 4096 * 8));
	const char* names[] = {"FFMA", "FFMA2", "FADD", "HFMA2", "IADD", "FFMA+IADD 1:1", "FFMA2+IADD 1:1", "FFMA2+FFMA 1:1", "HFMA2+FFMA 1:1", "FFMA2+IADD+XOR 1:2"};
	const int per[] = {1, 1, 1, 1, 1, 2, 2, 2, 2, 3};
	for (int warps = 16; warps <= 32; warps += 16) {
		for (int m = 0; m < 10; m++) {
			const int iters = 4096;
			void (*k)(float*, int, long long*) = nullptr;
			switch (m) { case 0: k = k_pipe<0>; break; case 1: k = k_pipe<1>; break; case 2: k = k_pipe<2>; break; case 3: k = k_pipe<3>; break; case 4: k = k_pipe<4>; break;
			             case 5: k = k_pipe<5>; break; case 6: k = k_pipe<6>; break; case 7: k = k_pipe<7>; break; case 8: k = k_pipe<8>; break; case 9: k = k_pipe<9>; break; }
			k<<<sms, warps * 32>>>(out, 16, cyc); CK(cudaDeviceSynchronize());
			k<<<sms, warps * 32>>>(out, iters, cyc); CK(cudaDeviceSynchronize());
			const long long c = maxcyc(cyc, sms);
			const double winst = (double) iters * NCH * per[m] * warps;
			printf("pipe  warps/SM=%2d  %-20s  %.3f warp-instr/clk/SM  (%.3f clk per instr-group per SMSP)\n", warps, names[m], winst / c, (double) c / (iters * NCH * warps / 4.0));
		}
	}
	CK(cudaFuncSetAttribute(k_lds<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
	CK(cudaFuncSetAttribute(k_lds<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
	CK(cudaFuncSetAttribute(k_lds<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
	CK(cudaFuncSetAttribute(k_sts, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
	const int ks[] = {1, 2, 3, 4, 6, 8, 11, 16, 32};
	const int strides[] = {176, 208, 272};
	for (int wb = 4; wb <= 16; wb *= 2) for (int si = 0; si < 3; si++) for (int ki = 0; ki < 9; ki++) {
		const int iters = 2048, warps = 16;
		void (*k)(float*, int, int, int, long long*) = wb == 4 ? k_lds<4> : wb == 8 ? k_lds<8> : k_lds<16>;
		k<<<sms, warps * 32, 64 * 1024>>>(out, 16, ks[ki], strides[si], cyc); CK(cudaDeviceSynchronize());
		k<<<sms, warps * 32, 64 * 1024>>>(out, iters, ks[ki], strides[si], cyc); CK(cudaDeviceSynchronize());
		const long long c = maxcyc(cyc, sms);
		printf("lds   width=%2d B  group stride=%3d B  distinct=%2d  %.2f clk per LDS per SM\n", wb, strides[si], ks[ki], (double) c / ((double) iters * 8 * warps));
	}
	for (int am = 1; am <= 32; am *= 2) {
		const int iters = 2048, warps = 16;
		k_sts<<<sms, warps * 32, 64 * 1024>>>(out, 16, am, cyc); CK(cudaDeviceSynchronize());
		k_sts<<<sms, warps * 32, 64 * 1024>>>(out, iters, am, cyc); CK(cudaDeviceSynchronize());
		const long long c = maxcyc(cyc, sms);
		printf("sts16 one lane in %2d active  %.2f clk per STS per SM\n", am, (double) c / ((double) iters * 8 * warps));
	}
	unsigned char* src; CK(cudaMalloc(&src, (size_t) sms * 5 * 65536 * 4)); CK(cudaMemset(src, 1, (size_t) sms * 5 * 65536 * 4));
	CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
	const int cfgs[][3] = {{9, 2304, 1}, {9, 2304, 9}, {27, 768, 27}, {117, 176, 117}, {117, 176, 32}, {117, 176, 1}, {468, 48, 128}, {468, 48, 32}, {234, 96, 117}};
	for (int ctas = 1; ctas <= 4; ctas *= 4) for (auto& cf : cfgs) {
		const int iters = 2000;
		k_tma<<<sms * ctas, 128, 36 * 1024>>>(src, out, 10, cf[0], cf[1], cf[2], cyc); CK(cudaDeviceSynchronize());
		k_tma<<<sms * ctas, 128, 36 * 1024>>>(src, out, iters, cf[0], cf[1], cf[2], cyc); CK(cudaDeviceSynchronize());
		const long long c = maxcyc(cyc, sms * ctas);
		printf("tma   CTAs/SM=%d  %3d copies x %4d B, %3d issuing threads: %.0f clk per round (%.1f B/clk/SM)\n", ctas, cf[0], cf[1], cf[2], (double) c / iters, (double) cf[0] * cf[1] * ctas / ((double) c / iters));
	}
	return 0;
}
