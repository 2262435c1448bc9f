// tc_probe.cu -- stand-alone check of the tcgen05 building blocks used by force_1clj_tc.cu: shared-memory operand
// descriptors (K-major, no swizzle), the kind::f16 instruction descriptor, TMEM allocation, commit -> mbarrier, and
// tcgen05.ld 32x32b. Computes D[128][N] = A[128][16] * B[N][16]^T (fp16 in, fp32 out) and compares with the host.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
	return (uint64_t) ((addr >> 4) & 0x3FFFu) | ((uint64_t) ((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t) ((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// canonical K-major no-swizzle layout: (row, k) -> (row/8)*256 + (k/8)*128 + (row%8)*16 + (k%8)*2 bytes
__device__ __forceinline__ uint32_t op_off(int row, int k) { return (uint32_t) ((row >> 3) * 256 + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2); }

template <int N>
__global__ void __launch_bounds__(128) k_probe(const __half* A, const __half* B, float* D) {
	extern __shared__ __align__(128) unsigned char sm[];
	unsigned char* sa = sm;                  // 128 rows * 32 B
	unsigned char* sb = sm + 4096;           // N rows * 32 B
	__shared__ __align__(8) unsigned long long bar;
	__shared__ uint32_t tmem_base_s;
	const int tid = threadIdx.x, wid = tid >> 5;
	for (int i = tid; i < 128 * 16; i += 128) *(__half*) (sa + op_off(i / 16, i % 16)) = A[i];
	for (int i = tid; i < N * 16; i += 128) *(__half*) (sb + op_off(i / 16, i % 16)) = B[i];
	if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); }
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	if (wid == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(smem_u32(&tmem_base_s)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = tmem_base_s;
	if (tid == 0) {
		const uint64_t da = make_desc(smem_u32(sa), 128, 256), db = make_desc(smem_u32(sb), 128, 256);
		// idesc: c_format F32 (1 << 4), a/b format F16 (0), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
		const uint32_t idesc = (1u << 4) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
		             :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
	}
	{
		uint32_t ok; const uint32_t b = smem_u32(&bar);
		do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0u) : "memory"); } while (!ok);
	}
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	for (int c0 = 0; c0 < N; c0 += 32) {
		uint32_t v[32];
		const uint32_t taddr = tmem + ((uint32_t) (32 * wid) << 16) + (uint32_t) c0;
		asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
		             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
		               "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
		             : "r"(taddr));
		asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
		for (int k = 0; k < 32; k++) D[(size_t) tid * N + c0 + k] = __uint_as_float(v[k]);
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tmem) : "memory");
}

int main() {
	constexpr int N = 128;
	__half hA[128 * 16], hB[N * 16];
	srand(1);
	for (int i = 0; i < 128 * 16; i++) hA[i] = __float2half((float) (rand() % 65 - 32) / 8.f);
	for (int i = 0; i < N * 16; i++) hB[i] = __float2half((float) (rand() % 65 - 32) / 8.f);
	__half *dA, *dB; float* dD;
	cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(float) * 128 * N);
	cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
	cudaMemset(dD, 0xFF, sizeof(float) * 128 * N);
	const int smem = 4096 + N * 32;
	k_probe<N><<<1, 128, smem>>>(dA, dB, dD);
	cudaError_t e = cudaDeviceSynchronize();
	printf("kernel: %s\n", cudaGetErrorString(e));
	static float hD[128 * N];
	cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
	int bad = 0; double maxerr = 0;
	for (int i = 0; i < 128; i++) for (int j = 0; j < N; j++) {
		float ref = 0; for (int k = 0; k < 16; k++) ref += __half2float(hA[i * 16 + k]) * __half2float(hB[j * 16 + k]);
		const double err = fabs((double) ref - hD[i * N + j]);
		if (err > maxerr) maxerr = err;
		if (err > 1e-3 && bad++ < 8) printf("mismatch D[%d][%d] = %g, expected %g\n", i, j, hD[i * N + j], ref);
	}
	printf("tc_probe: %d mismatches of %d, max |err| = %g\n", bad, 128 * N, maxerr);
	return bad != 0;
}
