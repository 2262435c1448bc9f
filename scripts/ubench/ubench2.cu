// ubench2.cu -- second round of micro-benchmarks (B200): (1) which lane-group layouts let LDS.64 / LDS.128 run in the fast
// mode (groups of lanes reading the same record: sizes aligned to 4 / 2 lanes or not); (2) IDP.4A, I2F, F2F pack and
// predicated-STS issue rates.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

struct Pattern { int gid[32]; };

template <int WB, int MODE>      // MODE 0: all lanes of a group read the same record; 1: lanes of a group read consecutive records (lane-in-group * WB)
__global__ void __launch_bounds__(1024) k_lds(float* out, int iters, Pattern pat, int gstride_bytes, long long* cyc) {
	extern __shared__ __align__(16) unsigned char sm[];
	for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = (float) i;
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const int g = pat.gid[lane];
	int first = lane; while (first > 0 && pat.gid[first - 1] == g) first--;
	uint32_t base = (uint32_t) __cvta_generic_to_shared(sm) + (uint32_t) g * gstride_bytes + (threadIdx.x >> 5) * 64 + (MODE == 1 ? (lane - first) * WB : 0);
	float acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
		#pragma unroll
		for (int u = 0; u < 8; u++) {
			const uint32_t ad = base + (uint32_t) u * WB * (MODE == 1 ? 16 : 1);
			if (WB == 4) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(ad) : "memory"); acc0 += v; }
			if (WB == 8) { float v, w; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v), "=f"(w) : "r"(ad) : "memory"); acc0 += v; acc1 += w; }
			if (WB == 16) { float v, w, x, y; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v), "=f"(w), "=f"(x), "=f"(y) : "r"(ad) : "memory"); acc0 += v; acc1 += w; acc2 += x; acc3 += y; }
		}
		base += 16; if ((it & 15) == 15) base -= 16 * 16;
	}
	long long t1 = clock64();
	float s = acc0 + acc1 + acc2 + acc3;
	if (s == 12345.678f) out[0] = s;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
__global__ void __launch_bounds__(1024) k_pipe(float* out, int iters, long long* cyc) {
	constexpr int NCH = 8;
	unsigned a[NCH], b = 0x01020304u + iters, c[NCH]; float f[NCH]; unsigned h[NCH];
	for (int i = 0; i < NCH; i++) { a[i] = threadIdx.x * 0x01010101u + i; c[i] = i; f[i] = threadIdx.x + i * 0.5f; h[i] = 0; }
	__syncthreads();
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
		#pragma unroll
		for (int i = 0; i < NCH; i++) {
			if (MODE == 0) asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(c[i]) : "r"(a[i]), "r"(b));
			if (MODE == 1) { asm volatile("dp4a.s32.s32 %0, %1, %1, %0;" : "+r"(c[i]) : "r"(a[i])); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c[(i + 1) % NCH])); asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(b)); }
			if (MODE == 2) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[i]) : "r"(a[i] + (unsigned) f[(i + 1) % NCH]));
			if (MODE == 3) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(f[i] + (float) h[(i + 1) % NCH]), "f"(f[(i + 2) % NCH]));
			if (MODE == 4) asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(a[i]) : "f"(f[i] + (float) a[(i + 1) % NCH]));
			if (MODE == 5) { asm volatile("{.reg .pred p; setp.lt.f32 p, %1, %2; @p add.u32 %0, %0, 256;}" : "+r"(a[i]) : "f"(f[i]), "f"(f[(i + 1) % NCH])); }
		}
	}
	long long t1 = clock64();
	float s = 0; for (int i = 0; i < NCH; i++) s += (float) a[i] + (float) c[i] + f[i] + (float) h[i];
	if (s == 12345.678f) out[0] = s;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static long long maxcyc(long long* d, int n) { static long long h[4096]; cudaMemcpy(h, d, n * sizeof(long long), cudaMemcpyDeviceToHost); long long m = 0; for (int i = 0; i < n; i++) if (h[i] > m) m = h[i]; return m; }

static Pattern make(const int* sizes, int n) { Pattern p; int l = 0; for (int g = 0; g < n; g++) for (int k = 0; k < sizes[g] && l < 32; k++) p.gid[l++] = g; while (l < 32) p.gid[l++] = n - 1; return p; }

int main() {
	cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
	const int sms = p.multiProcessorCount;
	float* out; long long* cyc; CK(cudaMalloc(&out, 1024)); CK(cudaMalloc(&cyc, 4096 * 8));
	const int pats[][8] = {{11, 10, 11}, {12, 8, 12}, {10, 12, 10}, {12, 12, 8}, {16, 16}, {20, 12}, {8, 8, 8, 8}, {6, 10, 6, 10}, {9, 14, 9}, {4, 12, 12, 4}, {2, 12, 12, 6}, {1, 11, 11, 9}, {3, 11, 11, 7}, {11, 21}, {13, 19}, {14, 18}};
	const int npat[] = {3, 3, 3, 3, 2, 2, 4, 4, 3, 4, 4, 4, 4, 2, 2, 2};
	CK(cudaFuncSetAttribute(k_lds<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
	CK(cudaFuncSetAttribute(k_lds<8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
	CK(cudaFuncSetAttribute(k_lds<16, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
	CK(cudaFuncSetAttribute(k_lds<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
	CK(cudaFuncSetAttribute(k_lds<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
	for (int mode = 0; mode < 2; mode++) for (int wb = (mode ? 8 : 4); wb <= 16; wb *= 2) for (int pi = 0; pi < 16; pi++) for (int stride = 176; stride <= 208; stride += 32) {
		const int iters = 2048, warps = 16;
		Pattern pt = make(pats[pi], npat[pi]);
		void (*k)(float*, int, Pattern, int, long long*) = mode == 0 ? (wb == 4 ? k_lds<4, 0> : wb == 8 ? k_lds<8, 0> : k_lds<16, 0>) : (wb == 8 ? k_lds<8, 1> : k_lds<16, 1>);
		k<<<sms, warps * 32, 64 * 1024>>>(out, 16, pt, stride, cyc); CK(cudaDeviceSynchronize());
		k<<<sms, warps * 32, 64 * 1024>>>(out, iters, pt, stride, cyc); CK(cudaDeviceSynchronize());
		const long long c = maxcyc(cyc, sms);
		char buf[64]; int o = 0; for (int g = 0; g < npat[pi]; g++) o += snprintf(buf + o, 64 - o, "%d%s", pats[pi][g], g + 1 < npat[pi] ? "," : "");
		printf("lds   %s  width=%2d B  group stride=%4d B  groups=(%s)  %.2f clk per LDS per SM\n", mode ? "consecutive-in-group" : "broadcast-in-group  ", wb, stride, buf, (double) c / ((double) iters * 8 * warps));
	}
	const char* names[] = {"IDP.4A", "IDP.4A+IADD+LOP (3 instr)", "I2F", "F2F.F16x2 pack", "F2I", "FSETP+@p IADD (2 instr)"};
	for (int m = 0; m < 6; m++) {
		const int iters = 4096, warps = 16;
		void (*k)(float*, int, long long*) = m == 0 ? k_pipe<0> : m == 1 ? k_pipe<1> : m == 2 ? k_pipe<2> : m == 3 ? k_pipe<3> : m == 4 ? k_pipe<4> : k_pipe<5>;
		k<<<sms, warps * 32>>>(out, 16, cyc); CK(cudaDeviceSynchronize());
		k<<<sms, warps * 32>>>(out, iters, cyc); CK(cudaDeviceSynchronize());
		const long long c = maxcyc(cyc, sms);
		printf("pipe  %-28s %.3f clk per group per SMSP\n", names[m], (double) c / (iters * 8 * warps / 4.0));
	}
	return 0;
}
