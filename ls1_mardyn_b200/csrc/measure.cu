// measure.cu -- measurement helpers of include/ls1b200.h: the reference's FlopCounter counts for the current state
// (adapter/FlopCounter.cpp:20-74, FlopCounter.h:263-270) and two device micro-benchmarks that give the roofline
// denominators MEASURED_PEAKS.json does not hold (FP32 FMA rate) or that bench.py wants to re-check (HBM copy).
#include "pairs.cuh"
#include "types.cuh"

namespace ls1b200 {

// Counts what the reference's C08 traversal evaluates (Newton-3 counting, each cell pair once, halo rules of
// VectorizedCellProcessor::processCellPair): an owned-owned molecule pair once (slot i < slot j), an owned-halo pair once
// from its owned side (the mirrored computation is the other periodic image and is counted at its own owned side).
// out[0] = molecule distance evaluations, out[1+p] = site pairs of potential p, out[8+p] = macroscopic pairs.
template <class C>
__global__ void __launch_bounds__(128) k_count_flops(const typename Vec4<C>::type* __restrict__ mpos, const uint32_t* __restrict__ cell_start,
                                                      const uint32_t* __restrict__ cell_of, const CompDev* __restrict__ comps, uint32_t n_slots,
                                                      int nx, int ny, C rc2, C ljrc2, unsigned long long* __restrict__ out) {
	typedef typename Vec4<C>::type V4;
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long cnt[15];
	for (int k = 0; k < 15; k++) cnt[k] = 0;
	if (t < n_slots) {
		const V4 mi = mpos[t];
		const uint32_t bits = wbits<C>(mi.w);
		if (!(bits & 0x80000000u)) {
			const CompDev& ci = comps[bits & 0xFFFFu];
			const uint32_t c = cell_of[t];
			const int cx = c % nx, cy = (c / nx) % ny, cz = c / (nx * ny);
			const uint32_t cb = cell_start[c], ce = cell_start[c + 1];
			for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
				const uint32_t cj = (uint32_t) (((cz + dz) * ny + (cy + dy)) * nx + (cx + dx));
				const uint32_t jb = cell_start[cj], je = cell_start[cj + 1];
				for (uint32_t j = jb; j < je; j++) {
					const V4 mj = mpos[j];
					const uint32_t bj = wbits<C>(mj.w);
					const bool halo_j = (bj & 0x80000000u) != 0;
					if (!halo_j && j <= t) continue;
					cnt[0]++;
					const C r2 = molpair_r2<C>(mi.x, mi.y, mi.z, mj.x, mj.y, mj.z);
					if (j >= cb && j < ce && !(r2 != C(0))) continue;
					const bool macro = !halo_j || c < cj;
					const CompDev& q = comps[bj & 0xFFFFu];
					unsigned long long np[7] = {0, 0, 0, 0, 0, 0, 0};
					if (r2 < ljrc2) np[0] = (unsigned long long) ci.n[0] * q.n[0];
					if (r2 < rc2) {
						np[1] = (unsigned long long) ci.n[1] * q.n[1];
						np[2] = (unsigned long long) ci.n[1] * q.n[2] + (unsigned long long) ci.n[2] * q.n[1];
						np[3] = (unsigned long long) ci.n[2] * q.n[2];
						np[4] = (unsigned long long) ci.n[1] * q.n[3] + (unsigned long long) ci.n[3] * q.n[1];
						np[5] = (unsigned long long) ci.n[2] * q.n[3] + (unsigned long long) ci.n[3] * q.n[2];
						np[6] = (unsigned long long) ci.n[3] * q.n[3];
					}
					for (int p = 0; p < 7; p++) { cnt[1 + p] += np[p]; if (macro) cnt[8 + p] += np[p]; }
				}
			}
		}
	}
	for (int k = 0; k < 15; k++) {
		unsigned long long v = cnt[k];
		for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
		if ((threadIdx.x & 31) == 0 && v) atomicAdd(&out[k], v);
	}
}

// FP32 FMA peak: 8 independent chains per thread, enough resident warps to saturate all four sub-partitions
__global__ void __launch_bounds__(256) k_fma_peak(float* out, int iters, float a, float b) {
	float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
	for (int i = 0; i < iters; i++) {
		#pragma unroll
		for (int u = 0; u < 8; u++) {
			x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
			x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
		}
	}
	const float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
	if (s == 123.456f) out[0] = s;
}

// FP64 FMA peak (the roofline denominator of the double build): same structure with DFMA
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b) {
	double x0 = threadIdx.x * 1e-3, x1 = x0 + 1., x2 = x0 + 2., x3 = x0 + 3., x4 = x0 + 4., x5 = x0 + 5., x6 = x0 + 6., x7 = x0 + 7.;
	for (int i = 0; i < iters; i++) {
		#pragma unroll
		for (int u = 0; u < 8; u++) {
			x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
			x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
		}
	}
	const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
	if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) k_copy16(const int4* __restrict__ src, int4* __restrict__ dst, size_t n) {
	const size_t stride = (size_t) gridDim.x * blockDim.x;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

}  // namespace ls1b200

using namespace ls1b200;

extern "C" {

int ls1b200_count_flops(ls1b200_ctx* c, ls1b200_flopcount* out) {
	if (!c || !out) return fail("ls1b200_count_flops: null argument");
	if (!c->sorted) return fail("ls1b200_count_flops: call ls1b200_update first");
	LS1_CUDA(cudaSetDevice(c->device));
	memset(out, 0, sizeof(*out));
	if (c->n_owned == 0 && !c->dd.active) return 0;
	LS1_CUDA(cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(DevScalars), cudaMemcpyDeviceToHost, c->stream));
	LS1_CUDA(cudaStreamSynchronize(c->stream));
	const uint32_t ns = (uint32_t) std::min<uint64_t>(c->h_scal->n_slots2[c->parity], c->cap);
	unsigned long long* d = nullptr;
	LS1_CUDA(cudaMalloc((void**) &d, 15 * sizeof(unsigned long long)));
	LS1_CUDA(cudaMemsetAsync(d, 0, 15 * sizeof(unsigned long long), c->stream));
	const double rc = c->cfg.cutoff, rclj = c->cfg.cutoff_lj;
	const uint32_t g = (ns + 127) / 128;
	if (c->mixed)
		k_count_flops<float><<<g, 128, 0, c->stream>>>((const float4*) c->mpos, c->cell_start, c->cell_of2, c->d_comps, ns, c->geom.cells[0], c->geom.cells[1],
		                                              (float) (rc * rc), (float) (rclj * rclj), d);
	else
		k_count_flops<double><<<g, 128, 0, c->stream>>>((const double4*) c->mpos, c->cell_start, c->cell_of2, c->d_comps, ns, c->geom.cells[0], c->geom.cells[1],
		                                               rc * rc, rclj * rclj, d);
	g_launches++;
	unsigned long long h[15];
	cudaError_t e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	cudaFree(d);
	LS1_CUDA(e);
	double* k = &out->lj; double* m = &out->lj_macro;
	out->mol_dist = (double) h[0];
	for (int p = 0; p < 7; p++) { k[p] = (double) h[1 + p]; m[p] = (double) h[8 + p]; }
	// FlopCounter.cpp:27-74: per potential (kernel flops, macroscopic flops, sum-forces flops, sum-macroscopic flops); 8 = distance
	static const double kk[7] = {15, 10, 34, 101, 52, 121, 131}, mm[7] = {4, 2, 3, 5, 2, 2, 2}, ss[7] = {12, 12, 15, 18, 15, 18, 18}, sm[7] = {2, 2, 2, 3, 2, 2, 2};
	out->flops = 8. * out->mol_dist;
	for (int p = 0; p < 7; p++) out->flops += (8. + kk[p] + ss[p]) * k[p] + (mm[p] + sm[p]) * m[p];
	return 0;
}

int ls1b200_measure_fp32_peak(int device, double* tflops) {
	if (!tflops) return fail("ls1b200_measure_fp32_peak: null argument");
	LS1_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop; LS1_CUDA(cudaGetDeviceProperties(&prop, device));
	float* d = nullptr; LS1_CUDA(cudaMalloc((void**) &d, 64));
	cudaEvent_t a, b; LS1_CUDA(cudaEventCreate(&a)); LS1_CUDA(cudaEventCreate(&b));
	const int blocks = prop.multiProcessorCount * 8, iters = 4096;
	double best = 0.;
	for (int rep = 0; rep < 6; rep++) {
		LS1_CUDA(cudaEventRecord(a));
		k_fma_peak<<<blocks, 256>>>(d, iters, 0.999f, 1e-3f);
		g_launches++;
		LS1_CUDA(cudaEventRecord(b));
		LS1_CUDA(cudaEventSynchronize(b));
		float ms = 0.f; LS1_CUDA(cudaEventElapsedTime(&ms, a, b));
		const double fl = 2.0 * 64.0 * (double) iters * 256.0 * (double) blocks;
		if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
	}
	cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
	*tflops = best;
	return 0;
}

int ls1b200_measure_fp64_peak(int device, double* tflops) {
	if (!tflops) return fail("ls1b200_measure_fp64_peak: null argument");
	LS1_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop; LS1_CUDA(cudaGetDeviceProperties(&prop, device));
	double* d = nullptr; LS1_CUDA(cudaMalloc((void**) &d, 64));
	cudaEvent_t a, b; LS1_CUDA(cudaEventCreate(&a)); LS1_CUDA(cudaEventCreate(&b));
	const int blocks = prop.multiProcessorCount * 8, iters = 1024;
	double best = 0.;
	for (int rep = 0; rep < 6; rep++) {
		LS1_CUDA(cudaEventRecord(a));
		k_dfma_peak<<<blocks, 256>>>(d, iters, 0.999, 1e-3);
		g_launches++;
		LS1_CUDA(cudaEventRecord(b));
		LS1_CUDA(cudaEventSynchronize(b));
		float ms = 0.f; LS1_CUDA(cudaEventElapsedTime(&ms, a, b));
		const double fl = 2.0 * 64.0 * (double) iters * 256.0 * (double) blocks;
		if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
	}
	cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
	*tflops = best;
	return 0;
}

int ls1b200_measure_hbm_copy(int device, double* gbs) {
	if (!gbs) return fail("ls1b200_measure_hbm_copy: null argument");
	LS1_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop; LS1_CUDA(cudaGetDeviceProperties(&prop, device));
	const size_t bytes = 1ull << 30, n = bytes / 16;
	int4 *s = nullptr, *t = nullptr;
	LS1_CUDA(cudaMalloc((void**) &s, bytes)); LS1_CUDA(cudaMalloc((void**) &t, bytes));
	LS1_CUDA(cudaMemset(s, 1, bytes));
	cudaEvent_t a, b; LS1_CUDA(cudaEventCreate(&a)); LS1_CUDA(cudaEventCreate(&b));
	double best = 0.;
	for (int rep = 0; rep < 6; rep++) {
		LS1_CUDA(cudaEventRecord(a));
		k_copy16<<<prop.multiProcessorCount * 16, 256>>>(s, t, n);
		g_launches++;
		LS1_CUDA(cudaEventRecord(b));
		LS1_CUDA(cudaEventSynchronize(b));
		float ms = 0.f; LS1_CUDA(cudaEventElapsedTime(&ms, a, b));
		if (rep > 0) best = std::max(best, 2.0 * (double) bytes / (ms * 1e-3) / 1e9);
	}
	cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(s); cudaFree(t);
	*gbs = best;
	return 0;
}

}  // extern "C"
