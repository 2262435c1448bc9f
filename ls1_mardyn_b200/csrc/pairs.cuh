// pairs.cuh -- the seven site-pair interactions of ls1-MarDyn as device functions, templated on the calc type C
// (float: mixed build, double: double build).
//
// Reference: VectorizedCellProcessor::_loopBodyLJ ... _loopBodyQuadrupole (particleContainer/adapter/
// VectorizedCellProcessor.cpp:171-792) and the scalar formulas of molecules/potforce.h:18-263. Conventions:
// every function returns f = force on site A with d = rA - rB; torques are returned per oriented site; the potential
// energy term is returned as written into the reference's accumulators (6*U for LJ, U for multipoles).
// Operation order follows the reference where it affects which pairs interact (the cutoff test, molpair_r2 below);
// elsewhere it is free: results are compared with relative tolerances (1e-5 mixed, 1e-10 double), not bit for bit.
#pragma once
#include "common.cuh"

namespace ls1b200 {

template <class C> __device__ __forceinline__ C c_rcp(C x);
// MUFU.RCP (1 ulp): the AVX2 reference itself uses rcp_ps + one Newton step (RealVecFloat.h:432-457), not a correctly rounded quotient
template <> __device__ __forceinline__ float c_rcp<float>(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
template <> __device__ __forceinline__ double c_rcp<double>(double x) { return 1.0 / x; }
template <class C> __device__ __forceinline__ C c_rsqrt(C x);
// MUFU.RSQ + one Newton step, like the reference's rsqrt_ps + NR (RealVecFloat.h:459-480)
template <> __device__ __forceinline__ float c_rsqrt<float>(float x) {
	float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r * fmaf(-0.5f * x * r, r, 1.5f);
}
template <> __device__ __forceinline__ double c_rsqrt<double>(double x) { return 1.0 / sqrt(x); }
template <class C> __device__ __forceinline__ C c_fma(C a, C b, C c);
template <> __device__ __forceinline__ float c_fma<float>(float a, float b, float c) { return __fmaf_rn(a, b, c); }
template <> __device__ __forceinline__ double c_fma<double>(double a, double b, double c) { return __fma_rn(a, b, c); }
template <class C> __device__ __forceinline__ C c_mul(C a, C b);
template <> __device__ __forceinline__ float c_mul<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double c_mul<double>(double a, double b) { return __dmul_rn(a, b); }
template <class C> __device__ __forceinline__ C c_sub(C a, C b);
template <> __device__ __forceinline__ float c_sub<float>(float a, float b) { return __fsub_rn(a, b); }
template <> __device__ __forceinline__ double c_sub<double>(double a, double b) { return __dsub_rn(a, b); }

// Squared centre-of-mass distance EXACTLY as RealCalcVec::scal_prod of the AVX2 reference build evaluates it
// (vectorization/RealVecFloat.h:219-223, SIMD_VectorizedCellProcessorHelpers.h:443-449):
//   fma(dx, dx, fma(dy, dy, dz*dz)) with dx = xi - xj in calc precision. The cutoff decision r2 < rc^2 is bit-identical.
template <class C>
__device__ __forceinline__ C molpair_r2(C xi, C yi, C zi, C xj, C yj, C zj) {
	const C dx = c_sub<C>(xi, xj), dy = c_sub<C>(yi, yj), dz = c_sub<C>(zi, zj);
	return c_fma<C>(dx, dx, c_fma<C>(dy, dy, c_mul<C>(dz, dz)));
}

template <class C> struct SiteC { C x, y, z, p, ex, ey, ez; };

template <class C> __device__ __forceinline__ C dot3(C a1, C a2, C a3, C b1, C b2, C b3) { return a1 * b1 + a2 * b2 + a3 * b3; }

// Lennard-Jones 12-6: f, 6*U (+ 6*shift).  VectorizedCellProcessor.cpp:171-224
template <class C>
__device__ __forceinline__ void pair_lj(const SiteC<C>& a, const SiteC<C>& b, C eps24, C sig2, C shift6, C* f, C& upot6) {
	const C dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
	const C r2inv = c_rcp<C>(dx * dx + dy * dy + dz * dz);
	const C lj2 = sig2 * r2inv, lj6 = lj2 * lj2 * lj2, lj12 = lj6 * lj6, lj12m6 = lj12 - lj6;
	const C scale = eps24 * r2inv * (lj12 + lj12m6);
	f[0] = dx * scale; f[1] = dy * scale; f[2] = dz * scale;
	upot6 = eps24 * lj12m6 + shift6;
}

// charge - charge.  VectorizedCellProcessor.cpp:227-277
template <class C>
__device__ __forceinline__ void pair_charge(const SiteC<C>& a, const SiteC<C>& b, C* f, C& upot) {
	const C dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
	const C r2 = dx * dx + dy * dy + dz * dz;
	const C r2inv = c_rcp<C>(r2), rinv = c_rsqrt<C>(r2);
	upot = a.p * b.p * rinv;
	const C fac = upot * r2inv;
	f[0] = dx * fac; f[1] = dy * fac; f[2] = dz * fac;
}

// a = charge, b = dipole; M = torque on the dipole.  VectorizedCellProcessor.cpp:279-347
template <class C>
__device__ __forceinline__ void pair_charge_dipole(const SiteC<C>& a, const SiteC<C>& b, C* f, C* M, C& upot) {
	const C dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
	const C r2 = dx * dx + dy * dy + dz * dz;
	const C r2inv = c_rcp<C>(r2), rinv = c_rsqrt<C>(r2);
	const C re = dot3<C>(dx, dy, dz, b.ex, b.ey, b.ez);
	const C qpr3 = a.p * b.p * r2inv * rinv;
	const C fac = C(3) * r2inv * re;
	f[0] = qpr3 * (b.ex - dx * fac); f[1] = qpr3 * (b.ey - dy * fac); f[2] = qpr3 * (b.ez - dz * fac);
	upot = -(qpr3 * re);
	M[0] = qpr3 * (b.ey * dz - b.ez * dy);
	M[1] = qpr3 * (b.ez * dx - b.ex * dz);
	M[2] = qpr3 * (b.ex * dy - b.ey * dx);
}

// dipole - dipole with reaction field.  VectorizedCellProcessor.cpp:349-442
template <class C>
__device__ __forceinline__ void pair_dipole(const SiteC<C>& a, const SiteC<C>& b, C eps_rf_invrc3, C* f, C* M1, C* M2, C& upot, C& my_rf) {
	const C dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
	const C r2 = dx * dx + dy * dy + dz * dz;
	const C r2inv = c_rcp<C>(r2), rinv = c_rsqrt<C>(r2);
	const C r2three = C(3) * r2inv;
	const C p1p2 = a.p * b.p;
	const C rffac = p1p2 * eps_rf_invrc3;
	const C p1p2r3 = p1p2 * rinv * r2inv;
	const C p1p2three_r5 = p1p2r3 * r2three;
	const C e1e2 = dot3<C>(a.ex, a.ey, a.ez, b.ex, b.ey, b.ez);
	const C re1 = dot3<C>(dx, dy, dz, a.ex, a.ey, a.ez), re2 = dot3<C>(dx, dy, dz, b.ex, b.ey, b.ez);
	const C re1re2perr2 = r2inv * re1 * re2;
	const C t = e1e2 - C(5) * re1re2perr2;
	f[0] = p1p2three_r5 * (dx * t + a.ex * re2 + b.ex * re1);
	f[1] = p1p2three_r5 * (dy * t + a.ey * re2 + b.ey * re1);
	f[2] = p1p2three_r5 * (dz * t + a.ez * re2 + b.ez * re1);
	upot = p1p2r3 * (e1e2 - C(3) * re1re2perr2);
	my_rf = rffac * e1e2;
	const C cx = a.ey * b.ez - a.ez * b.ey, cy = a.ez * b.ex - a.ex * b.ez, cz = a.ex * b.ey - a.ey * b.ex;   // e1 x e2
	const C re1three = re1 * r2three, re2three = re2 * r2three;
	M1[0] = p1p2r3 * (re2three * (a.ey * dz - a.ez * dy) - cx) + rffac * cx;
	M1[1] = p1p2r3 * (re2three * (a.ez * dx - a.ex * dz) - cy) + rffac * cy;
	M1[2] = p1p2r3 * (re2three * (a.ex * dy - a.ey * dx) - cz) + rffac * cz;
	M2[0] = p1p2r3 * (re1three * (b.ey * dz - b.ez * dy) + cx) - rffac * cx;
	M2[1] = p1p2r3 * (re1three * (b.ez * dx - b.ex * dz) + cy) - rffac * cy;
	M2[2] = p1p2r3 * (re1three * (b.ex * dy - b.ey * dx) + cz) - rffac * cz;
}

// a = charge, b = quadrupole; M = torque on the quadrupole.  VectorizedCellProcessor.cpp:444-524
template <class C>
__device__ __forceinline__ void pair_charge_quadrupole(const SiteC<C>& a, const SiteC<C>& b, C* f, C* M, C& upot) {
	const C dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
	const C r2 = dx * dx + dy * dy + dz * dz;
	const C invdr2 = c_rcp<C>(r2), invdr = c_rsqrt<C>(r2);
	const C qQ05 = C(0.5) * a.p * b.p;
	const C costj = dot3<C>(b.ex, b.ey, b.ez, dx, dy, dz) * invdr;
	const C qQinv4dr3 = qQ05 * invdr * invdr2;
	upot = qQinv4dr3 * (C(3) * costj * costj - C(1));
	const C minus_dUdr = C(3) * upot * invdr2;
	const C dUdt = C(6) * costj * qQinv4dr3 * invdr;
	const C fac = costj * dUdt * invdr + minus_dUdr;
	f[0] = fac * dx - dUdt * b.ex; f[1] = fac * dy - dUdt * b.ey; f[2] = fac * dz - dUdt * b.ez;
	M[0] = dUdt * (b.ez * dy - b.ey * dz);
	M[1] = dUdt * (b.ex * dz - b.ez * dx);
	M[2] = dUdt * (b.ey * dx - b.ex * dy);
}

// a = dipole, b = quadrupole.  VectorizedCellProcessor.cpp:526-648
template <class C>
__device__ __forceinline__ void pair_dipole_quadrupole(const SiteC<C>& a, const SiteC<C>& b, C* f, C* M1, C* M2, C& upot) {
	const C dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
	const C r2 = dx * dx + dy * dy + dz * dz;
	const C invdr2 = c_rcp<C>(r2), invdr = c_rsqrt<C>(r2);
	const C myqfac = C(1.5) * a.p * b.p * invdr2 * invdr2;
	const C costi = dot3<C>(a.ex, a.ey, a.ez, dx, dy, dz) * invdr;
	const C costj = dot3<C>(b.ex, b.ey, b.ez, dx, dy, dz) * invdr;
	const C cos2tj = costj * costj;
	const C cosgij = dot3<C>(a.ex, a.ey, a.ez, b.ex, b.ey, b.ez);
	const C f5 = C(5) * cos2tj - C(1);
	const C c2 = C(2) * costj;
	upot = myqfac * (c2 * cosgij - costi * f5);
	const C myqfacXinvdr = myqfac * invdr;
	const C mpR = C(4) * upot * invdr2;
	const C mpTi = myqfacXinvdr * f5;
	const C mpTj = myqfacXinvdr * C(2) * (C(5) * costi * costj - cosgij);
	const C pG = myqfac * c2;
	const C fac = mpR - (costi * mpTi + costj * mpTj) * invdr;
	f[0] = fac * dx + mpTi * a.ex + mpTj * b.ex;
	f[1] = fac * dy + mpTi * a.ey + mpTj * b.ey;
	f[2] = fac * dz + mpTi * a.ez + mpTj * b.ez;
	const C gx = pG * (a.ey * b.ez - a.ez * b.ey), gy = pG * (a.ez * b.ex - a.ex * b.ez), gz = pG * (a.ex * b.ey - a.ey * b.ex);
	M1[0] = mpTi * (a.ey * dz - a.ez * dy) - gx;
	M1[1] = mpTi * (a.ez * dx - a.ex * dz) - gy;
	M1[2] = mpTi * (a.ex * dy - a.ey * dx) - gz;
	M2[0] = mpTj * (b.ey * dz - b.ez * dy) + gx;
	M2[1] = mpTj * (b.ez * dx - b.ex * dz) + gy;
	M2[2] = mpTj * (b.ex * dy - b.ey * dx) + gz;
}

// quadrupole - quadrupole.  VectorizedCellProcessor.cpp:650-792
template <class C>
__device__ __forceinline__ void pair_quadrupole(const SiteC<C>& a, const SiteC<C>& b, C* f, C* M1, C* M2, C& upot) {
	const C dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
	const C r2 = dx * dx + dy * dy + dz * dz;
	const C invdr2 = c_rcp<C>(r2), invdr = c_rsqrt<C>(r2);
	const C qfac = C(0.75) * invdr * (a.p * b.p) * (invdr2 * invdr2);
	const C costi = dot3<C>(a.ex, a.ey, a.ez, dx, dy, dz) * invdr;
	const C costj = dot3<C>(b.ex, b.ey, b.ez, dx, dy, dz) * invdr;
	const C cos2ti = costi * costi, cos2tj = costj * costj;
	const C cosgij = dot3<C>(a.ex, a.ey, a.ez, b.ex, b.ey, b.ez);
	const C term = cosgij - C(5) * (costi * costj);
	upot = qfac * ((C(1) + C(2) * term * term) - (C(5) * (cos2ti + cos2tj) + C(15) * cos2ti * cos2tj));
	const C mpR = C(5) * upot * invdr2;
	const C part1 = qfac * C(10) * invdr, term2 = C(2) * term;
	const C mpTi = part1 * (costi + term2 * costj + C(3) * costi * cos2tj);
	const C mpTj = part1 * (costj + term2 * costi + C(3) * costj * cos2ti);
	const C pG = qfac * C(4) * term;
	const C fac = mpR - (mpTi * costi + mpTj * costj) * invdr;
	f[0] = fac * dx + mpTi * a.ex + mpTj * b.ex;
	f[1] = fac * dy + mpTi * a.ey + mpTj * b.ey;
	f[2] = fac * dz + mpTi * a.ez + mpTj * b.ez;
	const C gx = pG * (a.ey * b.ez - a.ez * b.ey), gy = pG * (a.ez * b.ex - a.ex * b.ez), gz = pG * (a.ex * b.ey - a.ey * b.ex);
	M1[0] = mpTi * (a.ey * dz - a.ez * dy) - gx;
	M1[1] = mpTi * (a.ez * dx - a.ex * dz) - gy;
	M1[2] = mpTi * (a.ex * dy - a.ey * dx) - gz;
	M2[0] = mpTj * (b.ey * dz - b.ez * dy) + gx;
	M2[1] = mpTj * (b.ez * dx - b.ex * dz) + gy;
	M2[2] = mpTj * (b.ex * dy - b.ey * dx) + gz;
}

// ---- block reduction of up to 8 doubles per thread into partials[blockIdx][8] ---------------------------------------
template <int NV>
__device__ __forceinline__ void block_reduce_store(double* v, double* __restrict__ partials, uint32_t block_linear) {
	__shared__ double sh[32][NV];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	#pragma unroll
	for (int k = 0; k < NV; k++) {
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
	}
	if (lane == 0) {
		#pragma unroll
		for (int k = 0; k < NV; k++) sh[w][k] = v[k];
	}
	__syncthreads();
	if (threadIdx.x < NV) {
		double s = 0.;
		for (int i = 0; i < nw; i++) s += sh[i][threadIdx.x];
		partials[(size_t) block_linear * 8 + threadIdx.x] = s;
	}
}

}  // namespace ls1b200
