// force_1clj_pf.cu -- single-component 1CLJ cell-pair kernel, FP32 pair arithmetic (the mixed-precision build and the
// reduced-memory mode): the tiled kernel of force_1clj_tiled.cu with a cheaper candidate test and conflict-free lists.
//
// Reference: VectorizedCellProcessor::_calculatePairs LJ block (adapter/VectorizedCellProcessor.cpp:1029-1178), _loopBodyLJ
// (:171-224), calcDistLookup (vectorization/SIMD_VectorizedCellProcessorHelpers.h:427-472), FullMolecule::calcFM
// (molecules/FullMolecule.cpp:548-626); cell pairs of the C08 traversal (C08BasedTraversals.h:48-99) visited from both ends.
//
// What differs from force_1clj_tiled.cu (same chunks, same TMA windows, same two phases):
//   * Window frame. After the TMA copies land, the CTA rewrites its 9 windows in place: (x, y, z, cid) -> (x - ox, y - oy,
//     z - oz, |r'|^2). The origin o is the lower corner of the window rounded DOWN to FP32 (0 where the window touches 0),
//     chosen so that the subtraction is EXACT (every coordinate is a multiple of ulp(o) and the difference needs < 24 bits;
//     checked per dimension, o = 0 otherwise): differences of translated coordinates are bit-identical to differences of the
//     reference's FP32 coordinates.
//   * Phase 1 is a CONSERVATIVE prefilter, 3 FFMA + 1 compare per candidate instead of 3 FADD + 1 FMUL + 2 FFMA + compare:
//     |ri - rj|^2 = |rj|^2 - 2 ri.rj + |ri|^2, so  t = fma(-2xi, xj, fma(-2yi, yj, fma(-2zi, zj, wj))) < rc^2 + margin - |ri|^2.
//     In the window frame |r'|^2 is a few thousand, the rounding error of t is bounded by 18 * 2^-24 * R^2 (R^2: the frame's
//     largest |r'|^2; derivation in DESIGN.md 4.1), and margin = 32 * 2^-24 * R^2 makes the filter rigorous: no pair inside the
//     cutoff is lost; ~0.1 % more list entries.
//   * Phase 2 evaluates the reference's bit-identical molecule-distance test (fma(dx,dx,fma(dy,dy,dz*dz)) < rc^2, and != 0 --
//     the molecule itself and coincident molecules, SingleCellPolicy_) on every list entry and masks the pair through 1/r^2.
//   * List entries are 32-bit words, one word per thread and entry: every append and every list read is bank-conflict free
//     (16-bit entries put two threads into one bank: 32 % of the tiled kernel's shared-memory wavefronts were conflicts).
#include "tile.cuh"

namespace ls1b200 {

#ifndef LS1_PF_LCAP
#define LS1_PF_LCAP 32
#endif
#ifndef LS1_PF_SOFT
#define LS1_PF_SOFT 14
#endif
#ifndef LS1_PF_TPB
#define LS1_PF_TPB 128
#endif
#ifndef LS1_PF_W
#define LS1_PF_W 176
#endif
constexpr int PF_TPB = LS1_PF_TPB, PF_W = LS1_PF_W, PF_LCAP = LS1_PF_LCAP, PF_SOFT = LS1_PF_SOFT;

struct PFArgs {
	const float4* mpos;
	const uint32_t* cell_start;
	const uint32_t* cell_of;
	const double* border; uint32_t border_stride;      // [3][stride] cell borders (context.cu)
	double* F; double* M; double* Vi;
	double* partials;
	const Chunk* chunks;
	const uint32_t* n_chunks;
	uint32_t* work_counter;
	DevScalars* scal;
	uint32_t cap;
	int nx, ny, nz;
	float ljrc2, eps24, sig2, shift6;
	float* rmm_v; float rmm_dt_m;
};

__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_v4f(uint32_t addr, float4 v) {
	asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Frame of one dimension of a window whose cells are [c_lo, c_hi]: origin o (exactly subtractable from every FP32 coordinate
// of the window, or 0) and an upper bound X of |coordinate - o|. Coordinates lie in [border(c_lo), border(c_hi + 1)) in FP64
// (that is how they were binned) and were rounded to nearest, so they lie in [RD(lo), RU(hi)] in FP32.
__device__ __forceinline__ void pf_frame(const double* __restrict__ tab, int c_lo, int c_hi, float& o, float& X) {
	const double lo = tab[c_lo], hi = tab[c_hi + 1];
	const float l = __double2float_rd(lo), h = __double2float_ru(hi);
	float of = 0.f;
	if (l >= 1.f) {
		const float p2 = __int_as_float(__float_as_int(l) & 0x7F800000);      // 2^e, e the exponent of l: ulp(l) = 2^(e-23)
		// x - l is a non-negative multiple of ulp(l) for every FP32 x >= l; it is representable iff it is < 2^(e+1)
		if (__fsub_ru(h, l) < 2.f * p2) of = l;
	}
	o = of;
	X = fmaxf(fabsf(__fsub_ru(h, of)), fabsf(__fsub_rd(l, of)));
}

template <bool RMM, bool CHECK>
__global__ void __launch_bounds__(PF_TPB, 4) k_force_1clj_pf(PFArgs a) {
	constexpr int TPB = PF_TPB, W = PF_W, LCAP = PF_LCAP;
	constexpr uint32_t VB = 16, LSTRIDE = (uint32_t) TPB * 4u;                      // bytes between consecutive entries of one list
	extern __shared__ __align__(128) unsigned char smem_raw[];
	float4* win = reinterpret_cast<float4*>(smem_raw);                                  // [9][W] (+1 sentinel record at [9*W])
	uint32_t* list = reinterpret_cast<uint32_t*>(smem_raw + (size_t) (9 * W + 1) * VB);  // [LCAP+1][TPB]
	uint32_t* cs_all = list + (size_t) (LCAP + 1) * TPB;                                // [2][12][CH_MAXC+3]
	constexpr int CSB = 12 * (CH_MAXC + 3);
	__shared__ uint32_t s_rowbase_all[2][9];
	__shared__ float s_frame_all[2][4];                                                 // ox, oy, oz, R^2 of this chunk and of the next one
	__shared__ __align__(8) unsigned long long s_bar;
	__shared__ uint32_t s_next;
	const uint32_t win_s = smem_u32(win), list_s = smem_u32(list) + threadIdx.x * 4, bar = smem_u32(&s_bar);
	const uint32_t n_chunks = *a.n_chunks;
	const uint32_t sentinel = win_s + (uint32_t) (9 * W) * VB;
	if (threadIdx.x == 0) {
		mbar_init(bar, 1);
		win[9 * W] = make_float4(1e18f, 1e18f, 1e18f, 3e36f);      // pads odd lists: fails the distance test of phase 2
		s_next = atomicAdd(a.work_counter, 1u);
	}
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	fence_proxy_async();
	__syncthreads();
	auto fill_cs = [&](const Chunk& q, uint32_t* csq, uint32_t* rbq, float* frq) {
		const int qn = q.ncells;
		const uint32_t qR = q.cell0 / (uint32_t) a.nx;
		const int qxa = (int) (q.cell0 % (uint32_t) a.nx);
		for (int k = threadIdx.x; k < 9 * (qn + 3); k += TPB) {
			const int r = k / (qn + 3), c = k % (qn + 3);
			const uint32_t row = (uint32_t) ((int) qR + (r / 3 - 1) * a.ny + (r % 3 - 1)) * a.nx;
			const uint32_t b0 = a.cell_start[row + qxa - 1];
			csq[r * (CH_MAXC + 3) + c] = a.cell_start[row + qxa - 1 + c] - b0;
			if (c == 0) rbq[r] = b0;
		}
		if (threadIdx.x == TPB - 1) {      // the frame of the window: cells qxa-1 .. qxa+qn, rows cy-1 .. cy+1, cz-1 .. cz+1
			const int cy = (int) (qR % (uint32_t) a.ny), cz = (int) (qR / (uint32_t) a.ny);
			float o[3], X[3];
			pf_frame(a.border, qxa - 1, qxa + qn, o[0], X[0]);
			pf_frame(a.border + a.border_stride, cy - 1, cy + 1, o[1], X[1]);
			pf_frame(a.border + 2 * (size_t) a.border_stride, cz - 1, cz + 1, o[2], X[2]);
			frq[0] = o[0]; frq[1] = o[1]; frq[2] = o[2];
			frq[3] = __fmaf_ru(X[0], X[0], __fmaf_ru(X[1], X[1], __fmul_ru(X[2], X[2])));
		}
	};
	uint32_t ch = s_next;
	Chunk cd; cd.cell0 = 0; cd.ncells = 0; cd.flags = 0;
	int cur = 0;
	if (ch < n_chunks) { cd = a.chunks[ch]; if (!(cd.flags & CHUNK_BIG)) fill_cs(cd, cs_all, s_rowbase_all[0], s_frame_all[0]); }
	__syncthreads();
	uint32_t parity = 0;
	while (ch < n_chunks) {
		uint32_t* cs = cs_all + cur * CSB;
		uint32_t* s_rowbase = s_rowbase_all[cur];
		const float* s_frame = s_frame_all[cur];
		const bool big = (cd.flags & CHUNK_BIG) != 0;
		const int nc = cd.ncells;
		const uint32_t R = cd.cell0 / (uint32_t) a.nx;
		if (threadIdx.x == 0) {
			if (!big) {
				uint32_t total = 0;
				for (int r = 0; r < 9; r++) total += cs[r * (CH_MAXC + 3) + nc + 2] * VB;
				if (total) {
					fence_proxy_async();
					mbar_expect_tx(bar, total);
					for (int r = 0; r < 9; r++) {
						const uint32_t b = s_rowbase[r], len = cs[r * (CH_MAXC + 3) + nc + 2];
						if (len) tma_bulk_g2s(win_s + (uint32_t) (r * W) * VB, a.mpos + b, len * VB, bar);
					}
				}
			}
			s_next = atomicAdd(a.work_counter, 1u);
		}
		__syncthreads();
		const uint32_t nch = s_next;
		Chunk ncd; ncd.cell0 = 0; ncd.ncells = 0; ncd.flags = 0;
		if (nch < n_chunks) { ncd = a.chunks[nch]; if (!(ncd.flags & CHUNK_BIG)) fill_cs(ncd, cs_all + (cur ^ 1) * CSB, s_rowbase_all[cur ^ 1], s_frame_all[cur ^ 1]); }
		double acc[2] = {0., 0.};
		if (big) {
			// ---- slow generic path: a cell denser than the tile. Global loads, the reference's test, pair evaluated at once. ----
			const uint32_t ib = a.cell_start[cd.cell0], ie = a.cell_start[cd.cell0 + nc];
			for (uint32_t t = ib + threadIdx.x; t < ie; t += TPB) {
				const float4 mi = a.mpos[t];
				const int cx = (int) (a.cell_of[t] % (uint32_t) a.nx);
				double Fd[3] = {0., 0., 0.}, Vd[3] = {0., 0., 0.}, Ud = 0.;
				for (int r = 0; r < 9; r++) {
					const uint32_t row = (uint32_t) ((int) R + (r / 3 - 1) * a.ny + (r % 3 - 1)) * a.nx + cx - 1;
					const uint32_t jb = a.cell_start[row], je = a.cell_start[row + 3];
					for (uint32_t j0 = jb; j0 < je; j0 += 32) {
						float fx = 0.f, fy = 0.f, fz = 0.f, vx = 0.f, vy = 0.f, vz = 0.f, up = 0.f;
						int nh = 0;
						for (uint32_t j = j0; j < min(je, j0 + 32); j++) {
							const float4 mj = a.mpos[j];
							const float dx = __fsub_rn(mi.x, mj.x), dy = __fsub_rn(mi.y, mj.y), dz = __fsub_rn(mi.z, mj.z);
							const float r2 = __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, __fmul_rn(dz, dz)));
							if ((r2 < a.ljrc2) && (r2 != 0.f)) {
								const float r2inv = rcp_fast<float>(r2);
								const float lj2 = a.sig2 * r2inv, lj6 = lj2 * lj2 * lj2, lj12 = lj6 * lj6, lj12m6 = lj12 - lj6;
								const float scale = r2inv * (lj12 + lj12m6);
								fx = fmaf(dx, scale, fx); fy = fmaf(dy, scale, fy); fz = fmaf(dz, scale, fz);
								vx = fmaf(dx * dx, scale, vx); vy = fmaf(dy * dy, scale, vy); vz = fmaf(dz * dz, scale, vz);
								up += lj12m6; nh++;
							}
						}
						Fd[0] += (double) (a.eps24 * fx); Fd[1] += (double) (a.eps24 * fy); Fd[2] += (double) (a.eps24 * fz);
						Vd[0] += (double) (a.eps24 * vx); Vd[1] += (double) (a.eps24 * vy); Vd[2] += (double) (a.eps24 * vz);
						Ud += (double) (a.eps24 * up + (float) nh * a.shift6);
					}
				}
				acc[0] += 0.5 * Ud; acc[1] += 0.5 * (Vd[0] + Vd[1] + Vd[2]);
				const size_t cp = a.cap;
				if (RMM) {
					a.rmm_v[t] += (float) Fd[0] * a.rmm_dt_m; a.rmm_v[cp + t] += (float) Fd[1] * a.rmm_dt_m; a.rmm_v[2 * cp + t] += (float) Fd[2] * a.rmm_dt_m;
				} else {
					a.F[t] = Fd[0]; a.F[cp + t] = Fd[1]; a.F[2 * cp + t] = Fd[2];
					a.Vi[t] = 0.5 * Vd[0]; a.Vi[cp + t] = 0.5 * Vd[1]; a.Vi[2 * cp + t] = 0.5 * Vd[2];
					if (a.M) { a.M[t] = 0.; a.M[cp + t] = 0.; a.M[2 * cp + t] = 0.; }
				}
			}
		} else {
		const uint32_t* csc = cs + 4 * (CH_MAXC + 3);          // centre row (relative record indices)
		const uint32_t n_mol = csc[nc + 1] - csc[1];
		const bool have = threadIdx.x < n_mol;
		const uint32_t trel = csc[1] + threadIdx.x;
		int kx = 0;                                             // cell (inside the chunk) of my molecule
		while (kx + 1 < nc && csc[kx + 2] <= trel) kx++;
		const float ox = s_frame[0], oy = s_frame[1], oz = s_frame[2];
		// margin of the prefilter: 32 ulp-units of the frame's largest |r'|^2 (error bound of t and of the threshold: 18) + 8 of rc^2
		const float margin = s_frame[3] * (32.f / 16777216.f) + a.ljrc2 * (8.f / 16777216.f);
		mbar_wait(bar, parity); parity ^= 1u;
		// ---- window frame: (x, y, z, cid) -> (x - ox, y - oy, z - oz, |r'|^2), in place ----
		#pragma unroll 1
		for (int r = 0; r < 9; r++) {
			const uint32_t len = cs[r * (CH_MAXC + 3) + nc + 2];
			for (uint32_t k = threadIdx.x; k < len; k += TPB) {
				const uint32_t ad = win_s + (uint32_t) (r * W + (int) k) * VB;
				float4 m = lds_v4<float>(ad);
				const float x = __fsub_rn(m.x, ox), y = __fsub_rn(m.y, oy), z = __fsub_rn(m.z, oz);
				if (CHECK) { if (__fadd_rn(x, ox) != m.x || __fadd_rn(y, oy) != m.y || __fadd_rn(z, oz) != m.z) a.scal->overflow = 6; }
				m.x = x; m.y = y; m.z = z; m.w = __fmaf_rn(x, x, __fmaf_rn(y, y, __fmul_rn(z, z)));
				sts_v4f(ad, m);
			}
		}
		__syncthreads();
		// ---- phase 1 (prefilter) + phase 2 (exact test, LJ 12-6) ----
		float4 mi = make_float4(0.f, 0.f, 0.f, 0.f);
		float Ti = -3e38f;                                      // an absent molecule: no candidate passes
		if (have) { mi = lds_v4<float>(win_s + (uint32_t) (4 * W + (int) trel) * VB); Ti = __fsub_rn(__fadd_rn(a.ljrc2, margin), mi.w); }
		float ax = -2.f * mi.x, ay = -2.f * mi.y, az = -2.f * mi.z;
		asm volatile("" : "+f"(ax), "+f"(ay), "+f"(az));      // keeps the three factors in registers (the compiler would recompute them inside the loop)
		double Fd[3] = {0., 0., 0.}, Vd[3] = {0., 0., 0.}, Ud = 0.;
		const uint32_t lbase = list_s, wlimit = lbase + (uint32_t) (LCAP - 4) * LSTRIDE, wsoft = lbase + (uint32_t) PF_SOFT * LSTRIDE;
		uint32_t w = lbase;
		uint32_t src = 0, end = 0;
		int r = -1;
		for (;;) {
			bool flush = true;
			if (r < 9) {
				while (src + 4 * VB <= end && w <= wlimit) {
					const float4 m0 = lds_v4<float>(src), m1 = lds_v4<float>(src + VB), m2 = lds_v4<float>(src + 2 * VB), m3 = lds_v4<float>(src + 3 * VB);
					const float t0 = __fmaf_rn(ax, m0.x, __fmaf_rn(ay, m0.y, __fmaf_rn(az, m0.z, m0.w)));
					const float t1 = __fmaf_rn(ax, m1.x, __fmaf_rn(ay, m1.y, __fmaf_rn(az, m1.z, m1.w)));
					const float t2 = __fmaf_rn(ax, m2.x, __fmaf_rn(ay, m2.y, __fmaf_rn(az, m2.z, m2.w)));
					const float t3 = __fmaf_rn(ax, m3.x, __fmaf_rn(ay, m3.y, __fmaf_rn(az, m3.z, m3.w)));
					if (t0 < Ti) { sts32(w, src); w += LSTRIDE; }
					if (t1 < Ti) { sts32(w, src + VB); w += LSTRIDE; }
					if (t2 < Ti) { sts32(w, src + 2 * VB); w += LSTRIDE; }
					if (t3 < Ti) { sts32(w, src + 3 * VB); w += LSTRIDE; }
					src += 4 * VB;
				}
				if (w <= wlimit) {                   // the row is down to < 4 candidates
					for (; src < end; src += VB) {
						const float4 mj = lds_v4<float>(src);
						const float t = __fmaf_rn(ax, mj.x, __fmaf_rn(ay, mj.y, __fmaf_rn(az, mj.z, mj.w)));
						if (t < Ti) { sts32(w, src); w += LSTRIDE; }
					}
					r++;
					if (r < 9) {
						uint32_t kb = 0, ke = 0;
						if (have) { kb = cs[r * (CH_MAXC + 3) + kx]; ke = cs[r * (CH_MAXC + 3) + kx + 3]; }      // cells cx-1 .. cx+1 of this row
						src = win_s + (uint32_t) (r * W + (int) kb) * VB; end = win_s + (uint32_t) (r * W + (int) ke) * VB;
						// the lists are emptied by the whole warp together, at a row boundary, as soon as one of them is PF_SOFT entries long:
						// a lane that runs out of room in the middle of a row would evaluate its list alone (the other lanes idle)
						flush = __any_sync(__activemask(), w > wsoft) != 0;
					}
				}
			}
			if (flush) {
				const int cnt = (int) ((w - lbase) / LSTRIDE);
				float fx[2] = {0.f, 0.f}, fy[2] = {0.f, 0.f}, fz[2] = {0.f, 0.f}, vx[2] = {0.f, 0.f}, vy[2] = {0.f, 0.f}, vz[2] = {0.f, 0.f}, up[2] = {0.f, 0.f};
				float nhf = 0.f;                     // number of pairs inside the cutoff: sum of r2 * (1 / r2) = 1 +- 2^-22 per pair, rounded at the end
				if (cnt & 1) sts32(w, sentinel);
				#pragma unroll 1
				for (uint32_t e = lbase; e < w; e += 2 * LSTRIDE) {
					const uint32_t i0 = lds_u32(e), i1 = lds_u32(e + LSTRIDE);
					const float4 m0 = lds_v4<float>(i0), m1 = lds_v4<float>(i1);
					#pragma unroll
					for (int u = 0; u < 2; u++) {
						const float4 mj = u ? m1 : m0;
						const float dx = __fsub_rn(mi.x, mj.x), dy = __fsub_rn(mi.y, mj.y), dz = __fsub_rn(mi.z, mj.z);
						const float dz2 = __fmul_rn(dz, dz);
						const float r2 = __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, dz2));      // the reference's sequence: bit-identical cutoff decision
						const bool hit = (r2 < a.ljrc2) && (r2 != 0.f);
						const float r2inv = hit ? rcp_fast<float>(r2) : 0.f;
						nhf = fmaf(r2, r2inv, nhf);
						const float lj2 = a.sig2 * r2inv, lj6 = lj2 * lj2 * lj2, lj12 = lj6 * lj6, lj12m6 = lj12 - lj6;
						const float scale = r2inv * (lj12 + lj12m6);
						fx[u] = fmaf(dx, scale, fx[u]); fy[u] = fmaf(dy, scale, fy[u]); fz[u] = fmaf(dz, scale, fz[u]);
						vx[u] = fmaf(dx * dx, scale, vx[u]); vy[u] = fmaf(dy * dy, scale, vy[u]); vz[u] = fmaf(dz2, scale, vz[u]);
						up[u] += lj12m6;
					}
				}
				const float f24 = a.eps24;
				Fd[0] += (double) (f24 * (fx[0] + fx[1])); Fd[1] += (double) (f24 * (fy[0] + fy[1])); Fd[2] += (double) (f24 * (fz[0] + fz[1]));
				Vd[0] += (double) (f24 * (vx[0] + vx[1])); Vd[1] += (double) (f24 * (vy[0] + vy[1])); Vd[2] += (double) (f24 * (vz[0] + vz[1]));
				Ud += (double) (f24 * (up[0] + up[1]) + rintf(nhf) * a.shift6);
				w = lbase;
				if (r >= 9) break;
			}
		}
		// ---- results ----
		if (have) {
			const size_t cp = a.cap;
			const uint32_t t = s_rowbase[4] + trel;
			acc[0] += 0.5 * Ud;                          // every pair is visited from both of its ends
			acc[1] += 0.5 * (Vd[0] + Vd[1] + Vd[2]);
			if (RMM) {
				a.rmm_v[t] += (float) Fd[0] * a.rmm_dt_m; a.rmm_v[cp + t] += (float) Fd[1] * a.rmm_dt_m; a.rmm_v[2 * cp + t] += (float) Fd[2] * a.rmm_dt_m;
			} else {
				a.F[t] = Fd[0]; a.F[cp + t] = Fd[1]; a.F[2 * cp + t] = Fd[2];
				a.Vi[t] = 0.5 * Vd[0]; a.Vi[cp + t] = 0.5 * Vd[1]; a.Vi[2 * cp + t] = 0.5 * Vd[2];
				if (a.M) { a.M[t] = 0.; a.M[cp + t] = 0.; a.M[2 * cp + t] = 0.; }
			}
		}
		}
		#pragma unroll
		for (int k = 0; k < 2; k++) {
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
		}
		if ((threadIdx.x & 31) == 0) *reinterpret_cast<double2*>(a.partials + (size_t) ch * 8 + 2 * (threadIdx.x >> 5)) = make_double2(acc[0], acc[1]);
		if (TPB < 128 && threadIdx.x >= TPB / 32 && threadIdx.x < 4) *reinterpret_cast<double2*>(a.partials + (size_t) ch * 8 + 2 * threadIdx.x) = make_double2(0., 0.);
		__syncthreads();      // windows and lists are reused by the next chunk; its cell_start cache and frame are complete
		ch = nch; cd = ncd; cur ^= 1;
	}
}

int reduce_force_chunks(ls1b200_ctx* c, uint32_t nrows, int per_warp, bool kin = false, bool accumulate = false);      // force_1clj_tiled.cu

template <bool RMM>
static int launch_pf(ls1b200_ctx* c, float* rmm_v, float rmm_dt_m) {
	const Geom& g = c->geom;
	if (c->comps.size() != 1 || (!c->mixed && !RMM)) return fail("force_1clj_pf: single-component FP32 pair arithmetic only");
	uint32_t nrows = 0;
	LS1_TRY(build_chunks(c, PF_TPB, PF_W, 1, &nrows));
	PFArgs a;
	a.mpos = (const float4*) c->mpos; a.cell_start = c->cell_start; a.cell_of = c->cell_of2;
	a.border = c->d_border; a.border_stride = c->border_stride;
	a.F = c->F; a.M = c->rotates ? c->M : nullptr; a.Vi = c->Vi; a.partials = c->partials;
	a.chunks = (const Chunk*) c->chunks; a.n_chunks = c->chunk_row_start + nrows; a.work_counter = c->work_counter; a.scal = c->d_scal;
	a.cap = (uint32_t) c->cap; a.nx = g.cells[0]; a.ny = g.cells[1]; a.nz = g.cells[2];
	const double rclj = c->cfg.cutoff_lj;
	a.ljrc2 = (float) (rclj * rclj);
	a.eps24 = (float) c->h_eps24[0]; a.sig2 = (float) c->h_sig2[0]; a.shift6 = (float) c->h_shift6[0];
	a.rmm_v = rmm_v; a.rmm_dt_m = rmm_dt_m;
	const size_t smem = (size_t) (9 * PF_W + 1) * 16 + (size_t) (PF_LCAP + 1) * PF_TPB * 4 + 2 * 12 * (CH_MAXC + 3) * 4;
	auto kern = c->pf_check ? k_force_1clj_pf<RMM, true> : k_force_1clj_pf<RMM, false>;
	LS1_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	int per_sm = 0;
	LS1_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PF_TPB, smem));
	if (per_sm < 1) return fail("force_1clj_pf: kernel does not fit on an SM");
	const uint32_t grid = (uint32_t) (c->sm_count * per_sm);
	LS1_LAUNCH(c, PH_FORCE, kern, grid, PF_TPB, smem, a);
	return reduce_force_chunks(c, nrows, 1);
}

int force_1clj_pf(ls1b200_ctx* c) { return launch_pf<false>(c, nullptr, 0.f); }
int force_1clj_pf_rmm(ls1b200_ctx* c, float* v, float dt_over_m) { return launch_pf<true>(c, v, dt_over_m); }

}  // namespace ls1b200
