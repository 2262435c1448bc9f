// force_1clj_mma.cu -- the production 1CLJ cell-pair kernel of the mixed-precision build: the molecule-distance tests of a
// whole 16 x 8 tile of (i, j) pairs are done by ONE pair of tensor-core instructions (a conservative prefilter), the
// reference's bit-exact FP32 cutoff test and the LJ 12-6 arithmetic then run on the survivors only.
//
// Reference: same arithmetic as force_1clj_tiled.cu (VectorizedCellProcessor::_calculatePairs LJ block,
// adapter/VectorizedCellProcessor.cpp:1029-1178; _loopBodyLJ :171-224; calcDistLookup,
// vectorization/SIMD_VectorizedCellProcessorHelpers.h:427-472; FullMolecule::calcFM, molecules/FullMolecule.cpp:548-626).
// The cell pairs are those of the C08 traversal (LinkedCellTraversals/C08BasedTraversals.h:48-99), visited from both ends.
//
// Why: in force_1clj_tiled.cu 70 % of the instructions are the ~300 distance tests per molecule (27 cells x 11 molecules, of
// which ~50 are inside the cutoff), 12 instructions each. Here the test of 128 pairs costs 2 MMA + 1 LDS.64 per thread:
//
//   |ri - rj|^2 - (rc^2 + margin) = sum_k A[i][k] * B[j][k],   k = 0 .. 15, evaluated as mma.sync.m16n8k8 (TF32) x 2
//
// with coordinates relative to the chunk centre, every FP32 value split into a TF32 high part and a low remainder
// (3-term products hi*hi + hi*lo + lo*hi, the squares |r|^2 as hi + lo): the error of the expansion is below 2^-16 R^2
// (R = half extent of the chunk's windows), the margin is 2^-14 R^2, so no pair inside the cutoff is ever missed. A negative
// result is a SURVIVOR; survivors (pairs in range + the molecule itself + < 1 % near misses) are appended to per-thread
// lists in shared memory and evaluated by phase 2, which repeats the reference's test bit for bit
// (fma(dx,dx,fma(dy,dy,dz*dz)) < rc^2, != 0) before it accumulates anything. Tensor cores are a filter here, not the
// arithmetic of record: forces, virial and energy are FP32 FMA work folded into FP64 exactly as in the tiled kernel.
//
// Work decomposition: chunks as in tile.cuh (<= 64 molecules of consecutive cells of one x-row, 9 windows <= 112 records,
// TMA bulk copies). A CTA has 4 warps; warp w owns the 16 consecutive molecules 16w .. 16w+15 of the chunk (one MMA "M" tile).
// Its candidates are, per window row, the records of the cells cell(first)-1 .. cell(last)+1, walked in tiles of 8 ("N").
// Thread (g = lane / 4, t = lane % 4) receives C[g][2t], C[g][2t+1], C[g+8][2t], C[g+8][2t+1]: it appends survivors to its
// two lists (rows g and g+8). Phase 2: thread (g, t) evaluates row (t < 2 ? g : g + 8) over the lists written by the lanes
// 4g + 2(t&1) and 4g + 2(t&1) + 1 -- one molecule per thread, ~25 pairs, then a two-lane shuffle reduction.
#include "tile.cuh"

namespace ls1b200 {

namespace {

constexpr int M_TPB = 128, M_NI = 64, M_W = 112, M_WP = M_W + 8, M_LCAP = 40;
constexpr uint32_t M_LS = 128;           // bytes between consecutive entries of one list: [entry][64 lists] u16 per warp
constexpr size_t M_OFF_PF = ((size_t) (9 * M_W + 1) * 16 + 31) & ~(size_t) 31;
constexpr size_t M_OFF_LIST = M_OFF_PF + (size_t) 9 * M_WP * 32;
constexpr size_t M_OFF_CS = M_OFF_LIST + (size_t) (M_TPB / 32) * M_LCAP * M_LS;
constexpr size_t M_SMEM = M_OFF_CS + (size_t) 12 * (CH_MAXC + 3) * 4;

struct MArgs {
	const float4* mpos;
	const uint32_t* cell_start;
	const uint32_t* cell_of;
	const float* eps24; const float* sig2; const float* shift6;
	double* F; double* M; double* Vi;
	double* partials;
	const Chunk* chunks;
	const uint32_t* n_chunks;
	uint32_t* work_counter;
	uint32_t cap, n_lj;
	int nx, ny, nz;
	float ljrc2;
	float eps24_0, sig2_0, shift6_0;
	float hmin[3], clen[3];      // window geometry in FP32: only the prefilter's origin and error bound use it
};

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
	asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
	             : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float2 lds_v2(uint32_t addr) {
	float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory"); return v;
}

template <bool SINGLE>
__global__ void __launch_bounds__(M_TPB, 3) k_force_1clj_mma(MArgs a) {
	typedef float C;
	typedef float4 V4;
	constexpr int TPB = M_TPB, W = M_W, WP = M_WP, LCAP = M_LCAP;
	constexpr uint32_t VB = 16, LS = M_LS;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	V4* win = reinterpret_cast<V4*>(smem_raw);                                   // [9][W] exact records (+ sentinel at [9*W])
	float* pf = reinterpret_cast<float*>(smem_raw + M_OFF_PF);                   // [9][WP][8] prefilter records
	uint32_t* cs = reinterpret_cast<uint32_t*>(smem_raw + M_OFF_CS);             // [9][CH_MAXC+3]
	__shared__ uint32_t s_rowbase[9];
	__shared__ __align__(8) unsigned long long s_bar;
	__shared__ uint32_t s_chunk;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
	const uint32_t win_s = smem_u32(win), pf_s = smem_u32(pf), bar = smem_u32(&s_bar);
	const uint32_t lw_s = smem_u32(smem_raw + M_OFF_LIST) + (uint32_t) wid * (LCAP * LS);      // this warp's lists
	const uint32_t sentinel = win_s + (uint32_t) (9 * W) * VB;
	const uint32_t n_chunks = *a.n_chunks;
	if (threadIdx.x == 0) {
		mbar_init(bar, 1);
		win[9 * W] = make_float4(1e18f, 1e18f, 1e18f, 0.f);
	}
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	fence_proxy_async();
	__syncthreads();
	uint32_t parity = 0;
	for (;;) {
		if (threadIdx.x == 0) s_chunk = atomicAdd(a.work_counter, 1u);
		__syncthreads();
		const uint32_t ch = s_chunk;
		if (ch >= n_chunks) break;
		const Chunk cd = a.chunks[ch];
		const int nc = cd.ncells;
		const uint32_t R = cd.cell0 / (uint32_t) a.nx;                 // row index cz*ny + cy
		const int xa = (int) (cd.cell0 % (uint32_t) a.nx);
		double acc[2] = {0., 0.};
		if (cd.flags & CHUNK_BIG) {
			// ---- slow generic path: a cell denser than the tile. Global loads, pair evaluated as soon as it is found. ----
			const uint32_t ib = a.cell_start[cd.cell0], ie = a.cell_start[cd.cell0 + nc];
			for (uint32_t s = ib + threadIdx.x; s < ie; s += TPB) {
				const V4 mi = a.mpos[s];
				const uint32_t cid_i = __float_as_uint(mi.w) & 0xFFFFu;
				const int cx = (int) (a.cell_of[s] % (uint32_t) a.nx);
				double Fd[3] = {0., 0., 0.}, Vd[3] = {0., 0., 0.}, Ud = 0.;
				for (int r = 0; r < 9; r++) {
					const uint32_t row = (uint32_t) ((int) R + (r / 3 - 1) * a.ny + (r % 3 - 1)) * a.nx + cx - 1;
					const uint32_t jb = a.cell_start[row], je = a.cell_start[row + 3];
					for (uint32_t j0 = jb; j0 < je; j0 += 32) {
						C fx = 0.f, fy = 0.f, fz = 0.f, vx = 0.f, vy = 0.f, vz = 0.f, up = 0.f;
						for (uint32_t j = j0; j < min(je, j0 + 32); j++) {
							const V4 mj = a.mpos[j];
							const C dx = mi.x - mj.x, dy = mi.y - mj.y, dz = mi.z - mj.z;
							const C r2 = molpair_r2<C>(mi.x, mi.y, mi.z, mj.x, mj.y, mj.z);
							if ((r2 < a.ljrc2) && (r2 != 0.f)) {
								C eps24 = a.eps24_0, sig2 = a.sig2_0, shift6 = a.shift6_0;
								if (!SINGLE) { const size_t ti = (size_t) cid_i * a.n_lj + (__float_as_uint(mj.w) & 0xFFFFu); eps24 = a.eps24[ti]; sig2 = a.sig2[ti]; shift6 = a.shift6[ti]; }
								const C r2inv = rcp_fast<C>(r2);
								const C lj2 = sig2 * r2inv, lj6 = lj2 * lj2 * lj2, lj12 = lj6 * lj6, lj12m6 = lj12 - lj6;
								const C scale = eps24 * r2inv * (lj12 + lj12m6);
								fx = fmaf(dx, scale, fx); fy = fmaf(dy, scale, fy); fz = fmaf(dz, scale, fz);
								vx = fmaf(dx * dx, scale, vx); vy = fmaf(dy * dy, scale, vy); vz = fmaf(dz * dz, scale, vz);
								up += eps24 * lj12m6 + shift6;
							}
						}
						Fd[0] += (double) fx; Fd[1] += (double) fy; Fd[2] += (double) fz;
						Vd[0] += (double) vx; Vd[1] += (double) vy; Vd[2] += (double) vz;
						Ud += (double) up;
					}
				}
				acc[0] += 0.5 * Ud; acc[1] += 0.5 * (Vd[0] + Vd[1] + Vd[2]);
				const size_t cp = a.cap;
				a.F[s] = Fd[0]; a.F[cp + s] = Fd[1]; a.F[2 * cp + s] = Fd[2];
				a.Vi[s] = 0.5 * Vd[0]; a.Vi[cp + s] = 0.5 * Vd[1]; a.Vi[2 * cp + s] = 0.5 * Vd[2];
				if (a.M) { a.M[s] = 0.; a.M[cp + s] = 0.; a.M[2 * cp + s] = 0.; }
			}
			block_reduce_store<2>(acc, a.partials, ch);
			__syncthreads();
			continue;
		}
		// ---- 1. cell_start cache of the 9 rows (cells xa-1 .. xb+2 -> nc+3 entries each, relative to the window) and the TMA loads ----
		for (int k = threadIdx.x; k < 9 * (nc + 3); k += TPB) {
			const int r = k / (nc + 3), c = k % (nc + 3);
			const uint32_t row = (uint32_t) ((int) R + (r / 3 - 1) * a.ny + (r % 3 - 1)) * a.nx;
			const uint32_t b0 = a.cell_start[row + xa - 1];
			cs[r * (CH_MAXC + 3) + c] = a.cell_start[row + xa - 1 + c] - b0;
			if (c == 0) s_rowbase[r] = b0;
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			uint32_t total = 0;
			for (int r = 0; r < 9; r++) total += cs[r * (CH_MAXC + 3) + nc + 2] * VB;
			if (total) {
				fence_proxy_async();      // the previous chunk's generic-proxy reads of the windows precede these async-proxy writes
				mbar_expect_tx(bar, total);
				for (int r = 0; r < 9; r++) {
					const uint32_t b = s_rowbase[r], len = cs[r * (CH_MAXC + 3) + nc + 2];
					if (len) tma_bulk_g2s(win_s + (uint32_t) (r * W) * VB, a.mpos + b, len * VB, bar);
				}
			}
		}
		// origin of the prefilter coordinates (chunk centre) and the error margin of the TF32 expansion
		const C ox = a.hmin[0] + ((C) xa + 0.5f * (C) nc) * a.clen[0];
		const C oy = a.hmin[1] + ((C) (R % (uint32_t) a.ny) + 0.5f) * a.clen[1];
		const C oz = a.hmin[2] + ((C) (R / (uint32_t) a.ny) + 0.5f) * a.clen[2];
		const C hx = (0.5f * (C) nc + 1.5f) * a.clen[0];
		const C thr = a.ljrc2 + 6.1035156e-5f * (hx * hx + 2.25f * (a.clen[1] * a.clen[1] + a.clen[2] * a.clen[2]));
		__syncthreads();      // (nothing to publish; keeps the single-thread TMA issue ahead of the wait of every warp)
		mbar_wait(bar, parity); parity ^= 1u;
		// ---- 2. prefilter records of every window record: (x'h, x'l, y'h, y'l, z'h, z'l, s'h, s'l), s' = |r'|^2 - thr;
		//         8 far-away records behind each row so that the last tile of a range may overhang ----
		for (int idx = threadIdx.x; idx < 9 * WP; idx += TPB) {
			const int r = idx / WP, k = idx - r * WP;
			const int len = (int) cs[r * (CH_MAXC + 3) + nc + 2];
			if (k >= len + 8) continue;
			float4 lo4 = make_float4(0.f, 0.f, 0.f, 0.f), hi4 = make_float4(0.f, 0.f, 1e30f, 0.f);      // (x'h, x'l, y'h, y'l), (z'h, z'l, s'h, s'l)
			if (k < len) {
				const V4 m = win[r * W + k];
				const C x = m.x - ox, y = m.y - oy, z = m.z - oz;
				const C s = fmaf(x, x, fmaf(y, y, z * z)) - thr;
				const C xh = tf32_hi(x), yh = tf32_hi(y), zh = tf32_hi(z), sh = tf32_hi(s);
				lo4 = make_float4(xh, x - xh, yh, y - yh); hi4 = make_float4(zh, z - zh, sh, s - sh);
			}
			float4* dst = reinterpret_cast<float4*>(pf + (size_t) idx * 8);
			dst[0] = lo4; dst[1] = hi4;
		}
		__syncthreads();
		// ---- 3. one i-tile (16 molecules) per warp ----
		const uint32_t* csc = cs + 4 * (CH_MAXC + 3);          // centre row (relative record indices)
		const uint32_t n_mol = csc[nc + 1] - csc[1];
		const uint32_t i0 = 16u * (uint32_t) wid;
		if (i0 < n_mol) {
			const uint32_t ia = i0 + (uint32_t) g, ib = ia + 8u;
			const bool have_a = ia < n_mol, have_b = ib < n_mol;
			V4 ma = make_float4(1e18f, 1e18f, 1e18f, 0.f), mb = ma;      // an absent molecule: nothing is ever in range
			if (have_a) ma = win[4 * W + csc[1] + ia];
			if (have_b) mb = win[4 * W + csc[1] + ib];
			// cells (inside the chunk) of the first and the last molecule of the tile
			const uint32_t first = csc[1] + i0, last = csc[1] + min(i0 + 15u, n_mol - 1u);
			int kxa = 0; while (kxa + 1 < nc && csc[kxa + 2] <= first) kxa++;
			int kxb = kxa; while (kxb + 1 < nc && csc[kxb + 2] <= last) kxb++;
			// A fragments (constant over the tile loop). Slots k: t = 0..2 -> -2hi*x'h, -2hi*x'l, -2lo*x'h, {sh, sl, 0}*1; t = 3 -> 1*s'h, 1*s'l, 0, 0
			uint32_t A1[4], A2[4];
			{
				const C xa_ = ma.x - ox, ya_ = ma.y - oy, za_ = ma.z - oz, xb_ = mb.x - ox, yb_ = mb.y - oy, zb_ = mb.z - oz;
				const C ca = t == 0 ? xa_ : t == 1 ? ya_ : za_, cb = t == 0 ? xb_ : t == 1 ? yb_ : zb_;
				const C cah = tf32_hi(ca), cbh = tf32_hi(cb);
				const C sa = fmaf(xa_, xa_, fmaf(ya_, ya_, za_ * za_)), sb = fmaf(xb_, xb_, fmaf(yb_, yb_, zb_ * zb_));
				const C sah = tf32_hi(sa), sbh = tf32_hi(sb);
				const C ea = t == 0 ? sah : t == 1 ? sa - sah : 0.f, eb = t == 0 ? sbh : t == 1 ? sb - sbh : 0.f;
				if (t < 3) {
					A1[0] = __float_as_uint(-2.f * cah); A1[1] = __float_as_uint(-2.f * cbh); A1[2] = A1[0]; A1[3] = A1[1];
					A2[0] = __float_as_uint(-2.f * (ca - cah)); A2[1] = __float_as_uint(-2.f * (cb - cbh)); A2[2] = __float_as_uint(ea); A2[3] = __float_as_uint(eb);
				} else {
					A1[0] = A1[1] = A1[2] = A1[3] = 0x3F800000u;
					A2[0] = A2[1] = A2[2] = A2[3] = 0u;
				}
			}
			const uint32_t wa0 = lw_s + (uint32_t) lane * 2u, wb0 = wa0 + 64u;      // lists lane (row g) and 32 + lane (row g + 8)
			uint32_t wa = wa0, wb = wb0;
			// phase 2 works on row (t < 2 ? g : g + 8) with the lists of the lanes L0 = 4g + 2(t & 1) and L0 + 1
			const int h = t >> 1, L0 = (lane & ~3) + 2 * (t & 1);
			const V4 mi = h ? mb : ma;
			const bool have = h ? have_b : have_a;
			const uint32_t cid_i = __float_as_uint(mi.w) & 0xFFFFu;
			const uint32_t p0 = lw_s + (uint32_t) (L0 + 32 * h) * 2u, p1 = p0 + 2u;
			double Fd[3] = {0., 0., 0.}, Vd[3] = {0., 0., 0.}, Ud = 0.;
			int r = 0;
			for (;;) {
				if (r < 9) {
					uint32_t kb = cs[r * (CH_MAXC + 3) + kxa], ke = cs[r * (CH_MAXC + 3) + kxb + 3];      // cells kxa-1 .. kxb+1 of this row
					const int ntr = (int) ((ke - kb + 7u) >> 3);
					const uint32_t room = (uint32_t) (LCAP - 2 * ntr) * LS;
					const bool fits = (wa - wa0 <= room) && (wb - wb0 <= room);
					if (__all_sync(0xffffffffu, fits)) {
						uint32_t pfa = pf_s + (uint32_t) (r * WP + (int) kb + g) * 32u + (uint32_t) t * 8u;
						uint32_t exa = win_s + (uint32_t) (r * W + (int) kb + 2 * t) * VB;
						#pragma unroll 2
						for (int k = 0; k < ntr; k++) {
							const float2 v = lds_v2(pfa);
							float c[4] = {0.f, 0.f, 0.f, 0.f};
							mma_tf32(c, A1, __float_as_uint(v.x), __float_as_uint(v.y));
							mma_tf32(c, A2, __float_as_uint(v.x), 0x3F800000u);
							if (__float_as_int(c[0]) < 0) { sts16(wa, exa); wa += LS; }
							if (__float_as_int(c[1]) < 0) { sts16(wa, exa + VB); wa += LS; }
							if (__float_as_int(c[2]) < 0) { sts16(wb, exa); wb += LS; }
							if (__float_as_int(c[3]) < 0) { sts16(wb, exa + VB); wb += LS; }
							pfa += 256u; exa += 8u * VB;
						}
						r++;
						continue;
					}
				}
				// ---- phase 2: the survivors of my row; the reference's test decides, LJ 12-6 on the pairs that pass ----
				const uint32_t na = (wa - wa0) / LS, nb = (wb - wb0) / LS;
				const uint32_t na0 = __shfl_sync(0xffffffffu, na, L0), nb0 = __shfl_sync(0xffffffffu, nb, L0);
				const uint32_t na1 = __shfl_sync(0xffffffffu, na, L0 + 1), nb1 = __shfl_sync(0xffffffffu, nb, L0 + 1);
				const uint32_t n0 = h ? nb0 : na0, n1 = h ? nb1 : na1, total = n0 + n1;
				__syncwarp();
				{
					C fx[2] = {0.f, 0.f}, fy[2] = {0.f, 0.f}, fz[2] = {0.f, 0.f}, vx[2] = {0.f, 0.f}, vy[2] = {0.f, 0.f}, vz[2] = {0.f, 0.f}, up[2] = {0.f, 0.f}, np[2] = {0.f, 0.f};
					C eps24 = a.eps24_0, sig2 = a.sig2_0, shift6 = a.shift6_0;
					#pragma unroll 1
					for (uint32_t e = 0; e < total; e += 2) {
						uint32_t j[2];
						#pragma unroll
						for (int u = 0; u < 2; u++) {
							const uint32_t ee = e + (uint32_t) u;
							const uint32_t la = ee < n0 ? p0 + ee * LS : p1 + (ee - n0) * LS;
							j[u] = ee < total ? lds16(la) : sentinel;
						}
						#pragma unroll
						for (int u = 0; u < 2; u++) {
							const V4 mj = lds_v4<C>(j[u]);
							const C dx = mi.x - mj.x, dy = mi.y - mj.y, dz = mi.z - mj.z;
							const C dz2 = __fmul_rn(dz, dz);
							const C r2 = __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, dz2));      // == molpair_r2: bit-identical cutoff decision
							const bool pass = (r2 < a.ljrc2) && (r2 != 0.f);
							if (!SINGLE) {
								const size_t ti = (size_t) cid_i * a.n_lj + (__float_as_uint(mj.w) & 0xFFFFu);
								eps24 = a.eps24[ti]; sig2 = a.sig2[ti]; shift6 = a.shift6[ti];
							}
							const C r2inv = rcp_fast<C>(r2);
							const C lj2 = sig2 * r2inv, lj6 = lj2 * lj2 * lj2, lj12 = lj6 * lj6, lj12m6 = lj12 - lj6;
							C scale = r2inv * (lj12 + lj12m6);
							if (!SINGLE) scale *= eps24;
							scale = pass ? scale : 0.f;
							const C dx2 = dx * dx, dy2 = dy * dy;
							fx[u] = fmaf(dx, scale, fx[u]); fy[u] = fmaf(dy, scale, fy[u]); fz[u] = fmaf(dz, scale, fz[u]);
							vx[u] = fmaf(dx2, scale, vx[u]); vy[u] = fmaf(dy2, scale, vy[u]); vz[u] = fmaf(dz2, scale, vz[u]);
							if (SINGLE) { up[u] += pass ? lj12m6 : 0.f; np[u] += pass ? 1.f : 0.f; }
							else up[u] += pass ? fmaf(eps24, lj12m6, shift6) : 0.f;
						}
					}
					if (SINGLE) {
						const C f24 = eps24;
						fx[0] = f24 * (fx[0] + fx[1]); fy[0] = f24 * (fy[0] + fy[1]); fz[0] = f24 * (fz[0] + fz[1]);
						vx[0] = f24 * (vx[0] + vx[1]); vy[0] = f24 * (vy[0] + vy[1]); vz[0] = f24 * (vz[0] + vz[1]);
						up[0] = f24 * (up[0] + up[1]) + (np[0] + np[1]) * shift6;
						fx[1] = fy[1] = fz[1] = vx[1] = vy[1] = vz[1] = up[1] = 0.f;
					}
					Fd[0] += (double) fx[0] + (double) fx[1]; Fd[1] += (double) fy[0] + (double) fy[1]; Fd[2] += (double) fz[0] + (double) fz[1];
					Vd[0] += (double) vx[0] + (double) vx[1]; Vd[1] += (double) vy[0] + (double) vy[1]; Vd[2] += (double) vz[0] + (double) vz[1];
					Ud += (double) up[0] + (double) up[1];
				}
				__syncwarp();      // the lists are rewritten by other lanes from here on
				wa = wa0; wb = wb0;
				if (r >= 9) break;
			}
			// ---- 4. results: the two lanes of a row add up; every pair is visited from both of its ends ----
			acc[0] += 0.5 * Ud;
			acc[1] += 0.5 * (Vd[0] + Vd[1] + Vd[2]);
			#pragma unroll
			for (int d = 0; d < 3; d++) { Fd[d] += __shfl_xor_sync(0xffffffffu, Fd[d], 1); Vd[d] += __shfl_xor_sync(0xffffffffu, Vd[d], 1); }
			if (have && (t & 1) == 0) {
				const size_t cp = a.cap;
				const uint32_t s = s_rowbase[4] + csc[1] + (h ? ib : ia);
				a.F[s] = Fd[0]; a.F[cp + s] = Fd[1]; a.F[2 * cp + s] = Fd[2];
				a.Vi[s] = 0.5 * Vd[0]; a.Vi[cp + s] = 0.5 * Vd[1]; a.Vi[2 * cp + s] = 0.5 * Vd[2];
				if (a.M) { a.M[s] = 0.; a.M[cp + s] = 0.; a.M[2 * cp + s] = 0.; }
			}
		}
		block_reduce_store<2>(acc, a.partials, ch);
		__syncthreads();      // windows, prefilter records, lists and cs are reused by the next chunk
	}
}

}  // namespace

int reduce_force_chunks(ls1b200_ctx* c, uint32_t nrows, int per_warp, bool kin = false, bool accumulate = false);      // force_1clj_tiled.cu

// mixed-precision ensembles of 1CLJ components whose cells are wider than the cutoff by more than FP32 rounding (otherwise a
// molecule two cells away could pass the FP32 test although the reference never pairs those cells): everything else takes
// the tiled kernel.
int force_1clj_mma(ls1b200_ctx* c) {
	const Geom& g = c->geom;
	bool ok = c->mixed;
	for (int d = 0; d < 3; d++) ok = ok && (g.clen[d] > c->cfg.cutoff_lj * (1. + 1e-5));
	if (!ok) return force_1clj_tiled(c);
	uint32_t nrows = 0;
	LS1_TRY(build_chunks(c, M_NI, M_W, 1, &nrows));
	MArgs a;
	a.mpos = (const float4*) c->mpos; a.cell_start = c->cell_start; a.cell_of = c->cell_of2;
	a.eps24 = c->d_eps24f; a.sig2 = c->d_sig2f; a.shift6 = c->d_shift6f;
	a.F = c->F; a.M = c->rotates ? c->M : nullptr; a.Vi = c->Vi; a.partials = c->partials;
	a.chunks = (const Chunk*) c->chunks; a.n_chunks = c->chunk_row_start + nrows; a.work_counter = c->work_counter;
	a.cap = (uint32_t) c->cap; a.n_lj = c->n_lj; a.nx = g.cells[0]; a.ny = g.cells[1]; a.nz = g.cells[2];
	const double rclj = c->cfg.cutoff_lj;
	a.ljrc2 = (float) (rclj * rclj);
	a.eps24_0 = (float) c->h_eps24[0]; a.sig2_0 = (float) c->h_sig2[0]; a.shift6_0 = (float) c->h_shift6[0];
	for (int d = 0; d < 3; d++) { a.hmin[d] = (float) g.hmin[d]; a.clen[d] = (float) g.clen[d]; }
	auto kern = c->comps.size() == 1 ? k_force_1clj_mma<true> : k_force_1clj_mma<false>;
	LS1_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) M_SMEM));
	int per_sm = 0;
	LS1_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, M_TPB, M_SMEM));
	if (per_sm < 1) return fail("force_1clj_mma: kernel does not fit on an SM");
	const uint32_t grid = (uint32_t) (c->sm_count * per_sm);
	LS1_LAUNCH(c, PH_FORCE, kern, grid, M_TPB, M_SMEM, a);
	return reduce_force_chunks(c, nrows, 0);
}

}  // namespace ls1b200
