// force_1clj_tc.cu -- the production 1CLJ cell-pair kernel of the mixed-precision build: the ~300 molecule-distance tests
// per molecule are done by the 5th-generation tensor cores (tcgen05.mma, accumulator in TMEM) as a conservative PREFILTER;
// the reference's bit-exact FP32 cutoff test and the LJ 12-6 arithmetic run on the survivors only.
//
// Reference: same arithmetic as force_1clj_tiled.cu (VectorizedCellProcessor::_calculatePairs LJ block,
// adapter/VectorizedCellProcessor.cpp:1029-1178; _loopBodyLJ :171-224; calcDistLookup,
// vectorization/SIMD_VectorizedCellProcessorHelpers.h:427-472; FullMolecule::calcFM, molecules/FullMolecule.cpp:548-626).
// The cell pairs are those of the C08 traversal (LinkedCellTraversals/C08BasedTraversals.h:48-99), visited from both ends.
//
// Why: in force_1clj_tiled.cu 70 % of the issued instructions are distance tests (27 cells x 11 molecules per molecule,
// 12 instructions each, of which ~50 pass). Here
//
//   D[i][j] = |ri - rj|^2 - (rc^2 + margin) = sum_k A[i][k] * B[j][k],   k = 0 .. 15  (one tcgen05.mma kind::f16, K = 16)
//
// for the 128 molecules i of a chunk (the M dimension = the 128 TMEM lanes) against 128 window records j per instruction.
// Coordinates are taken relative to the chunk centre and scaled by a power of two; every FP32 value is split into an FP16
// high part and an FP16 remainder (3-term products hi*hi + hi*lo + lo*hi, the squares |r|^2 as hi + lo), products are exact
// in the FP32 accumulator: the error of the expansion is below 4e-6 R^2 (R = half extent of the chunk's windows), the margin
// is 2^-14 R^2 = 6.1e-5 R^2, so no pair inside the cutoff is ever missed. A thread reads the row of its molecule from TMEM
// (tcgen05.ld 32x32b: lane = molecule, 32 columns = 32 candidates), turns the signs into a bit mask, restricts it to the
// 3 cells of its own neighbourhood and appends the survivors (pairs in range, the molecule itself, < 1 % near misses) to its
// neighbour list in shared memory. Phase 2 is the one of the tiled kernel: it repeats the reference's test bit for bit
// (fma(dx,dx,fma(dy,dy,dz*dz)) < rc^2, != 0) before anything is accumulated. Tensor cores are a filter, not the arithmetic
// of record: forces, virial and energy are FP32 FMA work folded into FP64 exactly as before.
//
// Per chunk (tile.cuh: <= 128 molecules of consecutive cells of one x-row, 9 windows <= 160 records, TMA bulk copies):
//   1. windows by TMA; A (128 x 16 halves) and B (9 rows x 160 columns x 16 halves) operands written to shared memory in
//      the canonical K-major no-swizzle layout (8-row core matrices: 16 B per row, K chunks 128 B apart, groups 256 B apart);
//   2. 12 passes of 128 columns, double buffered in TMEM (2 x 128 columns): lane 0 of a fifth warp issues tcgen05.mma + commit for pass
//      p + 2 as soon as the four molecule warps have read pass p (mbarriers, no CTA-wide barrier inside the pass loop);
//   3. phase 2 over the lists (whenever a list could overflow, and at the end); F, Vi, partial sums as in the tiled kernel.
#include <cuda_fp16.h>

#include "tile.cuh"

namespace ls1b200 {

namespace {

// 128 molecule threads + one warp whose lane 0 issues the MMAs; windows of <= 160 records = 5 blocks of 32 columns per row
constexpr int T_NI = 128, T_TPB = 160, T_W = 160, T_CS = 160, T_LCAP = 64, T_NBLK = 9 * T_CS / 32, T_NPASS = (T_NBLK + 3) / 4, T_MB = 24;
constexpr uint32_t T_VB = 16, T_LSTRIDE = T_NI * 2;
constexpr size_t T_OFF_A = ((size_t) (9 * T_W + 1) * 16 + 1023) & ~(size_t) 1023;
constexpr size_t T_OFF_B = T_OFF_A + 128 * 32;
constexpr size_t T_OFF_LIST = T_OFF_B + (size_t) T_NPASS * 128 * 32;
constexpr size_t T_OFF_MASK = T_OFF_LIST + (size_t) (T_LCAP + 1) * T_NI * 2;      // [T_MB][128] u32 survivor masks waiting to be decoded
constexpr size_t T_OFF_BASE = T_OFF_MASK + (size_t) T_MB * T_NI * 4;              // [4 warps][T_MB] u32: window address of column 0 of the block
constexpr size_t T_OFF_CS = T_OFF_BASE + (size_t) 4 * T_MB * 4;
constexpr size_t T_SMEM = T_OFF_CS + (size_t) 12 * (CH_MAXC + 3) * 4;
constexpr uint32_t T_TMEM_COLS = 256;

struct TcArgs {
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
	float scale;                 // power of two that brings the cell length into [1, 2)
};

// shared-memory matrix descriptor: K-major, no swizzle, LBO = 128 B (between the two 16-byte K chunks), SBO = 256 B
// (between 8-row groups), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t tc_desc(uint32_t addr) {
	return (uint64_t) ((addr >> 4) & 0x3FFFu) | ((uint64_t) (128u >> 4) << 16) | ((uint64_t) (256u >> 4) << 32) | (1ull << 46);
}
// byte offset of operand row `row` (its first K chunk; the second one is 128 B further)
__device__ __forceinline__ uint32_t tc_row_off(uint32_t row) { return (row >> 3) * 256u + (row & 7u) * 16u; }
// instruction descriptor of kind::f16: D = F32, A = B = F16, both K-major, M = 128
__device__ __forceinline__ uint32_t tc_idesc(uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
	             :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
	             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
	               "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
	             : "r"(taddr) : "memory");
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
	const __half2 h = __floats2half2_rn(lo, hi);
	return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float h_hi(float x) { return __half2float(__float2half_rn(x)); }
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
	asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// LJ 12-6 of one list entry (shared-memory address of the candidate); the reference's test decides whether it counts
#define TC_PAIR(u, addr)                                                                          \
	{                                                                                              \
		const float4 mj = lds_v4<float>(addr);                                                      \
		const float dx = mi.x - mj.x, dy = mi.y - mj.y, dz = mi.z - mj.z;                            \
		const float dz2 = __fmul_rn(dz, dz);                                                         \
		const float r2 = __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, dz2));                                  \
		const bool pass = (r2 < a.ljrc2) && (r2 != 0.f);                                             \
		if (!SINGLE) {                                                                               \
			const size_t ti = (size_t) cid_i * a.n_lj + (__float_as_uint(mj.w) & 0xFFFFu);            \
			eps24 = a.eps24[ti]; sig2 = a.sig2[ti]; shift6 = a.shift6[ti];                             \
		}                                                                                            \
		const float r2inv = rcp_fast<float>(r2);                                                     \
		const float lj2 = sig2 * r2inv, lj6 = lj2 * lj2 * lj2, lj12 = lj6 * lj6, lj12m6 = lj12 - lj6;  \
		float scale = r2inv * (lj12 + lj12m6);                                                       \
		if (!SINGLE) scale *= eps24;                                                                 \
		scale = pass ? scale : 0.f;                                                                  \
		const float dx2 = dx * dx, dy2 = dy * dy;                                                    \
		fx[u] = fmaf(dx, scale, fx[u]); fy[u] = fmaf(dy, scale, fy[u]); fz[u] = fmaf(dz, scale, fz[u]); \
		vx[u] = fmaf(dx2, scale, vx[u]); vy[u] = fmaf(dy2, scale, vy[u]); vz[u] = fmaf(dz2, scale, vz[u]); \
		if (SINGLE) { up[u] += pass ? lj12m6 : 0.f; np[u] += pass ? 1.f : 0.f; }                     \
		else up[u] += pass ? fmaf(eps24, lj12m6, shift6) : 0.f;                                      \
	}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }

template <bool SINGLE>
__global__ void __launch_bounds__(T_TPB, 2) k_force_1clj_tc(TcArgs a) {
	typedef float4 V4;
	constexpr int TPB = T_TPB, NI = T_NI, W = T_W, CS = T_CS, LCAP = T_LCAP, MB = T_MB, BPR = T_CS / 32;
	constexpr uint32_t VB = T_VB, LSTRIDE = T_LSTRIDE;
	extern __shared__ __align__(1024) unsigned char smem_raw[];
	V4* win = reinterpret_cast<V4*>(smem_raw);                                   // [9][W] exact records (+ sentinel at [9*W])
	uint32_t* cs = reinterpret_cast<uint32_t*>(smem_raw + T_OFF_CS);             // [9][CH_MAXC+3]
	__shared__ uint32_t s_rowbase[9];
	__shared__ __align__(8) unsigned long long s_bar, s_full[2], s_free[2];
	__shared__ uint32_t s_chunk, s_tmem;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const bool issuer = wid == NI / 32;                     // warp 4: lane 0 issues the MMAs, nothing else
	const uint32_t win_s = smem_u32(win), opa_s = smem_u32(smem_raw + T_OFF_A), opb_s = smem_u32(smem_raw + T_OFF_B);
	const uint32_t list_s = smem_u32(smem_raw + T_OFF_LIST) + threadIdx.x * 2u;
	const uint32_t mask_s = smem_u32(smem_raw + T_OFF_MASK) + threadIdx.x * 4u;                   // [MB][128] u32, my column
	const uint32_t base_s = smem_u32(smem_raw + T_OFF_BASE) + (uint32_t) (wid & 3) * (MB * 4);     // [MB] u32 of my warp
	const uint32_t bar = smem_u32(&s_bar);
	const uint32_t full0 = smem_u32(&s_full[0]), full1 = smem_u32(&s_full[1]), free0 = smem_u32(&s_free[0]), free1 = smem_u32(&s_free[1]);
	const uint32_t sentinel = win_s + (uint32_t) (9 * W) * VB;
	const uint32_t n_chunks = *a.n_chunks;
	if (threadIdx.x == 0) {
		mbar_init(bar, 1); mbar_init(full0, 1); mbar_init(full1, 1); mbar_init(free0, NI / 32); mbar_init(free1, NI / 32);
		win[9 * W] = make_float4(1e18f, 1e18f, 1e18f, 0.f);
	}
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	if (wid == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "n"(T_TMEM_COLS) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	fence_proxy_async();
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = s_tmem;
	const uint32_t tlane = tmem + ((uint32_t) (32 * (wid & 3)) << 16);      // this warp's quarter of the 128 TMEM lanes
	uint32_t parity = 0, gen = 0;      // gen: passes so far (all chunks); pass gen uses TMEM buffer gen & 1, barrier phase gen >> 1
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
						float fx = 0.f, fy = 0.f, fz = 0.f, vx = 0.f, vy = 0.f, vz = 0.f, up = 0.f;
						for (uint32_t j = j0; j < min(je, j0 + 32); j++) {
							const V4 mj = a.mpos[j];
							const float dx = mi.x - mj.x, dy = mi.y - mj.y, dz = mi.z - mj.z;
							const float r2 = molpair_r2<float>(mi.x, mi.y, mi.z, mj.x, mj.y, mj.z);
							if ((r2 < a.ljrc2) && (r2 != 0.f)) {
								float eps24 = a.eps24_0, sig2 = a.sig2_0, shift6 = a.shift6_0;
								if (!SINGLE) { const size_t ti = (size_t) cid_i * a.n_lj + (__float_as_uint(mj.w) & 0xFFFFu); eps24 = a.eps24[ti]; sig2 = a.sig2[ti]; shift6 = a.shift6[ti]; }
								const float r2inv = rcp_fast<float>(r2);
								const float lj2 = sig2 * r2inv, lj6 = lj2 * lj2 * lj2, lj12 = lj6 * lj6, lj12m6 = lj12 - lj6;
								const float scale = eps24 * r2inv * (lj12 + lj12m6);
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
		// origin of the prefilter coordinates (chunk centre), scale, threshold with the error margin of the FP16 expansion
		const float sc = a.scale;
		const float ox = a.hmin[0] + ((float) xa + 0.5f * (float) nc) * a.clen[0];
		const float oy = a.hmin[1] + ((float) (R % (uint32_t) a.ny) + 0.5f) * a.clen[1];
		const float oz = a.hmin[2] + ((float) (R / (uint32_t) a.ny) + 0.5f) * a.clen[2];
		const float hx = (0.5f * (float) nc + 1.5f) * a.clen[0];
		const float thr = (a.ljrc2 + 6.1035156e-5f * (hx * hx + 2.25f * (a.clen[1] * a.clen[1] + a.clen[2] * a.clen[2]))) * sc * sc;
		const uint32_t* csc = cs + 4 * (CH_MAXC + 3);          // centre row (relative record indices)
		const uint32_t n_mol = csc[nc + 1] - csc[1];
		const bool have = threadIdx.x < n_mol;                 // n_mol <= 128: never true for the issuer warp
		const uint32_t trel = csc[1] + threadIdx.x;
		int kx = 0;                                            // cell (inside the chunk) of my molecule
		while (kx + 1 < nc && csc[kx + 2] <= trel) kx++;
		// the cells of the first and the last molecule of my warp: the columns any of its lanes may need
		const int kxa_w = __shfl_sync(0xffffffffu, kx, 0), kxb_w = __shfl_sync(0xffffffffu, kx, 31);
		const bool warp_active = !issuer && (uint32_t) (32 * wid) < n_mol;
		// blocks (32 columns of one window row) some lane of my warp needs: bit gb of need
		unsigned long long need = 0ull;
		{
			bool n0 = false, n1 = false;
			if (warp_active) {
				int gb = lane;
				if (gb < T_NBLK) { const int r = gb / BPR, c0 = (gb - r * BPR) * 32; const uint32_t* csr = cs + r * (CH_MAXC + 3); n0 = (int) csr[kxa_w] < c0 + 32 && (int) csr[kxb_w + 3] > c0; }
				gb = lane + 32;
				if (gb < T_NBLK) { const int r = gb / BPR, c0 = (gb - r * BPR) * 32; const uint32_t* csr = cs + r * (CH_MAXC + 3); n1 = (int) csr[kxa_w] < c0 + 32 && (int) csr[kxb_w + 3] > c0; }
			}
			need = (unsigned long long) __ballot_sync(0xffffffffu, n0) | ((unsigned long long) __ballot_sync(0xffffffffu, n1) << 32);
		}
		mbar_wait(bar, parity); parity ^= 1u;
		// ---- 2. operands: A row of my molecule, B rows of every window record ----
		V4 mi = make_float4(1e18f, 1e18f, 1e18f, 0.f);      // an absent molecule: its list stays empty
		if (have) mi = win[4 * W + trel];
		if (!issuer) {
			uint32_t lo[4] = {0u, 0u, 0u, 0u}, hi[4] = {0u, 0u, 0u, 0u};
			if (have) {
				const float x = (mi.x - ox) * sc, y = (mi.y - oy) * sc, z = (mi.z - oz) * sc;
				const float s = fmaf(x, x, fmaf(y, y, z * z));
				const float xh = h_hi(x), yh = h_hi(y), zh = h_hi(z), sh = h_hi(s);
				lo[0] = pack_h2(-2.f * xh, -2.f * xh); lo[1] = pack_h2(-2.f * (x - xh), -2.f * yh); lo[2] = pack_h2(-2.f * yh, -2.f * (y - yh)); lo[3] = pack_h2(-2.f * zh, -2.f * zh);
				hi[0] = pack_h2(-2.f * (z - zh), sh); hi[1] = pack_h2(s - sh, 1.f); hi[2] = pack_h2(1.f, 0.f);
			}
			const uint32_t o = opa_s + tc_row_off(threadIdx.x);
			sts_v4(o, lo[0], lo[1], lo[2], lo[3]); sts_v4(o + 128u, hi[0], hi[1], hi[2], hi[3]);
		}
		for (int idx = threadIdx.x; idx < 9 * W; idx += TPB) {
			const int r = idx / W, k = idx - r * W;
			if (k >= (int) cs[r * (CH_MAXC + 3) + nc + 2]) continue;
			const V4 m = win[idx];
			const float x = (m.x - ox) * sc, y = (m.y - oy) * sc, z = (m.z - oz) * sc;
			const float s = fmaf(x, x, fmaf(y, y, z * z)) - thr;
			const float xh = h_hi(x), yh = h_hi(y), zh = h_hi(z), sh = h_hi(s);
			const uint32_t o = opb_s + tc_row_off((uint32_t) (r * CS + k));
			sts_v4(o, pack_h2(xh, x - xh), pack_h2(xh, yh), pack_h2(y - yh, yh), pack_h2(zh, z - zh));
			sts_v4(o + 128u, pack_h2(zh, 1.f), pack_h2(1.f, sh), pack_h2(s - sh, 0.f), 0u);
		}
		fence_proxy_async();      // generic-proxy operand writes -> async-proxy reads of the tensor core
		tc_fence_before();
		__syncthreads();
		if (issuer) {
			// ---- 3a. the MMA issuer: pass p goes to TMEM buffer p & 1 as soon as the four warps have read pass p - 2 from it ----
			if (lane == 0) {
				tc_fence_after();
				const uint64_t da = tc_desc(opa_s);
				for (int p = 0; p < T_NPASS; p++) {
					const uint32_t g = gen + (uint32_t) p;
					if (g >= 2u) { mbar_wait((g & 1u) ? free1 : free0, ((g >> 1) + 1u) & 1u); tc_fence_after(); }
					const uint32_t n = (p == T_NPASS - 1) ? (uint32_t) (T_NBLK * 32 - (T_NPASS - 1) * 128) : 128u;
					tc_mma(tmem + (g & 1u) * 128u, da, tc_desc(opb_s + (uint32_t) p * 4096u), tc_idesc(n));
					tc_commit((g & 1u) ? full1 : full0);
				}
			}
			gen += T_NPASS;
			__syncwarp();
		} else {
			// ---- 3b. the molecule warps: survivor masks of the blocks they need; decode into the lists; phase 2 ----
			double Fd[3] = {0., 0., 0.}, Vd[3] = {0., 0., 0.}, Ud = 0.;
			const uint32_t cid_i = __float_as_uint(mi.w) & 0xFFFFu;
			const uint32_t lbase = list_s;
			uint32_t w = lbase;
			uint32_t nmb = 0, pend = 0;      // mask blocks waiting in the buffer (warp uniform); survivors in them (mine)
			auto phase2 = [&]() {
				const int cnt = (int) ((w - lbase) / LSTRIDE);
				float fx[2] = {0.f, 0.f}, fy[2] = {0.f, 0.f}, fz[2] = {0.f, 0.f}, vx[2] = {0.f, 0.f}, vy[2] = {0.f, 0.f}, vz[2] = {0.f, 0.f}, up[2] = {0.f, 0.f}, np[2] = {0.f, 0.f};
				float eps24 = a.eps24_0, sig2 = a.sig2_0, shift6 = a.shift6_0;
				if (cnt & 1) sts16(w, sentinel);      // pad odd lists with the far-away record (fails the test)
				#pragma unroll 1
				for (uint32_t e = lbase; e < w; e += 2 * LSTRIDE) {
					const uint32_t i0 = lds16(e), i1 = lds16(e + LSTRIDE);
					TC_PAIR(0, i0);
					TC_PAIR(1, i1);
				}
				if (SINGLE) {
					const float f24 = eps24;
					fx[0] = f24 * (fx[0] + fx[1]); fy[0] = f24 * (fy[0] + fy[1]); fz[0] = f24 * (fz[0] + fz[1]);
					vx[0] = f24 * (vx[0] + vx[1]); vy[0] = f24 * (vy[0] + vy[1]); vz[0] = f24 * (vz[0] + vz[1]);
					up[0] = f24 * (up[0] + up[1]) + (np[0] + np[1]) * shift6;
					fx[1] = fy[1] = fz[1] = vx[1] = vy[1] = vz[1] = up[1] = 0.f;
				}
				Fd[0] += (double) fx[0] + (double) fx[1]; Fd[1] += (double) fy[0] + (double) fy[1]; Fd[2] += (double) fz[0] + (double) fz[1];
				Vd[0] += (double) vx[0] + (double) vx[1]; Vd[1] += (double) vy[0] + (double) vy[1]; Vd[2] += (double) vz[0] + (double) vz[1];
				Ud += (double) up[0] + (double) up[1];
				w = lbase;
			};
			// turns the buffered masks into list entries: every lane walks its own words, so lanes with few survivors in one block
			// do not wait for lanes with many
			auto decode = [&]() {
				__syncwarp();      // base_s entries written by lane 0
				uint32_t idx = 0, m = 0, base = 0;
				for (;;) {
					while (m == 0 && idx < nmb) { m = lds_u32(mask_s + idx * (NI * 4)); base = lds_u32(base_s + idx * 4); idx++; }
					if (m == 0) break;
					const int k = __clz(m);
					sts16(w, base + (uint32_t) k * VB); w += LSTRIDE;
					m &= ~(0x80000000u >> k);
				}
				__syncwarp();
				nmb = 0; pend = 0;
			};
			#pragma unroll 1
			for (int p = 0; p < T_NPASS; p++, gen++) {
				mbar_wait((gen & 1u) ? full1 : full0, (gen >> 1) & 1u);
				tc_fence_after();
				uint32_t nb4 = (uint32_t) (need >> (4 * p)) & 15u;
				while (nb4) {
					const int q = __ffs(nb4) - 1; nb4 &= nb4 - 1u;
					const int gb = 4 * p + q;
					const int r = gb / BPR, c0 = (gb - r * BPR) * 32;      // window row, first column (record) of the block
					const uint32_t* csr = cs + r * (CH_MAXC + 3);
					uint32_t v[32];
					tc_ld32(tlane + (gen & 1u) * 128u + (uint32_t) (q * 32), v);
					uint32_t m = 0;
					#pragma unroll
					for (int k = 0; k < 32; k++) m = __funnelshift_l(v[k], m, 1);      // column k -> bit 31 - k
					// my own neighbourhood: records [kb, ke) of this row
					int lo = 0, hi = 0;
					if (have) { lo = max((int) csr[kx] - c0, 0); hi = min((int) csr[kx + 3] - c0, 32); }
					const uint32_t rm = hi > lo ? ((hi - lo == 32 ? 0xFFFFFFFFu : ((1u << (hi - lo)) - 1u)) << (32 - hi)) : 0u;
					m &= rm;
					const uint32_t pc = (uint32_t) __popc(m);
					if (nmb == (uint32_t) MB || __any_sync(0xffffffffu, (w - lbase) / LSTRIDE + pend + pc > (uint32_t) LCAP)) {
						decode();
						if (__any_sync(0xffffffffu, (w - lbase) / LSTRIDE + pc > (uint32_t) LCAP)) phase2();
					}
					asm volatile("st.shared.u32 [%0], %1;" :: "r"(mask_s + nmb * (NI * 4)), "r"(m) : "memory");
					if (lane == 0) asm volatile("st.shared.u32 [%0], %1;" :: "r"(base_s + nmb * 4), "r"(win_s + (uint32_t) (r * W + c0) * VB) : "memory");
					nmb++; pend += pc;
				}
				tc_fence_before();
				__syncwarp();
				if (lane == 0) mbar_arrive((gen & 1u) ? free1 : free0);      // my warp has read this pass's TMEM buffer
			}
			decode();
			phase2();
			// ---- 4. results ----
			if (have) {
				const size_t cp = a.cap;
				const uint32_t s = s_rowbase[4] + trel;
				acc[0] += 0.5 * Ud;                          // every pair is visited from both of its ends
				acc[1] += 0.5 * (Vd[0] + Vd[1] + Vd[2]);
				a.F[s] = Fd[0]; a.F[cp + s] = Fd[1]; a.F[2 * cp + s] = Fd[2];
				a.Vi[s] = 0.5 * Vd[0]; a.Vi[cp + s] = 0.5 * Vd[1]; a.Vi[2 * cp + s] = 0.5 * Vd[2];
				if (a.M) { a.M[s] = 0.; a.M[cp + s] = 0.; a.M[2 * cp + s] = 0.; }
			}
		}
		block_reduce_store<2>(acc, a.partials, ch);
		__syncthreads();      // windows, operands, lists and cs are reused by the next chunk
	}
	tc_fence_before();
	__syncthreads();
	if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(T_TMEM_COLS) : "memory");
}

}  // namespace

int reduce_force_chunks(ls1b200_ctx* c, uint32_t nrows, int per_warp, bool kin = false, bool accumulate = false);      // force_1clj_tiled.cu

// mixed-precision ensembles of 1CLJ components whose cells are wider than the cutoff by more than FP32 rounding (otherwise a
// molecule two cells away could pass the FP32 test although the reference never pairs those cells): everything else takes
// the tiled kernel.
int force_1clj_tc(ls1b200_ctx* c) {
	const Geom& g = c->geom;
	bool ok = c->mixed;
	double lmax = 0.;
	for (int d = 0; d < 3; d++) { ok = ok && (g.clen[d] > c->cfg.cutoff_lj * (1. + 1e-5)); lmax = std::max(lmax, g.clen[d]); }
	// the FP16 operands hold |coordinate| * scale <= 14 cells * [1, 2): cells of very different edge lengths would overflow the margin estimate
	for (int d = 0; d < 3; d++) ok = ok && (g.clen[d] * 4. > lmax);
	if (!ok) return force_1clj_tiled(c);
	uint32_t nrows = 0;
	LS1_TRY(build_chunks(c, T_NI, T_W, 1, &nrows));
	TcArgs a;
	a.mpos = (const float4*) c->mpos; a.cell_start = c->cell_start; a.cell_of = c->cell_of2;
	a.eps24 = c->d_eps24f; a.sig2 = c->d_sig2f; a.shift6 = c->d_shift6f;
	a.F = c->F; a.M = c->rotates ? c->M : nullptr; a.Vi = c->Vi; a.partials = c->partials;
	a.chunks = (const Chunk*) c->chunks; a.n_chunks = c->chunk_row_start + nrows; a.work_counter = c->work_counter;
	a.cap = (uint32_t) c->cap; a.n_lj = c->n_lj; a.nx = g.cells[0]; a.ny = g.cells[1]; a.nz = g.cells[2];
	const double rclj = c->cfg.cutoff_lj;
	a.ljrc2 = (float) (rclj * rclj);
	a.eps24_0 = (float) c->h_eps24[0]; a.sig2_0 = (float) c->h_sig2[0]; a.shift6_0 = (float) c->h_shift6[0];
	for (int d = 0; d < 3; d++) { a.hmin[d] = (float) g.hmin[d]; a.clen[d] = (float) g.clen[d]; }
	a.scale = (float) exp2(-floor(log2(lmax)));
	auto kern = c->comps.size() == 1 ? k_force_1clj_tc<true> : k_force_1clj_tc<false>;
	LS1_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) T_SMEM));
	LS1_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int) cudaSharedmemCarveoutMaxShared));
	// two CTAs per SM: 2 x 105 KB of shared memory, 2 x 256 of the 512 TMEM columns (persistent CTAs pull chunks from a counter,
	// so the grid size is not a correctness matter)
	const uint32_t grid = (uint32_t) (c->sm_count * (int) (512 / T_TMEM_COLS));
	LS1_LAUNCH(c, PH_FORCE, kern, grid, T_TPB, T_SMEM, a);
	return reduce_force_chunks(c, nrows, 0);
}

}  // namespace ls1b200
