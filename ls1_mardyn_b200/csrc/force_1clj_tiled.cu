// force_1clj_tiled.cu -- the production 1CLJ cell-pair kernel: neighbour-cell tiles staged in shared memory by TMA bulk
// copies, per-thread neighbour lists in shared memory, no global-memory access inside the pair loops.
//
// Reference: same arithmetic as force_1clj.cu (VectorizedCellProcessor::_calculatePairs LJ block,
// adapter/VectorizedCellProcessor.cpp:1029-1178; _loopBodyLJ :171-224; calcDistLookup,
// vectorization/SIMD_VectorizedCellProcessorHelpers.h:427-472; FullMolecule::calcFM, molecules/FullMolecule.cpp:548-626).
// The cell pairs are those of the C08 traversal (LinkedCellTraversals/C08BasedTraversals.h:48-99), visited from both ends.
//
// Work decomposition ("chunks"): a chunk is a run of consecutive inner cells [xa, xb] of one x-row of the cell grid whose
// molecules (the "i" molecules, <= TPB) become the threads of one CTA pass. Because slots are sorted by cell index with x
// fastest, the candidates "j" of ALL molecules of the chunk are exactly 9 CONTIGUOUS slot ranges (one per (dy, dz) row
// offset, cells xa-1 .. xb+1), at most W records each. k_chunks_count / k_chunks_write build the chunk list on the device
// after every update (greedy, deterministic); chunks that do not fit (a cell denser than the tile) are flagged BIG and
// take a slow generic path inside the same kernel.
//
// Kernel (persistent CTAs pulling chunks from an atomic work counter):
//   1. one thread arms an mbarrier and issues 9 cp.async.bulk (TMA) copies global -> shared of the 9 windows;
//   2. phase 1, thread per i: molecule-distance test of the reference (bit-identical FP32 / FP64 sequence) against its
//      3x9 neighbour cells read from shared memory; hits are appended to a per-thread u16 list in shared memory;
//   3. phase 2: dense LJ 12-6 over the list (positions again from shared memory), FP32 partial sums folded into FP64;
//   4. F, Vi written coalesced; per-chunk partial sums of 6*U and the virial (deterministic final reduction).
#include "tile.cuh"

namespace ls1b200 {

#ifndef LS1_TILE_RI_F
// measured on B200, 16 M molecules (profiles/r01_tile_sweep.md): RI/TPB/W/LCAP = 1/128/176/64 -> 4.98 ms per traversal,
// 2/64/176/64 -> 6.07 (register blocking over molecules halves the instructions per test but not the shared-memory
// wavefronts -- a wavefront is spent per distinct cell range in the warp -- and leaves 10 warps per SM)
#define LS1_TILE_RI_F 1
#define LS1_TILE_TPB_F 128
#define LS1_TILE_W_F 176
#define LS1_TILE_LCAP_F 64
#endif
template <class C> struct TileCfg;
// TPB threads, each owning RI molecules of ONE cell (RI > 1 is register blocking: one shared-memory load of a candidate
// serves RI distance tests); W records per window row; LCAP list entries per molecule.
template <> struct TileCfg<float> { static constexpr int RI = LS1_TILE_RI_F, TPB = LS1_TILE_TPB_F, W = LS1_TILE_W_F, LCAP = LS1_TILE_LCAP_F; };
template <> struct TileCfg<double> { static constexpr int RI = 1, TPB = 64, W = 112, LCAP = 64; };

template <class C>
struct TArgs {
	const typename Vec4<C>::type* mpos;
	const uint32_t* cell_start;
	const uint32_t* cell_of;
	const C* eps24; const C* sig2; const C* shift6;
	double* F; double* M; double* Vi;
	double* partials;
	const Chunk* chunks;
	const uint32_t* n_chunks;     // -> row_start[nrows]
	uint32_t* work_counter;
	uint32_t cap, n_lj;
	int nx, ny, nz;
	C ljrc2;
	C eps24_0, sig2_0, shift6_0;
	// reduced-memory mode (rmm.cu; VCP1CLJRMM.cpp:238-369): the force goes straight into the FP32 velocities, v += f dt / m; F, Vi are not stored
	float* rmm_v; float rmm_dt_m;
	// fused step (ls1b200_simulate): the kinetic sum of Leapfrog::transition2to3 (Leapfrog.cpp:113-143), sum of m (v + dt/2m F)^2 over the
	// molecules this kernel visits, from the force still in registers; the kick itself is applied by the next step's first kernel
	const double* kin_v; double kin_dth; const CompDev* comps; double* partials_kin;      // partials_kin: [chunk][4 warps]
	const uint32_t* kin_perm;      // lazy permutation: v of slot t is still at kin_perm[t] of the previous order (kin_v points there)
	// overlapped multi-GPU step, boundary chunks (BND): the imports of this step's exchange are a second slot range per cell, sorted on
	// their own behind the slots of the main sort (update.cu: do_update_imports); cell_start_b is absolute
	const uint32_t* cell_start_b;
};

// LJ 12-6 of one pair, accumulated into FP32 partial sums (slot u). Force and virial are fused multiply-adds on the common
// factor; with a single component 24 eps is pair independent and is applied once per list (scale1 omits it).
#define LJ_PAIR_ACCUM(u, mj)                                                                   \
	{                                                                                           \
		const C dx = mi.x - (mj).x, dy = mi.y - (mj).y, dz = mi.z - (mj).z;                      \
		const C dx2 = dx * dx, dy2 = dy * dy, dz2 = dz * dz;                                      \
		const C r2 = dx2 + dy2 + dz2;                                                             \
		if (!SINGLE) {                                                                            \
			const size_t ti = (size_t) cid_i * a.n_lj + (wbits<C>((mj).w) & 0xFFFFu);              \
			eps24 = a.eps24[ti]; sig2 = a.sig2[ti]; shift6 = a.shift6[ti];                          \
		}                                                                                         \
		const C r2inv = rcp_fast<C>(r2);                                                          \
		const C lj2 = sig2 * r2inv, lj6 = lj2 * lj2 * lj2, lj12 = lj6 * lj6, lj12m6 = lj12 - lj6;  \
		C scale = r2inv * (lj12 + lj12m6);                                                        \
		if (!SINGLE) scale *= eps24;                                                              \
		fx[u] = fma(dx, scale, fx[u]); fy[u] = fma(dy, scale, fy[u]); fz[u] = fma(dz, scale, fz[u]); \
		vx[u] = fma(dx2, scale, vx[u]); vy[u] = fma(dy2, scale, vy[u]); vz[u] = fma(dz2, scale, vz[u]); \
		if (SINGLE) up[u] += lj12m6; else up[u] += eps24 * lj12m6 + shift6;                       \
	}

template <class C, bool SINGLE, bool RMM = false, bool BND = false>
#ifndef LS1_TILE_MINB
#define LS1_TILE_MINB 6      // measured (gpurun_out/minb, minb2): 1 / 5 / 6 / 7 / 8 resident CTAs asked for -> 89 / 85 / 78 / 71 / 64 registers -> 6.082 / 6.075 / 6.000 / 6.105 / 6.289 ms per step; without the second argument ptxas stops at 64 registers and spills
#endif
#ifdef LS1_TILE_NOMINB
__global__ void __launch_bounds__(TileCfg<C>::TPB) k_force_1clj_tiled(TArgs<C> a) {
#else
__global__ void __launch_bounds__(TileCfg<C>::TPB, LS1_TILE_MINB) k_force_1clj_tiled(TArgs<C> a) {
#endif
	typedef typename Vec4<C>::type V4;
	constexpr int TPB = TileCfg<C>::TPB, W = TileCfg<C>::W, LCAP = TileCfg<C>::LCAP, RI = TileCfg<C>::RI;
	constexpr uint32_t VB = sizeof(V4), LSTRIDE = (uint32_t) TPB * 2u * RI;      // bytes between consecutive entries of one list
	extern __shared__ __align__(128) unsigned char smem_raw[];
	V4* win = reinterpret_cast<V4*>(smem_raw);                                         // [9][W] (+1 sentinel record at [9*W])
	uint16_t* list = reinterpret_cast<uint16_t*>(smem_raw + (size_t) (9 * W + 1) * VB);   // [LCAP+1][RI][TPB]
	// cell_start caches (entries are record indices relative to the window row, <= W): 16-bit in the boundary kernel, whose two tables
	// then take the room of the inner kernel's one -- the same number of CTAs per SM (one less costs 12 %, gpurun_out/ctas)
	typedef typename std::conditional<BND, uint16_t, uint32_t>::type cs_t;
	cs_t* cs_all = reinterpret_cast<cs_t*>(list + (size_t) (LCAP + 1) * RI * TPB);           // [2][12][CH_MAXC+3]: this chunk's and the next one's
	constexpr int CSB = 12 * (CH_MAXC + 3);
	static_assert(!BND || RI == 1, "boundary chunks: one molecule per thread");
	cs_t* csb_all = cs_all + 2 * CSB;                                                        // BND: the same for the imports' ranges
	__shared__ uint32_t s_rowbase_all[2][9];
	// BND: the absolute starts of the imports' 9 rows live in the two spare rows (10, 11) of the imports' table: the footprint is the inner kernel's
	auto rowbaseB_of = [&](int buf) -> uint32_t* { return reinterpret_cast<uint32_t*>(csb_all + buf * CSB + 10 * (CH_MAXC + 3)); };
	constexpr int NR = BND ? 18 : 9;                                                         // candidate ranges per molecule: 9 rows (x 2 with imports)
	__shared__ __align__(8) unsigned long long s_bar;
	__shared__ uint32_t s_next;
	const uint32_t win_s = smem_u32(win), list_s = smem_u32(list) + threadIdx.x * 2, bar = smem_u32(&s_bar);
	const uint32_t n_chunks = *a.n_chunks;
	if (threadIdx.x == 0) {
		mbar_init(bar, 1);
		// sentinel record: a point far away from everything, pads odd lists (all LJ terms underflow to 0)
		win[9 * W] = make_c4<C>((C) 1e18, (C) 1e18, (C) 1e18, (C) 0);
		s_next = atomicAdd(a.work_counter, 1u);
	}
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	fence_proxy_async();
	__syncthreads();
	// cell_start cache of the 9 rows of a chunk (cells xa-1 .. xb+2 -> nc+3 entries each), stored RELATIVE to the first record of
	// the window (the absolute start goes to rowbase): the per-thread row set-up then needs two loads per row and no subtraction
	auto fill_cs = [&](const Chunk& q, cs_t* csq, uint32_t* rbq, cs_t* csbq, uint32_t* rbbq) {
		const int qn = q.ncells;
		const uint32_t qR = q.cell0 / (uint32_t) a.nx;
		const int qxa = (int) (q.cell0 % (uint32_t) a.nx);
		for (int k = threadIdx.x; k < 9 * (qn + 3); k += TPB) {
			const int r = k / (qn + 3), c = k % (qn + 3);
			const uint32_t row = (uint32_t) ((int) qR + (r / 3 - 1) * a.ny + (r % 3 - 1)) * a.nx;
			const uint32_t b0 = a.cell_start[row + qxa - 1];
			csq[r * (CH_MAXC + 3) + c] = (cs_t) (a.cell_start[row + qxa - 1 + c] - b0);
			if (c == 0) rbq[r] = b0;
			if (BND) {
				const uint32_t bb = a.cell_start_b[row + qxa - 1];
				csbq[r * (CH_MAXC + 3) + c] = (cs_t) (a.cell_start_b[row + qxa - 1 + c] - bb);
				if (c == 0) rbbq[r] = bb;
			}
		}
	};
	// The chunk loop is software pipelined: while chunk n is computed, the id of chunk n + 1 is claimed from the work counter,
	// its descriptor is loaded and its cell_start cache is filled (second buffer), so that at the chunk boundary only the TMA
	// load of the windows is exposed -- two barriers per chunk, and no dependent chain of four global round trips.
	uint32_t ch = s_next;
	Chunk cd; cd.cell0 = 0; cd.ncells = 0; cd.flags = 0;
	int cur = 0;
	if (ch < n_chunks) { cd = a.chunks[ch]; if (!(cd.flags & CHUNK_BIG)) fill_cs(cd, cs_all, s_rowbase_all[0], csb_all, rowbaseB_of(0)); }
	__syncthreads();
	uint32_t parity = 0;
	while (ch < n_chunks) {
		cs_t* cs = cs_all + cur * CSB;
		uint32_t* s_rowbase = s_rowbase_all[cur];
		cs_t* csb = csb_all + cur * CSB;
		uint32_t* s_rowbaseB = rowbaseB_of(cur);
		const bool big = (cd.flags & CHUNK_BIG) != 0;
		const int nc = cd.ncells;
		const uint32_t R = cd.cell0 / (uint32_t) a.nx;                 // row index cz*ny + cy
		const int xa = (int) (cd.cell0 % (uint32_t) a.nx);
		if (threadIdx.x == 0) {
			if (!big) {
				uint32_t total = 0;
				for (int r = 0; r < 9; r++) total += ((uint32_t) cs[r * (CH_MAXC + 3) + nc + 2] + (BND ? (uint32_t) csb[r * (CH_MAXC + 3) + nc + 2] : 0u)) * VB;
				if (total) {
					fence_proxy_async();      // the previous chunk's generic-proxy reads of the windows precede these async-proxy writes
					mbar_expect_tx(bar, total);
					for (int r = 0; r < 9; r++) {
						const uint32_t b = s_rowbase[r], len = cs[r * (CH_MAXC + 3) + nc + 2];
						if (len) tma_bulk_g2s(win_s + (uint32_t) (r * W) * VB, a.mpos + b, len * VB, bar);
						if (BND) {      // the imports of the row follow its sorted records in the window
							const uint32_t bb = s_rowbaseB[r], lenb = csb[r * (CH_MAXC + 3) + nc + 2];
							if (lenb) tma_bulk_g2s(win_s + (uint32_t) (r * W + (int) len) * VB, a.mpos + bb, lenb * VB, bar);
						}
					}
				}
			}
			s_next = atomicAdd(a.work_counter, 1u);      // the NEXT chunk of this CTA
		}
		__syncthreads();
		const uint32_t nch = s_next;
		Chunk ncd; ncd.cell0 = 0; ncd.ncells = 0; ncd.flags = 0;
		if (nch < n_chunks) { ncd = a.chunks[nch]; if (!(ncd.flags & CHUNK_BIG)) fill_cs(ncd, cs_all + (cur ^ 1) * CSB, s_rowbase_all[cur ^ 1], csb_all + (cur ^ 1) * CSB, rowbaseB_of(cur ^ 1)); }
		double acc[3] = {0., 0., 0.};
		if (big) {
			// ---- slow generic path: a cell denser than the tile. Global loads, pair evaluated as soon as it is found. ----
			auto big_i = [&](uint32_t t) {
				const V4 mi = a.mpos[t];
				const uint32_t cid_i = wbits<C>(mi.w) & 0xFFFFu;
				const int cx = (int) (a.cell_of[t] % (uint32_t) a.nx);
				C eps24 = a.eps24_0, sig2 = a.sig2_0, shift6 = a.shift6_0;
				double Fd[3] = {0., 0., 0.}, Vd[3] = {0., 0., 0.}, Ud = 0.;
				for (int rr = 0; rr < NR; rr++) {
					const int r = rr >= 9 ? rr - 9 : rr;
					const uint32_t* cst = rr >= 9 ? a.cell_start_b : a.cell_start;
					const uint32_t row = (uint32_t) ((int) R + (r / 3 - 1) * a.ny + (r % 3 - 1)) * a.nx + cx - 1;
					const uint32_t jb = cst[row], je = cst[row + 3];
					for (uint32_t j0 = jb; j0 < je; j0 += 32) {
						C fx[1] = {C(0)}, fy[1] = {C(0)}, fz[1] = {C(0)}, vx[1] = {C(0)}, vy[1] = {C(0)}, vz[1] = {C(0)}, up[1] = {C(0)};
						int nh = 0;
						for (uint32_t j = j0; j < min(je, j0 + 32); j++) {
							const V4 mj = a.mpos[j];
							const C r2t = molpair_r2<C>(mi.x, mi.y, mi.z, mj.x, mj.y, mj.z);
							if ((r2t < a.ljrc2) && (r2t != C(0))) { LJ_PAIR_ACCUM(0, mj); nh++; }
						}
						if (SINGLE) { fx[0] *= eps24; fy[0] *= eps24; fz[0] *= eps24; vx[0] *= eps24; vy[0] *= eps24; vz[0] *= eps24; }
						Fd[0] += (double) fx[0]; Fd[1] += (double) fy[0]; Fd[2] += (double) fz[0];
						Vd[0] += (double) vx[0]; Vd[1] += (double) vy[0]; Vd[2] += (double) vz[0];
						Ud += SINGLE ? (double) (eps24 * up[0] + (C) nh * shift6) : (double) up[0];
					}
				}
				acc[0] += 0.5 * Ud; acc[1] += 0.5 * (Vd[0] + Vd[1] + Vd[2]);
				const size_t cp = a.cap;
				if (!RMM && a.kin_v) {
					const double mass = a.comps[cid_i].mass, am = a.kin_dth / mass;
					const size_t tv = a.kin_perm ? a.kin_perm[t] : t;
					const double v0 = a.kin_v[tv] + am * Fd[0], v1 = a.kin_v[cp + tv] + am * Fd[1], v2 = a.kin_v[2 * cp + tv] + am * Fd[2];
					acc[2] += mass * (v0 * v0 + v1 * v1 + v2 * v2);
				}
				if (RMM) {
					a.rmm_v[t] += (float) Fd[0] * a.rmm_dt_m; a.rmm_v[cp + t] += (float) Fd[1] * a.rmm_dt_m; a.rmm_v[2 * cp + t] += (float) Fd[2] * a.rmm_dt_m;
				} else {
					a.F[t] = Fd[0]; a.F[cp + t] = Fd[1]; a.F[2 * cp + t] = Fd[2];
					a.Vi[t] = 0.5 * Vd[0]; a.Vi[cp + t] = 0.5 * Vd[1]; a.Vi[2 * cp + t] = 0.5 * Vd[2];
					if (a.M) { a.M[t] = 0.; a.M[cp + t] = 0.; a.M[2 * cp + t] = 0.; }
				}
			};
			for (uint32_t t = a.cell_start[cd.cell0] + threadIdx.x; t < a.cell_start[cd.cell0 + nc]; t += TPB) big_i(t);
			if (BND) for (uint32_t t = a.cell_start_b[cd.cell0] + threadIdx.x; t < a.cell_start_b[cd.cell0 + nc]; t += TPB) big_i(t);
		} else {
		const cs_t* csc = cs + 4 * (CH_MAXC + 3);          // centre row (relative record indices)
		// thread -> RI consecutive molecules of the chunk (slot order); they may lie in two neighbouring cells
		const cs_t* csbc = csb + 4 * (CH_MAXC + 3);
		const uint32_t n_molA = (uint32_t) csc[nc + 1] - (uint32_t) csc[1];
		const uint32_t n_mol = n_molA + (BND ? (uint32_t) csbc[nc + 1] - (uint32_t) csbc[1] : 0u);      // BND: the imports of the chunk's cells are molecules of the chunk too (arrivals)
		const bool isB = BND && threadIdx.x >= n_molA;                             // my molecule is one of them
		uint32_t trel[RI]; bool have[RI];
		#pragma unroll
		for (int l = 0; l < RI; l++) { have[l] = (uint32_t) RI * threadIdx.x + l < n_mol; trel[l] = csc[1] + (uint32_t) RI * threadIdx.x + l; }
		const cs_t* cst = isB ? csbc : csc;                                    // the table my molecule is listed in, and its index there
		const uint32_t tkey = isB ? (uint32_t) csbc[1] + (threadIdx.x - n_molA) : trel[0];
		if (isB) trel[0] = (uint32_t) csc[nc + 2] + tkey;                                     // its record in the centre row of the window (imports follow the sorted records)
		int kxa = 0;                                            // cell (inside the chunk) of my first molecule ...
		while (kxa + 1 < nc && cst[kxa + 2] <= tkey) kxa++;
		int kxb = kxa;                                          // ... and of my last one
		if (RI > 1 && have[RI - 1]) { while (kxb + 1 < nc && csc[kxb + 2] <= trel[RI - 1]) kxb++; }
		mbar_wait(bar, parity); parity ^= 1u;
		// ---- 2. phase 1 + 3. phase 2 -----------------------------------------------------------------------------------------
		// Phase 1 walks my 9 candidate rows (cells cx-1 .. cx+1 of each -- one more cell if my molecules straddle two cells --
		// one contiguous range of the row's window); every candidate is ONE LDS.128 that serves the reference's molecule-distance
		// test (strict <, bit-identical to calcDistLookup; in my own row SingleCellPolicy_ additionally masks pairs at distance 0,
		// i.e. the molecule itself) of all RI molecules of the thread (register blocking: shared-memory wavefronts per test
		// halve, the kernel is shared-memory bound -- profiles/r01r_force_full.md). A hit appends the candidate's shared-memory
		// ADDRESS (it fits 16 bits) to the molecule's list. When a list is nearly full, or after the last row, phase 2 evaluates
		// the lists densely (LJ 12-6; FP32 partial sums folded into FP64 like RealAccumVec) and empties them.
		V4 mi[RI]; uint32_t cid_i[RI];
		#pragma unroll
		for (int l = 0; l < RI; l++) {
			mi[l] = make_c4<C>((C) 1e18, (C) 1e18, (C) 1e18, (C) 0);      // an absent molecule: no candidate is ever in range
			if (have[l]) mi[l] = lds_v4<C>(win_s + (uint32_t) (4 * W + (int) trel[l]) * VB);
			cid_i[l] = wbits<C>(mi[l].w) & 0xFFFFu;
		}
		double Fd[RI][3], Vd[RI][3], Ud[RI];
		#pragma unroll
		for (int l = 0; l < RI; l++) { Fd[l][0] = Fd[l][1] = Fd[l][2] = Vd[l][0] = Vd[l][1] = Vd[l][2] = Ud[l] = 0.; }
		const uint32_t sentinel = win_s + (uint32_t) (9 * W) * VB;
		uint32_t w[RI], wlimit[RI], lbase[RI];
		#pragma unroll
		for (int l = 0; l < RI; l++) { lbase[l] = list_s + (uint32_t) l * (TPB * 2); w[l] = lbase[l]; wlimit[l] = lbase[l] + (uint32_t) (LCAP - 4) * LSTRIDE; }
		uint32_t src = 0, end = 0;              // shared-memory address of the next candidate / end of the row range
		int r = -1;
		auto room = [&]() -> bool { bool ok = true; _Pragma("unroll") for (int l = 0; l < RI; l++) ok = ok && (w[l] <= wlimit[l]); return ok; };
		for (;;) {
			bool flush = true;
			if (r < NR) {
				if (r == 4 || (BND && r == 13)) {
					while (src + 4 * VB <= end && room()) {
						const V4 m0 = lds_v4<C>(src), m1 = lds_v4<C>(src + VB), m2 = lds_v4<C>(src + 2 * VB), m3 = lds_v4<C>(src + 3 * VB);
						#pragma unroll
						for (int l = 0; l < RI; l++) {
							const C r0 = molpair_r2<C>(mi[l].x, mi[l].y, mi[l].z, m0.x, m0.y, m0.z), r1 = molpair_r2<C>(mi[l].x, mi[l].y, mi[l].z, m1.x, m1.y, m1.z);
							const C r2 = molpair_r2<C>(mi[l].x, mi[l].y, mi[l].z, m2.x, m2.y, m2.z), r3 = molpair_r2<C>(mi[l].x, mi[l].y, mi[l].z, m3.x, m3.y, m3.z);
							if (r0 < a.ljrc2 && r0 != C(0)) { sts16(w[l], src); w[l] += LSTRIDE; }
							if (r1 < a.ljrc2 && r1 != C(0)) { sts16(w[l], src + VB); w[l] += LSTRIDE; }
							if (r2 < a.ljrc2 && r2 != C(0)) { sts16(w[l], src + 2 * VB); w[l] += LSTRIDE; }
							if (r3 < a.ljrc2 && r3 != C(0)) { sts16(w[l], src + 3 * VB); w[l] += LSTRIDE; }
						}
						src += 4 * VB;
					}
				} else {
					while (src + 4 * VB <= end && room()) {
						const V4 m0 = lds_v4<C>(src), m1 = lds_v4<C>(src + VB), m2 = lds_v4<C>(src + 2 * VB), m3 = lds_v4<C>(src + 3 * VB);
						#pragma unroll
						for (int l = 0; l < RI; l++) {
							const C r0 = molpair_r2<C>(mi[l].x, mi[l].y, mi[l].z, m0.x, m0.y, m0.z), r1 = molpair_r2<C>(mi[l].x, mi[l].y, mi[l].z, m1.x, m1.y, m1.z);
							const C r2 = molpair_r2<C>(mi[l].x, mi[l].y, mi[l].z, m2.x, m2.y, m2.z), r3 = molpair_r2<C>(mi[l].x, mi[l].y, mi[l].z, m3.x, m3.y, m3.z);
							if (r0 < a.ljrc2) { sts16(w[l], src); w[l] += LSTRIDE; }
							if (r1 < a.ljrc2) { sts16(w[l], src + VB); w[l] += LSTRIDE; }
							if (r2 < a.ljrc2) { sts16(w[l], src + 2 * VB); w[l] += LSTRIDE; }
							if (r3 < a.ljrc2) { sts16(w[l], src + 3 * VB); w[l] += LSTRIDE; }
						}
						src += 4 * VB;
					}
				}
				if (room()) {                        // the row is down to < 4 candidates (the lists are not the reason the loop stopped)
					for (; src < end; src += VB) {
						const V4 mj = lds_v4<C>(src);
						#pragma unroll
						for (int l = 0; l < RI; l++) {
							const C rr = molpair_r2<C>(mi[l].x, mi[l].y, mi[l].z, mj.x, mj.y, mj.z);
							if (rr < a.ljrc2 && ((r != 4 && !(BND && r == 13)) || rr != C(0))) { sts16(w[l], src); w[l] += LSTRIDE; }
						}
					}
					r++;
					if (r < NR) {
						uint32_t kb = 0, ke = 0;
						const int row = (BND && r >= 9) ? r - 9 : r;
						if (BND && r >= 9) {      // the imports of the row: behind its sorted records
							const uint32_t off = cs[row * (CH_MAXC + 3) + nc + 2];
							if (have[0]) { kb = off + csb[row * (CH_MAXC + 3) + kxa]; ke = off + csb[row * (CH_MAXC + 3) + kxb + 3]; }
						} else
						if (have[0]) { kb = cs[r * (CH_MAXC + 3) + kxa]; ke = cs[r * (CH_MAXC + 3) + kxb + 3]; }      // cells cxa-1 .. cxb+1 of this row
						src = win_s + (uint32_t) (row * W + (int) kb) * VB; end = win_s + (uint32_t) (row * W + (int) ke) * VB;
						flush = false;
					}
				}
			}
			if (flush) {
				// ---- phase 2: dense LJ over the lists; two interleaved FP32 accumulator sets per list ----
				#pragma unroll
				for (int l = 0; l < RI; l++) {
					const int cnt = (int) ((w[l] - lbase[l]) / LSTRIDE);
					C fx[2] = {C(0), C(0)}, fy[2] = {C(0), C(0)}, fz[2] = {C(0), C(0)}, vx[2] = {C(0), C(0)}, vy[2] = {C(0), C(0)}, vz[2] = {C(0), C(0)}, up[2] = {C(0), C(0)};
					C eps24 = a.eps24_0, sig2 = a.sig2_0, shift6 = a.shift6_0;
					if (cnt & 1) sts16(w[l], sentinel);      // pad odd lists with the far-away record (all LJ terms underflow to 0)
					const V4 mi_ = mi[l]; const uint32_t cid_i_ = cid_i[l];
					#pragma unroll 1
					for (uint32_t e = lbase[l]; e < w[l]; e += 2 * LSTRIDE) {
						const uint32_t i0 = lds16(e), i1 = lds16(e + LSTRIDE);
						const V4 m0 = lds_v4<C>(i0), m1 = lds_v4<C>(i1);
						#define mi mi_
						#define cid_i cid_i_
						LJ_PAIR_ACCUM(0, m0);
						if (SINGLE || i1 != sentinel) LJ_PAIR_ACCUM(1, m1);
						#undef mi
						#undef cid_i
					}
					if (SINGLE) {
						// eps24 and the shift are pair independent: applied once per list
						const C f24 = eps24;
						fx[0] = f24 * (fx[0] + fx[1]); fy[0] = f24 * (fy[0] + fy[1]); fz[0] = f24 * (fz[0] + fz[1]);
						vx[0] = f24 * (vx[0] + vx[1]); vy[0] = f24 * (vy[0] + vy[1]); vz[0] = f24 * (vz[0] + vz[1]);
						up[0] = f24 * (up[0] + up[1]) + (C) cnt * shift6;
						fx[1] = fy[1] = fz[1] = vx[1] = vy[1] = vz[1] = up[1] = C(0);
					}
					Fd[l][0] += (double) fx[0] + (double) fx[1]; Fd[l][1] += (double) fy[0] + (double) fy[1]; Fd[l][2] += (double) fz[0] + (double) fz[1];
					Vd[l][0] += (double) vx[0] + (double) vx[1]; Vd[l][1] += (double) vy[0] + (double) vy[1]; Vd[l][2] += (double) vz[0] + (double) vz[1];
					Ud[l] += (double) up[0] + (double) up[1];
					w[l] = lbase[l];
				}
				if (r >= NR) break;
			}
		}
		// ---- 4. results ----
		#pragma unroll
		for (int l = 0; l < RI; l++) {
			if (have[l]) {
				const size_t cp = a.cap;
				const uint32_t t = isB ? s_rowbaseB[4] + tkey : s_rowbase[4] + trel[l];
				acc[0] += 0.5 * Ud[l];                          // every pair is visited from both of its ends
				acc[1] += 0.5 * (Vd[l][0] + Vd[l][1] + Vd[l][2]);
				if (!RMM && a.kin_v) {
					const double mass = a.comps[cid_i[l]].mass, am = a.kin_dth / mass;
					const size_t tv = a.kin_perm ? a.kin_perm[t] : t;
					const double v0 = a.kin_v[tv] + am * Fd[l][0], v1 = a.kin_v[cp + tv] + am * Fd[l][1], v2 = a.kin_v[2 * cp + tv] + am * Fd[l][2];
					acc[2] += mass * (v0 * v0 + v1 * v1 + v2 * v2);
				}
				if (RMM) {      // the kick in the epilogue: 12 B read + 12 B written per molecule instead of 48 B of F and Vi
					a.rmm_v[t] += (float) Fd[l][0] * a.rmm_dt_m; a.rmm_v[cp + t] += (float) Fd[l][1] * a.rmm_dt_m; a.rmm_v[2 * cp + t] += (float) Fd[l][2] * a.rmm_dt_m;
				} else {
				a.F[t] = Fd[l][0]; a.F[cp + t] = Fd[l][1]; a.F[2 * cp + t] = Fd[l][2];
				a.Vi[t] = 0.5 * Vd[l][0]; a.Vi[cp + t] = 0.5 * Vd[l][1]; a.Vi[2 * cp + t] = 0.5 * Vd[l][2];
				if (a.M) { a.M[t] = 0.; a.M[cp + t] = 0.; a.M[2 * cp + t] = 0.; }
				}
			}
		}
		}
		// per-warp partial sums of the chunk (slot 2 * warp of its 8 doubles; unused slots are zeroed): no block-wide reduction
		#pragma unroll
		for (int k = 0; k < 2; k++) {
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
		}
		if (!RMM && a.kin_v) {
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) acc[2] += __shfl_xor_sync(0xffffffffu, acc[2], o);
			if ((threadIdx.x & 31) == 0) a.partials_kin[(size_t) ch * 4 + (threadIdx.x >> 5)] = acc[2];
			if (TPB < 128 && threadIdx.x >= TPB / 32 && threadIdx.x < 4) a.partials_kin[(size_t) ch * 4 + threadIdx.x] = 0.;
		}
		if ((threadIdx.x & 31) == 0) *reinterpret_cast<double2*>(a.partials + (size_t) ch * 8 + 2 * (threadIdx.x >> 5)) = make_double2(acc[0], acc[1]);
		if (TPB < 128 && threadIdx.x >= TPB / 32 && threadIdx.x < 4) *reinterpret_cast<double2*>(a.partials + (size_t) ch * 8 + 2 * threadIdx.x) = make_double2(0., 0.);
		__syncthreads();      // windows and lists are reused by the next chunk; its cell_start cache is complete
		ch = nch; cd = ncd; cur ^= 1;
	}
}

// deterministic two-level reduction of the per-chunk partial sums (fixed assignment chunk -> block -> thread, fixed order)
constexpr int RED_BLOCKS = 128;
__global__ void __launch_bounds__(256) k_reduce_force_chunks1(const double* __restrict__ partials, const uint32_t* __restrict__ n_chunks, double* __restrict__ partials2, int per_warp,
                                                               const double* __restrict__ partials_kin) {
	double v[3] = {0., 0., 0.};
	const uint32_t n = *n_chunks;
	for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x) {
		const double2* q = reinterpret_cast<const double2*>(partials + (size_t) b * 8);
		if (per_warp) { const double2 p0 = q[0], p1 = q[1], p2 = q[2], p3 = q[3]; v[0] += (p0.x + p1.x) + (p2.x + p3.x); v[1] += (p0.y + p1.y) + (p2.y + p3.y); }
		else { const double2 p = q[0]; v[0] += p.x; v[1] += p.y; }
		if (partials_kin) { const double2* k = reinterpret_cast<const double2*>(partials_kin + (size_t) b * 4); const double2 k0 = k[0], k1 = k[1]; v[2] += (k0.x + k0.y) + (k1.x + k1.y); }
	}
	block_reduce_store<3>(v, partials2, blockIdx.x);
}
__global__ void __launch_bounds__(RED_BLOCKS) k_reduce_force_chunks2(const double* __restrict__ partials2, DevScalars* scal, int kin, int accumulate) {
	double v[3] = {partials2[(size_t) threadIdx.x * 8], partials2[(size_t) threadIdx.x * 8 + 1], partials2[(size_t) threadIdx.x * 8 + 2]};
	__shared__ double out[8];
	block_reduce_store<3>(v, out, 0);
	__syncthreads();
	if (threadIdx.x == 0) {
		if (accumulate) {      // second pass of the overlapped multi-GPU step (boundary chunks): adds to the sums of the inner chunks
			scal->upot6lj += out[0]; scal->virial += out[1];
			if (kin) { scal->summv2 += out[2]; scal->fault = scal->overflow ? 1. : 0.; }
		} else {
		scal->upot6lj = out[0]; scal->upot_xpoles = 0.; scal->virial = out[1]; scal->my_rf = 0.;
		if (kin) { scal->summv2 = out[2]; scal->sum_iw2 = 0.; scal->fault = scal->overflow ? 1. : 0.; }      // what k_reduce_kinetic delivers after leapfrog_post
		}
	}
}

// sums the per-chunk partials of a chunked force kernel into DevScalars (also used by force_1clj_mma.cu)
int reduce_force_chunks(ls1b200_ctx* c, uint32_t nrows, int per_warp, bool kin, bool accumulate) {
	LS1_LAUNCH(c, PH_FORCE, k_reduce_force_chunks1, RED_BLOCKS, 256, 0, c->partials, (accumulate ? c->chunk_row_start_b : c->chunk_row_start) + nrows, c->partials2, per_warp, kin ? (const double*) c->partials_kin : (const double*) nullptr);
	LS1_LAUNCH(c, PH_FORCE, k_reduce_force_chunks2, 1, RED_BLOCKS, 0, c->partials2, c->d_scal, (int) kin, (int) accumulate);
	return 0;
}

template <class C, bool RMM = false>
static int launch_tiled(ls1b200_ctx* c, float* rmm_v = nullptr, float rmm_dt_m = 0.f) {
	typedef typename Vec4<C>::type V4;
	constexpr int TPB = TileCfg<C>::TPB, W = TileCfg<C>::W, LCAP = TileCfg<C>::LCAP, RI = TileCfg<C>::RI;
	const Geom& g = c->geom;
	uint32_t nrows = 0;
	// overlapped multi-GPU step (c->force_pass): 1 = the chunks no import can touch (while the exchange runs), 2 = the others, with the
	// imports as a second slot range per cell (BND kernel), added to the sums of pass 1; 0 = everything in one pass
	const int pass = RMM ? 0 : c->force_pass;
	if (c->chunks_ready && pass != 2) {      // built on the side stream while the sort permuted the state (update.cu)
		LS1_CUDA(cudaStreamWaitEvent(c->stream, c->ev_c1, 0));
		nrows = c->chunks_nrows; c->chunks_ready = false;
	} else if (pass == 2 && c->chunks_b_ready) {      // built on the exchange stream behind the imports' sort (update.cu); the caller has waited for that stream
		nrows = c->chunks_b_nrows; c->chunks_b_ready = false;
	} else
	LS1_TRY(build_chunks(c, TPB * RI, W, 1, &nrows, pass, pass == 2 ? (const uint32_t*) c->cell_start_b : (const uint32_t*) nullptr, pass == 2 ? 1 : 0));      // a chunk holds at most TPB * RI molecules
	Chunk* chunks = (Chunk*) (pass == 2 ? c->chunks_b : c->chunks);
	uint32_t* row_start = pass == 2 ? c->chunk_row_start_b : c->chunk_row_start;
	TArgs<C> a;
	a.mpos = (const V4*) c->mpos; a.cell_start = c->cell_start; a.cell_of = c->cell_of2;
	if (sizeof(C) == 4) { a.eps24 = (const C*) c->d_eps24f; a.sig2 = (const C*) c->d_sig2f; a.shift6 = (const C*) c->d_shift6f; }
	else { a.eps24 = (const C*) c->d_eps24; a.sig2 = (const C*) c->d_sig2; a.shift6 = (const C*) c->d_shift6; }
	a.F = c->F; a.M = c->rotates ? c->M : nullptr; a.Vi = c->Vi; a.partials = c->partials;
	a.chunks = chunks; a.n_chunks = row_start + nrows; a.work_counter = pass == 2 ? c->work_counter_b : c->work_counter;
	a.cap = (uint32_t) c->cap; a.n_lj = c->n_lj; a.nx = g.cells[0]; a.ny = g.cells[1]; a.nz = g.cells[2];
	const double rclj = c->cfg.cutoff_lj;
	a.ljrc2 = (C) (rclj * rclj);
	a.eps24_0 = (C) c->h_eps24[0]; a.sig2_0 = (C) c->h_sig2[0]; a.shift6_0 = (C) c->h_shift6[0];
	a.rmm_v = rmm_v; a.rmm_dt_m = rmm_dt_m;
	const bool kin = !RMM && c->kin_dth != 0.;
	a.kin_v = kin ? (c->state_lazy ? c->nxt.v : c->cur.v) : nullptr; a.kin_perm = c->state_lazy ? c->perm : nullptr; a.kin_dth = c->kin_dth; a.comps = c->d_comps; a.partials_kin = c->partials_kin;
	if (kin && !c->partials_kin) return fail("force_1clj_tiled: no buffer for the kinetic partial sums");
	// the cell_start cache is padded by three rows: the compiler hoists the (guarded) loads of the next row's entries
	a.cell_start_b = c->cell_start_b;
	const size_t smem = (size_t) (9 * W + 1) * sizeof(V4) + (size_t) (LCAP + 1) * RI * TPB * 2 + 2 * 12 * (CH_MAXC + 3) * 4;      // (boundary kernel: two 16-bit tables in the room of the one 32-bit table)
	if (pass == 2 && (!c->cell_start_b || RI != 1)) return fail("force_1clj_tiled: boundary pass without the imports' cell table");
	auto kern = pass == 2 ? (c->comps.size() == 1 ? k_force_1clj_tiled<C, true, false, true> : k_force_1clj_tiled<C, false, false, true>)
	                      : (c->comps.size() == 1 ? k_force_1clj_tiled<C, true, RMM> : k_force_1clj_tiled<C, false, RMM>);
	LS1_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	int per_sm = 0;
	LS1_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TPB, smem));
	if (per_sm < 1) return fail("force_1clj_tiled: kernel does not fit on an SM");
	// overlapped multi-GPU step, inner pass: the persistent CTAs of this kernel would hold every byte of shared memory until they finish, and
	// the kernels of the exchange stream (record staging, scans, chunk lists: all need some) could not start beside them -- one CTA less per SM
	// leaves them room (LS1B200_TILE_MAXCTAS overrides, for measurements)
	if (pass == 1 && c->tile_pass1_ctas > 0) per_sm = std::min(per_sm, c->tile_pass1_ctas);
	if (c->tile_max_ctas > 0) per_sm = std::min(per_sm, c->tile_max_ctas);
	const uint32_t grid = (uint32_t) (c->sm_count * per_sm);
	LS1_LAUNCH(c, PH_FORCE, kern, grid, TPB, smem, a);
	return reduce_force_chunks(c, nrows, 1, kin, pass == 2);
}

int build_chunks_tiled_1clj(ls1b200_ctx* c, uint32_t* nrows, int mode, int set) {
	const uint32_t* csb = mode == 2 ? (const uint32_t*) c->cell_start_b : (const uint32_t*) nullptr;
	return c->mixed ? build_chunks(c, TileCfg<float>::TPB * TileCfg<float>::RI, TileCfg<float>::W, 1, nrows, mode, csb, set)
	                : build_chunks(c, TileCfg<double>::TPB * TileCfg<double>::RI, TileCfg<double>::W, 1, nrows, mode, csb, set);
}

int force_1clj_tiled(ls1b200_ctx* c) { return c->mixed ? launch_tiled<float>(c) : launch_tiled<double>(c); }
// reduced-memory mode: FP32 only, the kick v += f dt / m in the epilogue (v: [3][cap] planes of the slots the kernel visits)
int force_1clj_tiled_rmm(ls1b200_ctx* c, float* v, float dt_over_m) { return launch_tiled<float, true>(c, v, dt_over_m); }

}  // namespace ls1b200
