// force_1clj_v3.cu -- round-2 EXPERIMENT (LS1B200_LJ_KERNEL=v3, tested, measured slower than force_1clj_tiled.cu: 5.2 - 6.4 ms vs 4.97 ms at
// 16 M molecules, profiles/r02_force_kernel_experiments.md). Same arithmetic and the same structure as force_1clj_tiled.cu
// (TMA-staged neighbour windows in shared memory, per-thread u16 neighbour lists, dense LJ phase; reference:
// VectorizedCellProcessor::_calculatePairs LJ block, adapter/VectorizedCellProcessor.cpp:1029-1178, _loopBodyLJ :171-224,
// calcDistLookup vectorization/SIMD_VectorizedCellProcessorHelpers.h:427-472, FullMolecule::calcFM FullMolecule.cpp:548-626,
// cell pairs of C08BasedTraversals.h:48-99 visited from both ends), with two changes measured in profiles/r02*:
//
//  1. MOLECULE-granular chunks. A chunk is a run of up to TPB consecutive slots of one x-row (it may start and end in the
//     middle of a cell); the round-1 chunks were whole cells, which left 20 % of the lanes without a molecule.
//  2. x-TRIMMED candidate ranges. The counting sort orders the molecules of a reference cell by S fine x-bins (update.cu,
//     geom.cuh: fine_sub), so each of the 9 window rows is coarsely sorted by x and a thread scans only the fine bins that
//     intersect [x - rc, x + rc] instead of three whole cells: 25 % fewer distance tests at S = 4. The cutoff decision itself
//     is still the reference's bit-identical FP32 / FP64 sequence on every candidate that is scanned, and every candidate
//     the reference would accept is scanned (the trim bound is evaluated with the sort's own FP64 expression: monotone).
//     The lanes of a warp that share a fine bin share their ranges (shared-memory broadcast); ALIGN pads every bin's lane
//     group to a multiple of ALIGN lanes.
#include "tile.cuh"
#include "geom.cuh"

namespace ls1b200 {

struct Chunk2 { uint32_t m0; uint32_t cell0; uint16_t n; uint8_t ncells; uint8_t flags; };
static_assert(sizeof(Chunk2) == 12, "Chunk2 layout");

template <class C> struct V3Cfg;
template <> struct V3Cfg<float> { static constexpr int TPB = 128, W = 176, LCAP = 64; };
template <> struct V3Cfg<double> { static constexpr int TPB = 64, W = 112, LCAP = 64; };
constexpr int V3_MAXC = 24;        // cells per chunk

// ---------------------------------------------------------------------------------------------------------------------
// chunk list: one warp per x-row. Lanes compute, for every possible first cell, how far the 9 windows fit (monotone); lane 0
// then walks the row: a chunk takes molecules fine bin by fine bin until its TPB thread slots are used (a bin's lanes are
// padded to a multiple of ALIGN), its windows would overflow, or V3_MAXC cells are reached.
// ---------------------------------------------------------------------------------------------------------------------
struct Chunk2Geo { int nx, ny, nz; int tpb, w, align; uint32_t S; };
constexpr int CHUNK2_ROWS_PER_BLOCK = 4;
template <bool WRITE>
__global__ void __launch_bounds__(32 * CHUNK2_ROWS_PER_BLOCK) k_chunks2(Chunk2Geo g, const uint32_t* __restrict__ cs, const uint32_t* __restrict__ csf, uint32_t nrows,
                                                                        uint32_t* __restrict__ cnt, const uint32_t* __restrict__ start, Chunk2* __restrict__ chunks) {
	extern __shared__ uint32_t sm_cs[];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const uint32_t r = blockIdx.x * CHUNK2_ROWS_PER_BLOCK + w;
	if (r >= nrows) return;
	const int nx = g.nx, P = nx + 1, S = (int) g.S, PF = nx * S + 1;
	const int cy = 1 + (int) (r % (uint32_t) (g.ny - 2)), cz = 1 + (int) (r / (uint32_t) (g.ny - 2));
	uint32_t* rowcs = sm_cs + (size_t) w * (10 * P + PF);      // [9][P] cell_start of the neighbour rows
	uint32_t* reach = rowcs + 9 * P;                           // [P] last cell a chunk starting in cell xa may reach (windows), xa - 1 if none
	uint32_t* ff = reach + P;                                  // [PF] fine starts of the centre row
	for (int k = 0; k < 9; k++) {
		const uint32_t row = (uint32_t) ((cz + k / 3 - 1) * g.ny + (cy + k % 3 - 1)) * nx;
		for (int x = lane; x <= nx; x += 32) rowcs[k * P + x] = cs[row + x];
	}
	const uint32_t R = (uint32_t) (cz * g.ny + cy) * nx;
	for (int x = lane; x < PF; x += 32) ff[x] = csf[(size_t) R * S + x];
	__syncwarp();
	const uint32_t* c4 = rowcs + 4 * P;
	const int last_x = nx - 2;
	for (int xa = 1 + lane; xa <= last_x; xa += 32) {
		uint32_t lo[9];
		#pragma unroll
		for (int k = 0; k < 9; k++) lo[k] = rowcs[k * P + xa - 1];
		int xb = xa - 1;
		for (int x = xa; x <= last_x && x - xa + 1 <= V3_MAXC; x++) {
			bool f = true;
			#pragma unroll
			for (int k = 0; k < 9; k++) f = f && (rowcs[k * P + x + 2] - lo[k] <= (uint32_t) g.w);
			if (!f) break;
			xb = x;
		}
		reach[xa] = (uint32_t) xb;
	}
	__syncwarp();
	if (lane == 0) {
		Chunk2* out = WRITE ? chunks + start[r] : nullptr;
		uint32_t n_out = 0;
		uint32_t m = c4[1];
		const uint32_t mend = c4[nx - 1];
		int x = 1;
		const uint32_t tpb = (uint32_t) g.tpb, A = (uint32_t) g.align;
		while (m < mend) {
			while (c4[x + 1] <= m) x++;
			const int xa = x, xl = (int) reach[xa];
			uint32_t mm = m; int xb = xa; uint8_t flags = 0;
			if (xl < xa) {      // the windows of this single cell do not fit: slow path of the kernel, TPB molecules at a time
				mm = min(c4[xa + 1], m + tpb); flags = (uint8_t) CHUNK_BIG;
			} else {
				uint32_t slots = 0; bool stop = false;
				for (int xx = xa; xx <= xl && !stop; xx++) {
					for (int sb = 0; sb < S; sb++) {
						const uint32_t lo = max(ff[xx * S + sb], mm), hi = ff[xx * S + sb + 1];
						if (hi <= lo) continue;
						const uint32_t room = tpb - slots;
						if (room == 0) { stop = true; break; }
						const uint32_t take = min(hi - lo, room);
						slots += (take + A - 1) / A * A;
						mm = lo + take; xb = xx;
						if (take < hi - lo) { stop = true; break; }
					}
					if (slots >= tpb) stop = true;
				}
			}
			if (WRITE) { Chunk2 c; c.m0 = m; c.cell0 = R + xa; c.n = (uint16_t) (mm - m); c.ncells = (uint8_t) (xb - xa + 1); c.flags = flags; out[n_out] = c; }
			n_out++;
			m = mm;
		}
		if (!WRITE) cnt[r] = n_out;
	}
}

static inline int build_chunks2(ls1b200_ctx* c, int tpb, int w, int align, bool fine, uint32_t* nrows_out) {
	const Geom& g = c->geom;
	const uint32_t nrows = (uint32_t) (g.cells[1] - 2) * (uint32_t) (g.cells[2] - 2);
	const uint32_t S = fine ? c->fine_s : 1u;
	Chunk2Geo cg; cg.nx = g.cells[0]; cg.ny = g.cells[1]; cg.nz = g.cells[2]; cg.tpb = tpb; cg.w = w; cg.align = align; cg.S = S;
	const size_t need = ((size_t) g.ncells + c->cap / 32 + nrows + 64) * sizeof(Chunk2);
	if (c->chunks2_bytes < need) {
		if (c->chunks2) cudaFree(c->chunks2);
		c->chunks2 = nullptr; c->chunks2_bytes = 0;
		LS1_CUDA(cudaMalloc(&c->chunks2, need));
		c->chunks2_bytes = need;
	}
	if ((uint64_t) g.ncells + c->cap / 32 + nrows + 64 > c->partial_blocks) return fail("force_1clj_v3: partials buffer too small");
	const uint32_t* csf = S > 1 ? c->cell_start_fine : c->cell_start;
	const uint32_t gr = (nrows + CHUNK2_ROWS_PER_BLOCK - 1) / CHUNK2_ROWS_PER_BLOCK;
	const size_t csm = (size_t) CHUNK2_ROWS_PER_BLOCK * (10 * (cg.nx + 1) + (size_t) cg.nx * S + 1) * sizeof(uint32_t);
	if (csm > 200 * 1024) return fail("force_1clj_v3: rows of more than ~850 cells are not supported by the chunk builder");
	if (csm > 48 * 1024) {
		LS1_CUDA(cudaFuncSetAttribute(k_chunks2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) csm));
		LS1_CUDA(cudaFuncSetAttribute(k_chunks2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) csm));
	}
	LS1_LAUNCH(c, PH_FORCE, k_chunks2<false>, gr, 32 * CHUNK2_ROWS_PER_BLOCK, csm, cg, c->cell_start, csf, nrows, c->chunk_row_count, (const uint32_t*) nullptr, (Chunk2*) nullptr);
	LS1_LAUNCH(c, PH_FORCE, k_chunks_scan, 1, 1024, 0, c->chunk_row_count, c->chunk_row_start, nrows, c->work_counter);
	LS1_LAUNCH(c, PH_FORCE, k_chunks2<true>, gr, 32 * CHUNK2_ROWS_PER_BLOCK, csm, cg, c->cell_start, csf, nrows, (uint32_t*) nullptr, c->chunk_row_start, (Chunk2*) c->chunks2);
	*nrows_out = nrows;
	return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
template <class C>
struct V3Args {
	const typename Vec4<C>::type* mpos;
	const uint32_t* cell_start;   // reference cells
	const uint32_t* csf;          // fine bins (== cell_start if S == 1)
	const uint32_t* cell_of;
	const double* bx;             // x borders of the cells (the table of the sort)
	const C* eps24; const C* sig2; const C* shift6;
	double* F; double* M; double* Vi;
	double* partials;
	const Chunk2* chunks;
	const uint32_t* n_chunks;
	uint32_t* work_counter;
	uint32_t cap, n_lj;
	int nx, ny, nz;
	C ljrc2;
	C eps24_0, sig2_0, shift6_0;
	double finv, trim_rc;
};

#define V3_LJ_PAIR_ACCUM(u, mj)                                                                \
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

// S: fine bins per cell the kernel works with (1: whole cells, no trim); ALIGN: lanes per bin padded to a multiple of it
template <class C, bool SINGLE, int S, int ALIGN>
__global__ void __launch_bounds__(V3Cfg<C>::TPB) k_force_1clj_v3(V3Args<C> a) {
	typedef typename Vec4<C>::type V4;
	constexpr int TPB = V3Cfg<C>::TPB, W = V3Cfg<C>::W, LCAP = V3Cfg<C>::LCAP;
	constexpr uint32_t VB = sizeof(V4), LSTRIDE = (uint32_t) TPB * 2u;
	constexpr int CSF = (V3_MAXC + 2) * S + 4;          // fine starts of one window row (cells xa-1 .. xb+1) + padding
	constexpr int NBIN = V3_MAXC * S;                   // fine bins that can hold molecules of a chunk
	extern __shared__ __align__(128) unsigned char smem_raw[];
	V4* win = reinterpret_cast<V4*>(smem_raw);                                         // [9][W] (+1 sentinel record at [9*W])
	uint16_t* list = reinterpret_cast<uint16_t*>(smem_raw + (size_t) (9 * W + 1) * VB);   // [LCAP+1][TPB]
	uint32_t* cs = reinterpret_cast<uint32_t*>(list + (size_t) (LCAP + 1) * TPB);      // [9][CSF] fine starts relative to the window start
	__shared__ uint32_t s_rowbase[9];
	__shared__ uint16_t s_tb[NBIN + 1];      // first thread slot of every fine bin of the chunk's cells
	__shared__ uint16_t s_first[NBIN];       // first molecule (relative to the centre window) of the bin that belongs to this chunk
	__shared__ uint8_t s_cnt[NBIN];          // molecules of the bin in this chunk
	__shared__ uint8_t s_owner[TPB];         // thread slot -> fine bin
	__shared__ __align__(8) unsigned long long s_bar;
	__shared__ uint32_t s_next;
	const uint32_t win_s = smem_u32(win), list_s = smem_u32(list) + threadIdx.x * 2, bar = smem_u32(&s_bar);
	const uint32_t n_chunks = *a.n_chunks;
	if (threadIdx.x == 0) {
		mbar_init(bar, 1);
		win[9 * W] = make_c4<C>((C) 1e18, (C) 1e18, (C) 1e18, (C) 0);      // sentinel record: pads odd lists
	}
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	fence_proxy_async();
	uint32_t parity = 0;
	for (;;) {
		__syncthreads();      // windows, lists, tables of the previous chunk are free
		if (threadIdx.x == 0) s_next = atomicAdd(a.work_counter, 1u);
		__syncthreads();
		const uint32_t ch = s_next;
		if (ch >= n_chunks) break;
		const Chunk2 cd = a.chunks[ch];
		const bool big = (cd.flags & CHUNK_BIG) != 0;
		const int nc = cd.ncells;
		const uint32_t R = cd.cell0 / (uint32_t) a.nx;                 // row index cz*ny + cy
		const int xa = (int) (cd.cell0 % (uint32_t) a.nx);
		double acc[2] = {0., 0.};
		if (big) {
			// ---- slow generic path: a cell denser than the tile. Global loads, pair evaluated as soon as it is found. ----
			const uint32_t ib = cd.m0, ie = cd.m0 + cd.n;
			for (uint32_t t = ib + threadIdx.x; t < ie; t += TPB) {
				const V4 mi = a.mpos[t];
				const uint32_t cid_i = wbits<C>(mi.w) & 0xFFFFu;
				const int cx = (int) (a.cell_of[t] % (uint32_t) a.nx);
				C eps24 = a.eps24_0, sig2 = a.sig2_0, shift6 = a.shift6_0;
				double Fd[3] = {0., 0., 0.}, Vd[3] = {0., 0., 0.}, Ud = 0.;
				for (int r = 0; r < 9; r++) {
					const uint32_t row = (uint32_t) ((int) R + (r / 3 - 1) * a.ny + (r % 3 - 1)) * a.nx + cx - 1;
					const uint32_t jb = a.cell_start[row], je = a.cell_start[row + 3];
					for (uint32_t j0 = jb; j0 < je; j0 += 32) {
						C fx[1] = {C(0)}, fy[1] = {C(0)}, fz[1] = {C(0)}, vx[1] = {C(0)}, vy[1] = {C(0)}, vz[1] = {C(0)}, up[1] = {C(0)};
						int nh = 0;
						for (uint32_t j = j0; j < min(je, j0 + 32); j++) {
							const V4 mj = a.mpos[j];
							const C r2t = molpair_r2<C>(mi.x, mi.y, mi.z, mj.x, mj.y, mj.z);
							if ((r2t < a.ljrc2) && (r2t != C(0))) { V3_LJ_PAIR_ACCUM(0, mj); nh++; }
						}
						if (SINGLE) { fx[0] *= eps24; fy[0] *= eps24; fz[0] *= eps24; vx[0] *= eps24; vy[0] *= eps24; vz[0] *= eps24; }
						Fd[0] += (double) fx[0]; Fd[1] += (double) fy[0]; Fd[2] += (double) fz[0];
						Vd[0] += (double) vx[0]; Vd[1] += (double) vy[0]; Vd[2] += (double) vz[0];
						Ud += SINGLE ? (double) (eps24 * up[0] + (C) nh * shift6) : (double) up[0];
					}
				}
				acc[0] += 0.5 * Ud; acc[1] += 0.5 * (Vd[0] + Vd[1] + Vd[2]);
				const size_t cp = a.cap;
				a.F[t] = Fd[0]; a.F[cp + t] = Fd[1]; a.F[2 * cp + t] = Fd[2];
				a.Vi[t] = 0.5 * Vd[0]; a.Vi[cp + t] = 0.5 * Vd[1]; a.Vi[2 * cp + t] = 0.5 * Vd[2];
				if (a.M) { a.M[t] = 0.; a.M[cp + t] = 0.; a.M[2 * cp + t] = 0.; }
			}
		} else {
		// ---- 1. fine starts of the 9 window rows (cells xa-1 .. xb+1), relative to the first record of the window ----
		const int nf = (nc + 2) * S + 1;
		for (int k = threadIdx.x; k < 9 * nf; k += TPB) {
			const int r = k / nf, f = k - r * nf;
			const size_t row = (size_t) ((int) R + (r / 3 - 1) * a.ny + (r % 3 - 1)) * a.nx + xa - 1;
			const uint32_t b0 = a.csf[row * S];
			cs[r * CSF + f] = a.csf[row * S + f] - b0;
			if (f == 0) s_rowbase[r] = b0;
		}
		__syncthreads();
		// ---- 2. TMA loads of the windows; thread slots of the chunk's fine bins ----
		if (threadIdx.x == 0) {
			uint32_t total = 0;
			for (int r = 0; r < 9; r++) total += cs[r * CSF + nf - 1] * VB;
			if (total) {
				fence_proxy_async();      // the previous chunk's generic-proxy reads of the windows precede these async-proxy writes
				mbar_expect_tx(bar, total);
				for (int r = 0; r < 9; r++) {
					const uint32_t len = cs[r * CSF + nf - 1];
					if (len) tma_bulk_g2s(win_s + (uint32_t) (r * W) * VB, a.mpos + s_rowbase[r], len * VB, bar);
				}
			}
		}
		if (threadIdx.x < 32) {
			// one warp: molecules of every fine bin of the chunk's cells that lie in [m0, m0 + n), padded prefix -> thread slots
			const uint32_t m0r = cd.m0 - s_rowbase[4], m1r = m0r + cd.n;
			const uint32_t* csc = cs + 4 * CSF + S;                  // centre row, first bin of cell xa
			const int nb = nc * S;
			uint32_t carry = 0;
			for (int f0 = 0; f0 < nb; f0 += 32) {
				const int f = f0 + (int) threadIdx.x;
				uint32_t cnt = 0, first = 0;
				if (f < nb) { const uint32_t lo = max(csc[f], m0r), hi = min(csc[f + 1], m1r); first = lo; cnt = hi > lo ? hi - lo : 0u; }
				const uint32_t padded = (cnt + ALIGN - 1) / ALIGN * ALIGN;
				uint32_t inc = padded;
				#pragma unroll
				for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if ((int) threadIdx.x >= o) inc += t; }
				if (f < nb) { s_tb[f] = (uint16_t) (carry + inc - padded); s_first[f] = (uint16_t) first; s_cnt[f] = (uint8_t) cnt; }
				carry += __shfl_sync(0xffffffffu, inc, 31);
			}
			if (threadIdx.x == 0) s_tb[nb] = (uint16_t) carry;
		}
		__syncthreads();
		for (int f = threadIdx.x; f < nc * S; f += TPB) {
			const uint32_t b = s_tb[f], e = s_tb[f + 1];
			for (uint32_t t = b; t < e && t < (uint32_t) TPB; t++) s_owner[t] = (uint8_t) f;
		}
		__syncthreads();
		// thread slot -> molecule
		const bool in_slots = threadIdx.x < s_tb[nc * S];
		const int fb = in_slots ? (int) s_owner[threadIdx.x] : 0;
		const uint32_t jj = threadIdx.x - s_tb[fb];
		const bool have = in_slots && jj < s_cnt[fb];
		const uint32_t trel = s_first[fb] + jj;                      // relative to the centre window
		const int kx = fb / S;                                       // my cell inside the chunk (window cell kx + 1)
		mbar_wait(bar, parity); parity ^= 1u;
		V4 mi = make_c4<C>((C) 1e18, (C) 1e18, (C) 1e18, (C) 0);      // an absent molecule: no candidate is ever in range
		if (have) mi = lds_v4<C>(win_s + (uint32_t) (4 * W + (int) trel) * VB);
		const uint32_t cid_i = wbits<C>(mi.w) & 0xFFFFu;
		// candidate fine bins [fl, fh] of every row (window bin index; cell xa-1 holds bins 0 .. S-1)
		int fl = kx * S, fh = (kx + 3) * S - 1;
		if (S > 1 && have) {
			const double xi = (double) mi.x;
			fl = kx * S + fine_sub(xi - a.trim_rc, __ldg(a.bx + xa + kx - 1), a.finv, S);
			fh = (kx + 2) * S + fine_sub(xi + a.trim_rc, __ldg(a.bx + xa + kx + 1), a.finv, S);
		}
		double Fd[3] = {0., 0., 0.}, Vd[3] = {0., 0., 0.}, Ud = 0.;
		const uint32_t sentinel = win_s + (uint32_t) (9 * W) * VB;
		const uint32_t wlimit = list_s + (uint32_t) (LCAP - 4) * LSTRIDE;
		uint32_t w = list_s;
		uint32_t src = 0, end = 0;              // shared-memory address of the next candidate / end of the row range
		int r = -1;
		for (;;) {
			bool flush = true;
			if (r < 9) {
				if (r == 4) {
					while (src + 4 * VB <= end && w <= wlimit) {
						const V4 m0 = lds_v4<C>(src), m1 = lds_v4<C>(src + VB), m2 = lds_v4<C>(src + 2 * VB), m3 = lds_v4<C>(src + 3 * VB);
						const C r0 = molpair_r2<C>(mi.x, mi.y, mi.z, m0.x, m0.y, m0.z), r1 = molpair_r2<C>(mi.x, mi.y, mi.z, m1.x, m1.y, m1.z);
						const C r2 = molpair_r2<C>(mi.x, mi.y, mi.z, m2.x, m2.y, m2.z), r3 = molpair_r2<C>(mi.x, mi.y, mi.z, m3.x, m3.y, m3.z);
						if (r0 < a.ljrc2 && r0 != C(0)) { sts16(w, src); w += LSTRIDE; }
						if (r1 < a.ljrc2 && r1 != C(0)) { sts16(w, src + VB); w += LSTRIDE; }
						if (r2 < a.ljrc2 && r2 != C(0)) { sts16(w, src + 2 * VB); w += LSTRIDE; }
						if (r3 < a.ljrc2 && r3 != C(0)) { sts16(w, src + 3 * VB); w += LSTRIDE; }
						src += 4 * VB;
					}
				} else {
					while (src + 4 * VB <= end && w <= wlimit) {
						const V4 m0 = lds_v4<C>(src), m1 = lds_v4<C>(src + VB), m2 = lds_v4<C>(src + 2 * VB), m3 = lds_v4<C>(src + 3 * VB);
						const C r0 = molpair_r2<C>(mi.x, mi.y, mi.z, m0.x, m0.y, m0.z), r1 = molpair_r2<C>(mi.x, mi.y, mi.z, m1.x, m1.y, m1.z);
						const C r2 = molpair_r2<C>(mi.x, mi.y, mi.z, m2.x, m2.y, m2.z), r3 = molpair_r2<C>(mi.x, mi.y, mi.z, m3.x, m3.y, m3.z);
						if (r0 < a.ljrc2) { sts16(w, src); w += LSTRIDE; }
						if (r1 < a.ljrc2) { sts16(w, src + VB); w += LSTRIDE; }
						if (r2 < a.ljrc2) { sts16(w, src + 2 * VB); w += LSTRIDE; }
						if (r3 < a.ljrc2) { sts16(w, src + 3 * VB); w += LSTRIDE; }
						src += 4 * VB;
					}
				}
				if (w <= wlimit) {                   // the row is down to < 4 candidates (the list is not the reason the loop stopped)
					for (; src < end; src += VB) {
						const V4 mj = lds_v4<C>(src);
						const C rr = molpair_r2<C>(mi.x, mi.y, mi.z, mj.x, mj.y, mj.z);
						if (rr < a.ljrc2 && (r != 4 || rr != C(0))) { sts16(w, src); w += LSTRIDE; }
					}
					r++;
					if (r < 9) {
						uint32_t kb = 0, ke = 0;
						if (have) { kb = cs[r * CSF + fl]; ke = cs[r * CSF + fh + 1]; }
						src = win_s + (uint32_t) (r * W + (int) kb) * VB; end = win_s + (uint32_t) (r * W + (int) ke) * VB;
						flush = false;
					}
				}
			}
			if (flush) {
				// ---- phase 2: dense LJ over the list; two interleaved FP32 accumulator sets ----
				const int cnt = (int) ((w - list_s) / LSTRIDE);
				C fx[2] = {C(0), C(0)}, fy[2] = {C(0), C(0)}, fz[2] = {C(0), C(0)}, vx[2] = {C(0), C(0)}, vy[2] = {C(0), C(0)}, vz[2] = {C(0), C(0)}, up[2] = {C(0), C(0)};
				C eps24 = a.eps24_0, sig2 = a.sig2_0, shift6 = a.shift6_0;
				if (cnt & 1) sts16(w, sentinel);      // pad odd lists with the far-away record (all LJ terms underflow to 0)
				#pragma unroll 1
				for (uint32_t e = list_s; e < w; e += 2 * LSTRIDE) {
					const uint32_t i0 = lds16(e), i1 = lds16(e + LSTRIDE);
					const V4 m0 = lds_v4<C>(i0), m1 = lds_v4<C>(i1);
					V3_LJ_PAIR_ACCUM(0, m0);
					if (SINGLE || i1 != sentinel) V3_LJ_PAIR_ACCUM(1, m1);
				}
				if (SINGLE) {
					const C f24 = eps24;      // eps24 and the shift are pair independent: applied once per list
					fx[0] = f24 * (fx[0] + fx[1]); fy[0] = f24 * (fy[0] + fy[1]); fz[0] = f24 * (fz[0] + fz[1]);
					vx[0] = f24 * (vx[0] + vx[1]); vy[0] = f24 * (vy[0] + vy[1]); vz[0] = f24 * (vz[0] + vz[1]);
					up[0] = f24 * (up[0] + up[1]) + (C) cnt * shift6;
					fx[1] = fy[1] = fz[1] = vx[1] = vy[1] = vz[1] = up[1] = C(0);
				}
				Fd[0] += (double) fx[0] + (double) fx[1]; Fd[1] += (double) fy[0] + (double) fy[1]; Fd[2] += (double) fz[0] + (double) fz[1];
				Vd[0] += (double) vx[0] + (double) vx[1]; Vd[1] += (double) vy[0] + (double) vy[1]; Vd[2] += (double) vz[0] + (double) vz[1];
				Ud += (double) up[0] + (double) up[1];
				w = list_s;
				if (r >= 9) break;
			}
		}
		// ---- results ----
		if (have) {
			const size_t cp = a.cap;
			const uint32_t t = s_rowbase[4] + trel;
			acc[0] += 0.5 * Ud;                          // every pair is visited from both of its ends
			acc[1] += 0.5 * (Vd[0] + Vd[1] + Vd[2]);
			a.F[t] = Fd[0]; a.F[cp + t] = Fd[1]; a.F[2 * cp + t] = Fd[2];
			a.Vi[t] = 0.5 * Vd[0]; a.Vi[cp + t] = 0.5 * Vd[1]; a.Vi[2 * cp + t] = 0.5 * Vd[2];
			if (a.M) { a.M[t] = 0.; a.M[cp + t] = 0.; a.M[2 * cp + t] = 0.; }
		}
		}
		// per-warp partial sums of the chunk (slot 2 * warp of its 8 doubles; unused slots are zeroed)
		#pragma unroll
		for (int k = 0; k < 2; k++) {
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
		}
		if ((threadIdx.x & 31) == 0) *reinterpret_cast<double2*>(a.partials + (size_t) ch * 8 + 2 * (threadIdx.x >> 5)) = make_double2(acc[0], acc[1]);
		if (TPB < 128 && threadIdx.x >= TPB / 32 && threadIdx.x < 4) *reinterpret_cast<double2*>(a.partials + (size_t) ch * 8 + 2 * threadIdx.x) = make_double2(0., 0.);
	}
}

int reduce_force_chunks(ls1b200_ctx* c, uint32_t nrows, int per_warp, bool kin = false, bool accumulate = false);      // force_1clj_tiled.cu

template <class C, int S, int ALIGN>
static int launch_v3(ls1b200_ctx* c) {
	typedef typename Vec4<C>::type V4;
	constexpr int TPB = V3Cfg<C>::TPB, W = V3Cfg<C>::W, LCAP = V3Cfg<C>::LCAP;
	constexpr int CSF = (V3_MAXC + 2) * S + 4;
	const Geom& g = c->geom;
	uint32_t nrows = 0;
	LS1_TRY(build_chunks2(c, TPB, W, ALIGN, S > 1, &nrows));
	V3Args<C> a;
	a.mpos = (const V4*) c->mpos; a.cell_start = c->cell_start; a.csf = S > 1 ? c->cell_start_fine : c->cell_start; a.cell_of = c->cell_of2;
	a.bx = c->d_border;
	if (sizeof(C) == 4) { a.eps24 = (const C*) c->d_eps24f; a.sig2 = (const C*) c->d_sig2f; a.shift6 = (const C*) c->d_shift6f; }
	else { a.eps24 = (const C*) c->d_eps24; a.sig2 = (const C*) c->d_sig2; a.shift6 = (const C*) c->d_shift6; }
	a.F = c->F; a.M = c->rotates ? c->M : nullptr; a.Vi = c->Vi; a.partials = c->partials;
	a.chunks = (const Chunk2*) c->chunks2; a.n_chunks = c->chunk_row_start + nrows; a.work_counter = c->work_counter;
	a.cap = (uint32_t) c->cap; a.n_lj = c->n_lj; a.nx = g.cells[0]; a.ny = g.cells[1]; a.nz = g.cells[2];
	const double rclj = c->cfg.cutoff_lj;
	a.ljrc2 = (C) (rclj * rclj);
	a.eps24_0 = (C) c->h_eps24[0]; a.sig2_0 = (C) c->h_sig2[0]; a.shift6_0 = (C) c->h_shift6[0];
	a.finv = c->fine_inv;
	// Trim bound: a pair passes the calc-precision test only if |x_i - x_j| < rc (1 + 1e-5) in calc precision; the FP64 state the
	// sort binned differs from the calc-precision record by at most half an FP32 ulp of the coordinate.
	const double amax = std::max(std::fabs(g.hmin[0]), std::fabs(g.hmax[0]));
	a.trim_rc = rclj * (1. + 1e-5) + 2.4e-7 * amax;
	const size_t smem = (size_t) (9 * W + 1) * sizeof(V4) + (size_t) (LCAP + 1) * TPB * 2 + (size_t) 9 * CSF * 4;
	auto kern = c->comps.size() == 1 ? k_force_1clj_v3<C, true, S, ALIGN> : k_force_1clj_v3<C, false, S, ALIGN>;
	LS1_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	int per_sm = 0;
	LS1_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TPB, smem));
	if (per_sm < 1) return fail("force_1clj_v3: kernel does not fit on an SM");
	const uint32_t grid = (uint32_t) (c->sm_count * per_sm);
	LS1_LAUNCH(c, PH_FORCE, kern, grid, TPB, smem, a);
	return reduce_force_chunks(c, nrows, 1);
}

template <class C> static int launch_v3_variant(ls1b200_ctx* c) {
	// LS1B200_V3_ALIGN = 1 | 2 | 4 (lanes per fine bin padded to a multiple); LS1B200_V3_TRIM = 0 disables the x trim
	const int align = [] { const char* e = getenv("LS1B200_V3_ALIGN"); const int v = e ? atoi(e) : 1; return (v == 2 || v == 4) ? v : 1; }();
	const bool trim = [] { const char* e = getenv("LS1B200_V3_TRIM"); return !(e && atoi(e) == 0); }();
	if (trim && c->fine_s == 4) {
		if (align == 4) return launch_v3<C, 4, 4>(c);
		if (align == 2) return launch_v3<C, 4, 2>(c);
		return launch_v3<C, 4, 1>(c);
	}
	if (trim && c->fine_s == 8) {
		if (align == 2) return launch_v3<C, 8, 2>(c);
		return launch_v3<C, 8, 1>(c);
	}
	if (trim && c->fine_s == 2) {
		if (align == 2) return launch_v3<C, 2, 2>(c);
		return launch_v3<C, 2, 1>(c);
	}
	if (align == 4) return launch_v3<C, 1, 4>(c);
	if (align == 2) return launch_v3<C, 1, 2>(c);
	return launch_v3<C, 1, 1>(c);
}

int force_1clj_v3(ls1b200_ctx* c) { return c->mixed ? launch_v3_variant<float>(c) : launch_v3_variant<double>(c); }

}  // namespace ls1b200
