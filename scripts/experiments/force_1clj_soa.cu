// EXPERIMENT, NOT BUILT: see profiles/r01_notes_round1.md ("structure-of-arrays windows"). Counts pairs twice when cells at a row
// end are empty (needs sentinel gaps between rows of the SoA planes); kept as the starting point for that version.

// force_1clj_soa.cu -- the production 1CLJ cell-pair kernel of the mixed-precision build: force_1clj_tiled.cu with the
// neighbour windows held as structure-of-arrays planes (x[], y[], z[], cid[]) in shared memory.
//
// Reference: same arithmetic as force_1clj_tiled.cu (VectorizedCellProcessor::_calculatePairs LJ block,
// adapter/VectorizedCellProcessor.cpp:1029-1178; _loopBodyLJ :171-224; calcDistLookup,
// vectorization/SIMD_VectorizedCellProcessorHelpers.h:427-472; FullMolecule::calcFM, molecules/FullMolecule.cpp:548-626).
// The cell pairs are those of the C08 traversal (LinkedCellTraversals/C08BasedTraversals.h:48-99), visited from both ends.
// This is also the layout of the reference's own CellDataSoA (adapter/CellDataSoA.h:41-74: _mol_pos x | y | z).
//
// Why (profiles/r01_notes_round1.md): the test loop of the tiled kernel is bound by the shared-memory pipe -- every LDS.128
// of one float4 candidate costs a wavefront per distinct cell among the lanes (3.1 on average), 16.4 wavefront cycles against
// 13 issue cycles per 4 candidates. With SoA planes one LDS.128 delivers the x of FOUR candidates: 3 loads per 4 candidates
// instead of 4. Ranges are rounded to multiples of four candidates (the planes are 16-byte addressed), which also removes the
// single-candidate remainder loop; the up to three extra candidates at either end are molecules of a cell that is not a
// neighbour (or of the next row), so the reference's test rejects them -- guaranteed as long as a cell is wider than the
// cutoff by more than FP32 rounding, which the launcher checks (otherwise the tiled kernel runs).
//
// HBM side: `mpos_soa` = four float planes [cap + 64] written by k_gather next to the float4 records (16 B per molecule more);
// a window row is one TMA bulk copy per plane, source and size rounded to 16 bytes.
#include "tile.cuh"

namespace ls1b200 {

namespace {

constexpr int S_TPB = 128, S_W = 176, S_WS = S_W + 8, S_LCAP = 64;      // WS: plane row stride (window + alignment slack)
constexpr uint32_t S_LSTRIDE = S_TPB * 2;

struct SArgs {
	const float* soa;            // planes x, y, z, cid bits; plane stride `plane`
	size_t plane;
	const float4* mpos;          // float4 records (slow path of dense cells only)
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
};

__device__ __forceinline__ float lds_f32(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory"); return v; }

// LJ 12-6 of one pair, accumulated into FP32 partial sums (slot u); with a single component 24 eps is pair independent
// and is applied once per list
#define S_PAIR(u, jx, jy, jz, jw)                                                              \
	{                                                                                           \
		const float dx = mix - (jx), dy = miy - (jy), dz = miz - (jz);                            \
		const float dx2 = dx * dx, dy2 = dy * dy, dz2 = dz * dz;                                  \
		const float r2 = dx2 + dy2 + dz2;                                                         \
		if (!SINGLE) {                                                                            \
			const size_t ti = (size_t) cid_i * a.n_lj + (__float_as_uint(jw) & 0xFFFFu);           \
			eps24 = a.eps24[ti]; sig2 = a.sig2[ti]; shift6 = a.shift6[ti];                          \
		}                                                                                         \
		const float r2inv = rcp_fast<float>(r2);                                                  \
		const float lj2 = sig2 * r2inv, lj6 = lj2 * lj2 * lj2, lj12 = lj6 * lj6, lj12m6 = lj12 - lj6; \
		float scale = r2inv * (lj12 + lj12m6);                                                    \
		if (!SINGLE) scale *= eps24;                                                              \
		fx[u] = fmaf(dx, scale, fx[u]); fy[u] = fmaf(dy, scale, fy[u]); fz[u] = fmaf(dz, scale, fz[u]); \
		vx[u] = fmaf(dx2, scale, vx[u]); vy[u] = fmaf(dy2, scale, vy[u]); vz[u] = fmaf(dz2, scale, vz[u]); \
		if (SINGLE) up[u] += lj12m6; else up[u] += eps24 * lj12m6 + shift6;                       \
	}

template <bool SINGLE>
__global__ void __launch_bounds__(S_TPB) k_force_1clj_soa(SArgs a) {
	typedef float C;
	constexpr int TPB = S_TPB, WS = S_WS, LCAP = S_LCAP, NP = SINGLE ? 3 : 4;
	constexpr uint32_t LSTRIDE = S_LSTRIDE, PB = (uint32_t) (9 * WS + 4) * 4u;      // bytes per plane in shared memory (+ the sentinel)
	extern __shared__ __align__(128) unsigned char smem_raw[];
	uint16_t* list = reinterpret_cast<uint16_t*>(smem_raw + (size_t) NP * PB);                   // [LCAP+1][TPB]
	uint32_t* cs_all = reinterpret_cast<uint32_t*>(list + (size_t) (LCAP + 1) * TPB);           // [2][12][CH_MAXC+3]
	constexpr int CSB = 12 * (CH_MAXC + 3);
	__shared__ uint32_t s_rowbase_all[2][9];
	__shared__ __align__(8) unsigned long long s_bar;
	__shared__ uint32_t s_next;
	const uint32_t xs = smem_u32(smem_raw), ys = xs + PB, zs = ys + PB, ws = zs + PB;      // planes (ws only if !SINGLE)
	const uint32_t list_s = smem_u32(list) + threadIdx.x * 2, bar = smem_u32(&s_bar);
	const uint32_t n_chunks = *a.n_chunks;
	constexpr uint32_t SENT = (uint32_t) (9 * WS);      // index of the far-away record that pads odd lists
	if (threadIdx.x == 0) {
		mbar_init(bar, 1);
		float* p = reinterpret_cast<float*>(smem_raw);
		for (int k = 0; k < NP; k++) p[(size_t) k * (PB / 4) + SENT] = k < 3 ? 1e18f : 0.f;
		s_next = atomicAdd(a.work_counter, 1u);
	}
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	fence_proxy_async();
	__syncthreads();
	// cell_start cache of the 9 rows of a chunk (cells xa-1 .. xb+2 -> nc+3 entries each), RELATIVE to the 16-byte aligned start
	// of the row's window (rowbase = first record of cell xa-1 rounded down to a multiple of 4)
	auto fill_cs = [&](const Chunk& q, uint32_t* csq, uint32_t* rbq) {
		const int qn = q.ncells;
		const uint32_t qR = q.cell0 / (uint32_t) a.nx;
		const int qxa = (int) (q.cell0 % (uint32_t) a.nx);
		for (int k = threadIdx.x; k < 9 * (qn + 3); k += TPB) {
			const int r = k / (qn + 3), c = k % (qn + 3);
			const uint32_t row = (uint32_t) ((int) qR + (r / 3 - 1) * a.ny + (r % 3 - 1)) * a.nx;
			const uint32_t b0 = a.cell_start[row + qxa - 1] & ~3u;
			csq[r * (CH_MAXC + 3) + c] = a.cell_start[row + qxa - 1 + c] - b0;
			if (c == 0) rbq[r] = b0;
		}
	};
	uint32_t ch = s_next;
	Chunk cd; cd.cell0 = 0; cd.ncells = 0; cd.flags = 0;
	int cur = 0;
	if (ch < n_chunks) { cd = a.chunks[ch]; if (!(cd.flags & CHUNK_BIG)) fill_cs(cd, cs_all, s_rowbase_all[0]); }
	__syncthreads();
	uint32_t parity = 0;
	while (ch < n_chunks) {
		uint32_t* cs = cs_all + cur * CSB;
		uint32_t* s_rowbase = s_rowbase_all[cur];
		const bool big = (cd.flags & CHUNK_BIG) != 0;
		const int nc = cd.ncells;
		const uint32_t R = cd.cell0 / (uint32_t) a.nx;                 // row index cz*ny + cy
		if (threadIdx.x == 0) {
			if (!big) {
				// one bulk copy per plane and row: records [rowbase, rowbase + ceil4(end of cell xb+1)), whole 16-byte units
				uint32_t total = 0;
				for (int r = 0; r < 9; r++) total += ((cs[r * (CH_MAXC + 3) + nc + 2] + 3u) & ~3u) * 4u * NP;
				fence_proxy_async();      // the previous chunk's generic-proxy reads of the windows precede these async-proxy writes
				mbar_expect_tx(bar, total);
				for (int r = 0; r < 9; r++) {
					const uint32_t b = s_rowbase[r], n4 = (cs[r * (CH_MAXC + 3) + nc + 2] + 3u) & ~3u;
					if (n4 == 0) continue;
					#pragma unroll
					for (int k = 0; k < NP; k++) tma_bulk_g2s(xs + (uint32_t) k * PB + (uint32_t) (r * WS) * 4u, a.soa + (size_t) k * a.plane + b, n4 * 4u, bar);
				}
			}
			s_next = atomicAdd(a.work_counter, 1u);      // the NEXT chunk of this CTA
		}
		__syncthreads();
		const uint32_t nch = s_next;
		Chunk ncd; ncd.cell0 = 0; ncd.ncells = 0; ncd.flags = 0;
		if (nch < n_chunks) { ncd = a.chunks[nch]; if (!(ncd.flags & CHUNK_BIG)) fill_cs(ncd, cs_all + (cur ^ 1) * CSB, s_rowbase_all[cur ^ 1]); }
		double acc[2] = {0., 0.};
		if (big) {
			// ---- slow generic path: a cell denser than the tile. Global loads of the float4 records, pair evaluated as soon as it is found. ----
			const uint32_t ib = a.cell_start[cd.cell0], ie = a.cell_start[cd.cell0 + nc];
			for (uint32_t t = ib + threadIdx.x; t < ie; t += TPB) {
				const float4 mi = a.mpos[t];
				const float mix = mi.x, miy = mi.y, miz = mi.z;
				const uint32_t cid_i = __float_as_uint(mi.w) & 0xFFFFu;
				const int cx = (int) (a.cell_of[t] % (uint32_t) a.nx);
				C eps24 = a.eps24_0, sig2 = a.sig2_0, shift6 = a.shift6_0;
				double Fd[3] = {0., 0., 0.}, Vd[3] = {0., 0., 0.}, Ud = 0.;
				for (int r = 0; r < 9; r++) {
					const uint32_t row = (uint32_t) ((int) R + (r / 3 - 1) * a.ny + (r % 3 - 1)) * a.nx + cx - 1;
					const uint32_t jb = a.cell_start[row], je = a.cell_start[row + 3];
					for (uint32_t j0 = jb; j0 < je; j0 += 32) {
						C fx[1] = {0.f}, fy[1] = {0.f}, fz[1] = {0.f}, vx[1] = {0.f}, vy[1] = {0.f}, vz[1] = {0.f}, up[1] = {0.f};
						int nh = 0;
						for (uint32_t j = j0; j < min(je, j0 + 32); j++) {
							const float4 mj = a.mpos[j];
							const C r2t = molpair_r2<C>(mix, miy, miz, mj.x, mj.y, mj.z);
							if ((r2t < a.ljrc2) && (r2t != 0.f)) { S_PAIR(0, mj.x, mj.y, mj.z, mj.w); nh++; }
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
			const uint32_t* csc = cs + 4 * (CH_MAXC + 3);          // centre row (record indices relative to its aligned start)
			const uint32_t n_mol = csc[nc + 1] - csc[1];
			const bool have = threadIdx.x < n_mol;
			const uint32_t trel = csc[1] + threadIdx.x;
			int kx = 0;                                            // cell (inside the chunk) of my molecule
			while (kx + 1 < nc && csc[kx + 2] <= trel) kx++;
			mbar_wait(bar, parity); parity ^= 1u;
			// ---- phase 1 + phase 2 (see force_1clj_tiled.cu); candidates in groups of four ----
			float mix = 1e18f, miy = 1e18f, miz = 1e18f; uint32_t cid_i = 0;      // an absent molecule: no candidate is ever in range
			if (have) {
				const uint32_t o = (uint32_t) (4 * WS + (int) trel) * 4u;
				mix = lds_f32(xs + o); miy = lds_f32(ys + o); miz = lds_f32(zs + o);
				if (!SINGLE) cid_i = __float_as_uint(lds_f32(ws + o)) & 0xFFFFu;
			}
			double Fd[3] = {0., 0., 0.}, Vd[3] = {0., 0., 0.}, Ud = 0.;
			const uint32_t lbase = list_s, wlimit = lbase + (uint32_t) (LCAP - 4) * LSTRIDE;
			uint32_t w = lbase;
			uint32_t src = 0, end = 0;              // candidate INDEX (multiple of 4) of the next group / end of the row range
			int r = -1;
			for (;;) {
				bool flush = true;
				if (r < 9) {
					if (r == 4) {
						while (src < end && w <= wlimit) {
							const float4 X = lds_v4<float>(xs + src * 4u), Y = lds_v4<float>(ys + src * 4u), Z = lds_v4<float>(zs + src * 4u);
							const C r0 = molpair_r2<C>(mix, miy, miz, X.x, Y.x, Z.x), r1 = molpair_r2<C>(mix, miy, miz, X.y, Y.y, Z.y);
							const C r2 = molpair_r2<C>(mix, miy, miz, X.z, Y.z, Z.z), r3 = molpair_r2<C>(mix, miy, miz, X.w, Y.w, Z.w);
							if (r0 < a.ljrc2 && r0 != 0.f) { sts16(w, src); w += LSTRIDE; }
							if (r1 < a.ljrc2 && r1 != 0.f) { sts16(w, src + 1u); w += LSTRIDE; }
							if (r2 < a.ljrc2 && r2 != 0.f) { sts16(w, src + 2u); w += LSTRIDE; }
							if (r3 < a.ljrc2 && r3 != 0.f) { sts16(w, src + 3u); w += LSTRIDE; }
							src += 4u;
						}
					} else {
						while (src < end && w <= wlimit) {
							const float4 X = lds_v4<float>(xs + src * 4u), Y = lds_v4<float>(ys + src * 4u), Z = lds_v4<float>(zs + src * 4u);
							const C r0 = molpair_r2<C>(mix, miy, miz, X.x, Y.x, Z.x), r1 = molpair_r2<C>(mix, miy, miz, X.y, Y.y, Z.y);
							const C r2 = molpair_r2<C>(mix, miy, miz, X.z, Y.z, Z.z), r3 = molpair_r2<C>(mix, miy, miz, X.w, Y.w, Z.w);
							if (r0 < a.ljrc2) { sts16(w, src); w += LSTRIDE; }
							if (r1 < a.ljrc2) { sts16(w, src + 1u); w += LSTRIDE; }
							if (r2 < a.ljrc2) { sts16(w, src + 2u); w += LSTRIDE; }
							if (r3 < a.ljrc2) { sts16(w, src + 3u); w += LSTRIDE; }
							src += 4u;
						}
					}
					if (w <= wlimit) {                   // the row is finished (the list is not the reason the loop stopped)
						r++;
						if (r < 9) {
							uint32_t kb = 0, ke = 0;
							if (have) { kb = cs[r * (CH_MAXC + 3) + kx] & ~3u; ke = (cs[r * (CH_MAXC + 3) + kx + 3] + 3u) & ~3u; }      // cells cx-1 .. cx+1, whole groups of four
							src = (uint32_t) (r * WS) + kb; end = (uint32_t) (r * WS) + ke;
							flush = false;
						}
					}
				}
				if (flush) {
					// ---- phase 2: dense LJ over the list; two interleaved FP32 accumulator sets ----
					const int cnt = (int) ((w - lbase) / LSTRIDE);
					C fx[2] = {0.f, 0.f}, fy[2] = {0.f, 0.f}, fz[2] = {0.f, 0.f}, vx[2] = {0.f, 0.f}, vy[2] = {0.f, 0.f}, vz[2] = {0.f, 0.f}, up[2] = {0.f, 0.f};
					C eps24 = a.eps24_0, sig2 = a.sig2_0, shift6 = a.shift6_0;
					if (cnt & 1) sts16(w, SENT);      // pad odd lists with the far-away record (all LJ terms underflow to 0)
					#pragma unroll 1
					for (uint32_t e = lbase; e < w; e += 2 * LSTRIDE) {
						const uint32_t i0 = lds16(e) * 4u, i1 = lds16(e + LSTRIDE) * 4u;
						const float x0 = lds_f32(xs + i0), y0 = lds_f32(ys + i0), z0 = lds_f32(zs + i0);
						const float x1 = lds_f32(xs + i1), y1 = lds_f32(ys + i1), z1 = lds_f32(zs + i1);
						float w0 = 0.f, w1 = 0.f;
						if (!SINGLE) { w0 = lds_f32(ws + i0); w1 = lds_f32(ws + i1); }
						S_PAIR(0, x0, y0, z0, w0);
						if (SINGLE || i1 != SENT * 4u) S_PAIR(1, x1, y1, z1, w1);
					}
					if (SINGLE) {
						const C f24 = eps24;
						fx[0] = f24 * (fx[0] + fx[1]); fy[0] = f24 * (fy[0] + fy[1]); fz[0] = f24 * (fz[0] + fz[1]);
						vx[0] = f24 * (vx[0] + vx[1]); vy[0] = f24 * (vy[0] + vy[1]); vz[0] = f24 * (vz[0] + vz[1]);
						up[0] = f24 * (up[0] + up[1]) + (C) cnt * shift6;
						fx[1] = fy[1] = fz[1] = vx[1] = vy[1] = vz[1] = up[1] = 0.f;
					}
					Fd[0] += (double) fx[0] + (double) fx[1]; Fd[1] += (double) fy[0] + (double) fy[1]; Fd[2] += (double) fz[0] + (double) fz[1];
					Vd[0] += (double) vx[0] + (double) vx[1]; Vd[1] += (double) vy[0] + (double) vy[1]; Vd[2] += (double) vz[0] + (double) vz[1];
					Ud += (double) up[0] + (double) up[1];
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
				a.F[t] = Fd[0]; a.F[cp + t] = Fd[1]; a.F[2 * cp + t] = Fd[2];
				a.Vi[t] = 0.5 * Vd[0]; a.Vi[cp + t] = 0.5 * Vd[1]; a.Vi[2 * cp + t] = 0.5 * Vd[2];
				if (a.M) { a.M[t] = 0.; a.M[cp + t] = 0.; a.M[2 * cp + t] = 0.; }
			}
		}
		// per-warp partial sums of the chunk (slot 2 * warp of its 8 doubles): no block-wide reduction
		#pragma unroll
		for (int k = 0; k < 2; k++) {
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
		}
		if ((threadIdx.x & 31) == 0) *reinterpret_cast<double2*>(a.partials + (size_t) ch * 8 + 2 * (threadIdx.x >> 5)) = make_double2(acc[0], acc[1]);
		__syncthreads();      // windows and lists are reused by the next chunk; its cell_start cache is complete
		ch = nch; cd = ncd; cur ^= 1;
	}
}

}  // namespace

int reduce_force_chunks(ls1b200_ctx* c, uint32_t nrows, int per_warp);      // force_1clj_tiled.cu

// Mixed-precision ensembles of 1CLJ components whose cells are wider than the cutoff by more than FP32 rounding (the rounded-up
// candidate ranges rely on the test rejecting molecules of non-neighbour cells); everything else takes the tiled kernel.
int force_1clj_soa(ls1b200_ctx* c) {
	const Geom& g = c->geom;
	bool ok = c->mixed && c->mpos_soa != nullptr;
	for (int d = 0; d < 3; d++) ok = ok && (g.clen[d] > c->cfg.cutoff_lj * (1. + 1e-5));
	if (!ok) return force_1clj_tiled(c);
	uint32_t nrows = 0;
	LS1_TRY(build_chunks(c, S_TPB, S_W, 1, &nrows));
	SArgs a;
	a.soa = c->mpos_soa; a.plane = (size_t) c->cap + 64; a.mpos = (const float4*) c->mpos;
	a.cell_start = c->cell_start; a.cell_of = c->cell_of2;
	a.eps24 = c->d_eps24f; a.sig2 = c->d_sig2f; a.shift6 = c->d_shift6f;
	a.F = c->F; a.M = c->rotates ? c->M : nullptr; a.Vi = c->Vi; a.partials = c->partials;
	a.chunks = (const Chunk*) c->chunks; a.n_chunks = c->chunk_row_start + nrows; a.work_counter = c->work_counter;
	a.cap = (uint32_t) c->cap; a.n_lj = c->n_lj; a.nx = g.cells[0]; a.ny = g.cells[1]; a.nz = g.cells[2];
	const double rclj = c->cfg.cutoff_lj;
	a.ljrc2 = (float) (rclj * rclj);
	a.eps24_0 = (float) c->h_eps24[0]; a.sig2_0 = (float) c->h_sig2[0]; a.shift6_0 = (float) c->h_shift6[0];
	const bool single = c->comps.size() == 1;
	const size_t smem = (size_t) (single ? 3 : 4) * (9 * S_WS + 4) * 4 + (size_t) (S_LCAP + 1) * S_TPB * 2 + 2 * 12 * (CH_MAXC + 3) * 4;
	auto kern = single ? k_force_1clj_soa<true> : k_force_1clj_soa<false>;
	LS1_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	int per_sm = 0;
	LS1_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, S_TPB, smem));
	if (per_sm < 1) return fail("force_1clj_soa: kernel does not fit on an SM");
	const uint32_t grid = (uint32_t) (c->sm_count * per_sm);
	LS1_LAUNCH(c, PH_FORCE, kern, grid, S_TPB, smem, a);
	return reduce_force_chunks(c, nrows, 1);
}

}  // namespace ls1b200
