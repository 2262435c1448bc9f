// force_1clj.cu -- fast path of the cell-pair traversal for ensembles whose components are all single-centre
// Lennard-Jones molecules (BASELINE.json configs 1, 2 and 5): no site expansion, no per-site scratch, F / Vi written
// directly per molecule.
//
// Reference: VectorizedCellProcessor::_calculatePairs LJ block (adapter/VectorizedCellProcessor.cpp:1029-1178) with
//   _loopBodyLJ (:171-224) and calcDistLookup (vectorization/SIMD_VectorizedCellProcessorHelpers.h:427-472);
//   FullMolecule::calcFM for a molecule with one site at the origin: F = f, M = 0, Vi = V/2 (FullMolecule.cpp:548-626).
//
// Mapping: one thread per owned molecule i; the 27 neighbour cells are 9 contiguous slot ranges (rows). Per row a thread
//   1. evaluates the reference's molecule-distance test on every candidate j (bit-identical FP32/FP64 sequence) and
//      appends the hits to its private neighbour list in shared memory (one predicated store per candidate),
//   2. evaluates the LJ pair densely over that list: no LJ arithmetic is spent on candidates outside the cutoff and
//      all lanes of a warp run nearly the same number of iterations (list lengths vary far less than per-row hits).
// Mixed build: pair arithmetic FP32; partial sums FP32 over one list (<= 80 terms, two interleaved accumulators), folded
// into FP64 accumulators -- FP64 accumulation like RealAccumVec without a convert+DADD per pair.
#include "pairs.cuh"
#include "types.cuh"

namespace ls1b200 {

template <class C>
struct LJArgs {
	const typename Vec4<C>::type* mpos;
	const uint32_t* cell_start;
	const uint32_t* cell_of;
	const C* eps24; const C* sig2; const C* shift6;   // [n_lj*n_lj] (n_lj == number of components here)
	double* F; double* M; double* Vi;
	double* partials;
	const DevScalars* scal;
	uint32_t parity, cap, n_lj;
	int nx, ny, nz;
	C ljrc2;
	C eps24_0, sig2_0, shift6_0;   // the single-component constants
};

template <class C> __device__ __forceinline__ C fast_rcp(C x);
template <> __device__ __forceinline__ float fast_rcp<float>(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
template <> __device__ __forceinline__ double fast_rcp<double>(double x) { return 1.0 / x; }

// shared-memory access through 32-bit shared-window addresses (keeps ptxas from re-deriving the generic address per access)
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }

constexpr int LJ_THREADS = 128;
constexpr int LJ_LCAP = 56;      // neighbour-list entries per thread held in shared memory (flushed when full)

// Phase 2: LJ 12-6 for the listed (in-range) neighbours of molecule i; two interleaved FP32 partial sums (ILP, and
// half the summation length each), folded into the FP64 accumulators.
template <class C, bool SINGLE>
__device__ __forceinline__ void lj_process_list(const LJArgs<C>& a, uint32_t lst, int cnt, const typename Vec4<C>::type& mi, uint32_t cid_i,
                                                double* Fd, double* Vd, double& Ud) {
	typedef typename Vec4<C>::type V4;
	C fx[2] = {C(0), C(0)}, fy[2] = {C(0), C(0)}, fz[2] = {C(0), C(0)}, vx[2] = {C(0), C(0)}, vy[2] = {C(0), C(0)}, vz[2] = {C(0), C(0)}, up[2] = {C(0), C(0)};
	C eps24 = a.eps24_0, sig2 = a.sig2_0, shift6 = a.shift6_0;
	// an odd list is padded with the sentinel record mpos[cap] (a point ~1e18 away: every LJ term underflows to exactly 0)
	if (cnt & 1) { sts32(lst + (uint32_t) cnt * (LJ_THREADS * 4), a.cap); }
	#pragma unroll 1
	for (int e = 0; e < cnt; e += 2) {
		#pragma unroll
		for (int u = 0; u < 2; u++) {
			const uint32_t j = lds32(lst + (uint32_t) (e + u) * (LJ_THREADS * 4));
			const V4 mj = a.mpos[j];
			const C dx = mi.x - mj.x, dy = mi.y - mj.y, dz = mi.z - mj.z;
			const C r2 = dx * dx + dy * dy + dz * dz;
			if (!SINGLE) {
				const size_t ti = (size_t) cid_i * a.n_lj + (wbits<C>(mj.w) & 0xFFFFu);
				eps24 = a.eps24[ti]; sig2 = a.sig2[ti]; shift6 = a.shift6[ti];
			}
			const C r2inv = fast_rcp<C>(r2);
			const C lj2 = sig2 * r2inv, lj6 = lj2 * lj2 * lj2, lj12 = lj6 * lj6, lj12m6 = lj12 - lj6;
			const C scale = eps24 * r2inv * (lj12 + lj12m6);
			const C px = dx * scale, py = dy * scale, pz = dz * scale;
			fx[u] += px; fy[u] += py; fz[u] += pz;
			vx[u] += dx * px; vy[u] += dy * py; vz[u] += dz * pz;
			if (SINGLE) up[u] += lj12m6;                       // eps24 and the shift are applied once per list below
			else up[u] += (j == a.cap) ? C(0) : (eps24 * lj12m6 + shift6);
		}
	}
	if (SINGLE) { up[0] = eps24 * (up[0] + up[1]) + (C) cnt * shift6; up[1] = C(0); }
	Fd[0] += (double) fx[0] + (double) fx[1]; Fd[1] += (double) fy[0] + (double) fy[1]; Fd[2] += (double) fz[0] + (double) fz[1];
	Vd[0] += (double) vx[0] + (double) vx[1]; Vd[1] += (double) vy[0] + (double) vy[1]; Vd[2] += (double) vz[0] + (double) vz[1];
	Ud += (double) up[0] + (double) up[1];
}

template <class C, bool SINGLE>
__global__ void __launch_bounds__(LJ_THREADS) k_force_1clj(LJArgs<C> a) {
	typedef typename Vec4<C>::type V4;
	__shared__ uint32_t s_list[(LJ_LCAP + 1) * LJ_THREADS];   // [entry][thread]: conflict-free, one column per thread
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t n_slots = min(a.scal->n_slots2[a.parity], a.cap);
	double acc[2] = {0., 0.};   // 6*U_LJ, virial
	const bool active = t < n_slots;
	V4 mi = make_c4<C>(C(0), C(0), C(0), C(0)); uint32_t bits = 0x80000000u;
	if (active) { mi = a.mpos[t]; bits = wbits<C>(mi.w); }
	const bool inert = (bits & 0x80000000u) != 0;      // halo copy, halo-cell resident or beyond n_slots: no forces wanted
	const uint32_t cid_i = bits & 0xFFFFu;
	const uint32_t lst = (uint32_t) __cvta_generic_to_shared(s_list + threadIdx.x);   // this thread's column
	double Fd[3] = {0., 0., 0.}, Vd[3] = {0., 0., 0.}, Ud = 0.;
	int cx = 1, cy = 1, cz = 1;
	if (!inert) { const uint32_t c = a.cell_of[t]; cx = c % a.nx; cy = (c / a.nx) % a.ny; cz = c / (a.nx * a.ny); }
	int cnt = 0;
	// Phase 1 over the 9 neighbour rows; all lanes of a warp run the row loop together so that the flush vote is uniform
	#pragma unroll 1
	for (int rowi = 0; rowi < 9; rowi++) {
		const int dz = rowi / 3 - 1, dy = rowi % 3 - 1;
		uint32_t jb = 0, je = 0;
		if (!inert) {
			const uint32_t row = (uint32_t) (((cz + dz) * a.ny + (cy + dy)) * a.nx + (cx - 1));
			jb = a.cell_start[row]; je = a.cell_start[row + 3];
		}
		const bool self_row = rowi == 4;
		#pragma unroll 1
		while (__any_sync(0xffffffffu, jb < je)) {
			// make room: every lane may append up to min(len, LJ_LCAP) entries in this pass
			const uint32_t len = min(je - jb, (uint32_t) LJ_LCAP);
			if (__any_sync(0xffffffffu, cnt + (int) len > LJ_LCAP)) {
				lj_process_list<C, SINGLE>(a, lst, cnt, mi, cid_i, Fd, Vd, Ud);
				cnt = 0;
			}
			const uint32_t jend = jb + len;
			const V4* __restrict__ p = a.mpos + jb;
			uint32_t w = lst + (uint32_t) cnt * (LJ_THREADS * 4);
			if (self_row) {
				// the molecule's own cell is in this row: SingleCellPolicy_ masks pairs at distance 0 (the molecule itself)
				for (uint32_t j = jb; j < jend; j++, p++) {
					const V4 mj = *p;
					const C r2 = molpair_r2<C>(mi.x, mi.y, mi.z, mj.x, mj.y, mj.z);
					if ((r2 < a.ljrc2) && (r2 != C(0))) { sts32(w, j); w += LJ_THREADS * 4; }
				}
			} else {
				#pragma unroll 4
				for (uint32_t j = jb; j < jend; j++, p++) {
					const V4 mj = *p;
					const C r2 = molpair_r2<C>(mi.x, mi.y, mi.z, mj.x, mj.y, mj.z);
					if (r2 < a.ljrc2) { sts32(w, j); w += LJ_THREADS * 4; }      // strict <, bit-identical to calcDistLookup
				}
			}
			cnt = (int) ((w - lst) / (LJ_THREADS * 4));
			jb = jend;
		}
	}
	lj_process_list<C, SINGLE>(a, lst, cnt, mi, cid_i, Fd, Vd, Ud);
	if (active) {
		if (!inert) {
			acc[0] = 0.5 * Ud;                          // every pair is visited from both of its ends
			acc[1] = 0.5 * (Vd[0] + Vd[1] + Vd[2]);
		}
		const size_t cp = a.cap;
		a.F[t] = Fd[0]; a.F[cp + t] = Fd[1]; a.F[2 * cp + t] = Fd[2];
		a.Vi[t] = 0.5 * Vd[0]; a.Vi[cp + t] = 0.5 * Vd[1]; a.Vi[2 * cp + t] = 0.5 * Vd[2];
		if (a.M) { a.M[t] = 0.; a.M[cp + t] = 0.; a.M[2 * cp + t] = 0.; }
	}
	block_reduce_store<2>(acc, a.partials, blockIdx.x);
}

__global__ void __launch_bounds__(1024) k_reduce_force_lj(const double* __restrict__ partials, uint32_t nblocks, DevScalars* scal) {
	double v[2] = {0., 0.};
	for (uint32_t b = threadIdx.x; b < nblocks; b += blockDim.x) { v[0] += partials[(size_t) b * 8]; v[1] += partials[(size_t) b * 8 + 1]; }
	__shared__ double out[8];
	block_reduce_store<2>(v, out, 0);
	__syncthreads();
	if (threadIdx.x == 0) { scal->upot6lj = out[0]; scal->upot_xpoles = 0.; scal->virial = out[1]; scal->my_rf = 0.; }
}

template <class C>
static int launch_1clj(ls1b200_ctx* c) {
	LJArgs<C> a;
	a.mpos = (const typename Vec4<C>::type*) c->mpos; a.cell_start = c->cell_start; a.cell_of = c->cell_of2;
	if (sizeof(C) == 4) { a.eps24 = (const C*) c->d_eps24f; a.sig2 = (const C*) c->d_sig2f; a.shift6 = (const C*) c->d_shift6f; }
	else { a.eps24 = (const C*) c->d_eps24; a.sig2 = (const C*) c->d_sig2; a.shift6 = (const C*) c->d_shift6; }
	a.F = c->F; a.M = c->rotates ? c->M : nullptr; a.Vi = c->Vi; a.partials = c->partials; a.scal = c->d_scal;
	a.parity = c->parity; a.cap = (uint32_t) c->cap; a.n_lj = c->n_lj;
	a.nx = c->geom.cells[0]; a.ny = c->geom.cells[1]; a.nz = c->geom.cells[2];
	const double rclj = c->cfg.cutoff_lj;
	a.ljrc2 = (C) (rclj * rclj);
	a.eps24_0 = (C) c->h_eps24[0]; a.sig2_0 = (C) c->h_sig2[0]; a.shift6_0 = (C) c->h_shift6[0];
	const uint32_t gx = (uint32_t) ((c->cap + LJ_THREADS - 1) / LJ_THREADS);
	if (gx > c->partial_blocks) return fail("force_1clj: partials buffer too small");
	if (c->comps.size() == 1) LS1_LAUNCH(c, PH_FORCE, (k_force_1clj<C, true>), gx, LJ_THREADS, 0, a);
	else LS1_LAUNCH(c, PH_FORCE, (k_force_1clj<C, false>), gx, LJ_THREADS, 0, a);
	LS1_LAUNCH(c, PH_FORCE, k_reduce_force_lj, 1, 1024, 0, c->partials, gx, c->d_scal);
	return 0;
}

int force_1clj(ls1b200_ctx* c) { return c->mixed ? launch_1clj<float>(c) : launch_1clj<double>(c); }

}  // namespace ls1b200
