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
//      records the hits in a 64-bit mask held in registers (cheap, branch-free inner loop),
//   2. walks the set bits and evaluates the LJ pair only for pairs inside the cutoff (no predicated-off LJ work).
// Mixed build: pair arithmetic FP32; per-row partial sums FP32 (at most a few dozen terms), flushed into FP64
// accumulators after every row -- FP64 accumulation like RealAccumVec without a convert+DADD per pair.
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

template <class C, bool SINGLE>
__global__ void __launch_bounds__(128) k_force_1clj(LJArgs<C> a) {
	typedef typename Vec4<C>::type V4;
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t n_slots = min(a.scal->n_slots2[a.parity], a.cap);
	double acc[2] = {0., 0.};   // 6*U_LJ, virial
	bool active = t < n_slots;
	V4 mi; uint32_t bits = 0;
	if (active) { mi = a.mpos[t]; bits = wbits<C>(mi.w); }
	const bool inert = !active || (bits & 0x80000000u);
	if (active) {
		double Fd[3] = {0., 0., 0.}, Vd[3] = {0., 0., 0.};
		if (!inert) {
			const uint32_t c = a.cell_of[t];
			const int cx = c % a.nx, cy = (c / a.nx) % a.ny, cz = c / (a.nx * a.ny);
			const uint32_t cid_i = bits & 0xFFFFu;
			C eps24 = a.eps24_0, sig2 = a.sig2_0, shift6 = a.shift6_0;
			for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) {
				const uint32_t row = (uint32_t) (((cz + dz) * a.ny + (cy + dy)) * a.nx + (cx - 1));
				const uint32_t jb = a.cell_start[row], je = a.cell_start[row + 3];
				C fx = C(0), fy = C(0), fz = C(0), vx = C(0), vy = C(0), vz = C(0), up = C(0);
				for (uint32_t j0 = jb; j0 < je; j0 += 64) {
					const uint32_t n = min(64u, je - j0);
					unsigned long long mask = 0ull;
					for (uint32_t k = 0; k < n; k++) {
						const V4 mj = a.mpos[j0 + k];
						const C r2 = molpair_r2<C>(mi.x, mi.y, mi.z, mj.x, mj.y, mj.z);
						// strict < (calcDistLookup); distance 0 only occurs for the molecule itself (SingleCellPolicy_ masks it)
						const bool in = (r2 < a.ljrc2) && (r2 != C(0));
						mask |= (unsigned long long) in << k;
					}
					while (mask) {
						const int k = __ffsll((long long) mask) - 1;
						mask &= mask - 1ull;
						const V4 mj = a.mpos[j0 + k];
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
						fx += px; fy += py; fz += pz;
						vx += dx * px; vy += dy * py; vz += dz * pz;
						up += eps24 * lj12m6 + shift6;
					}
				}
				Fd[0] += (double) fx; Fd[1] += (double) fy; Fd[2] += (double) fz;
				Vd[0] += (double) vx; Vd[1] += (double) vy; Vd[2] += (double) vz;
				acc[0] += (double) up;
			}
			acc[0] *= 0.5;                          // every pair is visited from both ends
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
	const uint32_t gx = (uint32_t) ((c->cap + 127) / 128);
	if (gx > c->partial_blocks) return fail("force_1clj: partials buffer too small");
	if (c->comps.size() == 1) LS1_LAUNCH(c, PH_FORCE, (k_force_1clj<C, true>), gx, 128, 0, a);
	else LS1_LAUNCH(c, PH_FORCE, (k_force_1clj<C, false>), gx, 128, 0, a);
	LS1_LAUNCH(c, PH_FORCE, k_reduce_force_lj, 1, 1024, 0, c->partials, gx, c->d_scal);
	return 0;
}

int force_1clj(ls1b200_ctx* c) { return c->mixed ? launch_1clj<float>(c) : launch_1clj<double>(c); }

}  // namespace ls1b200
