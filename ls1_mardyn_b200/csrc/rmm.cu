// rmm.cu -- the reference's reduced-memory mode (RMM) on the device: the all-single-precision 1CLJ path in
// which a molecule is 32 bytes (FP32 position, FP32 velocity, id) and the force goes straight into the velocities.
//
// Reference (relative to /root/reference/src):
//   molecules/MoleculeRMM.h, particleContainer/adapter/CellDataSoARMM.h        state: r, v in FP32, id
//   integrators/LeapfrogRMM.cpp + PositionCellProcessorRMM.h:43-58             r = fma(dt, v, r) in FP32
//   parallel/DomainDecompBase.cpp:199-249, :317-376                            periodic wrap / halo copies, FP32 clamps
//   particleContainer/adapter/VCP1CLJRMM.cpp:195-369                           FP32 LJ 12-6, v_i += f dt/m, v_j -= f dt/m; FP32 sums of 6U, virial
//   integrators/VelocityCellProcessorRMM.h:119-127                             sum v^2 (not mass weighted)
//   Domain.cpp:157-252 + longRange/Homogeneous.cpp                             U = U6/6 + UpotCorr, T, beta, p (corrections supplied by the host)
//   thermostats/VelocityScalingThermostat.cpp:68-89                            v *= beta
// Checked by tests/test_gpu_rmm.py against the reference built in RMM configuration (tests/golden/rmm_1clj_2000.npz).
//
// Device layout: planes r[3][cap], v[3][cap] (FP32), id[cap] (bit 63 = periodic halo copy), double buffered by the counting sort;
// the force kernel is the production tiled kernel (force_1clj_tiled.cu) with the kick in its epilogue: it reads the 16-byte
// calc records the sort writes and adds f dt/m to the velocity planes -- no F, no Vi, no FP64 state anywhere. Every pair is
// visited from both ends (the reference applies Newton 3), FP32 partial sums are folded into FP64: same forces to FP32 rounding.
#include "geom.cuh"
#include "pairs.cuh"
#include "types.cuh"

namespace ls1b200 {

constexpr uint64_t RMM_HALO = 1ull << 63;
constexpr int RMM_SHIFT_BITS = 26;

struct RmmArgs {
	Geom g;
	uint32_t cap, parity;
	const double* bx; const double* by; const double* bz;
};

// DomainDecompBase::handleDomainLeavingParticles with FP32 storage
template <int D>
__device__ __forceinline__ float rmm_wrap(const Geom& g, float xf) {
	const double x = (double) xf, L = g.lbox[D];
	if (x < g.bmin[D]) {
		float y = (float) (x + L);
		if ((double) y >= g.bmax[D]) y = nextafterf((float) g.bmax[D], -INFINITY);
		return y;
	}
	if (x >= g.bmax[D]) {
		float y = (float) (x - L);
		if ((double) y < g.bmin[D]) y = (float) g.bmin[D];
		return y;
	}
	return xf;
}
// DomainDecompBase::populateHaloLayerWithCopies with FP32 storage: dir 1 = copy above the box, 2 = below
template <int D>
__device__ __forceinline__ float rmm_halo_coord(const Geom& g, float xf, int dir) {
	const double x = (double) xf, L = g.lbox[D];
	if (dir == 1) { float y = (float) (x + L); if ((double) y < g.bmax[D]) y = nextafterf((float) g.bmax[D], INFINITY); return y; }
	if (dir == 2) { float y = (float) (x - L); if ((double) y >= g.bmin[D]) y = nextafterf((float) g.bmin[D], -INFINITY); return y; }
	return xf;
}
__device__ __forceinline__ bool rmm_code_valid(uint32_t mask, int a, int b, int c) {
	return (a == 0 || ((mask >> (a - 1)) & 1u)) && (b == 0 || ((mask >> (2 + b - 1)) & 1u)) && (c == 0 || ((mask >> (4 + c - 1)) & 1u));
}

// positions (optionally after the pending thermostat factor), wrap, cell index, histogram. `advance` = 0 for the first binning.
__global__ void __launch_bounds__(256) k_rmm_move_key(RmmArgs a, float* __restrict__ r, float* __restrict__ v, const uint64_t* __restrict__ id, float dt, int advance, int use_beta,
                                                       uint32_t* __restrict__ key, uint32_t* __restrict__ count, uint32_t* __restrict__ count_halo, uint32_t* __restrict__ rank,
                                                       const DevScalars* __restrict__ scal) {
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	const Geom& g = a.g;
	const size_t cap = a.cap;
	uint32_t bin = KEY_NONE, mask = 0;
	float x = 0.f, y = 0.f, z = 0.f;
	if (s < min(scal->n_slots2[a.parity], a.cap)) {
		if (id[s] & RMM_HALO) key[s] = KEY_NONE;                 // the halo copies of the previous step are gone (deleteOuterParticles)
		else {
			x = r[s]; y = r[cap + s]; z = r[2 * cap + s];
			if (advance) {
				float vx = v[s], vy = v[cap + s], vz = v[2 * cap + s];
				if (use_beta) { const float b = (float) scal->beta_trans; vx *= b; vy *= b; vz *= b; v[s] = vx; v[cap + s] = vy; v[2 * cap + s] = vz; }
				x = __fmaf_rn(dt, vx, x); y = __fmaf_rn(dt, vy, y); z = __fmaf_rn(dt, vz, z);      // PositionCellProcessorRMM.h:43-58
			}
			if (g.periodic[0]) x = rmm_wrap<0>(g, x);
			if (g.periodic[1]) y = rmm_wrap<1>(g, y);
			if (g.periodic[2]) z = rmm_wrap<2>(g, z);
			r[s] = x; r[cap + s] = y; r[2 * cap + s] = z;
			const int cx = cell1d_tab<0>(g, a.bx, (double) x), cy = cell1d_tab<1>(g, a.by, (double) y), cz = cell1d_tab<2>(g, a.bz, (double) z);
			if (cx < 0 || cy < 0 || cz < 0 || cx >= g.cells[0] || cy >= g.cells[1] || cz >= g.cells[2]) { bin = g.ncells; key[s] = bin; }
			else {
				bin = cell_index(g, cx, cy, cz);
				mask = shift_mask<0>(g, (double) x) | (shift_mask<1>(g, (double) y) << 2) | (shift_mask<2>(g, (double) z) << 4);
				key[s] = bin | (mask << RMM_SHIFT_BITS);
			}
		}
	}
	{	// one histogram atomic per run of equal bins in the warp (slots are in the cell order of the previous step); its return value is the rank
		const int lane = threadIdx.x & 31;
		const uint32_t prev = __shfl_up_sync(0xffffffffu, bin, 1);
		const bool head = lane == 0 || prev != bin;
		const uint32_t heads = __ballot_sync(0xffffffffu, head);
		const uint32_t upto = 0xFFFFFFFFu >> (31 - lane);
		const int start = 31 - __clz(heads & upto);
		const uint32_t above = heads & ~upto;
		const int end = above ? __ffs(above) - 1 : 32;
		uint32_t base = 0;
		if (head && bin != KEY_NONE) base = atomicAdd(&count[bin], (uint32_t) (end - start));
		base = __shfl_sync(0xffffffffu, base, start);
		if (bin != KEY_NONE) rank[s] = base + (uint32_t) (lane - start);
	}
	if (mask) {
		for (int c3 = 0; c3 < 3; c3++) for (int b3 = 0; b3 < 3; b3++) for (int a3 = 0; a3 < 3; a3++) {
			if ((a3 | b3 | c3) == 0 || !rmm_code_valid(mask, a3, b3, c3)) continue;
			const int hx = cell1d_tab<0>(g, a.bx, (double) rmm_halo_coord<0>(g, x, a3)), hy = cell1d_tab<1>(g, a.by, (double) rmm_halo_coord<1>(g, y, b3)),
			          hz = cell1d_tab<2>(g, a.bz, (double) rmm_halo_coord<2>(g, z, c3));
			atomicAdd(&count_halo[cell_index(g, hx, hy, hz)], 1u);
		}
	}
}

__global__ void __launch_bounds__(256) k_rmm_scatter(RmmArgs a, const float* __restrict__ r, const uint32_t* __restrict__ key, const uint32_t* __restrict__ rank,
                                                      const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ cursor, uint32_t* __restrict__ entry,
                                                      uint32_t* __restrict__ cell_of, const DevScalars* __restrict__ scal) {
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= min(scal->n_slots2[a.parity], a.cap)) return;
	const uint32_t k = key[s];
	if (k == KEY_NONE) return;
	const Geom& g = a.g;
	const size_t cap = a.cap;
	const uint32_t bin = k & ((1u << RMM_SHIFT_BITS) - 1u), mask = k >> RMM_SHIFT_BITS;
	uint32_t p = cell_start[bin] + rank[s];
	if (p < a.cap) { entry[p] = s << 5; cell_of[p] = bin; }
	if (mask) {
		const float x = r[s], y = r[cap + s], z = r[2 * cap + s];
		for (int c3 = 0; c3 < 3; c3++) for (int b3 = 0; b3 < 3; b3++) for (int a3 = 0; a3 < 3; a3++) {
			if ((a3 | b3 | c3) == 0 || !rmm_code_valid(mask, a3, b3, c3)) continue;
			const int hx = cell1d_tab<0>(g, a.bx, (double) rmm_halo_coord<0>(g, x, a3)), hy = cell1d_tab<1>(g, a.by, (double) rmm_halo_coord<1>(g, y, b3)),
			          hz = cell1d_tab<2>(g, a.bz, (double) rmm_halo_coord<2>(g, z, c3));
			const uint32_t hb = cell_index(g, hx, hy, hz);
			p = atomicAdd(&cursor[hb], 1u);
			if (p < a.cap) { entry[p] = (s << 5) | (uint32_t) (a3 + 3 * b3 + 9 * c3); cell_of[p] = hb; }
		}
	}
}

// deterministic in-cell order (by source slot, shift code), permutation of the 32-byte state, calc records of the force kernel
__global__ void __launch_bounds__(256) k_rmm_gather(RmmArgs a, const float* __restrict__ r, const float* __restrict__ v, const uint64_t* __restrict__ id,
                                                     float* __restrict__ r2, float* __restrict__ v2, uint64_t* __restrict__ id2, const uint32_t* __restrict__ entry,
                                                     const uint32_t* __restrict__ cell_of, const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ cell_of2,
                                                     float4* __restrict__ mpos, const DevScalars* __restrict__ scal) {
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= min(scal->n_slots2[a.parity ^ 1u], a.cap)) return;
	const Geom& g = a.g;
	const size_t cap = a.cap;
	const uint32_t e = entry[p], c = cell_of[p];
	const uint32_t b = cell_start[c], en = min(cell_start[c + 1], a.cap);
	uint32_t rk = 0;
	for (uint32_t u = b; u < en; u++) rk += (entry[u] < e) ? 1u : 0u;
	const uint32_t t = b + rk, s = e >> 5, code = e & 31u;
	float x = r[s], y = r[cap + s], z = r[2 * cap + s];
	uint64_t idv = id[s];
	if (code) {
		x = rmm_halo_coord<0>(g, x, code % 3); y = rmm_halo_coord<1>(g, y, (code / 3) % 3); z = rmm_halo_coord<2>(g, z, code / 9);
		idv |= RMM_HALO;
	}
	r2[t] = x; r2[cap + t] = y; r2[2 * cap + t] = z;
	v2[t] = v[s]; v2[cap + t] = v[cap + s]; v2[2 * cap + t] = v[2 * cap + s];
	id2[t] = idv;
	cell_of2[t] = c;
	const bool inert = (idv & RMM_HALO) || c >= g.ncells || is_halo_cell(g, c);
	mpos[t] = make_float4(x, y, z, __uint_as_float(inert ? 0x80000000u : 0u));
}

// VelocityCellProcessorRMM: sum v^2 over the owned molecules (not mass weighted), per-block partials for k_rmm_globals
__global__ void __launch_bounds__(256) k_rmm_sumv2(const float* __restrict__ v, const uint64_t* __restrict__ id, uint32_t cap, uint32_t parity, const DevScalars* __restrict__ scal,
                                                    double* __restrict__ partials) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	double acc[1] = {0.};
	if (t < min(scal->n_slots2[parity], cap) && !(id[t] & RMM_HALO)) {
		const float vx = v[t], vy = v[(size_t) cap + t], vz = v[2 * (size_t) cap + t];
		acc[0] = (double) vx * vx + (double) vy * vy + (double) vz * vz;
	}
	block_reduce_store<1>(acc, partials, blockIdx.x);
}
// Domain::calculateGlobalValues with the long-range corrections of the host (Domain.cpp:186-187, :240-252, :142-146)
__global__ void __launch_bounds__(1024) k_rmm_globals(const double* __restrict__ partials, uint32_t nblocks, DevScalars* scal, double target_T, unsigned long long N, double volume,
                                                       double upot_corr, double virial_corr, ls1b200_step_record* rec) {
	double v[1] = {0.};
	for (uint32_t b = threadIdx.x; b < nblocks; b += blockDim.x) v[0] += partials[(size_t) b * 8];
	__shared__ double out[8];
	block_reduce_store<1>(v, out, 0);
	__syncthreads();
	if (threadIdx.x != 0) return;
	const double summv2 = out[0];
	const double upot = scal->upot6lj / 6.0 + upot_corr, virial = scal->virial + virial_corr;
	const double T = summv2 / (double) (3 * N);
	const double bt = target_T > 0. ? pow(3.0 * (double) N * target_T / summv2, 0.4) : 1.;
	const double rho = (double) N / volume;
	const double p = T * rho + rho * (virial / (double) N) / 3.;
	scal->summv2 = summv2; scal->sum_iw2 = 0.; scal->beta_trans = bt; scal->beta_rot = 1.;
	if (rec) { rec->upot = upot; rec->virial = virial; rec->summv2 = summv2; rec->sum_iw2 = 0.; rec->temperature = T; rec->pressure = p; rec->beta_trans = bt; rec->beta_rot = 1.; }
}
__global__ void __launch_bounds__(256) k_rmm_scale(float* __restrict__ v, const uint64_t* __restrict__ id, uint32_t cap, uint32_t parity, const DevScalars* __restrict__ scal) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= min(scal->n_slots2[parity], cap) || (id[t] & RMM_HALO)) return;
	const float b = (float) scal->beta_trans;
	v[t] *= b; v[(size_t) cap + t] *= b; v[2 * (size_t) cap + t] *= b;
}
__global__ void __launch_bounds__(256) k_rmm_upload(const uint64_t* __restrict__ id, const double* __restrict__ r, const double* __restrict__ v, float* __restrict__ dr, float* __restrict__ dv,
                                                     uint64_t* __restrict__ did, uint64_t n, uint64_t cap) {
	const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	for (int d = 0; d < 3; d++) { dr[d * cap + t] = (float) r[3 * t + d]; dv[d * cap + t] = (float) v[3 * t + d]; }
	did[t] = id[t] & ~RMM_HALO;
}
// owned molecules, compacted (order unspecified: the ids identify them)
__global__ void __launch_bounds__(256) k_rmm_download(const float* __restrict__ r, const float* __restrict__ v, const uint64_t* __restrict__ id, uint32_t cap, uint32_t parity,
                                                       const DevScalars* __restrict__ scal, uint64_t* __restrict__ oid, double* __restrict__ orr, double* __restrict__ ov,
                                                       unsigned long long* __restrict__ counter, unsigned long long limit) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= min(scal->n_slots2[parity], cap) || (id[t] & RMM_HALO)) return;
	const unsigned long long o = atomicAdd(counter, 1ull);
	if (o >= limit) return;
	oid[o] = id[t];
	for (int d = 0; d < 3; d++) { orr[3 * o + d] = (double) r[(size_t) d * cap + t]; ov[3 * o + d] = (double) v[(size_t) d * cap + t]; }
}

int launch_bin_scan(ls1b200_ctx* c, uint32_t nb, uint32_t* fstart, uint32_t parity);      // update.cu
int force_1clj_tiled_rmm(ls1b200_ctx* c, float* v, float dt_over_m);                        // force_1clj_tiled.cu

void rmm_release(ls1b200_ctx* c) {
	for (int k = 0; k < 2; k++) {
		if (c->rmm_r[k]) cudaFree(c->rmm_r[k]); if (c->rmm_v[k]) cudaFree(c->rmm_v[k]); if (c->rmm_id[k]) cudaFree(c->rmm_id[k]);
		c->rmm_r[k] = c->rmm_v[k] = nullptr; c->rmm_id[k] = nullptr;
	}
	c->rmm = false;
}

// one binning (+ optional position update before it): the RMM analogue of do_update
static int rmm_update(ls1b200_ctx* c, float dt, int advance, int use_beta) {
	LS1_TRY(span_begin(c, PH_UPDATE));
	RmmArgs a; a.g = c->geom; a.cap = (uint32_t) c->cap; a.parity = c->parity;
	a.bx = c->d_border; a.by = c->d_border + c->border_stride; a.bz = c->d_border + 2 * (size_t) c->border_stride;
	const int cur = c->rmm_cur, nxt = cur ^ 1;
	const uint32_t nb = c->geom.ncells + 1;
	LS1_CUDA(cudaMemsetAsync(c->cell_count, 0, sizeof(uint32_t) * 2 * ((size_t) c->geom.ncells * c->fine_s + 4), c->stream));
	const uint32_t gs = (uint32_t) ((c->cap + 255) / 256);
	LS1_LAUNCH(c, PH_UPDATE, k_rmm_move_key, gs, 256, 0, a, c->rmm_r[cur], c->rmm_v[cur], (const uint64_t*) c->rmm_id[cur], dt, advance, use_beta, c->key, c->cell_count, c->cell_count_halo, c->rank, c->d_scal);
	LS1_TRY(launch_bin_scan(c, nb, c->cell_start, a.parity));
	c->chunks_ready = false;      // the chunk list of an earlier ls1b200_update does not describe this order
	LS1_LAUNCH(c, PH_UPDATE, k_rmm_scatter, gs, 256, 0, a, (const float*) c->rmm_r[cur], (const uint32_t*) c->key, (const uint32_t*) c->rank, (const uint32_t*) c->cell_start, c->cursor, c->entry, c->cell_of, c->d_scal);
	LS1_LAUNCH(c, PH_UPDATE, k_rmm_gather, gs, 256, 0, a, (const float*) c->rmm_r[cur], (const float*) c->rmm_v[cur], (const uint64_t*) c->rmm_id[cur], c->rmm_r[nxt], c->rmm_v[nxt], c->rmm_id[nxt],
	           (const uint32_t*) c->entry, (const uint32_t*) c->cell_of, (const uint32_t*) c->cell_start, c->cell_of2, (float4*) c->mpos, c->d_scal);
	c->rmm_cur = nxt; c->parity ^= 1u;
	LS1_TRY(span_end(c));
	c->sorted = true;
	return 0;
}
static int rmm_forces_globals(ls1b200_ctx* c, float dt_force, double target_T, double upot_corr, double virial_corr, ls1b200_step_record* d_rec) {
	LS1_TRY(span_begin(c, PH_FORCE));
	const double mass = c->comps[0].mass;
	LS1_TRY(force_1clj_tiled_rmm(c, c->rmm_v[c->rmm_cur], (float) ((double) dt_force / mass)));
	LS1_TRY(span_end(c));
	LS1_TRY(span_begin(c, PH_POST));
	const uint32_t g = (uint32_t) ((c->cap + 255) / 256);
	LS1_LAUNCH(c, PH_POST, k_rmm_sumv2, g, 256, 0, (const float*) c->rmm_v[c->rmm_cur], (const uint64_t*) c->rmm_id[c->rmm_cur], (uint32_t) c->cap, c->parity, c->d_scal, c->partials);
	const double V = c->geom.lbox[0] * c->geom.lbox[1] * c->geom.lbox[2];
	LS1_LAUNCH(c, PH_POST, k_rmm_globals, 1, 1024, 0, (const double*) c->partials, g, c->d_scal, target_T, (unsigned long long) c->n_owned, V, upot_corr, virial_corr, d_rec);
	LS1_TRY(span_end(c));
	return 0;
}
static int rmm_records(ls1b200_ctx* c, uint64_t n) {
	if (c->records_cap < n) {
		if (c->d_records) cudaFree(c->d_records);
		c->d_records = nullptr; c->records_cap = 0;
		LS1_CUDA(cudaMalloc((void**) &c->d_records, sizeof(ls1b200_step_record) * n));
		c->records_cap = n;
	}
	return 0;
}
static int rmm_finish(ls1b200_ctx* c) {
	LS1_CUDA(cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(DevScalars), cudaMemcpyDeviceToHost, c->stream));
	LS1_CUDA(cudaStreamSynchronize(c->stream));
	c->n_slots = std::min<uint64_t>(c->h_scal->n_slots2[c->parity], c->cap);
	if (c->h_scal->overflow) return fail(std::string("ls1b200_rmm: ") + overflow_message(c->h_scal->overflow));
	return 0;
}

}  // namespace ls1b200

using namespace ls1b200;

extern "C" {

int ls1b200_rmm_upload(ls1b200_ctx* c, uint64_t n, const uint64_t* id, const double* r, const double* v) {
	if (!c) return fail("ls1b200_rmm_upload: null context");
	if (!c->components_set || !c->all_1clj || c->comps.size() != 1) return fail("ls1b200_rmm_upload: the reduced-memory mode is the single-component 1CLJ fluid (VCP1CLJRMM)");
	if (!c->mixed || c->dd.active) return fail("ls1b200_rmm_upload: needs a mixed-precision, non-decomposed context");
	if (n && (!id || !r || !v)) return fail("ls1b200_rmm_upload: null array");
	LS1_CUDA(cudaSetDevice(c->device));
	double fac = 1.;
	for (int d = 0; d < 3; d++) if (c->geom.periodic[d]) fac *= 1. + 2. * std::min(1., c->geom.cutoff / c->geom.lbox[d]);
	uint64_t cap = n + (uint64_t) ((fac - 1.) * 1.5 * (double) n) + 4096;
	cap = (cap + 63) & ~63ull;
	if (cap >= (1ull << 27)) return fail("ls1b200_rmm_upload: more than 2^27 slots per device are not supported");
	// what the mode is about: no FP64 state, no F / Vi planes -- only the 32-byte state (double buffered), the sort's scratch and the calc records
	ctx_free_molecules(c);      // also releases an earlier RMM state
	c->cap = cap; c->n_owned = n;
	for (int k = 0; k < 2; k++) {
		LS1_CUDA(cudaMalloc((void**) &c->rmm_r[k], sizeof(float) * 3 * cap)); LS1_CUDA(cudaMalloc((void**) &c->rmm_v[k], sizeof(float) * 3 * cap));
		LS1_CUDA(cudaMalloc((void**) &c->rmm_id[k], sizeof(uint64_t) * cap));
	}
	LS1_CUDA(cudaMalloc((void**) &c->rank, 4 * cap)); LS1_CUDA(cudaMalloc((void**) &c->key, 4 * cap)); LS1_CUDA(cudaMalloc((void**) &c->entry, 4 * cap));
	LS1_CUDA(cudaMalloc((void**) &c->cell_of, 4 * cap)); LS1_CUDA(cudaMalloc((void**) &c->cell_of2, 4 * cap));
	LS1_CUDA(cudaMalloc(&c->mpos, (cap + 64) * 16));
	{
		const float far_f[4] = {1e18f, 1e18f, 1e18f, 0.f};
		for (int k = 0; k < 64; k++) LS1_CUDA(cudaMemcpyAsync((char*) c->mpos + (cap + k) * 16, far_f, 16, cudaMemcpyHostToDevice, c->stream));
	}
	c->partial_blocks = (uint32_t) std::max<uint64_t>((cap + 63) / 64 + 64, (uint64_t) c->geom.ncells + 64);
	LS1_CUDA(cudaMalloc((void**) &c->partials, sizeof(double) * 8 * (size_t) c->partial_blocks));
	if (n) {
		uint64_t* sid = nullptr; double* sr = nullptr; double* sv = nullptr;
		LS1_CUDA(cudaMalloc((void**) &sid, 8 * n)); LS1_CUDA(cudaMalloc((void**) &sr, 24 * n)); LS1_CUDA(cudaMalloc((void**) &sv, 24 * n));
		LS1_CUDA(cudaMemcpyAsync(sid, id, 8 * n, cudaMemcpyHostToDevice, c->stream));
		LS1_CUDA(cudaMemcpyAsync(sr, r, 24 * n, cudaMemcpyHostToDevice, c->stream));
		LS1_CUDA(cudaMemcpyAsync(sv, v, 24 * n, cudaMemcpyHostToDevice, c->stream));
		k_rmm_upload<<<(uint32_t) ((n + 255) / 256), 256, 0, c->stream>>>(sid, sr, sv, c->rmm_r[0], c->rmm_v[0], c->rmm_id[0], n, cap); g_launches++;
		LS1_CUDA(cudaStreamSynchronize(c->stream));
		cudaFree(sid); cudaFree(sr); cudaFree(sv);
	}
	c->rmm = true; c->rmm_cur = 0;
	c->parity = 0; c->n_slots = (uint32_t) n;
	c->h_scal->n_slots2[0] = (uint32_t) n; c->h_scal->n_slots2[1] = 0; c->h_scal->overflow = 0; c->h_scal->fault = 0.;
	c->h_scal->beta_trans = 1.; c->h_scal->beta_rot = 1.;
	c->h_scal->clip_limit = 0.; c->h_scal->expl_counter = c->h_scal->expl_warning = c->h_scal->expl_error = 0; c->globals_initialised = false;      // a new state: Domain's explosion heuristics start over
	LS1_CUDA(cudaMemcpy(c->d_scal, c->h_scal, sizeof(DevScalars), cudaMemcpyHostToDevice));
	c->sorted = false; c->ever_sorted = false; c->forces_valid = false; c->keys_valid = false;
	return 0;
}

int ls1b200_rmm_start(ls1b200_ctx* c, double dt, double target_temperature, double upot_corr, double virial_corr, ls1b200_step_record* record) {
	if (!c || !c->rmm) return fail("ls1b200_rmm_start: call ls1b200_rmm_upload first");
	LS1_CUDA(cudaSetDevice(c->device));
	if (c->n_owned == 0) return 0;
	LS1_TRY(rmm_records(c, 1));
	// Simulation::prepare_start (Simulation.cpp:847-1000): wrap + bin + halo, the initial force applied with dt / 2, global values;
	// the thermostat factor is reported but not applied
	LS1_TRY(rmm_update(c, 0.f, 0, 0));
	LS1_TRY(rmm_forces_globals(c, (float) (0.5 * dt), target_temperature, upot_corr, virial_corr, c->d_records));
	const double one = 1.;
	LS1_CUDA(cudaMemcpyAsync(&c->d_scal->beta_trans, &one, sizeof(double), cudaMemcpyHostToDevice, c->stream));
	if (record) LS1_CUDA(cudaMemcpyAsync(record, c->d_records, sizeof(ls1b200_step_record), cudaMemcpyDeviceToHost, c->stream));
	c->forces_valid = true;
	return rmm_finish(c);
}

int ls1b200_rmm_simulate(ls1b200_ctx* c, uint64_t n_steps, double dt, double target_temperature, double upot_corr, double virial_corr, ls1b200_step_record* records) {
	if (!c || !c->rmm) return fail("ls1b200_rmm_simulate: call ls1b200_rmm_upload first");
	if (!c->forces_valid) return fail("ls1b200_rmm_simulate: call ls1b200_rmm_start first (the half kick of the initial force)");
	LS1_CUDA(cudaSetDevice(c->device));
	if (n_steps == 0 || c->n_owned == 0) return 0;
	LS1_TRY(rmm_records(c, n_steps));
	for (uint64_t s = 0; s < n_steps; s++) {
		// the thermostat factor of the previous step is applied on load by the position kernel (the first step of a call: factor 1)
		LS1_TRY(rmm_update(c, (float) dt, 1, target_temperature > 0.));
		LS1_TRY(rmm_forces_globals(c, (float) dt, target_temperature, upot_corr, virial_corr, c->d_records + s));
	}
	if (target_temperature > 0.) {      // the scaling the next position update would have applied; then the pending factor is 1
		const uint32_t g = (uint32_t) ((c->cap + 255) / 256);
		LS1_LAUNCH(c, PH_POST, k_rmm_scale, g, 256, 0, c->rmm_v[c->rmm_cur], (const uint64_t*) c->rmm_id[c->rmm_cur], (uint32_t) c->cap, c->parity, c->d_scal);
		const double one = 1.;
		LS1_CUDA(cudaMemcpyAsync(&c->d_scal->beta_trans, &one, sizeof(double), cudaMemcpyHostToDevice, c->stream));
	}
	if (records) LS1_CUDA(cudaMemcpyAsync(records, c->d_records, sizeof(ls1b200_step_record) * n_steps, cudaMemcpyDeviceToHost, c->stream));
	return rmm_finish(c);
}

int ls1b200_rmm_download(ls1b200_ctx* c, uint64_t* id, double* r, double* v, uint64_t n) {
	if (!c || !c->rmm) return fail("ls1b200_rmm_download: call ls1b200_rmm_upload first");
	if (!id || !r || !v) return fail("ls1b200_rmm_download: null array");
	if (n != c->n_owned) return fail("ls1b200_rmm_download: n does not match the number of molecules");
	LS1_CUDA(cudaSetDevice(c->device));
	if (n == 0) return 0;
	uint64_t* did = nullptr; double* dr = nullptr; double* dv = nullptr; unsigned long long* cnt = nullptr;
	LS1_CUDA(cudaMalloc((void**) &did, 8 * n)); LS1_CUDA(cudaMalloc((void**) &dr, 24 * n)); LS1_CUDA(cudaMalloc((void**) &dv, 24 * n)); LS1_CUDA(cudaMalloc((void**) &cnt, 8));
	LS1_CUDA(cudaMemsetAsync(cnt, 0, 8, c->stream));
	k_rmm_download<<<(uint32_t) ((c->cap + 255) / 256), 256, 0, c->stream>>>((const float*) c->rmm_r[c->rmm_cur], (const float*) c->rmm_v[c->rmm_cur], (const uint64_t*) c->rmm_id[c->rmm_cur],
	                                                                         (uint32_t) c->cap, c->parity, c->d_scal, did, dr, dv, cnt, (unsigned long long) n);
	g_launches++;
	unsigned long long got = 0;
	LS1_CUDA(cudaMemcpyAsync(id, did, 8 * n, cudaMemcpyDeviceToHost, c->stream));
	LS1_CUDA(cudaMemcpyAsync(r, dr, 24 * n, cudaMemcpyDeviceToHost, c->stream));
	LS1_CUDA(cudaMemcpyAsync(v, dv, 24 * n, cudaMemcpyDeviceToHost, c->stream));
	LS1_CUDA(cudaMemcpyAsync(&got, cnt, 8, cudaMemcpyDeviceToHost, c->stream));
	LS1_CUDA(cudaStreamSynchronize(c->stream));
	cudaFree(did); cudaFree(dr); cudaFree(dv); cudaFree(cnt);
	if (got != n) return fail("ls1b200_rmm_download: the number of owned molecules on the device differs from the upload (molecules left a non-periodic box?)");
	return 0;
}

uint64_t ls1b200_rmm_state_bytes(const ls1b200_ctx* c) { return (c && c->rmm) ? 32ull * c->n_owned : 0; }

}  // extern "C"
