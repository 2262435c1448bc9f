// update.cu -- ls1b200_update(): LinkedCells::update() + DomainDecompBase::exchangeMolecules() (periodic, one rank) +
// LinkedCells::updateMoleculeCaches() as a deterministic counting sort on the device.
//
// Reference: Simulation::updateParticleContainerAndDecomposition (Simulation.cpp:1417-1455);
//   LinkedCells::update_via_copies (particleContainer/LinkedCells.cpp:303-355) -- which cell a molecule belongs to;
//   DomainDecompBase::handleDomainLeavingParticles / populateHaloLayerWithCopies (parallel/DomainDecompBase.cpp:199-249,
//   :317-376) -- wrap + periodic halo copies;
//   FullParticleCell::buildSoACaches / FullMolecule::setupSoACache (FullParticleCell.cpp:141-191, FullMolecule.cpp:711-767)
//   -- calc-precision SoA (site expansion).
//
// Kernels (all HBM-bound; algorithmic bytes per molecule in DESIGN.md):
//   k_key      wrap, cell index (bit exact), histogram of owned molecules + periodic copies
//   k_scan_*   exclusive scan of the histogram -> cell_start
//   k_scatter  provisional slot per molecule / copy (atomic cursor, order not yet deterministic)
//   k_gather   deterministic rank inside the cell (by source slot, shift code), permutes the FP64 state cur -> nxt,
//              writes the calc-precision molecule record mpos
//   k_sites    quaternion -> site positions / orientations in calc precision (multi-site components only)
#include "geom.cuh"
#include "types.cuh"

namespace ls1b200 {

constexpr int SHIFT_BITS = 26;   // key = bin | shiftmask << 26  (bin <= ncells < 2^26)

constexpr uint32_t RANK_BITS = 28, RANK_MASK = (1u << RANK_BITS) - 1u;      // rank[] = in-bin rank | fine x-bin << 28

struct UpdArgs {
	Geom g;
	uint32_t cap;
	int cut[3];          // decomposed contexts: the block has a neighbour rank in that dimension (k_gather flags the boundary slots)
	uint32_t S, logS;    // fine x-bins per reference cell (1, 2, 4 or 8 = 1 << logS); histogram bin = cell * S + fine_sub (limbo: ncells * S)
	double finv;         // S / cell length in x
	uint32_t parity;     // DevScalars::n_slots2[parity] = source slots of this update, [parity^1] receives the new count
	const double* bx; const double* by; const double* bz;      // cell border tables (context.cu)
	// Overlapped multi-GPU step (the exchange runs beside this sort and converts leaving molecules in place): the gather
	// re-applies the conversion, which is idempotent, to the slots it finds in a halo cell of a cut dimension -- the result does not
	// depend on whether it read the slot before or after k_export touched it (halo.cu: k_export, LEAVING / HALO shift + clamp)
	int fix_leavers;
	int lo_wrap[3], hi_wrap[3];
	double glen[3], gmin[3], gmax[3], wrap_hi[3], halo_hi[3], halo_lo[3];
};

__device__ __forceinline__ uint32_t g_ncells(const UpdArgs& a) { return a.g.ncells; }

// Cells (and fine x-bins) of the periodic copies of one molecule: per dimension the cell coordinate of the +L copy (index 1) and of the
// -L copy (index 2) is evaluated ONCE (bit-exact binning of the shifted coordinate, three FP64 compares each) -- a copy with shift
// code (a3, b3, c3) lies in cell (hx[a3], hy[b3], hz[c3]); index 0 is the molecule's own cell coordinate.
struct CopyCells { int hx[3], hy[3], hz[3]; uint32_t fsub[3]; };
__device__ __forceinline__ void copy_cells(const UpdArgs& a, uint32_t mask, int cx, int cy, int cz, uint32_t sub, double x, double y, double z, CopyCells& h) {
	const Geom& g = a.g;
	h.hx[0] = cx; h.hy[0] = cy; h.hz[0] = cz; h.fsub[0] = sub;
	#pragma unroll
	for (int k = 1; k < 3; k++) {
		h.hx[k] = h.hy[k] = h.hz[k] = 0; h.fsub[k] = 0;
		if ((mask >> (k - 1)) & 1u) {
			const double xh = halo_coord<0>(g, x, k);
			h.hx[k] = cell1d_tab<0>(g, a.bx, xh);
			if (a.S > 1) h.fsub[k] = (uint32_t) fine_sub(xh, __ldg(a.bx + h.hx[k]), a.finv, (int) a.S);
		}
		if ((mask >> (2 + k - 1)) & 1u) h.hy[k] = cell1d_tab<1>(g, a.by, halo_coord<1>(g, y, k));
		if ((mask >> (4 + k - 1)) & 1u) h.hz[k] = cell1d_tab<2>(g, a.bz, halo_coord<2>(g, z, k));
	}
}
__device__ __forceinline__ bool code_valid(uint32_t mask, int a, int b, int c) {
	return (a == 0 || ((mask >> (a - 1)) & 1u)) && (b == 0 || ((mask >> (2 + b - 1)) & 1u)) && (c == 0 || ((mask >> (4 + c - 1)) & 1u));
}

// Binning of one slot (shared by k_key and by the fused integrate + key kernel k_pre_key): wrap, cell index (bit exact), key,
// histogram of owned molecules + periodic copies. `live`: the slot takes part in the binning (x, y, z hold its position);
// every thread of the warp must call this (shuffles).
__device__ __forceinline__ void key_slot(const UpdArgs& a, State& cur, uint32_t s, bool in_range, bool live, double x, double y, double z,
                                         uint32_t* __restrict__ key, uint32_t* __restrict__ count, uint32_t* __restrict__ count_halo, uint32_t* __restrict__ rank) {
	const Geom& g = a.g;
	uint32_t bin = KEY_NONE, mask = 0, sub = 0;      // bin: FINE bin (reference cell * S + x sub-bin)
	int mcx = 0, mcy = 0, mcz = 0;
	if (in_range && !live) key[s] = KEY_NONE;         // stale halo slot; fresh halo copies (halo.cu) are binned
	if (live) {
		if (g.periodic[0]) { double w = wrap_coord<0>(g, x); if (w != x) { x = w; cur.r[s] = w; } }
		if (g.periodic[1]) { double w = wrap_coord<1>(g, y); if (w != y) { y = w; cur.r[(size_t) a.cap + s] = w; } }
		if (g.periodic[2]) { double w = wrap_coord<2>(g, z); if (w != z) { z = w; cur.r[2 * (size_t) a.cap + s] = w; } }
		const int cx = cell1d_tab<0>(g, a.bx, x), cy = cell1d_tab<1>(g, a.by, y), cz = cell1d_tab<2>(g, a.bz, z);
		const bool out = cx < 0 || cy < 0 || cz < 0 || cx >= g.cells[0] || cy >= g.cells[1] || cz >= g.cells[2];
		if (out) { key[s] = g.ncells; bin = g.ncells * a.S; }
		else {
			const uint32_t cell = cell_index(g, cx, cy, cz);
			mcx = cx; mcy = cy; mcz = cz;
			mask = shift_mask<0>(g, x) | (shift_mask<1>(g, y) << 2) | (shift_mask<2>(g, z) << 4);
			key[s] = cell | (mask << SHIFT_BITS);
			if (a.S > 1) sub = (uint32_t) fine_sub(x, __ldg(a.bx + cx), a.finv, (int) a.S);      // S == 1: the reference cell is the bin
			bin = (cell << a.logS) + sub;
		}
	}
	// Histogram: slots are in the cell order of the previous step, so the lanes of a warp fall into a handful of bins -- one
	// atomic per distinct bin and warp instead of one per molecule. The value the atomic returns is the provisional rank of the
	// molecule inside its bin: k_scatter needs no second round of atomics for it.
	{
		// runs of equal bins among consecutive lanes (a shuffle, a ballot and bit arithmetic; __match_any_sync is far more expensive)
		const int lane = threadIdx.x & 31;
		const uint32_t prev = __shfl_up_sync(0xffffffffu, bin, 1);
		const bool head = lane == 0 || prev != bin;
		const uint32_t heads = __ballot_sync(0xffffffffu, head);
		const uint32_t upto = 0xFFFFFFFFu >> (31 - lane);                 // lanes <= mine
		const int start = 31 - __clz(heads & upto);                        // head of my run
		const uint32_t above = heads & ~upto;
		const int end = above ? __ffs(above) - 1 : 32;                     // one past its last lane
		uint32_t base = 0;
		if (head && bin != KEY_NONE) base = atomicAdd(&count[bin], (uint32_t) (end - start));
		base = __shfl_sync(0xffffffffu, base, start);
		if (bin != KEY_NONE) rank[s] = ((base + (uint32_t) (lane - start)) & RANK_MASK) | (sub << RANK_BITS);
	}
	if (mask) {
		CopyCells h;
		copy_cells(a, mask, mcx, mcy, mcz, sub, x, y, z, h);
		#pragma unroll
		for (int c3 = 0; c3 < 3; c3++)
		#pragma unroll
		for (int b3 = 0; b3 < 3; b3++)
		#pragma unroll
		for (int a3 = 0; a3 < 3; a3++) {
			if ((a3 | b3 | c3) == 0 || !code_valid(mask, a3, b3, c3)) continue;
			atomicAdd(&count_halo[(cell_index(g, h.hx[a3], h.hy[b3], h.hz[c3]) << a.logS) + h.fsub[a3]], 1u);
		}
	}
}

__global__ void __launch_bounds__(256) k_key(UpdArgs a, State cur, uint32_t* __restrict__ key, uint32_t* __restrict__ count, uint32_t* __restrict__ count_halo,
                                              uint32_t* __restrict__ rank, const DevScalars* __restrict__ scal, Region rg) {
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	// (overlapped multi-GPU step: two launches, inner slots before the imports have arrived, the rest after -- every slot in exactly one;
	//  fused step: mode 3, only the imports appended behind the sorted slots -- the others were binned by k_pre_key)
	const bool in_range = s < min(scal->n_slots2[a.parity], a.cap) && region_selected(rg, s, scal->n_sorted);      // the region test reads 4 bytes per slot, before anything else
	bool live = false;
	double x = 0., y = 0., z = 0.;
	if (in_range) {
		const uint64_t meta = cur.meta[s];
		live = !(meta_halo(meta) && !meta_fresh(meta));
		if (live) { x = cur.r[s]; y = cur.r[(size_t) a.cap + s]; z = cur.r[2 * (size_t) a.cap + s]; }
	}
	key_slot(a, cur, s, in_range, live, x, y, z, key, count, count_halo, rank);
}

// fused step, decomposed contexts: the imports the exchange appended behind the sorted slots (a few per cent of the slots) -- a small
// grid that starts at n_pre (the slots k_pre_key binned) instead of a launch over the whole capacity
__global__ void __launch_bounds__(256) k_key_tail(UpdArgs a, State cur, uint32_t* __restrict__ key, uint32_t* __restrict__ count, uint32_t* __restrict__ count_halo,
                                                   uint32_t* __restrict__ rank, const DevScalars* __restrict__ scal) {
	const uint32_t n = min(scal->n_slots2[a.parity], a.cap), first = min(scal->n_pre, n);
	for (uint32_t base = first + blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {      // uniform per block: every warp calls key_slot together
		const uint32_t s = base + threadIdx.x;
		const bool in_range = s < n;
		bool live = false;
		double x = 0., y = 0., z = 0.;
		if (in_range) {
			const uint64_t meta = cur.meta[s];
			live = !(meta_halo(meta) && !meta_fresh(meta));
			if (live) { x = cur.r[s]; y = cur.r[(size_t) a.cap + s]; z = cur.r[2 * (size_t) a.cap + s]; }
		}
		key_slot(a, cur, s, in_range, live, x, y, z, key, count, count_halo, rank);
	}
}

// k_pre_key's integration of one slot while the selective thermostat of Domain's explosion heuristics is pending (k_globals set
// DevScalars::clip_limit): v and F of a molecule above the limit are scaled to it before the thermostat factor and the next half kick
// (Domain.cpp:268-281). Returns the new position.
__device__ __noinline__ double3 pre_key_clipped_slot(State cur, const double* __restrict__ vs, const double* __restrict__ F, size_t c, uint32_t s, size_t so, double a0, double mass,
                                                     double bt, double dt, int post_kick, double limit, double rx, double ry, double rz) {
	double w[3], f[3], r[3] = {rx, ry, rz};
	#pragma unroll
	for (int d = 0; d < 3; d++) { w[d] = vs[d * c + so]; f[d] = F[d * c + s]; if (post_kick) w[d] += a0 * f[d]; }
	const double vc = clip_factor_trans(limit, mass, w[0], w[1], w[2]);
	#pragma unroll
	for (int d = 0; d < 3; d++) {
		double v = w[d] * vc;
		v *= bt;
		v += a0 * (f[d] * vc);
		cur.v[d * c + s] = v;
		r[d] += dt * v;
		cur.r[d * c + s] = r[d];
	}
	return make_double3(r[0], r[1], r[2]);
}

// The fused first kernel of a time step of ls1b200_simulate (ensembles without rotation): [the second half kick of the PREVIOUS
// step (Leapfrog::transition2to3 -> upd_postF, whose sum of m v^2 the force kernel already delivered) +] thermostat factor on
// load + upd_preF (half kick, drift) + the binning of k_key on the new position, which is still in registers: the slot state
// is read and written once, the FP64 compares of the binning hide behind the memory traffic of the integration.
// Same floating-point sequence per molecule as k_leapfrog_post -> k_scale -> k_leapfrog_pre -> k_key.
// `perm` (lazy permutation): the previous sort did not move r and v -- slot s of the current order still has them at perm[s] of the
// previous order (`old`); this kernel, which reads and writes them anyway, puts them in place.
__global__ void __launch_bounds__(256, 8) k_pre_key(UpdArgs a, State cur, const double* __restrict__ F, const CompDev* __restrict__ comps, double dt, int use_beta, int post_kick,
                                                  uint32_t* __restrict__ key, uint32_t* __restrict__ count, uint32_t* __restrict__ count_halo, uint32_t* __restrict__ rank,
                                                  DevScalars* __restrict__ scal, const uint32_t* __restrict__ perm, State old) {
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	const bool in_range = s < min(scal->n_slots2[a.parity], a.cap);
	if (s == 0) scal->n_pre = min(scal->n_slots2[a.parity], a.cap);      // the slots this kernel bins; what the exchange appends later lies behind
	bool live = false;
	double r[3] = {0., 0., 0.};
	if (in_range) {
		const uint64_t meta = cur.meta[s];
		const size_t c = a.cap;
		if (!meta_halo(meta)) {
			live = true;
			const double* rs = perm ? old.r : cur.r; const double* vs = perm ? old.v : cur.v;
			const size_t so = perm ? perm[s] : s;
			#pragma unroll
			for (int d = 0; d < 3; d++) r[d] = rs[d * c + so];
			if (!scal->overflow) {      // a failed step: the state stays as it is (k_leapfrog_pre)
				const double bt = use_beta ? scal->beta_trans : 1.;
				const double a0 = (.5 * dt) / comps[meta_cid(meta)].mass;
				if (use_beta && scal->clip_limit > 0.) {
					// pending selective thermostat of Domain's explosion heuristics (k_globals). Rare: a function of its own (not inlined),
					// so that the streaming loop below keeps its registers.
					const double3 rn = pre_key_clipped_slot(cur, vs, F, c, s, so, a0, comps[meta_cid(meta)].mass, bt, dt, post_kick, scal->clip_limit, r[0], r[1], r[2]);
					r[0] = rn.x; r[1] = rn.y; r[2] = rn.z;
				} else {
				#pragma unroll
				for (int d = 0; d < 3; d++) {
					double v = vs[d * c + so];
					const double f = F[d * c + s];
					if (post_kick) v += a0 * f;
					if (use_beta) v *= bt;
					v += a0 * f;
					cur.v[d * c + s] = v;
					r[d] += dt * v;
					cur.r[d * c + s] = r[d];
				}
				}
			} else if (perm) {
				#pragma unroll
				for (int d = 0; d < 3; d++) { cur.v[d * c + s] = vs[d * c + so]; cur.r[d * c + s] = r[d]; }
			}
		} else if (meta_fresh(meta)) {
			live = true;
			#pragma unroll
			for (int d = 0; d < 3; d++) r[d] = cur.r[d * c + s];
		}
	}
	key_slot(a, cur, s, in_range, live, r[0], r[1], r[2], key, count, count_halo, rank);
}

// ---- exclusive scan (three small kernels; n = ncells + 1 bins) ----------------------------------------------------
constexpr int SCAN_T = 256, SCAN_V = 8, SCAN_B = SCAN_T * SCAN_V;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* sh /*[32]*/) {
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t inc = v;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
	if (lane == 31) sh[w] = inc;
	__syncthreads();
	if (w == 0) {
		uint32_t x = (lane < (blockDim.x >> 5)) ? sh[lane] : 0u, xi = x;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += t; }
		sh[lane] = xi - x;
		if (lane == 31) *total = xi;
	}
	__syncthreads();
	uint32_t r = sh[w] + inc - v;
	__syncthreads();
	return r;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_block(const uint32_t* __restrict__ in, const uint32_t* __restrict__ in2, uint32_t* __restrict__ out, uint32_t* __restrict__ bsum, uint32_t n) {
	__shared__ uint32_t sh[32]; __shared__ uint32_t tot;
	const uint32_t base = blockIdx.x * SCAN_B + threadIdx.x * SCAN_V;
	uint32_t v[SCAN_V], sum = 0;
	#pragma unroll
	for (int k = 0; k < SCAN_V; k++) { v[k] = (base + k < n) ? in[base + k] + in2[base + k] : 0u; sum += v[k]; }
	uint32_t off = block_exclusive_scan(sum, &tot, sh);
	#pragma unroll
	for (int k = 0; k < SCAN_V; k++) { if (base + k < n) out[base + k] = off; off += v[k]; }
	if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(SCAN_T) k_scan_top(uint32_t* __restrict__ bsum, uint32_t nb, uint32_t* __restrict__ grand_total) {
	__shared__ uint32_t sh[32]; __shared__ uint32_t tot;
	uint32_t carry = 0;
	for (uint32_t b0 = 0; b0 < nb; b0 += SCAN_T) {
		const uint32_t i = b0 + threadIdx.x;
		const uint32_t v = (i < nb) ? bsum[i] : 0u;
		const uint32_t off = block_exclusive_scan(v, &tot, sh);
		if (i < nb) bsum[i] = carry + off;
		carry += tot;
		__syncthreads();
	}
	if (threadIdx.x == 0) *grand_total = carry;
}
// adds the block offsets; writes cell_start[0..n] (cell_start[n] = total), the scatter cursor, and the slot count
// (n = ncells * S + 1 FINE bins; `coarse` = the reference cells' cell_start: entry c is the start of fine bin c * S)
__global__ void __launch_bounds__(SCAN_T) k_scan_add(uint32_t* __restrict__ start, uint32_t* __restrict__ cursor, const uint32_t* __restrict__ count, const uint32_t* __restrict__ bsum,
                                                      const uint32_t* __restrict__ grand_total, uint32_t n, uint32_t cap, uint32_t parity, DevScalars* scal,
                                                      uint32_t* __restrict__ coarse, uint32_t S, int tail) {
	const uint32_t base = blockIdx.x * SCAN_B + threadIdx.x * SCAN_V;
	// tail: the bins of the imports of the overlapped step; their slots follow the n_sorted slots of the main sort
	const uint32_t off = bsum[blockIdx.x] + (tail ? scal->n_sorted : 0u);
	#pragma unroll
	for (int k = 0; k < SCAN_V; k++) if (base + k < n) {
		uint32_t v = start[base + k] + off; start[base + k] = v; cursor[base + k] = v + count[base + k];      // halo copies follow the ranked entries
		if (coarse != start && ((base + k) & (S - 1u)) == 0) coarse[(base + k) / S] = v;
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		const uint32_t total = *grand_total + (tail ? scal->n_sorted : 0u);
		start[n] = total;
		if (coarse != start) coarse[(n - 1) / S + 1] = total;
		scal->n_slots2[parity ^ 1u] = total;
		if (!tail) scal->n_sorted = min(total, cap);
		if (total > cap) scal->overflow = 1;
	}
}

__device__ __forceinline__ void scatter_slot(const UpdArgs& a, const State& cur, uint32_t s, const uint32_t* __restrict__ key, const uint32_t* __restrict__ rank, const uint32_t* __restrict__ cell_start,
                                             uint32_t* __restrict__ cursor, uint32_t* __restrict__ entry, uint32_t* __restrict__ cell_of, uint64_t* __restrict__ ekey) {
	const uint32_t k = key[s];
	if (k == KEY_NONE) return;
	const Geom& g = a.g;
	const uint32_t cell = k & ((1u << SHIFT_BITS) - 1u), mask = k >> SHIFT_BITS;
	const uint32_t rk = rank[s];
	const uint32_t bin = cell >= g_ncells(a) ? (g_ncells(a) << a.logS) : (cell << a.logS) + (rk >> RANK_BITS);      // fine bin (cell_start is the FINE array here)
	// provisional slot: start of the bin + the rank k_key obtained from the histogram atomic (the in-bin order is made
	// deterministic by k_gather)
	uint32_t p = cell_start[bin] + (rk & RANK_MASK);
	const uint64_t idk = ekey ? (cur.id[s] << 5) : 0ull;      // decomposed mode: in-cell order by (molecule id, shift code)
	if (p < a.cap) { entry[p] = s << 5; cell_of[p] = bin; if (ekey) ekey[p] = idk; }
	if (mask) {
		const double x = cur.r[s], y = cur.r[(size_t) a.cap + s], z = cur.r[2 * (size_t) a.cap + s];
		CopyCells h;
		copy_cells(a, mask, (int) (cell % (uint32_t) g.cells[0]), (int) ((cell / (uint32_t) g.cells[0]) % (uint32_t) g.cells[1]), (int) (cell / (uint32_t) (g.cells[0] * g.cells[1])), rk >> RANK_BITS, x, y, z, h);
		#pragma unroll
		for (int c3 = 0; c3 < 3; c3++)
		#pragma unroll
		for (int b3 = 0; b3 < 3; b3++)
		#pragma unroll
		for (int a3 = 0; a3 < 3; a3++) {
			if ((a3 | b3 | c3) == 0 || !code_valid(mask, a3, b3, c3)) continue;
			const uint32_t hb = (cell_index(g, h.hx[a3], h.hy[b3], h.hz[c3]) << a.logS) + h.fsub[a3];
			p = atomicAdd(&cursor[hb], 1u);      // the periodic copies of a bin follow its ranked entries
			if (p < a.cap) { entry[p] = (s << 5) | (uint32_t) (a3 + 3 * b3 + 9 * c3); cell_of[p] = hb; if (ekey) ekey[p] = idk | (uint64_t) (a3 + 3 * b3 + 9 * c3); }
		}
	}
}

// seg 0: every slot; 1: the slots k_pre_key binned (overlapped multi-GPU step: the exchange appends imports behind them meanwhile)
__global__ void __launch_bounds__(256) k_scatter(UpdArgs a, State cur, const uint32_t* __restrict__ key, const uint32_t* __restrict__ rank, const uint32_t* __restrict__ cell_start,
                                                  uint32_t* __restrict__ cursor, uint32_t* __restrict__ entry, uint32_t* __restrict__ cell_of, const DevScalars* __restrict__ scal,
                                                  uint64_t* __restrict__ ekey, int seg) {
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= min(seg == 1 ? scal->n_pre : scal->n_slots2[a.parity], a.cap)) return;
	scatter_slot(a, cur, s, key, rank, cell_start, cursor, entry, cell_of, ekey);
}
// the imports of the overlapped step, [n_pre, n_slots): their own histogram / cell_start (absolute slots behind n_sorted), small grid
__global__ void __launch_bounds__(256) k_scatter_tail(UpdArgs a, State cur, const uint32_t* __restrict__ key, const uint32_t* __restrict__ rank, const uint32_t* __restrict__ cell_start,
                                                       uint32_t* __restrict__ cursor, uint32_t* __restrict__ entry, uint32_t* __restrict__ cell_of, const DevScalars* __restrict__ scal,
                                                       uint64_t* __restrict__ ekey) {
	const uint32_t n = min(scal->n_slots2[a.parity], a.cap);
	for (uint32_t s = min(scal->n_pre, n) + blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x)
		scatter_slot(a, cur, s, key, rank, cell_start, cursor, entry, cell_of, ekey);
}

template <class C>
__device__ __forceinline__ void gather_slot(const UpdArgs& a, const State& cur, const State& nxt, uint32_t p, const uint32_t* __restrict__ entry, const uint32_t* __restrict__ cell_of,
                                            const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ cell_of2,
                                            typename Vec4<C>::type* __restrict__ mpos, int rotates,
                                            double* __restrict__ F, double* __restrict__ Vi, double* __restrict__ M, const uint64_t* __restrict__ ekey,
                                            uint8_t* __restrict__ region_flag, uint32_t* __restrict__ perm) {
	const Geom& g = a.g;
	const uint32_t e = entry[p], fb = cell_of[p];              // fine bin; cell_start is the FINE array
	const uint32_t c = min(fb >> a.logS, g.ncells);            // reference cell (limbo: ncells)
	const uint32_t b = cell_start[fb], en = min(cell_start[fb + 1], a.cap);
	uint32_t rank = 0;
	if (ekey) { const uint64_t ek = ekey[p]; for (uint32_t u = b; u < en; u++) rank += (ekey[u] < ek) ? 1u : 0u; }
	else for (uint32_t u = b; u < en; u++) rank += (entry[u] < e) ? 1u : 0u;
	const uint32_t t = b + rank, s = e >> 5, code = e & 31u;
	const size_t cap = a.cap;
	double x = cur.r[s], y = cur.r[cap + s], z = cur.r[2 * cap + s];
	uint64_t meta = cur.meta[s];
	if (code) {
		const int a3 = code % 3, b3 = (code / 3) % 3, c3 = code / 9;
		x = halo_coord<0>(g, x, a3); y = halo_coord<1>(g, y, b3); z = halo_coord<2>(g, z, c3);
		meta |= META_HALO;
	}
	if (a.fix_leavers && c < g.ncells) {
		// A molecule this rank owned that now lies in the halo layer of a face shared with another rank has left: it stays behind as
		// the halo copy of the molecule its new owner holds. k_export does this conversion in place while this kernel runs
		// (overlapped step); applied again here it gives the same value whichever state this thread read.
		const int cc[3] = {(int) (c % (uint32_t) g.cells[0]), (int) ((c / (uint32_t) g.cells[0]) % (uint32_t) g.cells[1]), (int) (c / (uint32_t) (g.cells[0] * g.cells[1]))};
		double* rr[3] = {&x, &y, &z};
		#pragma unroll
		for (int d = 0; d < 3; d++) {
			if (!a.cut[d]) continue;
			double& v = *rr[d];
			if (cc[d] == 0 && v < g.bmin[d]) {
				meta |= META_HALO;
				if (a.lo_wrap[d]) { double xl = __dadd_rn(v, a.glen[d]); if (xl >= a.gmax[d]) xl = a.wrap_hi[d]; double xh = __dsub_rn(xl, a.glen[d]); if (xh >= a.gmin[d]) xh = a.halo_lo[d]; v = xh; }
			} else if (cc[d] == g.cells[d] - 1 && v >= g.bmax[d]) {
				meta |= META_HALO;
				if (a.hi_wrap[d]) { double xl = __dsub_rn(v, a.glen[d]); if (xl < a.gmin[d]) xl = a.gmin[d]; double xh = __dadd_rn(xl, a.glen[d]); if (xh < a.gmax[d]) xh = a.halo_hi[d]; v = xh; }
			}
		}
	}
	meta &= ~META_FRESH;      // a fresh halo copy of this step's exchange is an ordinary (stale after this sort) halo slot from here on
	if (ekey) nxt.id[t] = cur.id[s];
	nxt.meta[t] = meta;
	if (perm) perm[t] = s;      // lazy permutation (fused step, owned slots are not halo copies of anything): r and v move in the next k_pre_key
	else {
	nxt.r[t] = x; nxt.r[cap + t] = y; nxt.r[2 * cap + t] = z;
	#pragma unroll
	for (int d = 0; d < 3; d++) nxt.v[d * cap + t] = cur.v[d * cap + s];
	}
	if (rotates) {
		#pragma unroll
		for (int d = 0; d < 4; d++) nxt.q[d * cap + t] = cur.q[d * cap + s];
		#pragma unroll
		for (int d = 0; d < 3; d++) nxt.L[d * cap + t] = cur.L[d * cap + s];
	}
	cell_of2[t] = c;
	if (region_flag) region_flag[t] = cell_is_boundary(c, g.ncells, g.cells[0], g.cells[1], g.cells[2], a.cut) ? 1 : 0;
	// the limbo bin and halo cells never hold "i" molecules: flag them like halo copies in the calc record
	const bool inert = meta_halo(meta) || c >= g.ncells || is_halo_cell(g, c);
	mpos[t] = make_mrec<C>(x, y, z, meta_cid(meta) | (inert ? 0x80000000u : 0u));
	if (inert && !meta_halo(meta)) {   // an owned molecule no force kernel visits (halo-cell resident, limbo): its force is zero
		for (int d = 0; d < 3; d++) { F[d * cap + t] = 0.; Vi[d * cap + t] = 0.; if (M) M[d * cap + t] = 0.; }
	}
}

template <class C>
__global__ void __launch_bounds__(256) k_gather(UpdArgs a, State cur, State nxt, const uint32_t* __restrict__ entry, const uint32_t* __restrict__ cell_of,
                                                 const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ cell_of2,
                                                 typename Vec4<C>::type* __restrict__ mpos, const DevScalars* __restrict__ scal, int rotates,
                                                 double* __restrict__ F, double* __restrict__ Vi, double* __restrict__ M, const uint64_t* __restrict__ ekey,
                                                 uint8_t* __restrict__ region_flag, int seg, uint32_t* __restrict__ perm) {
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	// seg 1 (overlapped multi-GPU step): the slots of THIS sort only -- the imports' sort, on the exchange stream, may already have added its own to n_slots2
	if (p >= min(seg == 1 ? scal->n_sorted : scal->n_slots2[a.parity ^ 1u], a.cap)) return;
	gather_slot<C>(a, cur, nxt, p, entry, cell_of, cell_start, cell_of2, mpos, rotates, F, Vi, M, ekey, region_flag, perm);
}
// the imports of the overlapped step: destination slots [n_sorted, n_slots2[parity ^ 1]) (cell_start: the imports' own, absolute)
template <class C>
__global__ void __launch_bounds__(256) k_gather_tail(UpdArgs a, State cur, State nxt, const uint32_t* __restrict__ entry, const uint32_t* __restrict__ cell_of,
                                                      const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ cell_of2,
                                                      typename Vec4<C>::type* __restrict__ mpos, const DevScalars* __restrict__ scal, int rotates,
                                                      double* __restrict__ F, double* __restrict__ Vi, double* __restrict__ M, const uint64_t* __restrict__ ekey) {
	const uint32_t n = min(scal->n_slots2[a.parity ^ 1u], a.cap);
	for (uint32_t p = min(scal->n_sorted, n) + blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x)
		gather_slot<C>(a, cur, nxt, p, entry, cell_of, cell_start, cell_of2, mpos, rotates, F, Vi, M, ekey, nullptr, nullptr);
}

// ---- site expansion (FullMolecule::setupSoACache, molecules/FullMolecule.cpp:711-767; Quaternion::rotate, Quaternion.cpp:43-64)
__device__ __forceinline__ void q_rotate(const double* q, const double* d, double* o) {
	if (d[0] == 0. && d[1] == 0. && d[2] == 0.) { o[0] = o[1] = o[2] = 0.; return; }
	const double w = q[0], x = q[1], y = q[2], z = q[3];
	const double ww = w * w, xx = x * x, yy = y * y, zz = z * z, wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
	o[0] = (ww + xx - yy - zz) * d[0] + 2. * (xy - wz) * d[1] + 2. * (wy + xz) * d[2];
	o[1] = 2. * (wz + xy) * d[0] + (ww - xx + yy - zz) * d[1] + 2. * (yz - wx) * d[2];
	o[2] = 2. * (xz - wy) * d[0] + 2. * (wx + yz) * d[1] + (ww - xx - yy + zz) * d[2];
}

template <class C>
__global__ void __launch_bounds__(128) k_sites(State st, uint32_t cap, uint32_t smax, const CompDev* __restrict__ comps, const SiteDev* __restrict__ sites,
                                                typename Vec4<C>::type* __restrict__ spos, typename Vec4<C>::type* __restrict__ sori,
                                                const DevScalars* __restrict__ scal, uint32_t parity, int rotates) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= min(scal->n_slots2[parity], cap)) return;
	const uint64_t meta = st.meta[t];
	const CompDev& cp = comps[meta_cid(meta)];
	const double r[3] = {st.r[t], st.r[(size_t) cap + t], st.r[2 * (size_t) cap + t]};
	double qn[4] = {1., 0., 0., 0.};
	if (rotates) {
		for (int k = 0; k < 4; k++) qn[k] = st.q[(size_t) k * cap + t];
		const double sc = 1. / sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
		for (int k = 0; k < 4; k++) qn[k] *= sc;
	}
	for (uint32_t k = 0; k < cp.nsites; k++) {
		const SiteDev& sd = sites[cp.site_offset + k];
		double d3[3], e3[3] = {0., 0., 0.};
		q_rotate(qn, sd.r, d3);
		if (sd.type >= 2) q_rotate(qn, sd.e, e3);
		C w = (C) sd.param;
		if (sd.type == 0) {
			if (sizeof(C) == 4) w = (C) __uint_as_float(sd.ljid); else w = (C) __longlong_as_double((long long) sd.ljid);
		}
		spos[(size_t) t * smax + k] = make_c4<C>((C) (d3[0] + r[0]), (C) (d3[1] + r[1]), (C) (d3[2] + r[2]), w);
		sori[(size_t) t * smax + k] = make_c4<C>((C) e3[0], (C) e3[1], (C) e3[2], (C) 0);
	}
}

// exclusive scan of cell_count + cell_count_halo over nb bins -> fstart (+ the reference cells' cell_start, the scatter cursor and
// the slot count of the new order in DevScalars::n_slots2[parity ^ 1]); shared with the reduced-memory path (rmm.cu)
static int launch_bin_scan_ex(ls1b200_ctx* c, uint32_t nb, uint32_t* count, uint32_t* count_halo, uint32_t* fstart, uint32_t* scan_tmp, uint32_t* cursor, uint32_t* coarse, uint32_t parity, int tail) {
	const uint32_t sb = (nb + SCAN_B - 1) / SCAN_B;
	LS1_LAUNCH(c, PH_UPDATE, k_scan_block, sb, SCAN_T, 0, count, count_halo, fstart, scan_tmp, nb);
	LS1_LAUNCH(c, PH_UPDATE, k_scan_top, 1, SCAN_T, 0, scan_tmp, sb, scan_tmp + sb);
	LS1_LAUNCH(c, PH_UPDATE, k_scan_add, sb, SCAN_T, 0, fstart, cursor, count, scan_tmp, scan_tmp + sb, nb, (uint32_t) c->cap, parity, c->d_scal, coarse, c->fine_s, tail);
	return 0;
}
int launch_bin_scan(ls1b200_ctx* c, uint32_t nb, uint32_t* fstart, uint32_t parity) {
	return launch_bin_scan_ex(c, nb, c->cell_count, c->cell_count_halo, fstart, c->scan_tmp, c->cursor, c->cell_start, parity, 0);
}

Region make_region(const ls1b200_ctx* c, int mode);      // integrate.cu

static UpdArgs make_upd_args(const ls1b200_ctx* c) {
	UpdArgs a; a.g = c->geom; a.cap = (uint32_t) c->cap; a.parity = c->parity; for (int d = 0; d < 3; d++) a.cut[d] = c->dd.active && c->dd.grid[d] > 1;
	a.S = c->fine_s; a.logS = c->fine_s == 8 ? 3u : c->fine_s == 4 ? 2u : c->fine_s == 2 ? 1u : 0u; a.finv = c->fine_inv;
	a.bx = c->d_border; a.by = c->d_border + c->border_stride; a.bz = c->d_border + 2 * (size_t) c->border_stride;
	a.fix_leavers = 0;
	for (int d = 0; d < 3; d++) {
		const auto& dd = c->dd;
		a.lo_wrap[d] = dd.active && dd.coords[d] == 0; a.hi_wrap[d] = dd.active && dd.coords[d] == dd.grid[d] - 1;
		a.gmin[d] = dd.gmin[d]; a.gmax[d] = dd.gmax[d]; a.glen[d] = dd.gmax[d] - dd.gmin[d];
		a.wrap_hi[d] = dd.wrap_hi[d]; a.halo_hi[d] = dd.halo_hi[d]; a.halo_lo[d] = dd.halo_lo[d];
	}
	return a;
}

// fused step (ls1b200_simulate, ensembles without rotation): integration + binning of every slot in one kernel; do_update(c, 3) continues
int do_pre_key(ls1b200_ctx* c, double dt, bool use_beta, bool post_kick) {
	if (c->n_owned == 0 && !c->dd.active) return 0;
	if (c->rotates || c->fine_s != 1) return fail("do_pre_key: ensembles without rotation, one bin per cell");
	LS1_TRY(span_begin(c, PH_PRE));
	const UpdArgs a = make_upd_args(c);
	const uint32_t gs = (uint32_t) ((c->cap + 255) / 256);
	LS1_CUDA(cudaMemsetAsync(c->cell_count, 0, sizeof(uint32_t) * 2 * ((size_t) c->geom.ncells * c->fine_s + 4), c->stream));      // cell_count and cell_count_halo
	LS1_LAUNCH(c, PH_PRE, k_pre_key, gs, 256, 0, a, c->cur, c->F, c->d_comps, dt, (int) use_beta, (int) post_kick, c->key, c->cell_count, c->cell_count_halo, c->rank, c->d_scal,
	           c->state_lazy ? (const uint32_t*) c->perm : (const uint32_t*) nullptr, c->nxt);
	c->state_lazy = false;
	LS1_TRY(span_end(c));
	c->forces_valid = false;
	return 0;
}

int do_update(ls1b200_ctx* c, int part) {
	if (!c->components_set) return fail("ls1b200_update: components not set");
	if (c->n_owned == 0 && !c->dd.active) { c->sorted = true; return 0; }
	if (c->cap == 0) return fail("ls1b200_update: call ls1b200_upload first");
	LS1_TRY(span_begin(c, PH_UPDATE));
	UpdArgs a = make_upd_args(c);
	a.fix_leavers = part == 4;
	const uint32_t nb = c->geom.ncells * c->fine_s + 1;   // fine bins + limbo bin
	// slot counts are only known on the device: every per-slot kernel is launched over the capacity, threads beyond exit
	const uint32_t gs = (uint32_t) ((c->cap + 255) / 256);
	if (part == 0 || part == 1) LS1_CUDA(cudaMemsetAsync(c->cell_count, 0, sizeof(uint32_t) * 2 * ((size_t) c->geom.ncells * c->fine_s + 4), c->stream));      // cell_count and cell_count_halo
	// part 3: the sorted slots were binned by k_pre_key; only the imports the exchange appended behind them are left (decomposed contexts)
	// part 4: the overlapped multi-GPU step -- as part 3, but the imports are NOT part of this sort (the exchange is still appending them
	// on its own stream; do_update_imports sorts them behind these slots later)
	if (part == 4) {
	} else if (part == 3) {
		if (c->dd.active) LS1_LAUNCH(c, PH_UPDATE, k_key_tail, 592, 256, 0, a, c->cur, c->key, c->cell_count, c->cell_count_halo, c->rank, c->d_scal);
	} else
		LS1_LAUNCH(c, PH_UPDATE, k_key, gs, 256, 0, a, c->cur, c->key, c->cell_count, c->cell_count_halo, c->rank, c->d_scal, make_region(c, part == 0 ? 0 : part == 1 ? 2 : part == 2 ? 1 : 3));
	if (part == 1) { LS1_TRY(span_end(c)); return 0; }
	// cell_start_fine: exclusive scan over the fine bins; cell_start (the reference cells) is every S-th entry of it
	uint32_t* fstart = c->fine_s > 1 ? c->cell_start_fine : c->cell_start;
	if (part == 4) {
		// the exchange is still walking the PREVIOUS order on its own stream, with the cell_start it was launched with: this sort writes the other buffer
		if (c->fine_s != 1) return fail("ls1b200_update: the overlapped step needs one bin per cell");
		if (!c->cell_start_alt) LS1_CUDA(cudaMalloc((void**) &c->cell_start_alt, sizeof(uint32_t) * ((size_t) c->geom.ncells + 4)));
		std::swap(c->cell_start, c->cell_start_alt);
		fstart = c->cell_start;
		LS1_TRY(launch_bin_scan_ex(c, nb, c->cell_count, c->cell_count_halo, fstart, c->scan_tmp, c->cursor, fstart, a.parity, 0));
		// from here the imports' sort may run (exchange stream): it needs n_sorted and writes behind the slots this sort fills
		LS1_CUDA(cudaEventRecord(c->ev_m, c->stream));
	} else
	LS1_TRY(launch_bin_scan(c, nb, fstart, a.parity));
	// The chunk list of the production 1CLJ kernel needs cell_start only: it is built on a side stream while scatter and gather
	// (HBM bound) permute the state; the force kernel waits for it (force_1clj_tiled.cu). The reduced-memory path has its own sort.
	c->chunks_ready = false;
	if (c->all_1clj && c->lj_kernel == 0 && c->fine_s == 1 && c->chunks_early) {
		if (!c->stream_c) {
			int prio_lo = 0, prio_hi = 0;      // high priority: a few small latency-bound kernels beside the 67 k blocks of the gather
			LS1_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
			LS1_CUDA(cudaStreamCreateWithPriority(&c->stream_c, cudaStreamNonBlocking, prio_hi));
			LS1_CUDA(cudaEventCreateWithFlags(&c->ev_c0, cudaEventDisableTiming)); LS1_CUDA(cudaEventCreateWithFlags(&c->ev_c1, cudaEventDisableTiming));
		}
		LS1_CUDA(cudaEventRecord(c->ev_c0, c->stream));
		LS1_CUDA(cudaStreamWaitEvent(c->stream_c, c->ev_c0, 0));
		cudaStream_t main_stream = c->stream;
		c->stream = c->stream_c;
		const int rc = build_chunks_tiled_1clj(c, &c->chunks_nrows, part == 4 ? 1 : 0);
		c->stream = main_stream;
		if (rc) return rc;
		LS1_CUDA(cudaEventRecord(c->ev_c1, c->stream_c));
		c->chunks_ready = true;
	}
	LS1_LAUNCH(c, PH_UPDATE, k_scatter, gs, 256, 0, a, c->cur, c->key, c->rank, (const uint32_t*) fstart, c->cursor, c->entry, c->cell_of, c->d_scal, c->dd.active ? c->ekey : (uint64_t*) nullptr, part == 4 ? 1 : 0);
	const uint32_t gc = gs;
	// lazy permutation: asked for by ls1b200_simulate for the steps whose successor starts with k_pre_key (it moves r and v itself)
	uint32_t* lazy_perm = nullptr;
	if (c->lazy_next && part == 3 && !c->rotates) {
		if (!c->perm) LS1_CUDA(cudaMalloc((void**) &c->perm, sizeof(uint32_t) * c->cap));
		lazy_perm = c->perm;
	}
	if (c->mixed)
		LS1_LAUNCH(c, PH_UPDATE, k_gather<float>, gc, 256, 0, a, c->cur, c->nxt, c->entry, c->cell_of, (const uint32_t*) fstart, c->cell_of2, (float4*) c->mpos, c->d_scal, (int) c->rotates, c->F, c->Vi, c->M, c->dd.active ? c->ekey : (const uint64_t*) nullptr, c->region_flag, part == 4 ? 1 : 0, lazy_perm);
	else
		LS1_LAUNCH(c, PH_UPDATE, k_gather<double>, gc, 256, 0, a, c->cur, c->nxt, c->entry, c->cell_of, (const uint32_t*) fstart, c->cell_of2, (double4*) c->mpos, c->d_scal, (int) c->rotates, c->F, c->Vi, c->M, c->dd.active ? c->ekey : (const uint64_t*) nullptr, c->region_flag, part == 4 ? 1 : 0, lazy_perm);
	std::swap(c->cur, c->nxt);
	c->state_lazy = lazy_perm != nullptr;
	c->parity ^= 1u;
	if (!c->all_1clj) {
		const uint32_t g2 = (uint32_t) ((c->cap + 127) / 128);
		if (c->mixed)
			LS1_LAUNCH(c, PH_UPDATE, k_sites<float>, g2, 128, 0, c->cur, (uint32_t) c->cap, c->smax, c->d_comps, c->d_sites, (float4*) c->spos, (float4*) c->sori, c->d_scal, c->parity, (int) c->rotates);
		else
			LS1_LAUNCH(c, PH_UPDATE, k_sites<double>, g2, 128, 0, c->cur, (uint32_t) c->cap, c->smax, c->d_comps, c->d_sites, (double4*) c->spos, (double4*) c->sori, c->d_scal, c->parity, (int) c->rotates);
	}
	LS1_TRY(span_end(c));
	c->sorted = true; c->ever_sorted = true;
	c->forces_valid = false;
	return 0;
}

// Overlapped multi-GPU step: the imports the exchange appended behind the n_pre slots of the OLD order (c->nxt after do_update(c, 4)
// swapped the buffers) get a sort of their own -- histogram, scan, scatter, gather with their own bin arrays -- into the slots behind
// the n_sorted slots of the main sort. cell_start_b[cell] is absolute: the force kernel reads a cell's molecules as two ranges.
// Runs on c->stream (the caller puts the exchange stream there) after the main gather has finished.
int do_update_imports(ls1b200_ctx* c) {
	if (c->fine_s != 1) return fail("do_update_imports: one bin per cell");
	const uint32_t nb = c->geom.ncells + 1;
	if (!c->cell_count_b) {
		const size_t nbb = (size_t) c->geom.ncells + 4;
		uint32_t* p = nullptr;
		LS1_CUDA(cudaMalloc((void**) &p, sizeof(uint32_t) * (4 * nbb + nbb / 2048 + 16)));
		c->cell_count_b = p; c->cell_start_b = p + 2 * nbb; c->cursor_b = p + 3 * nbb; c->scan_tmp_b = p + 4 * nbb;
	}
	const size_t nbb = (size_t) c->geom.ncells + 4;
	UpdArgs a = make_upd_args(c);
	a.parity = c->parity ^ 1u;      // the order the imports were appended to
	LS1_CUDA(cudaMemsetAsync(c->cell_count_b, 0, sizeof(uint32_t) * 2 * nbb, c->stream));
	uint32_t* count = c->cell_count_b; uint32_t* count_halo = c->cell_count_b + nbb;
	LS1_LAUNCH(c, PH_XCHG, k_key_tail, 296, 256, 0, a, c->nxt, c->key, count, count_halo, c->rank, c->d_scal);
	LS1_TRY(launch_bin_scan_ex(c, nb, count, count_halo, c->cell_start_b, c->scan_tmp_b, c->cursor_b, c->cell_start_b, a.parity, 1));
	LS1_LAUNCH(c, PH_XCHG, k_scatter_tail, 296, 256, 0, a, c->nxt, c->key, c->rank, (const uint32_t*) c->cell_start_b, c->cursor_b, c->entry, c->cell_of, c->d_scal, c->ekey);
	if (c->mixed)
		LS1_LAUNCH(c, PH_XCHG, k_gather_tail<float>, 296, 256, 0, a, c->nxt, c->cur, c->entry, c->cell_of, (const uint32_t*) c->cell_start_b, c->cell_of2, (float4*) c->mpos, c->d_scal, (int) c->rotates, c->F, c->Vi, c->M, (const uint64_t*) c->ekey);
	else
		LS1_LAUNCH(c, PH_XCHG, k_gather_tail<double>, 296, 256, 0, a, c->nxt, c->cur, c->entry, c->cell_of, (const uint32_t*) c->cell_start_b, c->cell_of2, (double4*) c->mpos, c->d_scal, (int) c->rotates, c->F, c->Vi, c->M, (const uint64_t*) c->ekey);
	// the chunk list of the boundary pass (its own buffers: the inner chunks are being traversed on the main stream)
	c->chunks_b_ready = false;
	if (c->all_1clj && c->lj_kernel == 0) {
		LS1_TRY(build_chunks_tiled_1clj(c, &c->chunks_b_nrows, 2, 1));
		c->chunks_b_ready = true;
	}
	return 0;
}

}  // namespace ls1b200
