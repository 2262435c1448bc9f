// generate.cu -- ls1b200_generate_cubic_grid(): the reference's CubicGridGenerator on the device, so that 10^7 - 10^9 molecules
// never exist as host arrays.
//
// Reference: io/CubicGridGeneratorInternal.cpp:125-187 (molecules per dimension, one body-centred cubic lattice over the
// whole simulation box), particleContainer/ParticleCellBase.cpp:51-159 (initCubicGrid: spacing = L / n, two sub-lattices at
// 1/4 and 3/4 of the spacing, coordinates created in vcp_real_calc, membership tested with min <= r < max in FP64),
// utils/generator/EqualVelocityAssigner.h:16-29 (|v| = sqrt((3 + rotDOF) T / m), direction from two uniform numbers),
// io/IOHelpers.cpp:8-60 (removeMomentum: v -= sum(m v) / sum(m) over all ranks); Molecule defaults q = (1, 0, 0, 0), L = 0.
//
// Two velocity modes. Default: the uniform numbers come from a counter-based generator keyed by the molecule's GLOBAL lattice
// index, so the generated state does not depend on how the box is decomposed. LS1B200_GEN_REFERENCE_VELOCITIES: the reference's
// own sequence -- initCubicGrid constructs a fresh EqualVelocityAssigner, i.e. std::mt19937(0), in EVERY cell, so the k-th molecule
// a cell inserts (loop order i, j, k; sub-lattice 1 before sub-lattice 2) gets the k-th direction of one fixed sequence. The
// directions are tabulated on the host with the same std::mt19937 / std::uniform_real_distribution / libm calls, the device computes
// each molecule's LinkedCells cell (bit-exact borders, geom.cuh) and its insertion rank in closed form. The state then equals what
// the unmodified reference binary generates for the same XML (tests/golden/generator_1clj_2000.npz pins it). ids are lattice index + 1.
#include <random>

#include "geom.cuh"
#include "pairs.cuh"
#include "types.cuh"

namespace ls1b200 {

struct GenArgs {
	double spacing[3], origin[2][3];      // origin[s][d] = spacing * (0.25, 0.75)
	long long lo[2][3];                    // first lattice index of sub-lattice s inside the box in dimension d
	unsigned long long cnt[2][3];          // number of lattice indices inside
	unsigned long long n_glob[3];          // molecules per dimension of the whole lattice
	unsigned long long n1, n;              // molecules of sub-lattice 0, total
	unsigned long long seed;
	double v_abs;
	int mixed;
	int ref_vel;                           // reference mode: velocity = vtab[insertion rank inside the LinkedCells cell]
	unsigned int vtab_n;
	const double* vtab;                    // [vtab_n][3]
	Geom g;
};

__host__ __device__ inline unsigned long long splitmix64(unsigned long long x) {
	x += 0x9E3779B97F4A7C15ull;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}
// coordinate of lattice index i: "origin + i * spacing" created in the calc type (ParticleCellBase.cpp:95-97). ONE fused
// multiply-add: the reference binary (g++ -mfma) contracts the expression -- the double build's coordinates in
// tests/golden/generator_1clj_2662.npz are the FMA's, one ulp off the two-rounding result in 9 of 22 planes
__host__ __device__ inline double lattice_coord(const GenArgs& a, int s, int d, long long i) {
	const double x = fma((double) i, a.spacing[d], a.origin[s][d]);
	return a.mixed ? (double) (float) x : x;
}

// ---- reference mode: the rank of a lattice molecule among the molecules ParticleCellBase::initCubicGrid inserts into its cell ----
struct IdxRange { long long first, end; };      // lattice indices [first, end) of one sub-lattice whose coordinate lies in [lo, hi)
__device__ inline IdxRange idx_range(const GenArgs& a, int s, int d, double lo, double hi, long long guess) {
	long long i = guess;
	while (lattice_coord(a, s, d, i - 1) >= lo) i--;
	while (lattice_coord(a, s, d, i) < lo) i++;
	long long j = i;
	while (lattice_coord(a, s, d, j) < hi) j++;
	IdxRange r; r.first = i; r.end = j;
	return r;
}
__device__ inline long long cnt_before(const IdxRange& r, long long i) { return min(max(i - r.first, 0ll), r.end - r.first); }
__device__ inline bool in_range(const IdxRange& r, long long i) { return i >= r.first && i < r.end; }
// molecules of one sub-lattice (ranges R) the loops i, j, k visit strictly before lattice point (i, j, k)
__device__ inline long long lex_before(const IdxRange R[3], long long i, long long j, long long k) {
	const long long ny = R[1].end - R[1].first, nz = R[2].end - R[2].first;
	long long n = cnt_before(R[0], i) * ny * nz;
	if (in_range(R[0], i)) { n += cnt_before(R[1], j) * nz; if (in_range(R[1], j)) n += cnt_before(R[2], k); }
	return n;
}
template <int D>
__device__ inline void cell_interval(const Geom& g, double x, double& lo, double& hi) {
	const int c = cell1d<D>(g, x);      // inside the box: 1 <= c <= cells - 2
	lo = cell_border<D>(g, c); hi = cell_border<D>(g, c + 1);
}
__device__ inline long long insertion_rank(const GenArgs& a, int s, const long long gi[3], const double x[3]) {
	double lo[3], hi[3];
	cell_interval<0>(a.g, x[0], lo[0], hi[0]); cell_interval<1>(a.g, x[1], lo[1], hi[1]); cell_interval<2>(a.g, x[2], lo[2], hi[2]);
	IdxRange A[3], B[3];
	#pragma unroll
	for (int d = 0; d < 3; d++) { A[d] = idx_range(a, 0, d, lo[d], hi[d], gi[d]); B[d] = idx_range(a, 1, d, lo[d], hi[d], gi[d]); }
	long long n = lex_before(A, gi[0], gi[1], gi[2]) + lex_before(B, gi[0], gi[1], gi[2]);
	if (s == 1 && in_range(A[0], gi[0]) && in_range(A[1], gi[1]) && in_range(A[2], gi[2])) n++;      // sub-lattice 1 of the same lattice point comes first
	return n;
}

__global__ void __launch_bounds__(256) k_generate(GenArgs a, State st, uint64_t* __restrict__ d_id, uint64_t cap, double* __restrict__ partials) {
	const unsigned long long t = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
	double mom[3] = {0., 0., 0.};
	if (t < a.n) {
		const int s = t < a.n1 ? 0 : 1;
		unsigned long long u = s ? t - a.n1 : t;
		const unsigned long long k = u % a.cnt[s][2]; u /= a.cnt[s][2];
		const unsigned long long j = u % a.cnt[s][1]; u /= a.cnt[s][1];
		const long long gi = a.lo[s][0] + (long long) u, gj = a.lo[s][1] + (long long) j, gk = a.lo[s][2] + (long long) k;
		st.r[t] = lattice_coord(a, s, 0, gi); st.r[cap + t] = lattice_coord(a, s, 1, gj); st.r[2 * cap + t] = lattice_coord(a, s, 2, gk);
		const unsigned long long gidx = (((unsigned long long) s * a.n_glob[0] + (unsigned long long) gi) * a.n_glob[1] + (unsigned long long) gj) * a.n_glob[2] + (unsigned long long) gk;
		const unsigned long long h1 = splitmix64(a.seed ^ (2ull * gidx)), h2 = splitmix64(a.seed ^ (2ull * gidx + 1ull));
		const double u1 = (double) (h1 >> 11) * (1.0 / 9007199254740992.0), u2 = (double) (h2 >> 11) * (1.0 / 9007199254740992.0);
		const double phi = 2.0 * 3.14159265358979323846 * u1, theta = acos(2.0 * u2 - 1.0);
		double vx = a.v_abs * sin(phi), vy = a.v_abs * cos(phi) * sin(theta), vz = a.v_abs * cos(phi) * cos(theta);
		if (a.ref_vel) {
			const long long gidx3[3] = {gi, gj, gk};
			const double x3[3] = {st.r[t], st.r[cap + t], st.r[2 * cap + t]};
			const long long rk = insertion_rank(a, s, gidx3, x3);
			if (rk < 0 || rk >= (long long) a.vtab_n) { vx = vy = vz = __longlong_as_double(0x7ff8000000000000ll); }      // cannot happen (table sized from the cell and lattice geometry); NaN would fail every check
			else { vx = a.vtab[3 * rk]; vy = a.vtab[3 * rk + 1]; vz = a.vtab[3 * rk + 2]; }
		}
		st.v[t] = vx; st.v[cap + t] = vy; st.v[2 * cap + t] = vz;
		st.meta[t] = make_meta((uint32_t) t, 0u, false);
		d_id[t] = gidx + 1ull;
		if (st.id) st.id[t] = gidx + 1ull;
		mom[0] = vx; mom[1] = vy; mom[2] = vz;
	}
	block_reduce_store<3>(mom, partials, blockIdx.x);
}
__global__ void __launch_bounds__(1024) k_gen_reduce(const double* __restrict__ partials, uint32_t nblocks, double* __restrict__ out) {
	double v[3] = {0., 0., 0.};
	for (uint32_t b = threadIdx.x; b < nblocks; b += blockDim.x) { v[0] += partials[(size_t) b * 8]; v[1] += partials[(size_t) b * 8 + 1]; v[2] += partials[(size_t) b * 8 + 2]; }
	__shared__ double sh[8];
	block_reduce_store<3>(v, sh, 0);
	__syncthreads();
	if (threadIdx.x == 0) { out[0] = sh[0]; out[1] = sh[1]; out[2] = sh[2]; }
}
// v -= shift; partial sums of v^2 of the shifted velocities (IOHelpers::removeMomentum rescales to the target temperature afterwards)
__global__ void __launch_bounds__(256) k_gen_shift(State st, uint64_t n, uint64_t cap, double sx, double sy, double sz, double* __restrict__ partials) {
	const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	double acc[3] = {0., 0., 0.};
	if (t < n) {
		const double vx = st.v[t] - sx, vy = st.v[cap + t] - sy, vz = st.v[2 * cap + t] - sz;
		st.v[t] = vx; st.v[cap + t] = vy; st.v[2 * cap + t] = vz;
		acc[0] = vx * vx + vy * vy + vz * vz;
	}
	block_reduce_store<3>(acc, partials, blockIdx.x);
}
__global__ void __launch_bounds__(256) k_gen_scale(State st, uint64_t n, uint64_t cap, double f) {
	const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	st.v[t] *= f; st.v[cap + t] *= f; st.v[2 * cap + t] *= f;
}
__global__ void k_gen_fill_q(double* __restrict__ q, uint64_t n, uint64_t cap) {
	const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	q[t] = 1.; q[cap + t] = 0.; q[2 * cap + t] = 0.; q[3 * cap + t] = 0.;
}

int allreduce_doubles(ls1b200_ctx* c, double* dev, int n);      // halo.cu: sum over the ranks of a decomposed context with a communicator (no-op otherwise)

}  // namespace ls1b200

using namespace ls1b200;

extern "C" int ls1b200_generate_cubic_grid(ls1b200_ctx* c, const uint64_t mols_per_dim[3], const double sim_box[3], double temperature, uint64_t seed,
                                           int32_t flags) {
	const bool remove_momentum = (flags & LS1B200_GEN_REMOVE_MOMENTUM) != 0, ref_vel = (flags & LS1B200_GEN_REFERENCE_VELOCITIES) != 0;
	if (!c || !mols_per_dim || !sim_box) return fail("ls1b200_generate_cubic_grid: null argument");
	if (!c->components_set) return fail("ls1b200_generate_cubic_grid: call ls1b200_set_components first");
	LS1_CUDA(cudaSetDevice(c->device));
	GenArgs a;
	a.mixed = c->mixed ? 1 : 0; a.seed = seed;
	const CompDev& comp = c->comps[0];      // like the reference's generator: component 0
	a.v_abs = std::sqrt((3.0 + (double) comp.rot_dof) * temperature / comp.mass);
	a.ref_vel = ref_vel ? 1 : 0; a.vtab = nullptr; a.vtab_n = 0; a.g = c->geom;
	uint64_t cnt_total[2] = {1, 1};
	for (int d = 0; d < 3; d++) {
		if (mols_per_dim[d] == 0 || !(sim_box[d] > 0.)) return fail("ls1b200_generate_cubic_grid: empty lattice");
		a.n_glob[d] = mols_per_dim[d];
		a.spacing[d] = sim_box[d] / (double) mols_per_dim[d];
		a.origin[0][d] = a.spacing[d] * 0.25; a.origin[1][d] = a.spacing[d] * 0.25 * 3.;
		// lattice indices whose coordinate lies in [bmin, bmax) of THIS context (ParticleCellBase.cpp:67-75, :95-103); the global box starts at 0
		const double bmin = c->geom.bmin[d], bmax = c->geom.bmax[d];
		for (int s = 0; s < 2; s++) {
			long long i0 = (long long) std::floor(bmin / sim_box[d] * (double) mols_per_dim[d]) - 2, i1 = (long long) std::ceil(bmax / sim_box[d] * (double) mols_per_dim[d]) + 2;
			i0 = std::max<long long>(i0, 0); i1 = std::min<long long>(i1, (long long) mols_per_dim[d]);
			long long lo = i1, hi = i1;      // [lo, hi)
			bool in_run = false;
			for (long long i = i0; i < i1; i++) {
				const double x = lattice_coord(a, s, d, i);
				const bool in = x >= bmin && x < bmax;
				if (in && !in_run) { lo = i; in_run = true; }
				if (!in && in_run) { hi = i; in_run = false; break; }
			}
			if (in_run) hi = i1;
			if (lo > hi) lo = hi;
			a.lo[s][d] = lo; a.cnt[s][d] = (unsigned long long) (hi - lo);
			cnt_total[s] *= a.cnt[s][d];
		}
	}
	for (int s = 0; s < 2; s++) for (int d = 0; d < 3; d++) if (a.cnt[s][d] == 0) a.cnt[s][d] = 1;      // the divisions of an empty sub-lattice
	a.n1 = cnt_total[0]; a.n = cnt_total[0] + cnt_total[1];
	const uint64_t n = a.n;
	if (n >= (1ull << 27)) return fail("ls1b200_generate_cubic_grid: more than 2^27 molecules per context are not supported");
	double fac = 1.;
	for (int d = 0; d < 3; d++) if (c->geom.periodic[d] || c->cfg.reserved[0] || c->dd.active) fac *= 1. + 2. * std::min(1., c->geom.cutoff / c->geom.lbox[d]);
	uint64_t cap = n + (uint64_t) ((fac - 1.) * 1.5 * (double) n) + (c->dd.active ? n / 10 + 16384 : 0) + 4096;
	cap = (cap + 63) & ~63ull;
	LS1_TRY(ctx_alloc_molecules(c, n, cap));
	c->rot_dof_total = (uint64_t) comp.rot_dof * n;
	struct DevBuf { double* p = nullptr; ~DevBuf() { if (p) cudaFree(p); } } vtab_guard;      // freed on every way out
	double*& d_vtab = vtab_guard.p;
	if (ref_vel && n) {
		// The reference's velocity sequence (ParticleCellBase.cpp:66-67, EqualVelocityAssigner.h:11-29): a fresh std::mt19937(0) per cell,
		// phi and theta from two draws of std::uniform_real_distribution<double>(0, 1) per molecule; the ensemble's temperature passes
		// through a vcp_real_calc variable first. Same library calls as the reference => the same bits. One entry per molecule a cell can hold.
		uint64_t m = 2;
		for (int d = 0; d < 3; d++) m *= (uint64_t) std::ceil(c->geom.clen[d] / a.spacing[d]) + 2;
		if (m > (1ull << 22)) return fail("ls1b200_generate_cubic_grid: more than 4 M lattice molecules per cell");
		const double T_calc = c->mixed ? (double) (float) temperature : temperature;
		const double v_abs = std::sqrt((3 + (double) comp.rot_dof) * T_calc / comp.mass);
		std::vector<double> tab(3 * m);
		std::mt19937 mt(0);
		std::uniform_real_distribution<double> uni(0, 1);
		for (uint64_t k = 0; k < m; k++) {
			const double phi = 2 * M_PI * uni(mt);
			const double theta = std::acos(2 * uni(mt) - 1);
			tab[3 * k] = v_abs * std::sin(phi); tab[3 * k + 1] = v_abs * std::cos(phi) * std::sin(theta); tab[3 * k + 2] = v_abs * std::cos(phi) * std::cos(theta);
		}
		LS1_CUDA(cudaMalloc((void**) &d_vtab, sizeof(double) * 3 * m));
		LS1_CUDA(cudaMemcpy(d_vtab, tab.data(), sizeof(double) * 3 * m, cudaMemcpyHostToDevice));
		a.vtab = d_vtab; a.vtab_n = (unsigned int) m;
	}
	if (n) {
		const uint32_t g = (uint32_t) ((n + 255) / 256);
		if ((uint64_t) g > c->partial_blocks) return fail("ls1b200_generate_cubic_grid: partials buffer too small");
		k_generate<<<g, 256, 0, c->stream>>>(a, c->cur, c->d_id, cap, c->partials); g_launches++;
		LS1_CUDA(cudaGetLastError());
		if (c->rotates) {
			k_gen_fill_q<<<g, 256, 0, c->stream>>>(c->cur.q, n, cap); g_launches++;
			LS1_CUDA(cudaMemsetAsync(c->cur.L, 0, sizeof(double) * 3 * cap, c->stream));
		}
		if (remove_momentum) {
			// one component: sum(m v) / sum(m) = sum(v) / N over ALL ranks
			double* dsum = nullptr;
			LS1_CUDA(cudaMalloc((void**) &dsum, 4 * sizeof(double)));
			k_gen_reduce<<<1, 1024, 0, c->stream>>>(c->partials, g, dsum); g_launches++;
			const double nd = (double) n;
			LS1_CUDA(cudaMemcpyAsync(dsum + 3, &nd, sizeof(double), cudaMemcpyHostToDevice, c->stream));
			int rc = allreduce_doubles(c, dsum, 4);
			double h[4] = {0., 0., 0., 1.};
			if (rc == 0) { cudaMemcpyAsync(h, dsum, sizeof(h), cudaMemcpyDeviceToHost, c->stream); cudaStreamSynchronize(c->stream); }
			if (rc) { cudaFree(dsum); return rc; }
			k_gen_shift<<<g, 256, 0, c->stream>>>(c->cur, n, cap, h[0] / h[3], h[1] / h[3], h[2] / h[3], c->partials); g_launches++;
			// "After subtraction, the temperature was effectively changed": the velocities are scaled to the target temperature again
			// (IOHelpers.cpp:75-115; T_current = 2 sum(m v^2 / 2) / sum(3 + rotDOF) over all ranks; L = 0 on a fresh lattice)
			k_gen_reduce<<<1, 1024, 0, c->stream>>>(c->partials, g, dsum); g_launches++;
			LS1_CUDA(cudaMemcpyAsync(dsum + 3, &nd, sizeof(double), cudaMemcpyHostToDevice, c->stream));
			rc = allreduce_doubles(c, dsum, 4);
			if (rc == 0) { cudaMemcpyAsync(h, dsum, sizeof(h), cudaMemcpyDeviceToHost, c->stream); cudaStreamSynchronize(c->stream); }
			cudaFree(dsum);
			if (rc) return rc;
			const double t_current = comp.mass * h[0] / ((3.0 + (double) comp.rot_dof) * h[3]);
			if (t_current > 0.) { k_gen_scale<<<g, 256, 0, c->stream>>>(c->cur, n, cap, std::sqrt(temperature / t_current)); g_launches++; }
		}
		LS1_CUDA(cudaStreamSynchronize(c->stream));
	}
	c->n_owned_local = n; c->rot_dof_local = c->rot_dof_total; c->dd.rot_dof_valid = false;
	c->parity = 0; c->n_slots = (uint32_t) n;
	c->h_scal->n_slots2[0] = (uint32_t) n; c->h_scal->n_slots2[1] = 0; c->h_scal->overflow = 0;
	c->h_scal->beta_trans = 1.; c->h_scal->beta_rot = 1.;
	c->h_scal->clip_limit = 0.; c->h_scal->expl_counter = c->h_scal->expl_warning = c->h_scal->expl_error = 0; c->globals_initialised = false;      // a new state: Domain's explosion heuristics start over
	LS1_CUDA(cudaMemcpy(c->d_scal, c->h_scal, sizeof(DevScalars), cudaMemcpyHostToDevice));
	c->sorted = false; c->ever_sorted = false; c->forces_valid = false; c->keys_valid = false;
	return 0;
}
