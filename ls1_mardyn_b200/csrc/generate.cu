// generate.cu -- ls1b200_generate_cubic_grid(): the reference's CubicGridGenerator on the device, so that 10^7 - 10^9 molecules
// never exist as host arrays.
//
// Reference: io/CubicGridGeneratorInternal.cpp:125-187 (molecules per dimension, one body-centred cubic lattice over the
// whole simulation box), particleContainer/ParticleCellBase.cpp:51-159 (initCubicGrid: spacing = L / n, two sub-lattices at
// 1/4 and 3/4 of the spacing, coordinates created in vcp_real_calc, membership tested with min <= r < max in FP64),
// utils/generator/EqualVelocityAssigner.h:16-29 (|v| = sqrt((3 + rotDOF) T / m), direction from two uniform numbers),
// io/IOHelpers.cpp:8-60 (removeMomentum: v -= sum(m v) / sum(m) over all ranks); Molecule defaults q = (1, 0, 0, 0), L = 0.
//
// Differences from the reference, both by construction: the uniform numbers come from a counter-based generator keyed by the
// molecule's GLOBAL lattice index (the reference re-seeds std::mt19937(0) in every cell, which makes all cells share one
// velocity sequence), so the generated state does not depend on how the box is decomposed; ids are lattice index + 1.
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
};

__host__ __device__ inline unsigned long long splitmix64(unsigned long long x) {
	x += 0x9E3779B97F4A7C15ull;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}
// coordinate of lattice index i: "origin + i * spacing" created in the calc type (ParticleCellBase.cpp:95-97)
__host__ __device__ inline double lattice_coord(const GenArgs& a, int s, int d, long long i) {
#ifdef __CUDA_ARCH__
	const double x = __dadd_rn(a.origin[s][d], __dmul_rn((double) i, a.spacing[d]));      // no contraction: the host evaluates the same two roundings
#else
	const double x = a.origin[s][d] + (double) i * a.spacing[d];
#endif
	return a.mixed ? (double) (float) x : x;
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
		const double vx = a.v_abs * sin(phi), vy = a.v_abs * cos(phi) * sin(theta), vz = a.v_abs * cos(phi) * cos(theta);
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
__global__ void __launch_bounds__(256) k_gen_shift(State st, uint64_t n, uint64_t cap, double sx, double sy, double sz) {
	const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	st.v[t] -= sx; st.v[cap + t] -= sy; st.v[2 * cap + t] -= sz;
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
                                           int32_t remove_momentum) {
	if (!c || !mols_per_dim || !sim_box) return fail("ls1b200_generate_cubic_grid: null argument");
	if (!c->components_set) return fail("ls1b200_generate_cubic_grid: call ls1b200_set_components first");
	LS1_CUDA(cudaSetDevice(c->device));
	GenArgs a;
	a.mixed = c->mixed ? 1 : 0; a.seed = seed;
	const CompDev& comp = c->comps[0];      // like the reference's generator: component 0
	a.v_abs = std::sqrt((3.0 + (double) comp.rot_dof) * temperature / comp.mass);
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
			cudaFree(dsum);
			if (rc) return rc;
			k_gen_shift<<<g, 256, 0, c->stream>>>(c->cur, n, cap, h[0] / h[3], h[1] / h[3], h[2] / h[3]); g_launches++;
		}
		LS1_CUDA(cudaStreamSynchronize(c->stream));
	}
	c->n_owned_local = n; c->rot_dof_local = c->rot_dof_total; c->dd.rot_dof_valid = false;
	c->parity = 0; c->n_slots = (uint32_t) n;
	c->h_scal->n_slots2[0] = (uint32_t) n; c->h_scal->n_slots2[1] = 0; c->h_scal->overflow = 0;
	c->h_scal->beta_trans = 1.; c->h_scal->beta_rot = 1.;
	LS1_CUDA(cudaMemcpy(c->d_scal, c->h_scal, sizeof(DevScalars), cudaMemcpyHostToDevice));
	c->sorted = false; c->ever_sorted = false; c->forces_valid = false; c->keys_valid = false;
	return 0;
}
