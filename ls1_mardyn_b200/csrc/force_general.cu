// force_general.cu -- the general (multi-site, multi-component) cell-pair force traversal:
// == LinkedCells::traverseCells(VectorizedCellProcessor&) + Simulation::updateForces() (FullMolecule::calcFM).
//
// Reference: C08BasedTraversals::processBaseCell (LinkedCellTraversals/C08BasedTraversals.h:48-99: which cell pairs),
//   VectorizedCellProcessor::_calculatePairs (adapter/VectorizedCellProcessor.cpp:794-2730: which site pairs, signs),
//   processCell/processCellPair (:2732-2819: halo rules), endTraversal (:122-155),
//   FullMolecule::calcFM (molecules/FullMolecule.cpp:548-626).
//
// B200 mapping: one thread per (owned molecule, site). A thread walks the 3x3 neighbour ROWS of its cell -- each row of
// three x-neighbour cells is one contiguous slot range of the cell-sorted arrays -- tests the molecule-molecule cutoff
// exactly like the reference (pairs.cuh: molpair_r2) and accumulates everything that acts on ITS site in registers
// (FP64 accumulators in both builds, like RealAccumVec). Every physical site pair is therefore visited from both of its
// ends: no Newton-3 scatter, no atomics, no write conflicts, deterministic sums. The forces on halo copies, which the
// reference computes and then discards (LinkedCells::deleteOuterParticles), are never computed. Macroscopic sums get
// weight 1/2 per visit, which equals the reference's "count once" rule (VectorizedCellProcessor.cpp:2795-2817) including
// pairs across the periodic boundary.
#include "pairs.cuh"
#include "types.cuh"

namespace ls1b200 {

template <class C>
struct FArgs {
	const typename Vec4<C>::type* mpos;
	const typename Vec4<C>::type* spos;
	const typename Vec4<C>::type* sori;
	const uint32_t* cell_start;
	const uint32_t* cell_of;
	const CompDev* comps;
	const C* eps24; const C* sig2; const C* shift6;
	double* sf;            // [cap*smax][9] per-site f, V, M
	double* partials;
	const DevScalars* scal;
	uint32_t parity, cap, smax, n_lj;
	int nx, ny, nz;
	C rc2, ljrc2, eps_rf_invrc3;
};

template <class C>
__device__ __forceinline__ SiteC<C> load_site(const FArgs<C>& a, size_t idx, bool oriented) {
	SiteC<C> s;
	const typename Vec4<C>::type p = a.spos[idx];
	s.x = p.x; s.y = p.y; s.z = p.z; s.p = p.w;
	s.ex = s.ey = s.ez = C(0);
	if (oriented) { const typename Vec4<C>::type e = a.sori[idx]; s.ex = e.x; s.ey = e.y; s.ez = e.z; }
	return s;
}

template <class C>
__global__ void __launch_bounds__(128) k_force_general(FArgs<C> a) {
	typedef typename Vec4<C>::type V4;
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t k = blockIdx.y;
	double acc[4] = {0., 0., 0., 0.};   // 6*U_LJ, U_xpoles, virial, myRF
	const uint32_t n_slots = min(a.scal->n_slots2[a.parity], a.cap);
	bool active = t < n_slots;
	V4 mi; uint32_t bits = 0;
	if (active) { mi = a.mpos[t]; bits = wbits<C>(mi.w); active = (bits & 0x80000000u) == 0; }
	if (active) active = k < a.comps[bits & 0xFFFFu].nsites;
	if (active) {
		const CompDev& ci = a.comps[bits & 0xFFFFu];
		const int type = (k < ci.n[0]) ? 0 : (k < ci.n[0] + ci.n[1]) ? 1 : (k < ci.n[0] + ci.n[1] + ci.n[2]) ? 2 : 3;
		const SiteC<C> me = load_site<C>(a, (size_t) t * a.smax + k, type >= 2);
		const uint32_t my_ljid = wbits<C>(me.p);
		double f[3] = {0., 0., 0.}, V[3] = {0., 0., 0.}, M[3] = {0., 0., 0.};
		const uint32_t c = a.cell_of[t];
		const int cx = c % a.nx, cy = (c / a.nx) % a.ny, cz = c / (a.nx * a.ny);
		const uint32_t cb = a.cell_start[c], ce = a.cell_start[c + 1];
		const C cut2 = (type == 0) ? a.ljrc2 : a.rc2;
		for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) {
			const uint32_t row = (uint32_t) (((cz + dz) * a.ny + (cy + dy)) * a.nx + (cx - 1));
			const uint32_t jb = a.cell_start[row], je = a.cell_start[row + 3];
			for (uint32_t j = jb; j < je; j++) {
				const V4 mj = a.mpos[j];
				const C r2 = molpair_r2<C>(mi.x, mi.y, mi.z, mj.x, mj.y, mj.z);
				// SingleCellPolicy_: inside one cell pairs at distance 0 (the molecule itself) are masked out
				// (SIMD_VectorizedCellProcessorHelpers.h:367-376); strict < like calcDistLookup (:443-449)
				if (!(r2 < cut2)) continue;
				if (j >= cb && j < ce && !(r2 != C(0))) continue;
				const CompDev& cj = a.comps[wbits<C>(mj.w) & 0xFFFFu];
				const size_t sj = (size_t) j * a.smax;
				const C dmx = mi.x - mj.x, dmy = mi.y - mj.y, dmz = mi.z - mj.z;   // COM difference (i - j) for the virial
				C ff[3], M1[3], M2[3], u, rf;
				// sign: +1 if this site is "A" of the pair function (f acts on it), -1 if it is "B"
				#define ACC_F(sign, dxm, dym, dzm)                                                           \
					{                                                                                         \
						f[0] += (sign) * (double) ff[0]; f[1] += (sign) * (double) ff[1]; f[2] += (sign) * (double) ff[2]; \
						const double v0 = (double) ((dxm) * ff[0]), v1 = (double) ((dym) * ff[1]), v2 = (double) ((dzm) * ff[2]); \
						V[0] += v0; V[1] += v1; V[2] += v2;                                                     \
						acc[2] += 0.5 * (v0 + v1 + v2);                                                         \
					}
				#define ACC_M(Mx) { M[0] += (double) Mx[0]; M[1] += (double) Mx[1]; M[2] += (double) Mx[2]; }
				if (type == 0) {
					for (uint32_t b = 0; b < cj.n[0]; b++) {
						const SiteC<C> o = load_site<C>(a, sj + b, false);
						const size_t ti = (size_t) my_ljid * a.n_lj + wbits<C>(o.p);
						pair_lj<C>(me, o, a.eps24[ti], a.sig2[ti], a.shift6[ti], ff, u);
						ACC_F(1., dmx, dmy, dmz);
						acc[0] += 0.5 * (double) u;
					}
				} else if (type == 1) {
					uint32_t b = cj.n[0];
					for (uint32_t e = b + cj.n[1]; b < e; b++) {          // charge - charge
						const SiteC<C> o = load_site<C>(a, sj + b, false);
						pair_charge<C>(me, o, ff, u);
						ACC_F(1., dmx, dmy, dmz); acc[1] += 0.5 * (double) u;
					}
					for (uint32_t e = b + cj.n[2]; b < e; b++) {          // charge(me) - dipole(j)
						const SiteC<C> o = load_site<C>(a, sj + b, true);
						pair_charge_dipole<C>(me, o, ff, M2, u);
						ACC_F(1., dmx, dmy, dmz); acc[1] += 0.5 * (double) u;
					}
					for (uint32_t e = b + cj.n[3]; b < e; b++) {          // charge(me) - quadrupole(j)
						const SiteC<C> o = load_site<C>(a, sj + b, true);
						pair_charge_quadrupole<C>(me, o, ff, M2, u);
						ACC_F(1., dmx, dmy, dmz); acc[1] += 0.5 * (double) u;
					}
				} else if (type == 2) {
					uint32_t b = cj.n[0];
					for (uint32_t e = b + cj.n[1]; b < e; b++) {          // charge(j) - dipole(me): f acts on the charge
						const SiteC<C> o = load_site<C>(a, sj + b, false);
						pair_charge_dipole<C>(o, me, ff, M1, u);
						ACC_F(-1., -dmx, -dmy, -dmz); ACC_M(M1); acc[1] += 0.5 * (double) u;
					}
					for (uint32_t e = b + cj.n[2]; b < e; b++) {          // dipole - dipole
						const SiteC<C> o = load_site<C>(a, sj + b, true);
						pair_dipole<C>(me, o, a.eps_rf_invrc3, ff, M1, M2, u, rf);
						ACC_F(1., dmx, dmy, dmz); ACC_M(M1); acc[1] += 0.5 * (double) u; acc[3] += 0.5 * (double) rf;
					}
					for (uint32_t e = b + cj.n[3]; b < e; b++) {          // dipole(me) - quadrupole(j)
						const SiteC<C> o = load_site<C>(a, sj + b, true);
						pair_dipole_quadrupole<C>(me, o, ff, M1, M2, u);
						ACC_F(1., dmx, dmy, dmz); ACC_M(M1); acc[1] += 0.5 * (double) u;
					}
				} else {
					uint32_t b = cj.n[0];
					for (uint32_t e = b + cj.n[1]; b < e; b++) {          // charge(j) - quadrupole(me)
						const SiteC<C> o = load_site<C>(a, sj + b, false);
						pair_charge_quadrupole<C>(o, me, ff, M1, u);
						ACC_F(-1., -dmx, -dmy, -dmz); ACC_M(M1); acc[1] += 0.5 * (double) u;
					}
					for (uint32_t e = b + cj.n[2]; b < e; b++) {          // dipole(j) - quadrupole(me)
						const SiteC<C> o = load_site<C>(a, sj + b, true);
						pair_dipole_quadrupole<C>(o, me, ff, M1, M2, u);
						ACC_F(-1., -dmx, -dmy, -dmz); ACC_M(M2); acc[1] += 0.5 * (double) u;
					}
					for (uint32_t e = b + cj.n[3]; b < e; b++) {          // quadrupole - quadrupole
						const SiteC<C> o = load_site<C>(a, sj + b, true);
						pair_quadrupole<C>(me, o, ff, M1, M2, u);
						ACC_F(1., dmx, dmy, dmz); ACC_M(M1); acc[1] += 0.5 * (double) u;
					}
				}
				#undef ACC_F
				#undef ACC_M
			}
		}
		double* o = a.sf + ((size_t) t * a.smax + k) * 9;
		o[0] = f[0]; o[1] = f[1]; o[2] = f[2]; o[3] = V[0]; o[4] = V[1]; o[5] = V[2]; o[6] = M[0]; o[7] = M[1]; o[8] = M[2];
	}
	block_reduce_store<4>(acc, a.partials, blockIdx.y * gridDim.x + blockIdx.x);
}

// Quaternion::rotate (molecules/Quaternion.cpp:43-64)
__device__ __forceinline__ void q_rot(const double* q, const double* d, double* o) {
	if (d[0] == 0. && d[1] == 0. && d[2] == 0.) { o[0] = o[1] = o[2] = 0.; return; }
	const double w = q[0], x = q[1], y = q[2], z = q[3];
	const double ww = w * w, xx = x * x, yy = y * y, zz = z * z, wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
	o[0] = (ww + xx - yy - zz) * d[0] + 2. * (xy - wz) * d[1] + 2. * (wy + xz) * d[2];
	o[1] = 2. * (wz + xy) * d[0] + (ww - xx + yy - zz) * d[1] + 2. * (yz - wx) * d[2];
	o[2] = 2. * (xz - wy) * d[0] + 2. * (wx + yz) * d[1] + (ww - xx - yy + zz) * d[2];
}

// FullMolecule::calcFM (molecules/FullMolecule.cpp:548-626): F = sum f, M = sum d x f + sum M_site, Vi = 1/2 sum V
__global__ void __launch_bounds__(128) k_calcfm(State st, uint32_t cap, uint32_t smax, const CompDev* __restrict__ comps, const SiteDev* __restrict__ sites,
                                                 const double* __restrict__ sf, const uint32_t* __restrict__ inert_src, int mixed,
                                                 double* __restrict__ F, double* __restrict__ M, double* __restrict__ Vi,
                                                 const void* __restrict__ mpos, const DevScalars* __restrict__ scal, uint32_t parity, int rotates) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= min(scal->n_slots2[parity], cap)) return;
	const uint32_t bits = mixed ? __float_as_uint(((const float4*) mpos)[t].w) : (uint32_t) __double_as_longlong(((const double4*) mpos)[t].w);
	double Fs[3] = {0., 0., 0.}, Ms[3] = {0., 0., 0.}, Vs[3] = {0., 0., 0.};
	if (!(bits & 0x80000000u)) {
		const CompDev& cp = comps[bits & 0xFFFFu];
		double qn[4] = {1., 0., 0., 0.};
		if (rotates) {
			for (int k = 0; k < 4; k++) qn[k] = st.q[(size_t) k * cap + t];
			const double sc = 1. / sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
			for (int k = 0; k < 4; k++) qn[k] *= sc;
		}
		for (uint32_t k = 0; k < cp.nsites; k++) {
			const double* s = sf + ((size_t) t * smax + k) * 9;
			const SiteDev& sd = sites[cp.site_offset + k];
			double d[3]; q_rot(qn, sd.r, d);
			Fs[0] += s[0]; Fs[1] += s[1]; Fs[2] += s[2];
			Vs[0] += s[3]; Vs[1] += s[4]; Vs[2] += s[5];
			Ms[0] += d[1] * s[2] - d[2] * s[1];
			Ms[1] += d[2] * s[0] - d[0] * s[2];
			Ms[2] += d[0] * s[1] - d[1] * s[0];
			if (sd.type >= 2) { Ms[0] += s[6]; Ms[1] += s[7]; Ms[2] += s[8]; }
		}
	}
	const size_t c = cap;
	F[t] = Fs[0]; F[c + t] = Fs[1]; F[2 * c + t] = Fs[2];
	M[t] = Ms[0]; M[c + t] = Ms[1]; M[2 * c + t] = Ms[2];
	Vi[t] = 0.5 * Vs[0]; Vi[c + t] = 0.5 * Vs[1]; Vi[2 * c + t] = 0.5 * Vs[2];
}

// deterministic final reduction of the per-block partial sums: one block, fixed order
__global__ void __launch_bounds__(1024) k_reduce_force(const double* __restrict__ partials, uint32_t nblocks, DevScalars* scal) {
	double v[4] = {0., 0., 0., 0.};
	for (uint32_t b = threadIdx.x; b < nblocks; b += blockDim.x) {
		#pragma unroll
		for (int k = 0; k < 4; k++) v[k] += partials[(size_t) b * 8 + k];
	}
	__shared__ double out[8];
	block_reduce_store<4>(v, out, 0);
	__syncthreads();
	if (threadIdx.x == 0) {
		scal->upot6lj = out[0]; scal->upot_xpoles = out[1]; scal->virial = out[2];
		scal->my_rf = -out[3];   // sign flip of VectorizedCellProcessor.cpp:2727-2728
	}
}

int reduce_force_partials(ls1b200_ctx* c, uint32_t nblocks) {
	LS1_LAUNCH(c, PH_FORCE, k_reduce_force, 1, 1024, 0, c->partials, nblocks, c->d_scal);
	return 0;
}

template <class C>
static int launch_general(ls1b200_ctx* c) {
	FArgs<C> a;
	a.mpos = (const typename Vec4<C>::type*) c->mpos; a.spos = (const typename Vec4<C>::type*) c->spos; a.sori = (const typename Vec4<C>::type*) c->sori;
	a.cell_start = c->cell_start; a.cell_of = c->cell_of2; a.comps = c->d_comps;
	if (sizeof(C) == 4) { a.eps24 = (const C*) c->d_eps24f; a.sig2 = (const C*) c->d_sig2f; a.shift6 = (const C*) c->d_shift6f; }
	else { a.eps24 = (const C*) c->d_eps24; a.sig2 = (const C*) c->d_sig2; a.shift6 = (const C*) c->d_shift6; }
	a.sf = c->sf; a.partials = c->partials; a.scal = c->d_scal;
	a.parity = c->parity; a.cap = (uint32_t) c->cap; a.smax = c->smax; a.n_lj = c->n_lj;
	a.nx = c->geom.cells[0]; a.ny = c->geom.cells[1]; a.nz = c->geom.cells[2];
	const double rc = c->cfg.cutoff, rclj = c->cfg.cutoff_lj, erf = c->cfg.epsilon_rf;
	a.rc2 = (C) (rc * rc); a.ljrc2 = (C) (rclj * rclj);
	a.eps_rf_invrc3 = (C) (2. * (erf - 1.) / ((rc * rc * rc) * (2. * erf + 1.)));   // VectorizedCellProcessor.cpp:22
	const uint32_t gx = (uint32_t) ((c->cap + 127) / 128);
	if ((uint64_t) gx * c->smax > c->partial_blocks) return fail("force_general: partials buffer too small");
	dim3 grid(gx, c->smax);
	LS1_LAUNCH(c, PH_FORCE, k_force_general<C>, grid, 128, 0, a);
	LS1_LAUNCH(c, PH_FORCE, k_calcfm, gx, 128, 0, c->cur, (uint32_t) c->cap, c->smax, c->d_comps, c->d_sites, c->sf, (const uint32_t*) nullptr, (int) c->mixed,
	           c->F, c->M, c->Vi, (const void*) c->mpos, c->d_scal, c->parity, (int) c->rotates);
	return reduce_force_partials(c, gx * c->smax);
}

// FullMolecule::calcFM for all slots from the per-site records (also used by force_multi_tiled.cu)
int launch_calcfm(ls1b200_ctx* c) {
	const uint32_t gx = (uint32_t) ((c->cap + 127) / 128);
	LS1_LAUNCH(c, PH_FORCE, k_calcfm, gx, 128, 0, c->cur, (uint32_t) c->cap, c->smax, c->d_comps, c->d_sites, c->sf, (const uint32_t*) nullptr, (int) c->mixed,
	           c->F, c->M, c->Vi, (const void*) c->mpos, c->d_scal, c->parity, (int) c->rotates);
	return 0;
}

int force_general(ls1b200_ctx* c) { return c->mixed ? launch_general<float>(c) : launch_general<double>(c); }

}  // namespace ls1b200
