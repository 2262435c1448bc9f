// integrate.cu -- Leapfrog integrator, velocity-scaling thermostat, Domain's global values and the fused multi-step loop.
//
// Reference: Leapfrog::transition1to2 / transition2to3 (integrators/Leapfrog.cpp:45-61, :63-147);
//   FullMolecule::upd_preF / upd_postF (molecules/FullMolecule.cpp:331-361, :363-386);
//   Quaternion::rotateinv / differentiate (molecules/Quaternion.cpp:66-103);
//   Domain::calculateGlobalValues, getGlobalPressure (Domain.cpp:157-252, :142-146);
//   VelocityScalingThermostat::apply (thermostats/VelocityScalingThermostat.cpp:68-89);
//   Simulation::simulateOneTimestep for the order of the phases (Simulation.cpp:1076-1261).
// All kernels here are HBM-bound streaming kernels over the FP64 state planes.
#include "pairs.cuh"
#include "types.cuh"

namespace ls1b200 {

__device__ __forceinline__ void q_rotateinv(const double* q, const double* d, double* o) {
	if (d[0] == 0. && d[1] == 0. && d[2] == 0.) { o[0] = o[1] = o[2] = 0.; return; }
	const double w = q[0], x = q[1], y = q[2], z = q[3];
	const double ww = w * w, xx = x * x, yy = y * y, zz = z * z, wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
	o[0] = (ww + xx - yy - zz) * d[0] + 2. * (xy + wz) * d[1] + 2. * (xz - wy) * d[2];
	o[1] = 2. * (xy - wz) * d[0] + (ww - xx + yy - zz) * d[1] + 2. * (yz + wx) * d[2];
	o[2] = 2. * (xz + wy) * d[0] + 2. * (yz - wx) * d[1] + (ww - xx - yy + zz) * d[2];
}
__device__ __forceinline__ void q_differentiate(const double* q, const double* w, double* dq) {
	dq[0] = .5 * (-q[1] * w[0] - q[2] * w[1] - q[3] * w[2]);
	dq[1] = .5 * (q[0] * w[0] - q[3] * w[1] + q[2] * w[2]);
	dq[2] = .5 * (q[3] * w[0] + q[0] * w[1] - q[1] * w[2]);
	dq[3] = .5 * (-q[2] * w[0] + q[1] * w[1] + q[0] * w[2]);
}

// upd_preF, with the thermostat factors of the previous step applied on load (v *= beta_trans, L *= beta_rot is what
// VelocityScalingThermostat::apply does between two steps; nothing reads v or L in between).
__global__ void __launch_bounds__(256) k_leapfrog_pre(State st, const double* __restrict__ F, const double* __restrict__ M, uint32_t cap,
                                                       const CompDev* __restrict__ comps, double dt, const DevScalars* __restrict__ scal,
                                                       uint32_t parity, int use_beta, int rotates, Region rg) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (scal->overflow) return;      // a failed step (lost message, peer time-out, another rank's fault): the state stays as it is
	if (t >= min(scal->n_slots2[parity], cap)) return;
	if (!region_selected(rg, t, scal->n_sorted)) return;
	const uint64_t meta = st.meta[t];
	if (meta_halo(meta)) return;
	const CompDev& cp = comps[meta_cid(meta)];
	const double bt = use_beta ? scal->beta_trans : 1., br = use_beta ? scal->beta_rot : 1.;
	const double dth = .5 * dt, dtInv2m = dth / cp.mass;
	const size_t c = cap;
	const double clip = use_beta ? scal->clip_limit : 0.;      // pending selective thermostat (explosion heuristics): v, F (and L, M) scaled first
	double vv[3] = {st.v[t], st.v[c + t], st.v[2 * c + t]};
	const double vc = clip > 0. ? clip_factor_trans(clip, cp.mass, vv[0], vv[1], vv[2]) : 1.;
	#pragma unroll
	for (int d = 0; d < 3; d++) {
		double v = vv[d] * vc;
		if (use_beta) v *= bt;
		v += dtInv2m * (F[d * c + t] * vc);
		st.v[d * c + t] = v;
		st.r[d * c + t] += dt * v;
	}
	if (rotates) {
		double q[4], L[3], w[3], qh[4], qi[4];
		for (int k = 0; k < 4; k++) q[k] = st.q[k * c + t];
		for (int d = 0; d < 3; d++) L[d] = st.L[d * c + t];
		double dc = 1.;
		if (clip > 0. && cp.rot_dof > 0) {      // U_rot > 3 rotDOF T: D and M scaled to the limit (Domain.cpp:283-292)
			q_rotateinv(q, L, w);
			double Iw2 = 0.;
			for (int d = 0; d < 3; d++) { w[d] *= cp.invI[d]; Iw2 += cp.I[d] * w[d] * w[d]; }
			const double urot = 0.5 * Iw2, limit_rot = 3.0 * (double) cp.rot_dof * scal->clip_ti;
			if (urot > limit_rot) dc = sqrt(limit_rot / urot);
		}
		for (int d = 0; d < 3; d++) { L[d] *= dc; if (use_beta) L[d] *= br; }
		q_rotateinv(q, L, w);
		for (int d = 0; d < 3; d++) w[d] *= cp.invI[d];
		q_differentiate(q, w, qh);
		for (int k = 0; k < 4; k++) qh[k] = qh[k] * dth + q[k];
		double qc = 1. / sqrt(qh[0] * qh[0] + qh[1] * qh[1] + qh[2] * qh[2] + qh[3] * qh[3]);
		for (int k = 0; k < 4; k++) qh[k] *= qc;
		for (int d = 0; d < 3; d++) L[d] += dth * (M[d * c + t] * dc);
		q_rotateinv(qh, L, w);
		for (int d = 0; d < 3; d++) w[d] *= cp.invI[d];
		q_differentiate(qh, w, qi);
		for (int k = 0; k < 4; k++) q[k] += qi[k] * dt;
		qc = 1. / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
		for (int k = 0; k < 4; k++) st.q[k * c + t] = q[k] * qc;
		for (int d = 0; d < 3; d++) st.L[d * c + t] = L[d];
	}
}

// The same two kernels for ensembles without rotation, two slots per thread with 16-byte loads / stores (the slot planes are
// 16-byte aligned and the capacity is even): more bytes in flight per thread, the streaming kernels get closer to the HBM roofline.
__global__ void __launch_bounds__(256) k_leapfrog_pre_v2(State st, const double* __restrict__ F, uint32_t cap, const CompDev* __restrict__ comps, double dt,
                                                          const DevScalars* __restrict__ scal, uint32_t parity, int use_beta, Region rg) {
	const uint32_t t = 2u * (blockIdx.x * blockDim.x + threadIdx.x);
	const uint32_t n = min(scal->n_slots2[parity], cap);
	if (t >= n || scal->overflow) return;      // overflow: see k_leapfrog_pre
	const uint32_t ns = scal->n_sorted;
	const bool sel0 = region_selected(rg, t, ns), sel1 = (t + 1 < n) && region_selected(rg, t + 1, ns);      // 4 bytes per slot: the other region's slots cost nothing more
	if (!sel0 && !sel1) return;
	const ulonglong2 m2 = *reinterpret_cast<const ulonglong2*>(st.meta + t);
	const bool on0 = sel0 && !meta_halo(m2.x), on1 = sel1 && !meta_halo(m2.y);
	if (!on0 && !on1) return;
	const double bt = use_beta ? scal->beta_trans : 1.;
	const double dth = .5 * dt;
	const double a0 = on0 ? dth / comps[meta_cid(m2.x)].mass : 0., a1 = on1 ? dth / comps[meta_cid(m2.y)].mass : 0.;
	const size_t c = cap;
	double2 v3[3];
	#pragma unroll
	for (int d = 0; d < 3; d++) v3[d] = *reinterpret_cast<const double2*>(st.v + d * c + t);
	const double clip = use_beta ? scal->clip_limit : 0.;      // pending selective thermostat (explosion heuristics): v and F scaled first
	double vc0 = 1., vc1 = 1.;
	if (clip > 0.) {
		if (on0) vc0 = clip_factor_trans(clip, comps[meta_cid(m2.x)].mass, v3[0].x, v3[1].x, v3[2].x);
		if (on1) vc1 = clip_factor_trans(clip, comps[meta_cid(m2.y)].mass, v3[0].y, v3[1].y, v3[2].y);
	}
	#pragma unroll
	for (int d = 0; d < 3; d++) {
		double2 v = v3[d];
		const double2 f = *reinterpret_cast<const double2*>(F + d * c + t);
		double2 r = *reinterpret_cast<const double2*>(st.r + d * c + t);
		if (on0) { v.x *= vc0; if (use_beta) v.x *= bt; v.x += a0 * (f.x * vc0); r.x += dt * v.x; }
		if (on1) { v.y *= vc1; if (use_beta) v.y *= bt; v.y += a1 * (f.y * vc1); r.y += dt * v.y; }
		*reinterpret_cast<double2*>(st.v + d * c + t) = v;
		*reinterpret_cast<double2*>(st.r + d * c + t) = r;
	}
}
__global__ void __launch_bounds__(256) k_leapfrog_post_v2(State st, const double* __restrict__ F, uint32_t cap, const CompDev* __restrict__ comps, double dt,
                                                           const DevScalars* __restrict__ scal, uint32_t parity, double* __restrict__ partials) {
	const uint32_t t = 2u * (blockIdx.x * blockDim.x + threadIdx.x);
	const uint32_t n = min(scal->n_slots2[parity], cap);
	double acc[2] = {0., 0.};
	if (t < n && !scal->overflow) {
		const ulonglong2 m2 = *reinterpret_cast<const ulonglong2*>(st.meta + t);
		const bool on0 = !meta_halo(m2.x), on1 = (t + 1 < n) && !meta_halo(m2.y);
		if (on0 || on1) {
			const double dth = .5 * dt;
			const double mass0 = on0 ? comps[meta_cid(m2.x)].mass : 1., mass1 = on1 ? comps[meta_cid(m2.y)].mass : 1.;
			const double a0 = dth / mass0, a1 = dth / mass1;
			const size_t c = cap;
			double v20 = 0., v21 = 0.;
			#pragma unroll
			for (int d = 0; d < 3; d++) {
				double2 v = *reinterpret_cast<const double2*>(st.v + d * c + t);
				const double2 f = *reinterpret_cast<const double2*>(F + d * c + t);
				if (on0) { v.x += a0 * f.x; v20 += v.x * v.x; }
				if (on1) { v.y += a1 * f.y; v21 += v.y * v.y; }
				*reinterpret_cast<double2*>(st.v + d * c + t) = v;
			}
			// same summation order as one molecule per thread followed by the pairwise tree: (m0 v0^2) + (m1 v1^2)
			acc[0] = (on0 ? mass0 * v20 : 0.) + (on1 ? mass1 * v21 : 0.);
		}
	}
	block_reduce_store<2>(acc, partials, blockIdx.x);
}

// upd_postF + the sums of Leapfrog::transition2to3 (Leapfrog.cpp:113-143)
__global__ void __launch_bounds__(256) k_leapfrog_post(State st, const double* __restrict__ F, const double* __restrict__ M, uint32_t cap,
                                                        const CompDev* __restrict__ comps, double dt, const DevScalars* __restrict__ scal,
                                                        uint32_t parity, int rotates, double* __restrict__ partials) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	double acc[2] = {0., 0.};
	if (t < min(scal->n_slots2[parity], cap) && !scal->overflow) {
		const uint64_t meta = st.meta[t];
		if (!meta_halo(meta)) {
			const CompDev& cp = comps[meta_cid(meta)];
			const double dth = .5 * dt, dtInv2m = dth / cp.mass;
			const size_t c = cap;
			double v2 = 0.;
			#pragma unroll
			for (int d = 0; d < 3; d++) {
				const double v = st.v[d * c + t] + dtInv2m * F[d * c + t];
				st.v[d * c + t] = v;
				v2 += v * v;
			}
			acc[0] = cp.mass * v2;
			if (rotates) {
				double q[4], L[3], w[3];
				for (int k = 0; k < 4; k++) q[k] = st.q[k * c + t];
				for (int d = 0; d < 3; d++) { L[d] = st.L[d * c + t] + dth * M[d * c + t]; st.L[d * c + t] = L[d]; }
				q_rotateinv(q, L, w);
				double Iw2 = 0.;
				for (int d = 0; d < 3; d++) { w[d] *= cp.invI[d]; Iw2 += cp.I[d] * w[d] * w[d]; }
				acc[1] = Iw2;
			}
		}
	}
	block_reduce_store<2>(acc, partials, blockIdx.x);
}

__global__ void __launch_bounds__(1024) k_reduce_kinetic(const double* __restrict__ partials, uint32_t nblocks, DevScalars* scal) {
	double v[2] = {0., 0.};
	for (uint32_t b = threadIdx.x; b < nblocks; b += blockDim.x) { v[0] += partials[(size_t) b * 8]; v[1] += partials[(size_t) b * 8 + 1]; }
	__shared__ double out[8];
	block_reduce_store<2>(v, out, 0);
	__syncthreads();
	if (threadIdx.x == 0) { scal->summv2 = out[0]; scal->sum_iw2 = out[1]; scal->fault = scal->overflow ? 1. : 0.; }      // fault travels with the allreduce of the sums
}

// from_scal: what the next leapfrog_pre of a running ls1b200_simulate would have applied on load -- the pending selective-thermostat
// scaling of the explosion heuristics (v, F and L, M of the molecules over the limit) and the thermostat factors
__global__ void __launch_bounds__(256) k_scale(State st, uint32_t cap, double bt, double br, const DevScalars* __restrict__ scal, uint32_t parity,
                                               int from_scal, int rotates, double* __restrict__ F, double* __restrict__ M, const CompDev* __restrict__ comps,
                                               double clip_arg, double ti_arg) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= min(scal->n_slots2[parity], cap)) return;
	const uint64_t meta = st.meta[t];
	if (meta_halo(meta)) return;
	if (from_scal) { bt = scal->beta_trans; br = scal->beta_rot; }
	const size_t c = cap;
	double vc = 1., dc = 1.;
	const double clip = from_scal ? scal->clip_limit : clip_arg, clip_ti = from_scal ? scal->clip_ti : ti_arg;
	if (clip > 0.) {
		const CompDev& cp = comps[meta_cid(meta)];
		vc = clip_factor_trans(clip, cp.mass, st.v[t], st.v[c + t], st.v[2 * c + t]);
		if (rotates && cp.rot_dof > 0) {
			double q[4], L[3], w[3];
			for (int k = 0; k < 4; k++) q[k] = st.q[k * c + t];
			for (int d = 0; d < 3; d++) L[d] = st.L[d * c + t];
			q_rotateinv(q, L, w);
			double Iw2 = 0.;
			for (int d = 0; d < 3; d++) { w[d] *= cp.invI[d]; Iw2 += cp.I[d] * w[d] * w[d]; }
			const double urot = 0.5 * Iw2, limit_rot = 3.0 * (double) cp.rot_dof * clip_ti;
			if (urot > limit_rot) dc = sqrt(limit_rot / urot);
		}
		if (vc != 1.) for (int d = 0; d < 3; d++) F[d * c + t] *= vc;
		if (dc != 1.) for (int d = 0; d < 3; d++) M[d * c + t] *= dc;
	}
	#pragma unroll
	for (int d = 0; d < 3; d++) st.v[d * c + t] = st.v[d * c + t] * vc * bt;
	if (rotates) for (int d = 0; d < 3; d++) st.L[d * c + t] = st.L[d * c + t] * dc * br;
}

// Domain::calculateGlobalValues + getGlobalPressure for one rank (Domain.cpp:157-314, :363-368, :142-146), incl. the heuristic handling
// of an "explosion": a thermostat factor below MIN_BETA = 0.9 while no error is pending makes the NEXT kernel that loads the velocities
// scale every molecule above 10 T (3 rotDOF T) of translational (rotational) kinetic energy down to that limit (clip_limit), beta_trans
// becomes 1 and beta_rot its 0.0091-th power; the counters decide for how long (still there after 40 steps, or back within 4000: warning;
// a second time: error -- the plain factors apply again). `init`: the evaluation of Simulation::prepare_start (Simulation.cpp:991) --
// the heuristics see the state and start their counters, the factors are not applied.
__global__ void k_globals(DevScalars* scal, double target_T, unsigned long long N, unsigned long long rot_dof, double volume, ls1b200_step_record* rec, int init) {
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	if (scal->fault != 0. && !scal->overflow) scal->overflow = 5;      // another rank failed: every rank freezes in the same step
	const double upot = scal->upot6lj / 6.0 + scal->upot_xpoles + scal->my_rf;
	const double virial = scal->virial + 3.0 * scal->my_rf;
	const double summv2 = scal->summv2;
	const double sumIw2 = rot_dof > 0 ? scal->sum_iw2 : 0.0;
	const double T = (N > 0) ? (summv2 + sumIw2) / (double) (3 * N + rot_dof) : target_T;
	double bt = 1., br = 1.;
	if (target_T > 0. && N > 0) {
		bt = pow(3.0 * (double) N * target_T / summv2, 0.4);
		br = (sumIw2 == 0.0) ? 1.0 : pow((double) rot_dof * target_T / sumIw2, 0.4);
	}
	double clip = 0.;
	if ((bt < 0.9 || br < 0.9) && scal->expl_error <= 0 && !scal->overflow) {
		clip = 10.0 * target_T;                                          // KINLIMIT_PER_T * Ti
		const int reappearance_limit = 4000, vanish_grace_period = 40;
		if (reappearance_limit - scal->expl_counter >= vanish_grace_period) {
			if (scal->expl_warning > 0) scal->expl_error = scal->expl_warning;
			if (scal->expl_counter > 0) scal->expl_warning = scal->expl_counter;
			scal->expl_counter = reappearance_limit;
		}
		bt = 1.0;
		br = pow(br, 0.0091);
	}
	if (scal->expl_counter > 0) scal->expl_counter--;
	if (scal->expl_warning > 0) scal->expl_warning--;
	if (scal->expl_error > 0) scal->expl_error--;
	scal->clip_limit = clip; scal->clip_ti = target_T;
	if (init) { bt = 1.; br = 1.; }
	const double rho = (double) N / volume;
	const double p = T * rho + rho * (virial / (double) N) / 3.;
	scal->beta_trans = bt; scal->beta_rot = br;
	if (rec) { rec->upot = upot; rec->virial = virial; rec->summv2 = summv2; rec->sum_iw2 = sumIw2; rec->temperature = T; rec->pressure = p; rec->beta_trans = bt; rec->beta_rot = br; }
}

int do_forces(ls1b200_ctx* c) {
	if (c->n_owned == 0 && !c->dd.active) { c->forces_valid = true; return 0; }
	LS1_TRY(span_begin(c, PH_FORCE));
	LS1_TRY(c->all_1clj ? (c->lj_kernel == 6 && c->mixed && c->comps.size() == 1 ? force_1clj_pf(c) : c->lj_kernel == 5 ? force_1clj_v3(c) : (c->lj_kernel == 0 || c->lj_kernel == 6) ? force_1clj_tiled(c) : c->lj_kernel == 1 ? force_1clj(c) : c->lj_kernel == 3 ? force_1clj_mma(c) : force_1clj_tc(c)) : (c->multi_kernel != 1 ? force_multi_tiled(c) : force_general(c)));
	LS1_TRY(span_end(c));
	c->forces_valid = true;
	return 0;
}

Region make_region(const ls1b200_ctx* c, int mode) {
	Region rg;
	rg.flag = c->region_flag; rg.mode = (c->region_flag || mode == 3) ? mode : 0;
	return rg;
}

int do_leapfrog_pre(ls1b200_ctx* c, double dt, bool use_beta, int region) {
	if (c->n_owned == 0 && !c->dd.active) return 0;
	LS1_TRY(span_begin(c, PH_PRE));
	const uint32_t g = (uint32_t) ((c->cap + 255) / 256);
	const Region rg = make_region(c, region);
	if (!c->rotates && c->cap % 2 == 0)
		LS1_LAUNCH(c, PH_PRE, k_leapfrog_pre_v2, (g + 1) / 2, 256, 0, c->cur, c->F, (uint32_t) c->cap, c->d_comps, dt, c->d_scal, c->parity, (int) use_beta, rg);
	else
	LS1_LAUNCH(c, PH_PRE, k_leapfrog_pre, g, 256, 0, c->cur, c->F, c->M, (uint32_t) c->cap, c->d_comps, dt, c->d_scal, c->parity, (int) use_beta, (int) c->rotates, rg);
	LS1_TRY(span_end(c));
	c->forces_valid = false;
	return 0;
}

int do_leapfrog_post(ls1b200_ctx* c, double dt) {
	if (c->n_owned == 0 && !c->dd.active) return 0;
	LS1_TRY(span_begin(c, PH_POST));
	const uint32_t g = (uint32_t) ((c->cap + 255) / 256);
	if (g > c->partial_blocks) return fail("leapfrog_post: partials buffer too small");
	if (!c->rotates && c->cap % 2 == 0) {
		LS1_LAUNCH(c, PH_POST, k_leapfrog_post_v2, (g + 1) / 2, 256, 0, c->cur, c->F, (uint32_t) c->cap, c->d_comps, dt, c->d_scal, c->parity, c->partials);
		LS1_LAUNCH(c, PH_POST, k_reduce_kinetic, 1, 1024, 0, c->partials, (g + 1) / 2, c->d_scal);
	} else {
		LS1_LAUNCH(c, PH_POST, k_leapfrog_post, g, 256, 0, c->cur, c->F, c->M, (uint32_t) c->cap, c->d_comps, dt, c->d_scal, c->parity, (int) c->rotates, c->partials);
		LS1_LAUNCH(c, PH_POST, k_reduce_kinetic, 1, 1024, 0, c->partials, g, c->d_scal);
	}
	LS1_TRY(span_end(c));
	return 0;
}

int do_globals(ls1b200_ctx* c, double target_T, uint64_t global_n, double global_volume, ls1b200_step_record* d_rec, bool init) {
	const uint64_t N = global_n ? global_n : c->n_owned;
	double V = global_volume;
	if (!(V > 0.)) V = c->geom.lbox[0] * c->geom.lbox[1] * c->geom.lbox[2];
	const uint64_t rdof = c->dd.active ? c->dd.rot_dof_global : c->rot_dof_total;
	LS1_LAUNCH(c, PH_POST, k_globals, 1, 32, 0, c->d_scal, target_T, (unsigned long long) N, (unsigned long long) rdof, V, d_rec, (int) init);
	return 0;
}

}  // namespace ls1b200

using namespace ls1b200;

extern "C" {

int ls1b200_leapfrog_pre(ls1b200_ctx* c, double dt) {
	if (!c) return fail("ls1b200_leapfrog_pre: null context");
	LS1_CUDA(cudaSetDevice(c->device));
	if (c->split_test && c->dd.active && c->region_flag && c->ever_sorted) {      // LS1B200_SPLIT_TEST: the two launches of the overlapped step, back to back
		LS1_TRY(do_leapfrog_pre(c, dt, false, 1));
		return do_leapfrog_pre(c, dt, false, 2);
	}
	// ensembles without rotation: the binning of the next ls1b200_update rides on this kernel (the new position is in registers).
	// Not on decomposed contexts: binning wraps the coordinates of the dimensions a rank is periodic in by itself, and the staged
	// exchange that follows must see them as the reference's messages carry them -- unwrapped until the receiver bins them
	// (tests/test_gpu_decomposition.py compares the messages record by record; ls1b200_simulate, whose messages nobody sees, wraps first:
	// the receiver ends up with the same bits either way)
	if (c->fused && !c->rotates && !c->dd.active && c->fine_s == 1 && c->components_set && c->cap) {
		LS1_TRY(do_pre_key(c, dt, false, false));
		c->keys_valid = true;
		return 0;
	}
	return do_leapfrog_pre(c, dt, false);
}

int ls1b200_leapfrog_post(ls1b200_ctx* c, double dt, ls1b200_kinetic* out) {
	if (!c) return fail("ls1b200_leapfrog_post: null context");
	LS1_CUDA(cudaSetDevice(c->device));
	LS1_TRY(do_leapfrog_post(c, dt));
	if (out) {
		LS1_CUDA(cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(DevScalars), cudaMemcpyDeviceToHost, c->stream));
		LS1_CUDA(cudaStreamSynchronize(c->stream));
		out->summv2 = c->h_scal->summv2; out->sum_iw2 = c->h_scal->sum_iw2;
		out->n_molecules = c->dd.active ? c->n_owned_local : c->n_owned;
		out->rot_dof = c->dd.active ? c->rot_dof_local : c->rot_dof_total;      // decomposed: over the slots owned NOW (count_owned)
	}
	return 0;
}

int ls1b200_scale_velocities(ls1b200_ctx* c, double beta_trans, double beta_rot) {
	if (!c) return fail("ls1b200_scale_velocities: null context");
	LS1_CUDA(cudaSetDevice(c->device));
	if (c->n_owned == 0 && !c->dd.active) return 0;
	const uint32_t g = (uint32_t) ((c->cap + 255) / 256);
	LS1_LAUNCH(c, PH_POST, k_scale, g, 256, 0, c->cur, (uint32_t) c->cap, beta_trans, beta_rot, c->d_scal, c->parity, 0, (int) c->rotates, c->F, c->M, c->d_comps, 0., 0.);
	return 0;
}

int ls1b200_selective_thermostat(ls1b200_ctx* c, double target_temperature) {
	if (!c) return fail("ls1b200_selective_thermostat: null context");
	if (!(target_temperature > 0.)) return fail("ls1b200_selective_thermostat: target temperature must be positive");
	LS1_CUDA(cudaSetDevice(c->device));
	if (c->n_owned == 0 && !c->dd.active) return 0;
	const uint32_t g = (uint32_t) ((c->cap + 255) / 256);
	LS1_LAUNCH(c, PH_POST, k_scale, g, 256, 0, c->cur, (uint32_t) c->cap, 1., 1., c->d_scal, c->parity, 0, (int) c->rotates, c->F, c->M, c->d_comps, 10.0 * target_temperature, target_temperature);
	return 0;
}

int ls1b200_simulate(ls1b200_ctx* c, uint64_t n_steps, double dt, double target_temperature, uint64_t global_n, double global_volume,
                     ls1b200_step_record* records) {
	if (!c) return fail("ls1b200_simulate: null context");
	if (!c->forces_valid) return fail("ls1b200_simulate: forces are not valid (call ls1b200_update + ls1b200_traverse_cells first)");
	LS1_CUDA(cudaSetDevice(c->device));
	if (n_steps == 0) return 0;
	c->keys_valid = false;
	if (c->dd.active) {
		if (global_n == 0 || !(global_volume > 0.)) return fail("ls1b200_simulate: a decomposed context needs global_n and global_volume");
		if (!c->dd.comm) return fail("ls1b200_simulate: a decomposed context needs use_nccl = 1 (or drive the phases and move the messages yourself)");
		if (!c->dd.rot_dof_valid) {      // rotational degrees of freedom of the whole system (Domain.cpp:220-230), once per upload
			double v = (double) c->rot_dof_total;
			LS1_CUDA(cudaMemcpyAsync(&c->d_scal->summv2, &v, sizeof(double), cudaMemcpyHostToDevice, c->stream));
			double zero[5] = {0., 0., 0., 0., 0.};
			LS1_CUDA(cudaMemcpyAsync(&c->d_scal->upot6lj, zero, 4 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
			LS1_CUDA(cudaMemcpyAsync(&c->d_scal->sum_iw2, zero, sizeof(double), cudaMemcpyHostToDevice, c->stream));
			LS1_TRY(do_allreduce_scalars(c));
			LS1_CUDA(cudaMemcpyAsync(&v, &c->d_scal->summv2, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
			LS1_CUDA(cudaStreamSynchronize(c->stream));
			c->dd.rot_dof_global = (uint64_t) (v + 0.5); c->dd.rot_dof_valid = true;
		}
	}
	if (c->records_cap < n_steps) {
		if (c->d_records) cudaFree(c->d_records);
		c->d_records = nullptr; c->records_cap = 0;
		LS1_CUDA(cudaMalloc((void**) &c->d_records, sizeof(ls1b200_step_record) * n_steps));
		c->records_cap = n_steps;
	}
	// thermostat factors (and the selective-thermostat scaling of the explosion heuristics) of the previous call have already been applied: start from 1
	c->h_scal->beta_trans = 1.; c->h_scal->beta_rot = 1.; c->h_scal->clip_limit = 0.;
	LS1_CUDA(cudaMemcpyAsync(&c->d_scal->beta_trans, &c->h_scal->beta_trans, 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	LS1_CUDA(cudaMemcpyAsync(&c->d_scal->clip_limit, &c->h_scal->clip_limit, sizeof(double), cudaMemcpyHostToDevice, c->stream));
	// The first call after an upload evaluates the global values once on the state as it is, like Simulation::prepare_start
	// (Simulation.cpp:986-992: calculateVelocitySums + calculateGlobalValues): the factors are not applied there, but Domain's explosion
	// heuristics see the initial state and start their counters. The sums come from a "second half kick" of length 0.
	if (!c->globals_initialised && target_temperature > 0.) {
		LS1_TRY(do_leapfrog_post(c, 0.));
		LS1_TRY(do_allreduce_scalars(c));
		LS1_TRY(do_globals(c, target_temperature, global_n, global_volume, nullptr, true));
	}
	c->globals_initialised = true;
	// Multi-GPU: the exchange is overlapped with the work on the INNER slots (Region, common.cuh) -- the analogue of the reference's
	// overlapping scheme (parallel/NonBlockingMPIMultiStepHandler.cpp:28-93: inner cells are processed while the messages travel), at
	// the stages that do not need the imports: boundary slots are integrated first, then the three export / transport / import
	// stages run on their own stream while the main stream integrates the inner slots and builds their part of the cell
	// histogram; the sort proper waits for the imports. (The force traversal cannot start earlier: one cell-sorted slot array.)
	const bool overlap = c->dd.active && c->overlap && c->region_flag && c->ever_sorted && (c->dd.comm || c->peer_ok);
	if ((overlap || (c->dd.active && c->overlap2 && c->fused)) && !c->stream_x) {
		// highest priority: the exchange is a chain of small kernels (scan, publish, wait, append) whose blocks must not queue behind
		// the thousands of blocks of the streaming kernels it runs beside
		int prio_lo = 0, prio_hi = 0;
		LS1_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
		LS1_CUDA(cudaStreamCreateWithPriority(&c->stream_x, cudaStreamNonBlocking, prio_hi));
		LS1_CUDA(cudaEventCreateWithFlags(&c->ev_xa, cudaEventDisableTiming)); LS1_CUDA(cudaEventCreateWithFlags(&c->ev_xb, cudaEventDisableTiming));
	}
	// Fused step (ensembles without rotation, the tiled 1CLJ kernel): the second half kick of a step moves into the first kernel of
	// the NEXT step, which also bins the molecules (k_pre_key), and the kinetic sum the thermostat needs comes out of the force
	// kernel's epilogue -- per step the slot state is read and written once instead of three times, and three launches go away.
	// Same floating-point sequence per molecule; the last step of the call runs the ordinary leapfrog_post.
	bool fused = c->fused && !overlap && !c->rotates && c->all_1clj && c->lj_kernel == 0 && c->fine_s == 1;
	for (int d = 0; d < 3; d++) if (!c->geom.periodic[d] && !(c->dd.active && c->dd.grid[d] > 1)) fused = false;      // open faces: owned molecules may sit in halo cells, where no force kernel visits them
	const bool overlap2 = fused && c->dd.active && c->overlap2 && c->ever_sorted && (c->dd.comm || c->peer_ok);
	if (overlap2 && !c->ev_m) LS1_CUDA(cudaEventCreateWithFlags(&c->ev_m, cudaEventDisableTiming));
	for (uint64_t s = 0; s < n_steps; s++) {
		const bool kin_here = fused && s + 1 < n_steps;      // this step's kinetic sum comes from the force kernel, its kick from the next step
		if (overlap) {
			LS1_TRY(do_leapfrog_pre(c, dt, target_temperature > 0., 1));      // boundary slots: what the exchange reads
			LS1_CUDA(cudaEventRecord(c->ev_xa, c->stream));
			LS1_CUDA(cudaStreamWaitEvent(c->stream_x, c->ev_xa, 0));
			cudaStream_t main_stream = c->stream;
			c->stream = c->stream_x; c->xchg_phase = PH_XCHG;
			const int rc = do_exchange(c);                                      // export -> NVLink / NCCL -> import, three stages, on stream_x
			c->stream = main_stream; c->xchg_phase = PH_UPDATE;
			if (rc) return rc;
			LS1_CUDA(cudaEventRecord(c->ev_xb, c->stream_x));
			LS1_TRY(do_leapfrog_pre(c, dt, target_temperature > 0., 2));      // inner slots, beside the exchange
			LS1_TRY(do_update(c, 1));                                           // ... and their keys / histogram
			LS1_CUDA(cudaStreamWaitEvent(c->stream, c->ev_xb, 0));
			LS1_TRY(do_update(c, 2));                                           // boundary slots + imports, scan, scatter, gather
		} else if (overlap2) {
			// Exchange overlapped with the inner-cell force computation (NonBlockingMPIMultiStepHandler.cpp:28-93). The first kernel has
			// integrated every slot; from here the three export -> NVLink / NCCL -> import stages run on their own stream and append the
			// imports behind the previous order, while the main stream sorts the slots it already has and traverses the chunks no
			// import can touch. The imports get a sort of their own behind the sorted slots; the boundary chunks, which read every cell
			// as two slot ranges, follow.
			LS1_TRY(do_pre_key(c, dt, target_temperature > 0., s > 0));
			LS1_CUDA(cudaEventRecord(c->ev_xa, c->stream));
			LS1_CUDA(cudaStreamWaitEvent(c->stream_x, c->ev_xa, 0));
			cudaStream_t main_stream = c->stream;
			c->stream = c->stream_x; c->xchg_phase = PH_XCHG;
			int rc = do_exchange(c);
			c->stream = main_stream; c->xchg_phase = PH_UPDATE;
			if (rc) return rc;
			LS1_TRY(do_update(c, 4));                                   // records ev_m behind its scan
			LS1_CUDA(cudaStreamWaitEvent(c->stream_x, c->ev_m, 0));
			c->stream = c->stream_x;
			rc = do_update_imports(c);
			c->stream = main_stream;
			if (rc) return rc;
			LS1_CUDA(cudaEventRecord(c->ev_xb, c->stream_x));
		} else if (fused) {
			LS1_TRY(do_pre_key(c, dt, target_temperature > 0., s > 0));
			LS1_TRY(do_exchange(c));
			c->lazy_next = kin_here && c->lazy;      // the next step starts with k_pre_key again: this sort leaves r and v where they are
			const int urc = do_update(c, 3);
			c->lazy_next = false;
			if (urc) return urc;
		} else {
		LS1_TRY(do_leapfrog_pre(c, dt, target_temperature > 0.));
		LS1_TRY(do_exchange(c));                 // multi-GPU: leaving molecules + halo copies over NCCL (no-op on one GPU)
		LS1_TRY(do_update(c));
		}
		c->kin_dth = kin_here ? .5 * dt : 0.;
		int frc;
		if (overlap2) {
			// LS1B200_OV2_PROFILE=1: CUDA-event breakdown of the two passes and of the wait between them (synchronises every step)
			static const bool prof2 = getenv("LS1B200_OV2_PROFILE") != nullptr;
			static cudaEvent_t pe[4]; static double pacc[3] = {0., 0., 0.}; static int pcalls = 0;
			if (prof2 && pcalls == 0) for (int k = 0; k < 4; k++) cudaEventCreate(&pe[k]);
			if (prof2) cudaEventRecord(pe[0], c->stream);
			c->force_pass = 1;
			frc = do_forces(c);                                        // the chunks no import can touch, beside the exchange
			if (prof2) cudaEventRecord(pe[1], c->stream);
			if (!frc) frc = cudaStreamWaitEvent(c->stream, c->ev_xb, 0) == cudaSuccess ? 0 : fail("ls1b200_simulate: cudaStreamWaitEvent failed");
			if (prof2) cudaEventRecord(pe[2], c->stream);
			c->force_pass = 2;
			if (!frc) frc = do_forces(c);                              // the boundary chunks, imports included
			c->force_pass = 0;
			if (prof2) {
				cudaEventRecord(pe[3], c->stream); cudaEventSynchronize(pe[3]);
				++pcalls;
				for (int k = 0; k < 3; k++) { float ms = 0.f; cudaEventElapsedTime(&ms, pe[k], pe[k + 1]); if (pcalls > 5) pacc[k] += ms; }
				if (pcalls > 5 && pcalls % 10 == 0 && c->dd.rank == 0)
					fprintf(stderr, "[ls1b200 overlap] inner pass %.3f ms, wait for the imports %.3f ms, boundary pass %.3f ms\n", pacc[0] / (pcalls - 5), pacc[1] / (pcalls - 5), pacc[2] / (pcalls - 5));
			}
		} else
		frc = do_forces(c);
		c->kin_dth = 0.;
		if (frc) return frc;
		if (!kin_here) LS1_TRY(do_leapfrog_post(c, dt));
		LS1_TRY(do_allreduce_scalars(c));        // multi-GPU: Upot, virial, sum m v^2, sum I w^2 over the ranks
		LS1_TRY(do_globals(c, target_temperature, global_n, global_volume, c->d_records + s));
	}
	if (target_temperature > 0.) {   // the scaling the next leapfrog_pre would have applied on load
		const uint32_t g = (uint32_t) ((c->cap + 255) / 256);
		LS1_LAUNCH(c, PH_POST, k_scale, g, 256, 0, c->cur, (uint32_t) c->cap, 1., 1., c->d_scal, c->parity, 1, (int) c->rotates, c->F, c->M, c->d_comps, 0., 0.);
	}
	if (records) LS1_CUDA(cudaMemcpyAsync(records, c->d_records, sizeof(ls1b200_step_record) * n_steps, cudaMemcpyDeviceToHost, c->stream));
	LS1_CUDA(cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(DevScalars), cudaMemcpyDeviceToHost, c->stream));
	LS1_CUDA(cudaStreamSynchronize(c->stream));
	c->n_slots = std::min<uint64_t>(c->h_scal->n_slots2[c->parity], c->cap);
	if (c->h_scal->overflow) return fail(std::string("ls1b200_simulate: ") + overflow_message(c->h_scal->overflow));
	if (c->dd.active) LS1_TRY(count_owned(c, false));
	return 0;
}

}  // extern "C"
