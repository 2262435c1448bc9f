// force_multi_tiled.cu -- the production cell-pair kernel for multi-site / multi-component ensembles (LJ, charge, dipole,
// quadrupole sites): == LinkedCells::traverseCells(VectorizedCellProcessor&) + Simulation::updateForces().
//
// Reference: VectorizedCellProcessor::_calculatePairs (adapter/VectorizedCellProcessor.cpp:794-2730: one cutoff test per
// MOLECULE pair on the centres of mass, :1011-1027, then all site pairs of the two molecules by type, signs as SURVEY.md A.3),
// the seven _loopBody* kernels (:171-792; pairs.cuh), processCell / processCellPair halo rules (:2732-2819),
// FullMolecule::calcFM (molecules/FullMolecule.cpp:548-626; k_calcfm in force_general.cu).
//
// Same tiling as force_1clj_tiled.cu (tile.cuh): a chunk of cells of one x-row per CTA pass, its 9 candidate windows of
// molecule records (COM + component id) brought to shared memory by TMA bulk copies, one THREAD PER MOLECULE.
//   phase 1  molecule-distance test (bit-identical FP32 / FP64 sequence, pairs.cuh: molpair_r2) of my molecule against the
//            molecules of its 27 cells from shared memory -> per-thread list of in-range neighbours (shared memory, u16);
//   phase 2  when the list is nearly full and after the last row: for each of MY sites (outer loop, so all lanes of a warp
//            work on the same site type), dense loop over the list, inner loop over the neighbour's sites by type; site data
//            (position + parameter, orientation) of the neighbours is read through L1 (it does not fit a shared-memory tile:
//            32 B per site). FP32 partial sums per (site, list) are folded into the FP64 per-site records f, V, M that
//            k_calcfm turns into F, M, Vi exactly like calcFM.
// Every pair is visited from both of its ends (no scatter, no atomics, deterministic); macroscopic sums get weight 1/2 per
// visit == the reference's "count once" rule for halo pairs (VectorizedCellProcessor.cpp:2795-2817).
// The v1 kernel (force_general.cu: thread per site, evaluation inside the search loop) ran at 10 of 32 active lanes and
// 1.6 % of the FP32 peak on the water-like benchmark (profiles/r01i_general_full.md); it stays selectable with
// LS1B200_MULTI_KERNEL=general.
#include "tile.cuh"

namespace ls1b200 {

template <class C> struct MultiCfg;
// TPB threads (= molecules) per CTA pass, LCAP list entries per molecule; the window row length W = 2^LOGW records is chosen
// per launch from the mean cell population (128 for liquids at rc = 2.5 sigma, 256 for 5 sigma), so that the tile leaves room
// for more resident CTAs when cells are sparsely populated
template <> struct MultiCfg<float> { static constexpr int TPB = 96, LCAP = 64; };
template <> struct MultiCfg<double> { static constexpr int TPB = 64, LCAP = 64; };

template <class C>
struct MArgs {
	const typename Vec4<C>::type* mpos;
	const typename Vec4<C>::type* spos;
	const typename Vec4<C>::type* sori;
	const uint32_t* cell_start;
	const uint32_t* cell_of;
	const CompDev* comps;
	const C* eps24; const C* sig2; const C* shift6;
	double* sf;                   // [cap*smax][9] per-site f, V, M (FP64)
	double* partials;
	const Chunk* chunks;
	const uint32_t* n_chunks;
	uint32_t* work_counter;
	uint32_t cap, smax, n_lj;
	int nx, ny, nz;
	C rc2, ljrc2, eps_rf_invrc3;
	int lj_shorter;               // rcLJ < rc: LJ sites need their own cutoff test
};

template <class C> struct SiteAcc { C f[3], V[3], M[3], u6, uxp, rf; };

template <class C> __device__ __forceinline__ C m_rcp(C x);
template <> __device__ __forceinline__ float m_rcp<float>(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
template <> __device__ __forceinline__ double m_rcp<double>(double x) { return 1.0 / x; }

__device__ __forceinline__ float4 ldg_v4(const float4* p) { return __ldg(p); }
__device__ __forceinline__ double4 ldg_v4(const double4* p) {
	const double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p) + 1);
	return make_double4(a.x, a.y, b.x, b.y);
}

template <class C>
__device__ __forceinline__ SiteC<C> m_load_site(const MArgs<C>& a, size_t idx, bool oriented) {
	SiteC<C> s;
	const typename Vec4<C>::type p = ldg_v4(&a.spos[idx]);
	s.x = p.x; s.y = p.y; s.z = p.z; s.p = p.w;
	s.ex = s.ey = s.ez = C(0);
	if (oriented) { const typename Vec4<C>::type e = ldg_v4(&a.sori[idx]); s.ex = e.x; s.ey = e.y; s.ez = e.z; }
	return s;
}

// All interactions of MY site `me` (type 0 LJ, 1 charge, 2 dipole, 3 quadrupole) with the sites of molecule j.
// (dmx, dmy, dmz) = r_i^COM - r_j^COM enters the virial; argument order / signs of the pair functions: SURVEY.md A.3.
template <class C>
__device__ __forceinline__ void eval_site_vs_molecule(const MArgs<C>& a, const int type, const SiteC<C>& me, const uint32_t my_ljid,
                                                       const C dmx, const C dmy, const C dmz, const bool lj_in_range,
                                                       const size_t sj, const CompDev& cj, SiteAcc<C>& s) {
	C ff[3], M1[3], M2[3], u, rf;
	#define ACC_F(sign, dxm, dym, dzm)                                                         \
		{                                                                                       \
			s.f[0] += (sign) * ff[0]; s.f[1] += (sign) * ff[1]; s.f[2] += (sign) * ff[2];         \
			s.V[0] = fma((dxm), ff[0], s.V[0]); s.V[1] = fma((dym), ff[1], s.V[1]); s.V[2] = fma((dzm), ff[2], s.V[2]); \
		}
	#define ACC_M(Mx) { s.M[0] += Mx[0]; s.M[1] += Mx[1]; s.M[2] += Mx[2]; }
	if (type == 0) {
		if (!lj_in_range) return;
		for (uint32_t b = 0; b < cj.n[0]; b++) {
			const SiteC<C> o = m_load_site<C>(a, sj + b, false);
			const size_t ti = (size_t) my_ljid * a.n_lj + wbits<C>(o.p);
			pair_lj<C>(me, o, a.eps24[ti], a.sig2[ti], a.shift6[ti], ff, u);
			ACC_F(C(1), dmx, dmy, dmz); s.u6 += u;
		}
	} else if (type == 1) {
		uint32_t b = cj.n[0];
		for (uint32_t e = b + cj.n[1]; b < e; b++) {          // charge - charge
			const SiteC<C> o = m_load_site<C>(a, sj + b, false);
			pair_charge<C>(me, o, ff, u);
			ACC_F(C(1), dmx, dmy, dmz); s.uxp += u;
		}
		for (uint32_t e = b + cj.n[2]; b < e; b++) {          // charge(me) - dipole(j)
			const SiteC<C> o = m_load_site<C>(a, sj + b, true);
			pair_charge_dipole<C>(me, o, ff, M2, u);
			ACC_F(C(1), dmx, dmy, dmz); s.uxp += u;
		}
		for (uint32_t e = b + cj.n[3]; b < e; b++) {          // charge(me) - quadrupole(j)
			const SiteC<C> o = m_load_site<C>(a, sj + b, true);
			pair_charge_quadrupole<C>(me, o, ff, M2, u);
			ACC_F(C(1), dmx, dmy, dmz); s.uxp += u;
		}
	} else if (type == 2) {
		uint32_t b = cj.n[0];
		for (uint32_t e = b + cj.n[1]; b < e; b++) {          // charge(j) - dipole(me): f acts on the charge
			const SiteC<C> o = m_load_site<C>(a, sj + b, false);
			pair_charge_dipole<C>(o, me, ff, M1, u);
			ACC_F(C(-1), -dmx, -dmy, -dmz); ACC_M(M1); s.uxp += u;
		}
		for (uint32_t e = b + cj.n[2]; b < e; b++) {          // dipole - dipole
			const SiteC<C> o = m_load_site<C>(a, sj + b, true);
			pair_dipole<C>(me, o, a.eps_rf_invrc3, ff, M1, M2, u, rf);
			ACC_F(C(1), dmx, dmy, dmz); ACC_M(M1); s.uxp += u; s.rf += rf;
		}
		for (uint32_t e = b + cj.n[3]; b < e; b++) {          // dipole(me) - quadrupole(j)
			const SiteC<C> o = m_load_site<C>(a, sj + b, true);
			pair_dipole_quadrupole<C>(me, o, ff, M1, M2, u);
			ACC_F(C(1), dmx, dmy, dmz); ACC_M(M1); s.uxp += u;
		}
	} else {
		uint32_t b = cj.n[0];
		for (uint32_t e = b + cj.n[1]; b < e; b++) {          // charge(j) - quadrupole(me)
			const SiteC<C> o = m_load_site<C>(a, sj + b, false);
			pair_charge_quadrupole<C>(o, me, ff, M1, u);
			ACC_F(C(-1), -dmx, -dmy, -dmz); ACC_M(M1); s.uxp += u;
		}
		for (uint32_t e = b + cj.n[2]; b < e; b++) {          // dipole(j) - quadrupole(me)
			const SiteC<C> o = m_load_site<C>(a, sj + b, true);
			pair_dipole_quadrupole<C>(o, me, ff, M1, M2, u);
			ACC_F(C(-1), -dmx, -dmy, -dmz); ACC_M(M2); s.uxp += u;
		}
		for (uint32_t e = b + cj.n[3]; b < e; b++) {          // quadrupole - quadrupole
			const SiteC<C> o = m_load_site<C>(a, sj + b, true);
			pair_quadrupole<C>(me, o, ff, M1, M2, u);
			ACC_F(C(1), dmx, dmy, dmz); ACC_M(M1); s.uxp += u;
		}
	}
	#undef ACC_F
	#undef ACC_M
}
// out-of-line copy for the slow path (a cell denser than the tile), so that the hot loop's code is not duplicated inline
template <class C>
__device__ __noinline__ void eval_site_vs_molecule_slow(const MArgs<C>& a, int type, const SiteC<C>& me, uint32_t my_ljid, C dmx, C dmy, C dmz,
                                                         bool lj_in_range, size_t sj, const CompDev& cj, SiteAcc<C>& s) {
	eval_site_vs_molecule<C>(a, type, me, my_ljid, dmx, dmy, dmz, lj_in_range, sj, cj, s);
}

template <class C> __device__ __forceinline__ void acc_zero(SiteAcc<C>& s) {
	s.f[0] = s.f[1] = s.f[2] = s.V[0] = s.V[1] = s.V[2] = s.M[0] = s.M[1] = s.M[2] = s.u6 = s.uxp = s.rf = C(0);
}
// folds the FP32 partial sums of one (site, list) into the FP64 per-site record and the thread's macroscopic sums
template <class C>
__device__ __forceinline__ void acc_fold(const SiteAcc<C>& s, double* __restrict__ o, bool first, double* mac) {
	const double v[9] = {(double) s.f[0], (double) s.f[1], (double) s.f[2], (double) s.V[0], (double) s.V[1], (double) s.V[2],
	                     (double) s.M[0], (double) s.M[1], (double) s.M[2]};
	if (first) { for (int k = 0; k < 9; k++) o[k] = v[k]; }
	else { for (int k = 0; k < 9; k++) o[k] += v[k]; }
	mac[0] += 0.5 * (double) s.u6; mac[1] += 0.5 * (double) s.uxp; mac[2] += 0.5 * (v[3] + v[4] + v[5]); mac[3] += 0.5 * (double) s.rf;
}
__device__ __forceinline__ int site_type(const CompDev& c, uint32_t k) {
	return (k < c.n[0]) ? 0 : (k < c.n[0] + c.n[1]) ? 1 : (k < c.n[0] + c.n[1] + c.n[2]) ? 2 : 3;
}

// Site signature of a single-component ensemble known at compile time: the specialised kernel keeps MY sites and the per-site
// force sums in registers and loads every neighbour's sites ONCE per molecule pair (the generic kernel re-reads them for each
// of my sites: 10 instead of 4 site loads per water pair, and with 230 KB of the SM's L1/shared array configured as shared
// memory those loads mostly go to L2 -- profiles/r01k_multi_full_water.md: L1 hit rate 28 %, long-scoreboard bound).
template <int NLJ_, int NCH_, int NDI_, int NQU_> struct Sig {
	static constexpr int NLJ = NLJ_, NCH = NCH_, NDI = NDI_, NQU = NQU_, NS = NLJ_ + NCH_ + NDI_ + NQU_;
	static constexpr bool GENERIC = false;
};
struct SigGeneric { static constexpr int NLJ = 0, NCH = 0, NDI = 0, NQU = 0, NS = 1; static constexpr bool GENERIC = true; };

// STAGE (site signatures without oriented sites, e.g. the water-like 1 LJ + 3 charges): the site records of the 9 window rows are
// brought to shared memory with the molecule records (9 more TMA bulk copies: spos is [slot][NS], a window row is contiguous), so
// phase 2 reads its neighbours' sites with LDS instead of gathering them through L1 / L2 (the kernel was bound by the latency of
// those gathers: long-scoreboard stalls 7.3 per issued instruction, profiles/r01k_multi_full_water.md).
// MINB: resident CTAs per SM the register allocation is tuned for (__launch_bounds__): the kernel is latency bound, registers
// decide its occupancy (5: what the compiler chose unconstrained, 128 registers; 6 -> 96 registers and 7 -> 80 registers spill
// and were measured 25 - 150 % slower, profiles/r02_multi_site.md)
template <class C, class SIG, int LOGW, bool STAGE = false, bool PD2 = false, int MINB = 5>
__global__ void __launch_bounds__(MultiCfg<C>::TPB, MINB) k_force_multi_tiled(MArgs<C> a) {
	typedef typename Vec4<C>::type V4;
	constexpr int TPB = MultiCfg<C>::TPB, W = 1 << LOGW, LCAP = MultiCfg<C>::LCAP;
	constexpr uint32_t VB = sizeof(V4), LSTRIDE = (uint32_t) TPB * 2u;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	V4* win = reinterpret_cast<V4*>(smem_raw);                                         // [9][W]
	constexpr int SNS = STAGE ? SIG::NS : 0;
	V4* swin = win + 9 * W;                                                            // [9][W][NS] site records (STAGE)
	uint16_t* list = reinterpret_cast<uint16_t*>(smem_raw + (size_t) (9 * W) * (1 + SNS) * VB);    // [LCAP+1][TPB]
	uint32_t* cs = reinterpret_cast<uint32_t*>(list + (size_t) (LCAP + 1) * TPB);      // [12][CH_MAXC+3] (3 rows of padding)
	__shared__ uint32_t s_tb[CH_MAXC + 1];
	__shared__ uint32_t s_rowbase[9];
	__shared__ __align__(8) unsigned long long s_bar;
	__shared__ uint32_t s_chunk;
	const uint32_t win_s = smem_u32(win), list_s = smem_u32(list) + threadIdx.x * 2, bar = smem_u32(&s_bar);
	const uint32_t swin_s = smem_u32(swin);
	const uint32_t n_chunks = *a.n_chunks;
	if (threadIdx.x == 0) mbar_init(bar, 1);
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	fence_proxy_async();
	__syncthreads();
	uint32_t parity = 0;
	for (;;) {
		if (threadIdx.x == 0) s_chunk = atomicAdd(a.work_counter, 1u);
		__syncthreads();
		const uint32_t ch = s_chunk;
		if (ch >= n_chunks) break;
		const Chunk cd = a.chunks[ch];
		const int nc = cd.ncells;
		const uint32_t R = cd.cell0 / (uint32_t) a.nx;                 // row index cz*ny + cy
		const int xa = (int) (cd.cell0 % (uint32_t) a.nx);
		double mac[4] = {0., 0., 0., 0.};                              // 6*U_LJ, U_xpoles, virial, myRF
		if (cd.flags & CHUNK_BIG) {
			// ---- slow path: one cell denser than the tile. Global loads, evaluation inside the search loop. ----
			const uint32_t ib = a.cell_start[cd.cell0], ie = a.cell_start[cd.cell0 + 1];
			for (uint32_t t = ib + threadIdx.x; t < ie; t += TPB) {
				const V4 mi = a.mpos[t];
				const CompDev& ci = a.comps[wbits<C>(mi.w) & 0xFFFFu];
				for (uint32_t k = 0; k < ci.nsites; k++) {
					const int type = site_type(ci, k);
					const SiteC<C> me = m_load_site<C>(a, (size_t) t * a.smax + k, type >= 2);
					SiteAcc<C> s; acc_zero(s);
					for (int r = 0; r < 9; r++) {
						const uint32_t row = (uint32_t) ((int) R + (r / 3 - 1) * a.ny + (r % 3 - 1)) * a.nx + xa - 1;
						const uint32_t jb = a.cell_start[row], je = a.cell_start[row + 3];
						for (uint32_t j = jb; j < je; j++) {
							const V4 mj = a.mpos[j];
							const C r2 = molpair_r2<C>(mi.x, mi.y, mi.z, mj.x, mj.y, mj.z);
							if (!(r2 < a.rc2) || (r == 4 && !(r2 != C(0)))) continue;
							eval_site_vs_molecule_slow<C>(a, type, me, wbits<C>(me.p), mi.x - mj.x, mi.y - mj.y, mi.z - mj.z, !a.lj_shorter || r2 < a.ljrc2,
							                              (size_t) j * a.smax, a.comps[wbits<C>(mj.w) & 0xFFFFu], s);
						}
					}
					acc_fold<C>(s, a.sf + ((size_t) t * a.smax + k) * 9, true, mac);
				}
			}
			block_reduce_store<4>(mac, a.partials, ch);
			__syncthreads();
			continue;
		}
		// ---- 1. cell_start cache of the 9 rows (relative to the window start) and the TMA loads ----
		for (int k = threadIdx.x; k < 9 * (nc + 3); k += TPB) {
			const int r = k / (nc + 3), c = k % (nc + 3);
			const uint32_t row = (uint32_t) ((int) R + (r / 3 - 1) * a.ny + (r % 3 - 1)) * a.nx;
			const uint32_t b0 = a.cell_start[row + xa - 1];
			cs[r * (CH_MAXC + 3) + c] = a.cell_start[row + xa - 1 + c] - b0;
			if (c == 0) s_rowbase[r] = b0;
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			uint32_t total = 0;
			for (int r = 0; r < 9; r++) total += cs[r * (CH_MAXC + 3) + nc + 2] * VB * (1 + SNS);
			if (total) {
				fence_proxy_async();      // the previous chunk's generic-proxy reads of the windows precede these async-proxy writes
				mbar_expect_tx(bar, total);
				for (int r = 0; r < 9; r++) {
					const uint32_t len = cs[r * (CH_MAXC + 3) + nc + 2];
					if (len) tma_bulk_g2s(win_s + (uint32_t) (r * W) * VB, a.mpos + s_rowbase[r], len * VB, bar);
					if (STAGE && len) tma_bulk_g2s(swin_s + (uint32_t) (r * W * SNS) * VB, a.spos + (size_t) s_rowbase[r] * SNS, len * SNS * VB, bar);
				}
			}
			uint32_t tb = 0;
			for (int c = 0; c < nc; c++) { s_tb[c] = tb; tb += cs[4 * (CH_MAXC + 3) + c + 2] - cs[4 * (CH_MAXC + 3) + c + 1]; }
			s_tb[nc] = tb;
		}
		__syncthreads();
		const uint32_t* csc = cs + 4 * (CH_MAXC + 3);          // centre row
		int kx = 0;
		while (kx + 1 < nc && s_tb[kx + 1] <= threadIdx.x) kx++;
		const bool have = threadIdx.x < s_tb[nc];
		const uint32_t t0rel = csc[kx + 1] + (threadIdx.x - s_tb[kx]);
		const uint32_t t0 = s_rowbase[4] + t0rel;
		mbar_wait(bar, parity); parity ^= 1u;
		V4 mi = make_c4<C>((C) 1e18, (C) 1e18, (C) 1e18, (C) 0);      // an absent molecule: no candidate is ever in range
		if (have) mi = lds_v4<C>(win_s + (uint32_t) (4 * W + (int) t0rel) * VB);
		uint32_t nsites_i = 0;
		if (have) nsites_i = a.comps[wbits<C>(mi.w) & 0xFFFFu].nsites;
		SiteC<C> my[SIG::NS];
		if constexpr (!SIG::GENERIC) {
			#pragma unroll
			for (int k = 0; k < SIG::NS; k++) {
				my[k].x = my[k].y = my[k].z = (C) 1e18; my[k].p = my[k].ex = my[k].ey = my[k].ez = C(0);
				if (have) my[k] = m_load_site<C>(a, (size_t) t0 * a.smax + k, k >= SIG::NLJ + SIG::NCH);
			}
		}
		const uint32_t wlimit = list_s + (uint32_t) (LCAP - 4) * LSTRIDE;
		uint32_t w = list_s, src = 0, end = 0;
		int r = -1;
		bool first = true;
		for (;;) {
			bool flush = true;
			if (r < 9) {
				// ---- phase 1: molecule-distance tests (strict <; pairs at distance 0 in my own row are the molecule itself) ----
				while (src + 4 * VB <= end && w <= wlimit) {
					const V4 m0 = lds_v4<C>(src), m1 = lds_v4<C>(src + VB), m2 = lds_v4<C>(src + 2 * VB), m3 = lds_v4<C>(src + 3 * VB);
					const C r0 = molpair_r2<C>(mi.x, mi.y, mi.z, m0.x, m0.y, m0.z), r1 = molpair_r2<C>(mi.x, mi.y, mi.z, m1.x, m1.y, m1.z);
					const C r2 = molpair_r2<C>(mi.x, mi.y, mi.z, m2.x, m2.y, m2.z), r3 = molpair_r2<C>(mi.x, mi.y, mi.z, m3.x, m3.y, m3.z);
					if (r0 < a.rc2 && (r != 4 || r0 != C(0))) { sts16(w, src); w += LSTRIDE; }
					if (r1 < a.rc2 && (r != 4 || r1 != C(0))) { sts16(w, src + VB); w += LSTRIDE; }
					if (r2 < a.rc2 && (r != 4 || r2 != C(0))) { sts16(w, src + 2 * VB); w += LSTRIDE; }
					if (r3 < a.rc2 && (r != 4 || r3 != C(0))) { sts16(w, src + 3 * VB); w += LSTRIDE; }
					src += 4 * VB;
				}
				if (w <= wlimit) {
					for (; src < end; src += VB) {
						const V4 mj = lds_v4<C>(src);
						const C rr = molpair_r2<C>(mi.x, mi.y, mi.z, mj.x, mj.y, mj.z);
						if (rr < a.rc2 && (r != 4 || rr != C(0))) { sts16(w, src); w += LSTRIDE; }
					}
					r++;
					if (r < 9) {
						uint32_t kb = 0, ke = 0;
						if (have) { kb = cs[r * (CH_MAXC + 3) + kx]; ke = cs[r * (CH_MAXC + 3) + kx + 3]; }      // cells cx-1 .. cx+1 of this row
						src = win_s + (uint32_t) (r * W + (int) kb) * VB; end = win_s + (uint32_t) (r * W + (int) ke) * VB;
						flush = false;
					}
				}
			}
			if (flush) {
				if constexpr (SIG::GENERIC) {
				// ---- phase 2: my sites (outer) x listed neighbours x their sites ----
					const CompDev& ci = a.comps[wbits<C>(mi.w) & 0xFFFFu];
					for (uint32_t k = 0; k < nsites_i; k++) {
						const int type = site_type(ci, k);
						const SiteC<C> me = m_load_site<C>(a, (size_t) t0 * a.smax + k, type >= 2);
						const uint32_t my_ljid = wbits<C>(me.p);
						SiteAcc<C> s; acc_zero(s);
						#pragma unroll 1
						for (uint32_t e = list_s; e < w; e += LSTRIDE) {
							const uint32_t addr = lds16(e);
							const V4 mj = lds_v4<C>(addr);
							const uint32_t idx = (addr - win_s) / VB, rr = idx / (uint32_t) W;
							const size_t slot = (size_t) s_rowbase[rr] + (idx - rr * (uint32_t) W);
							bool ljin = true;
							if (a.lj_shorter) ljin = molpair_r2<C>(mi.x, mi.y, mi.z, mj.x, mj.y, mj.z) < a.ljrc2;
							eval_site_vs_molecule<C>(a, type, me, my_ljid, mi.x - mj.x, mi.y - mj.y, mi.z - mj.z, ljin, slot * a.smax,
							                         a.comps[wbits<C>(mj.w) & 0xFFFFu], s);
						}
						acc_fold<C>(s, a.sf + ((size_t) t0 * a.smax + k) * 9, first, mac);
					}
				} else {
					// ---- phase 2, specialised: listed neighbours (outer) x all site pairs, everything in registers ----
					constexpr int NLJ = SIG::NLJ, NCH = SIG::NCH, NDI = SIG::NDI, NQU = SIG::NQU, NS = SIG::NS;
					constexpr int C0 = NLJ, D0 = NLJ + NCH, Q0 = NLJ + NCH + NDI;      // first charge / dipole / quadrupole site
					C fa[NS][3], Vv[3] = {C(0), C(0), C(0)}, Mm[3] = {C(0), C(0), C(0)}, u6 = C(0), uxp = C(0), rfs = C(0);
					#pragma unroll
					for (int k = 0; k < NS; k++) fa[k][0] = fa[k][1] = fa[k][2] = C(0);
					C e24[NLJ * NLJ + 1], sg2[NLJ * NLJ + 1], sh6[NLJ * NLJ + 1];
					#pragma unroll
					for (int k = 0; k < NLJ * NLJ; k++) { e24[k] = a.eps24[k]; sg2[k] = a.sig2[k]; sh6[k] = a.shift6[k]; }
					// software pipelined: the site records of the NEXT neighbour are requested (L1 / L2) before the current one is evaluated
					// (PD2: TWO neighbours ahead -- the gathers go to L2 more often than not, one evaluation does not cover their latency)
					V4 mj_n = mi, mj_n2 = mi; SiteC<C> oj_n[NS], oj_n2[NS];
					auto fetch_into = [&](uint32_t e, V4& mj_n, SiteC<C>* oj_n) {
						const uint32_t addr = lds16(e);
						mj_n = lds_v4<C>(addr);
						const uint32_t idx = (addr - win_s) / VB, rr = idx >> LOGW;
						if (STAGE) {
							#pragma unroll
							for (int k = 0; k < NS; k++) {
								const V4 p = lds_v4<C>(swin_s + (idx * (uint32_t) NS + (uint32_t) k) * VB);
								oj_n[k].x = p.x; oj_n[k].y = p.y; oj_n[k].z = p.z; oj_n[k].p = p.w; oj_n[k].ex = oj_n[k].ey = oj_n[k].ez = C(0);
							}
						} else {
						const size_t sj = ((size_t) s_rowbase[rr] + (idx & (uint32_t) (W - 1))) * a.smax;
						#pragma unroll
						for (int k = 0; k < NS; k++) oj_n[k] = m_load_site<C>(a, sj + k, k >= D0);
						}
					};
					if (list_s < w) fetch_into(list_s, mj_n, oj_n);
					if (PD2 && list_s + LSTRIDE < w) fetch_into(list_s + LSTRIDE, mj_n2, oj_n2);
					#pragma unroll 1
					for (uint32_t e = list_s; e < w; e += LSTRIDE) {
						const V4 mj = mj_n;
						SiteC<C> oj[NS];
						#pragma unroll
						for (int k = 0; k < NS; k++) oj[k] = oj_n[k];
						if (PD2) {
							mj_n = mj_n2;
							#pragma unroll
							for (int k = 0; k < NS; k++) oj_n[k] = oj_n2[k];
							if (e + 2 * LSTRIDE < w) fetch_into(e + 2 * LSTRIDE, mj_n2, oj_n2);
						} else if (e + LSTRIDE < w) fetch_into(e + LSTRIDE, mj_n, oj_n);
						const C dmx = mi.x - mj.x, dmy = mi.y - mj.y, dmz = mi.z - mj.z;
						bool ljin = true;
						if (a.lj_shorter) ljin = molpair_r2<C>(mi.x, mi.y, mi.z, mj.x, mj.y, mj.z) < a.ljrc2;
						C ff[3], M1[3], M2[3], u, rf;
						// force on MY site k (sign +1: my site is "A" of the pair function, -1: it is "B"), virial with the COM difference
						#define SP_F(k, sign)                                                                     \
							{                                                                                      \
								fa[k][0] += (sign) * ff[0]; fa[k][1] += (sign) * ff[1]; fa[k][2] += (sign) * ff[2];  \
								Vv[0] = fma((sign) * dmx, ff[0], Vv[0]); Vv[1] = fma((sign) * dmy, ff[1], Vv[1]); Vv[2] = fma((sign) * dmz, ff[2], Vv[2]); \
							}
						#define SP_M(Mx) { Mm[0] += Mx[0]; Mm[1] += Mx[1]; Mm[2] += Mx[2]; }
						if (ljin) {
							#pragma unroll
							for (int p = 0; p < NLJ; p++) {
								#pragma unroll
								for (int q = 0; q < NLJ; q++) { pair_lj<C>(my[p], oj[q], e24[p * NLJ + q], sg2[p * NLJ + q], sh6[p * NLJ + q], ff, u); SP_F(p, C(1)); u6 += u; }
							}
						}
						#pragma unroll
						for (int p = C0; p < D0; p++) {                                   // my charges
							#pragma unroll
							for (int q = C0; q < D0; q++) { pair_charge<C>(my[p], oj[q], ff, u); SP_F(p, C(1)); uxp += u; }
							#pragma unroll
							for (int q = D0; q < Q0; q++) { pair_charge_dipole<C>(my[p], oj[q], ff, M2, u); SP_F(p, C(1)); uxp += u; }
							#pragma unroll
							for (int q = Q0; q < NS; q++) { pair_charge_quadrupole<C>(my[p], oj[q], ff, M2, u); SP_F(p, C(1)); uxp += u; }
						}
						#pragma unroll
						for (int p = D0; p < Q0; p++) {                                   // my dipoles
							#pragma unroll
							for (int q = C0; q < D0; q++) { pair_charge_dipole<C>(oj[q], my[p], ff, M1, u); SP_F(p, C(-1)); SP_M(M1); uxp += u; }
							#pragma unroll
							for (int q = D0; q < Q0; q++) { pair_dipole<C>(my[p], oj[q], a.eps_rf_invrc3, ff, M1, M2, u, rf); SP_F(p, C(1)); SP_M(M1); uxp += u; rfs += rf; }
							#pragma unroll
							for (int q = Q0; q < NS; q++) { pair_dipole_quadrupole<C>(my[p], oj[q], ff, M1, M2, u); SP_F(p, C(1)); SP_M(M1); uxp += u; }
						}
						#pragma unroll
						for (int p = Q0; p < NS; p++) {                                   // my quadrupoles
							#pragma unroll
							for (int q = C0; q < D0; q++) { pair_charge_quadrupole<C>(oj[q], my[p], ff, M1, u); SP_F(p, C(-1)); SP_M(M1); uxp += u; }
							#pragma unroll
							for (int q = D0; q < Q0; q++) { pair_dipole_quadrupole<C>(oj[q], my[p], ff, M1, M2, u); SP_F(p, C(-1)); SP_M(M2); uxp += u; }
							#pragma unroll
							for (int q = Q0; q < NS; q++) { pair_quadrupole<C>(my[p], oj[q], ff, M1, M2, u); SP_F(p, C(1)); SP_M(M1); uxp += u; }
						}
						#undef SP_F
						#undef SP_M
					}
					// fold into the FP64 per-site records: f per site; the molecule's virial rides on site 0, the sum of the site
					// torques on the first oriented site (k_calcfm sums V over all sites and M over the oriented ones)
					if (have) {
						#pragma unroll
						for (int k = 0; k < NS; k++) {
							double* o = a.sf + ((size_t) t0 * a.smax + k) * 9;
							const double v[9] = {(double) fa[k][0], (double) fa[k][1], (double) fa[k][2],
							                     k == 0 ? (double) Vv[0] : 0., k == 0 ? (double) Vv[1] : 0., k == 0 ? (double) Vv[2] : 0.,
							                     k == D0 ? (double) Mm[0] : 0., k == D0 ? (double) Mm[1] : 0., k == D0 ? (double) Mm[2] : 0.};
							if (first) { for (int c = 0; c < 9; c++) o[c] = v[c]; }
							else { for (int c = 0; c < 9; c++) o[c] += v[c]; }
						}
						mac[0] += 0.5 * (double) u6; mac[1] += 0.5 * (double) uxp;
						mac[2] += 0.5 * ((double) Vv[0] + (double) Vv[1] + (double) Vv[2]); mac[3] += 0.5 * (double) rfs;
					}
				}
				first = false;
				w = list_s;
				if (r >= 9) break;
			}
		}
		block_reduce_store<4>(mac, a.partials, ch);
		__syncthreads();      // windows, lists and cs are reused by the next chunk
	}
}

// deterministic two-level reduction of the per-chunk partial sums
constexpr int MRED_BLOCKS = 128;
__global__ void __launch_bounds__(256) k_reduce_multi1(const double* __restrict__ partials, const uint32_t* __restrict__ n_chunks, double* __restrict__ partials2) {
	double v[4] = {0., 0., 0., 0.};
	const uint32_t n = *n_chunks;
	for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x) {
		const double4 p = *reinterpret_cast<const double4*>(partials + (size_t) b * 8);
		v[0] += p.x; v[1] += p.y; v[2] += p.z; v[3] += p.w;
	}
	block_reduce_store<4>(v, partials2, blockIdx.x);
}
__global__ void __launch_bounds__(MRED_BLOCKS) k_reduce_multi2(const double* __restrict__ partials2, DevScalars* scal) {
	double v[4];
	for (int k = 0; k < 4; k++) v[k] = partials2[(size_t) threadIdx.x * 8 + k];
	__shared__ double out[8];
	block_reduce_store<4>(v, out, 0);
	__syncthreads();
	if (threadIdx.x == 0) {
		scal->upot6lj = out[0]; scal->upot_xpoles = out[1]; scal->virial = out[2];
		scal->my_rf = -out[3];   // sign flip of VectorizedCellProcessor.cpp:2727-2728
	}
}

int launch_calcfm(ls1b200_ctx* c);      // force_general.cu

template <class C, class SIG, int LOGW, bool STAGE = false, bool PD2 = false, int MINB = 5>
static int launch_multi(ls1b200_ctx* c) {
	typedef typename Vec4<C>::type V4;
	constexpr int TPB = MultiCfg<C>::TPB, W = 1 << LOGW, LCAP = MultiCfg<C>::LCAP;
	const Geom& g = c->geom;
	uint32_t nrows = 0;
	LS1_TRY(build_chunks(c, TPB, W, 1, &nrows));
	MArgs<C> a;
	a.mpos = (const V4*) c->mpos; a.spos = (const V4*) c->spos; a.sori = (const V4*) c->sori;
	a.cell_start = c->cell_start; a.cell_of = c->cell_of2; a.comps = c->d_comps;
	if (sizeof(C) == 4) { a.eps24 = (const C*) c->d_eps24f; a.sig2 = (const C*) c->d_sig2f; a.shift6 = (const C*) c->d_shift6f; }
	else { a.eps24 = (const C*) c->d_eps24; a.sig2 = (const C*) c->d_sig2; a.shift6 = (const C*) c->d_shift6; }
	a.sf = c->sf; a.partials = c->partials;
	a.chunks = (const Chunk*) c->chunks; a.n_chunks = c->chunk_row_start + nrows; a.work_counter = c->work_counter;
	a.cap = (uint32_t) c->cap; a.smax = c->smax; a.n_lj = c->n_lj; a.nx = g.cells[0]; a.ny = g.cells[1]; a.nz = g.cells[2];
	const double rc = c->cfg.cutoff, rclj = c->cfg.cutoff_lj, erf = c->cfg.epsilon_rf;
	a.rc2 = (C) (rc * rc); a.ljrc2 = (C) (rclj * rclj);
	a.lj_shorter = rclj < rc;
	a.eps_rf_invrc3 = (C) (2. * (erf - 1.) / ((rc * rc * rc) * (2. * erf + 1.)));   // VectorizedCellProcessor.cpp:22
	if ((uint64_t) c->geom.ncells + 64 > c->partial_blocks) return fail("force_multi_tiled: partials buffer too small");
	const size_t smem = (size_t) (9 * W) * (1 + (STAGE ? SIG::NS : 0)) * sizeof(V4) + (size_t) (LCAP + 1) * TPB * 2 + 12 * (CH_MAXC + 3) * 4;
	auto kern = k_force_multi_tiled<C, SIG, LOGW, STAGE, PD2, MINB>;
	LS1_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	int per_sm = 0;
	LS1_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TPB, smem));
	if (per_sm < 1) return fail("force_multi_tiled: kernel does not fit on an SM");
	// Neighbour site records are gathered through L1 (they do not fit the shared-memory tile): the kernel is latency bound on those
	// gathers, and every CTA's tile is carved out of the same 228 KB as L1. LS1B200_MULTI_MAXCTAS caps the resident CTAs per SM and
	// asks for a matching carve-out, leaving the rest to L1 (measured: profiles/r02_multi_site.md).
	if (c->multi_max_ctas > 0 && c->multi_max_ctas < per_sm) {
		per_sm = c->multi_max_ctas;
		const int pct = (int) std::min<size_t>(100, ((size_t) per_sm * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
		LS1_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
	}
	LS1_LAUNCH(c, PH_FORCE, kern, (uint32_t) (c->sm_count * per_sm), TPB, smem, a);
	LS1_TRY(launch_calcfm(c));
	LS1_LAUNCH(c, PH_FORCE, k_reduce_multi1, MRED_BLOCKS, 256, 0, c->partials, c->chunk_row_start + nrows, c->partials2);
	LS1_LAUNCH(c, PH_FORCE, k_reduce_multi2, 1, MRED_BLOCKS, 0, c->partials2, c->d_scal);
	return 0;
}

template <class SIG> static int launch_sig(ls1b200_ctx* c) {
	// mean population of a cell (halo layer included in both counts): short windows while a chunk's 9 candidate rows fit
	const double nbar = (double) c->n_slots / (double) c->geom.ncells;
	// W = 128 suffices while a chunk of about two thirds of the CTA's threads plus its two flanking cells fits (with 15 % head room)
	const bool narrow = 0.67 * 1.15 * (c->mixed ? 96. : 64.) + 2.3 * nbar <= 128.;
	// site records staged in shared memory: signatures without oriented sites, narrow windows (9 x 128 x (1 + NS) x 16 B), mixed build
	if constexpr (!SIG::GENERIC && SIG::NDI == 0 && SIG::NQU == 0 && SIG::NS <= 4) {
		if (c->mixed && narrow && c->multi_stage && c->smax == (uint32_t) SIG::NS) return launch_multi<float, SIG, 7, true>(c);
	}
	if constexpr (!SIG::GENERIC) {
		if (c->mixed && c->multi_pd2) return narrow ? launch_multi<float, SIG, 7, false, true>(c) : launch_multi<float, SIG, 8, false, true>(c);
	}
	if (c->mixed) return narrow ? launch_multi<float, SIG, 7>(c) : launch_multi<float, SIG, 8>(c);
	return narrow ? launch_multi<double, SIG, 7>(c) : launch_multi<double, SIG, 8>(c);
}

int force_multi_tiled(ls1b200_ctx* c) {
	// single-component ensembles with a known site signature take the register-resident specialisation
	if (c->comps.size() == 1 && c->multi_kernel == 0) {
		const CompDev& k = c->comps[0];
		const uint32_t sig = k.n[0] * 1000u + k.n[1] * 100u + k.n[2] * 10u + k.n[3];
		switch (sig) {
			case 1300: return launch_sig<Sig<1, 3, 0, 0>>(c);      // water-like: 1 LJ + 3 charges (test_input/VectorizationWater.inp)
			case 2001: return launch_sig<Sig<2, 0, 0, 1>>(c);      // 2CLJQ (examples/Adsorption/CO2/components_co2_stoll.xml)
			case 2010: return launch_sig<Sig<2, 0, 1, 0>>(c);      // 2CLJD
			case 1010: return launch_sig<Sig<1, 0, 1, 0>>(c);      // Stockmayer: 1 LJ + dipole
			case 2000: return launch_sig<Sig<2, 0, 0, 0>>(c);      // 2CLJ
			case 3300: return launch_sig<Sig<3, 3, 0, 0>>(c);      // 3 LJ + 3 charges (examples/CO2/components.xml)
			default: break;
		}
	}
	return launch_sig<SigGeneric>(c);
}

}  // namespace ls1b200
