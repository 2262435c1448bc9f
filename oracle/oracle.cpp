// oracle.cpp -- TEST INFRASTRUCTURE ONLY.  "port" oracle: a scalar CPU restatement of ls1-MarDyn's short-range
// force path, used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER for the CUDA
// path. Nothing in ls1_mardyn_b200/ includes, links or calls this file.
//
// Parity status: PINNED. tests/test_oracle_golden.py checks this restatement against golden vectors produced by
// the unmodified reference built from /root/reference by oracle/build_ref.sh (oracle/_ref/{mixed,double}/refdump and
// MarDyn): the reference's own known-answer fixtures (ForceCalculationTestU0/F0), all 11 Vectorization*.inp
// differential fixtures, periodic-halo dumps, the Comp2Param tables and 100-step U_pot / U_kin / p / beta traces of the reference
// binary on a 1CLJ, a rotating water-like and a rotating multi-potential two-component system (tests/golden/make_traces.py ->
// tests/golden/trace_*.npz; test_trace_vs_reference_binary).
//
// Every function cites the reference file:line (relative to /root/reference/src) it restates.
//
// Build: g++ -O2 -std=c++17 -shared -fPIC -mfma oracle.cpp -o liboracle.so   (see oracle/Makefile)
//   -mfma + std::fma keeps the fused-multiply-add semantics of the AVX2 reference build
//   (adapter/vectorization/RealVecFloat.h:160-168) so that the cutoff decision is bit-identical.
//   -ffp-contract=off is passed so that the compiler does not fuse anything we did not ask for.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct Site { double r[3]; double e[3]; double param; };
struct Component {
	uint32_t n_lj, n_charge, n_dipole, n_quadrupole, site_offset, lj_lookup_id, rot_dof, reserved;
	double mass; double inertia[3];
};
struct Macro { double upot, virial, upot6lj, upot_xpoles, virial_raw, my_rf; };
struct Kinetic { double summv2, sum_iw2; uint64_t n, rot_dof; };
struct StepRecord { double upot, virial, summv2, sum_iw2, temperature, pressure, beta_trans, beta_rot; };
struct FlopCount {
	double mol_dist;
	double k[7];      // lj, charge, charge_dipole, dipole, charge_quadrupole, dipole_quadrupole, quadrupole
	double m[7];      // macroscopic pairs, same order
	double flops;
};

// ---------------------------------------------------------------------------------------------------------
// Geometry: particleContainer/LinkedCells.cpp:47-58 (ctor) and :148-166 (rebuild, XML path) ;
// borders: particleContainer/CellBorderAndFlagManager.h:134-152
// ---------------------------------------------------------------------------------------------------------
struct Geometry {
	double bmin[3], bmax[3], hmin[3], hmax[3], clen[3], crecip[3];
	int box_cells[3], cells[3];  // cells incl. 1 halo layer per side
	void init(const double* mn, const double* mx, double cutoff, int mode_ctor) {
		for (int d = 0; d < 3; d++) {
			bmin[d] = mn[d]; bmax[d] = mx[d];
			double diff = bmax[d] - bmin[d];
			if (mode_ctor) {
				box_cells[d] = (int) std::floor(diff / cutoff);
				if (box_cells[d] == 0) box_cells[d] = 1;
			} else {
				float rc = (float) (cutoff / 1);  // cellsInCutoff == 1
				box_cells[d] = (int) std::floor(diff / rc);  // double / float -> double division, as in the reference
			}
			clen[d] = diff / box_cells[d];
			crecip[d] = box_cells[d] / diff;
			hmin[d] = bmin[d] - clen[d];
			hmax[d] = bmax[d] + clen[d];
			cells[d] = box_cells[d] + 2;
		}
	}
	double border(int i, int d) const {
		if (i == 0) return hmin[d];
		if (i == 1) return bmin[d];
		if (i == cells[d] - 1) return bmax[d];
		if (i == cells[d]) return hmax[d];
		// NB: the reference binary evaluates index*_cellLength+_haloBoundingBoxMin as ONE fused multiply-add (g++ -mfma with
		// the default -ffp-contract=fast; verified in the disassembly of ParticleCellBase::testInBox of oracle/_ref and by
		// the golden fixture periodic_border_fma, whose molecules sit exactly on borders where fma != mul+add).
		return std::fma((double) i, clen[d], hmin[d]);
	}
	// the unique i with border(i) <= x < border(i+1)  (FullParticleCell::testInBox via MoleculeInterface::inBox,
	// molecules/MoleculeInterface.h:299-309); guess as LinkedCells.cpp:899. Returns -1 / cells if outside halo box.
	int cell1d(double x, int d) const {
		int g = (int) std::floor((x - bmin[d]) * crecip[d]) + 1;
		g = std::min(std::max(g, 0), cells[d] - 1);
		while (g > 0 && x < border(g, d)) g--;
		while (g < cells[d] - 1 && x >= border(g + 1, d)) g++;
		if (x < border(0, d)) return -1;
		if (x >= border(cells[d], d)) return cells[d];
		return g;
	}
	long index(int ix, int iy, int iz) const { return ((long) iz * cells[1] + iy) * cells[0] + ix; }
	long ncells() const { return (long) cells[0] * cells[1] * cells[2]; }
	bool is_halo(int ix, int iy, int iz) const {
		return ix == 0 || iy == 0 || iz == 0 || ix == cells[0] - 1 || iy == cells[1] - 1 || iz == cells[2] - 1;
	}
};

// ---------------------------------------------------------------------------------------------------------
// Quaternion rotation: molecules/Quaternion.cpp:43-64 (rotate), :66-87 (rotateinv), :99-104 (differentiate)
// ---------------------------------------------------------------------------------------------------------
static void q_rotate(const double* q, const double* d, double* o) {
	if (d[0] == 0. && d[1] == 0. && d[2] == 0.) { o[0] = o[1] = o[2] = 0.; return; }
	double w = q[0], x = q[1], y = q[2], z = q[3];
	double ww = w * w, xx = x * x, yy = y * y, zz = z * z, wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
	o[0] = (ww + xx - yy - zz) * d[0] + 2. * (xy - wz) * d[1] + 2. * (wy + xz) * d[2];
	o[1] = 2. * (wz + xy) * d[0] + (ww - xx + yy - zz) * d[1] + 2. * (yz - wx) * d[2];
	o[2] = 2. * (xz - wy) * d[0] + 2. * (wx + yz) * d[1] + (ww - xx - yy + zz) * d[2];
}
static void q_rotateinv(const double* q, const double* d, double* o) {
	if (d[0] == 0. && d[1] == 0. && d[2] == 0.) { o[0] = o[1] = o[2] = 0.; return; }
	double w = q[0], x = q[1], y = q[2], z = q[3];
	double ww = w * w, xx = x * x, yy = y * y, zz = z * z, wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
	o[0] = (ww + xx - yy - zz) * d[0] + 2. * (xy + wz) * d[1] + 2. * (xz - wy) * d[2];
	o[1] = 2. * (xy - wz) * d[0] + (ww - xx + yy - zz) * d[1] + 2. * (yz + wx) * d[2];
	o[2] = 2. * (xz + wy) * d[0] + 2. * (yz - wx) * d[1] + (ww - xx - yy + zz) * d[2];
}
static void q_differentiate(const double* q, const double* w, double* dq) {
	dq[0] = .5 * (-q[1] * w[0] - q[2] * w[1] - q[3] * w[2]);
	dq[1] = .5 * (q[0] * w[0] - q[3] * w[1] + q[2] * w[2]);
	dq[2] = .5 * (q[3] * w[0] + q[0] * w[1] - q[1] * w[2]);
	dq[3] = .5 * (-q[2] * w[0] + q[1] * w[1] + q[0] * w[2]);
}

// ---------------------------------------------------------------------------------------------------------
// SoA site record (adapter/CellDataSoA.h:41-74): what FullMolecule::setupSoACache pushes
// (molecules/FullMolecule.cpp:711-767) -- calc-precision copies of COM, site position, orientation.
// ---------------------------------------------------------------------------------------------------------
template <class C> struct SoaSite {
	C mr[3];     // molecule COM (calc)
	C r[3];      // site position (calc)
	C e[3];      // orientation (dipole / quadrupole)
	C p;         // q, |mu|, |Q|
	uint32_t ljid;
	double f[3], V[3], M[3];   // accum
};
template <class C> struct SoaMol {
	C mr[3];
	int first[4], num[4];  // per type: index into cell's per-type site arrays
	long mol;              // index into the molecule arrays (owner for halo copies), -1 never
	bool halo;
};
template <class C> struct Cell {
	std::vector<SoaMol<C>> mols;
	std::vector<SoaSite<C>> sites[4];  // LJ, charge, dipole, quadrupole
	bool halo = false;
	long index = 0;
};

template <class C> static inline C fmadd(C a, C b, C c) { return std::fma(a, b, c); }
template <class C> static inline C fnmadd(C a, C b, C c) { return std::fma(-a, b, c); }  // c - a*b
template <class C> static inline C fmsub(C a, C b, C c) { return std::fma(a, b, -c); }   // a*b - c
// RealCalcVec::scal_prod (RealVecFloat.h:219-223): fmadd(a1,b1, fmadd(a2,b2, a3*b3))
template <class C> static inline C sp(C a1, C a2, C a3, C b1, C b2, C b3) { return fmadd(a1, b1, fmadd(a2, b2, a3 * b3)); }

struct Sums { double upot6lj = 0, upotXpoles = 0, virial = 0, myRF = 0; };

// Pair kernels: adapter/VectorizedCellProcessor.cpp:171-792 (_loopBodyLJ ... _loopBodyQuadrupole).
// Every kernel returns f = force on site A, V = (comA - comB) o f, torques; macroscopic sums are added iff macro.
template <class C> struct PK {
	static void lj(const SoaSite<C>& a, const SoaSite<C>& b, C eps24, C sig2, C shift6, bool macro, Sums& s, C* f, double* V) {
		C dx = a.r[0] - b.r[0], dy = a.r[1] - b.r[1], dz = a.r[2] - b.r[2];
		C r2 = sp(dx, dy, dz, dx, dy, dz);
		C r2inv = C(1) / r2;
		C lj2 = sig2 * r2inv, lj4 = lj2 * lj2, lj6 = lj4 * lj2, lj12 = lj6 * lj6, lj12m6 = lj12 - lj6;
		C scale = (eps24 * r2inv) * (lj12 + lj12m6);
		f[0] = dx * scale; f[1] = dy * scale; f[2] = dz * scale;
		for (int d = 0; d < 3; d++) V[d] = (double) ((a.mr[d] - b.mr[d]) * f[d]);
		if (macro) {
			s.upot6lj += (double) fmadd(eps24, lj12m6, shift6);
			s.virial += V[0] + V[1] + V[2];
		}
	}
	static void charge(const SoaSite<C>& a, const SoaSite<C>& b, bool macro, Sums& s, C* f, double* V) {
		C dx = a.r[0] - b.r[0], dy = a.r[1] - b.r[1], dz = a.r[2] - b.r[2];
		C r2 = sp(dx, dy, dz, dx, dy, dz);
		C r2inv = C(1) / r2, rinv = C(1) / std::sqrt(r2);
		C upot = (a.p * b.p) * rinv;
		C fac = upot * r2inv;
		f[0] = dx * fac; f[1] = dy * fac; f[2] = dz * fac;
		for (int d = 0; d < 3; d++) V[d] = (double) ((a.mr[d] - b.mr[d]) * f[d]);
		if (macro) { s.upotXpoles += (double) upot; s.virial += V[0] + V[1] + V[2]; }
	}
	// a = charge, b = dipole; M = torque on the dipole
	static void charge_dipole(const SoaSite<C>& a, const SoaSite<C>& b, bool macro, Sums& s, C* f, double* V, double* M) {
		C dx = a.r[0] - b.r[0], dy = a.r[1] - b.r[1], dz = a.r[2] - b.r[2];
		C r2 = sp(dx, dy, dz, dx, dy, dz);
		C r2inv = C(1) / r2, rinv = C(1) / std::sqrt(r2);
		C r3inv = r2inv * rinv;
		C re = sp(dx, dy, dz, b.e[0], b.e[1], b.e[2]);
		C qp = a.p * b.p;
		C qpr3 = qp * r3inv;
		C fac = r2inv * C(3) * re;
		f[0] = qpr3 * fnmadd(dx, fac, b.e[0]);
		f[1] = qpr3 * fnmadd(dy, fac, b.e[1]);
		f[2] = qpr3 * fnmadd(dz, fac, b.e[2]);
		for (int d = 0; d < 3; d++) V[d] = (double) ((a.mr[d] - b.mr[d]) * f[d]);
		if (macro) { s.upotXpoles -= (double) (qpr3 * re); s.virial += V[0] + V[1] + V[2]; }
		C exy = fmsub(b.e[0], dy, b.e[1] * dx), eyz = fmsub(b.e[1], dz, b.e[2] * dy), ezx = fmsub(b.e[2], dx, b.e[0] * dz);
		M[0] = (double) (qpr3 * eyz); M[1] = (double) (qpr3 * ezx); M[2] = (double) (qpr3 * exy);
	}
	static void dipole(const SoaSite<C>& a, const SoaSite<C>& b, C epsRFInvrc3, bool macro, Sums& s, C* f, double* V, double* M1, double* M2) {
		C dx = a.r[0] - b.r[0], dy = a.r[1] - b.r[1], dz = a.r[2] - b.r[2];
		C r2 = sp(dx, dy, dz, dx, dy, dz);
		C r2inv = C(1) / r2, rinv = C(1) / std::sqrt(r2);
		C r2three = C(3) * r2inv;
		C p1p2 = a.p * b.p;
		C rffac = p1p2 * epsRFInvrc3;
		C p1p2r3 = p1p2 * rinv * r2inv;
		C p1p2three_r5 = p1p2r3 * r2three;
		const C* e1 = a.e; const C* e2 = b.e;
		C e1e2 = sp(e1[0], e1[1], e1[2], e2[0], e2[1], e2[2]);
		C re1 = sp(dx, dy, dz, e1[0], e1[1], e1[2]);
		C re2 = sp(dx, dy, dz, e2[0], e2[1], e2[2]);
		C re1three = re1 * r2three, re2three = re2 * r2three;
		C re1re2perr2 = r2inv * re1 * re2;
		C t = fnmadd(C(5), re1re2perr2, e1e2);
		f[0] = p1p2three_r5 * sp(dx, e1[0], e2[0], t, re2, re1);
		f[1] = p1p2three_r5 * sp(dy, e1[1], e2[1], t, re2, re1);
		f[2] = p1p2three_r5 * sp(dz, e1[2], e2[2], t, re2, re1);
		for (int d = 0; d < 3; d++) V[d] = (double) ((a.mr[d] - b.mr[d]) * f[d]);
		if (macro) {
			s.upotXpoles += (double) (p1p2r3 * fnmadd(C(3), re1re2perr2, e1e2));
			s.virial += V[0] + V[1] + V[2];
			s.myRF += (double) (rffac * e1e2);
		}
		C c_xy = fmsub(e1[0], e2[1], e1[1] * e2[0]);  // (e1 x e2)_z
		C c_yz = fmsub(e1[1], e2[2], e1[2] * e2[1]);  // (e1 x e2)_x
		C c_zx = fmsub(e1[2], e2[0], e1[0] * e2[2]);  // (e1 x e2)_y
		M1[0] = (double) fmadd(p1p2r3, fmsub(re2three, fmsub(e1[1], dz, e1[2] * dy), c_yz), rffac * c_yz);
		M1[1] = (double) fmadd(p1p2r3, fmsub(re2three, fmsub(e1[2], dx, e1[0] * dz), c_zx), rffac * c_zx);
		M1[2] = (double) fmadd(p1p2r3, fmsub(re2three, fmsub(e1[0], dy, e1[1] * dx), c_xy), rffac * c_xy);
		M2[0] = (double) fmsub(p1p2r3, fmadd(re1three, fmsub(e2[1], dz, e2[2] * dy), c_yz), rffac * c_yz);
		M2[1] = (double) fmsub(p1p2r3, fmadd(re1three, fmsub(e2[2], dx, e2[0] * dz), c_zx), rffac * c_zx);
		M2[2] = (double) fmsub(p1p2r3, fmadd(re1three, fmsub(e2[0], dy, e2[1] * dx), c_xy), rffac * c_xy);
	}
	// a = charge, b = quadrupole; M = torque on the quadrupole
	static void charge_quadrupole(const SoaSite<C>& a, const SoaSite<C>& b, bool macro, Sums& s, C* f, double* V, double* M) {
		C dx = a.r[0] - b.r[0], dy = a.r[1] - b.r[1], dz = a.r[2] - b.r[2];
		C r2 = sp(dx, dy, dz, dx, dy, dz);
		C invdr2 = C(1) / r2, invdr = C(1) / std::sqrt(r2);
		C qQ05 = C(0.5) * a.p * b.p;
		C costj = sp(b.e[0], b.e[1], b.e[2], dx, dy, dz) * invdr;
		C qQinv4dr3 = qQ05 * invdr * invdr2;
		C upot = qQinv4dr3 * (C(3) * costj * costj - C(1));
		C minus_partialRijInvdr = C(3) * upot * invdr2;
		C partialTjInvdr = C(6) * costj * qQinv4dr3 * invdr;
		C fac = fmadd(costj * partialTjInvdr, invdr, minus_partialRijInvdr);
		f[0] = fmsub(fac, dx, partialTjInvdr * b.e[0]);
		f[1] = fmsub(fac, dy, partialTjInvdr * b.e[1]);
		f[2] = fmsub(fac, dz, partialTjInvdr * b.e[2]);
		for (int d = 0; d < 3; d++) V[d] = (double) ((a.mr[d] - b.mr[d]) * f[d]);
		if (macro) { s.upotXpoles += (double) upot; s.virial += V[0] + V[1] + V[2]; }
		M[0] = (double) (partialTjInvdr * fmsub(b.e[2], dy, b.e[1] * dz));
		M[1] = (double) (partialTjInvdr * fmsub(b.e[0], dz, b.e[2] * dx));
		M[2] = (double) (partialTjInvdr * fmsub(b.e[1], dx, b.e[0] * dy));
	}
	// a = dipole, b = quadrupole
	static void dipole_quadrupole(const SoaSite<C>& a, const SoaSite<C>& b, bool macro, Sums& s, C* f, double* V, double* M1, double* M2) {
		C dx = a.r[0] - b.r[0], dy = a.r[1] - b.r[1], dz = a.r[2] - b.r[2];
		C r2 = sp(dx, dy, dz, dx, dy, dz);
		C invdr2 = C(1) / r2, invdr = C(1) / std::sqrt(r2);
		const C* ei = a.e; const C* ej = b.e;
		C myqfac = C(1.5) * a.p * b.p * invdr2 * invdr2;
		C costi = sp(ei[0], ei[1], ei[2], dx, dy, dz) * invdr;
		C costj = sp(ej[0], ej[1], ej[2], dx, dy, dz) * invdr;
		C cos2tj = costj * costj;
		C cosgij = sp(ei[0], ei[1], ei[2], ej[0], ej[1], ej[2]);
		C f5 = fmsub(C(5), cos2tj, C(1));
		C c2 = C(2) * costj;
		C upot = myqfac * (c2 * cosgij - costi * f5);
		C myqfacXinvdr = myqfac * invdr;
		C mpR = C(4) * upot * invdr2;
		C mpTi = myqfacXinvdr * f5;
		C part1 = fmsub(C(5), costi * costj, cosgij);
		C mpTj = myqfacXinvdr * C(2) * part1;
		C pG = myqfac * c2;
		C part3 = fmadd(costi, mpTi, costj * mpTj);
		C fac = fnmadd(part3, invdr, mpR);
		f[0] = sp(fac, mpTi, mpTj, dx, ei[0], ej[0]);
		f[1] = sp(fac, mpTi, mpTj, dy, ei[1], ej[1]);
		f[2] = sp(fac, mpTi, mpTj, dz, ei[2], ej[2]);
		for (int d = 0; d < 3; d++) V[d] = (double) ((a.mr[d] - b.mr[d]) * f[d]);
		if (macro) { s.upotXpoles += (double) upot; s.virial += V[0] + V[1] + V[2]; }
		C cz = fmsub(ei[0], ej[1], ei[1] * ej[0]), cx = fmsub(ei[1], ej[2], ei[2] * ej[1]), cy = fmsub(ei[2], ej[0], ei[0] * ej[2]);
		C gx = pG * cx, gy = pG * cy, gz = pG * cz;
		M1[0] = (double) fmsub(mpTi, fmsub(ei[1], dz, ei[2] * dy), gx);
		M1[1] = (double) fmsub(mpTi, fmsub(ei[2], dx, ei[0] * dz), gy);
		M1[2] = (double) fmsub(mpTi, fmsub(ei[0], dy, ei[1] * dx), gz);
		M2[0] = (double) fmadd(mpTj, fmsub(ej[1], dz, ej[2] * dy), gx);
		M2[1] = (double) fmadd(mpTj, fmsub(ej[2], dx, ej[0] * dz), gy);
		M2[2] = (double) fmadd(mpTj, fmsub(ej[0], dy, ej[1] * dx), gz);
	}
	static void quadrupole(const SoaSite<C>& a, const SoaSite<C>& b, bool macro, Sums& s, C* f, double* V, double* M1, double* M2) {
		C dx = a.r[0] - b.r[0], dy = a.r[1] - b.r[1], dz = a.r[2] - b.r[2];
		C r2 = sp(dx, dy, dz, dx, dy, dz);
		C invdr2 = C(1) / r2, invdr = C(1) / std::sqrt(r2);
		const C* ei = a.e; const C* ej = b.e;
		C qfac = C(0.75) * invdr;
		qfac = qfac * (a.p * b.p);
		qfac = qfac * (invdr2 * invdr2);
		C costi = sp(ei[0], ei[1], ei[2], dx, dy, dz) * invdr;
		C costj = sp(ej[0], ej[1], ej[2], dx, dy, dz) * invdr;
		C cos2ti = costi * costi, cos2tj = costj * costj;
		C cosgij = sp(ei[0], ei[1], ei[2], ej[0], ej[1], ej[2]);
		C term = cosgij - C(5) * (costi * costj);
		C part2 = C(15) * cos2ti * cos2tj;
		C part3 = C(2) * term * term;
		C upot = fmadd(C(5), (cos2ti + cos2tj), part2);
		upot = (C(1) + part3) - upot;
		upot = qfac * upot;
		C mpR = C(5) * upot * invdr2;
		C part1 = qfac * C(10) * invdr;
		part2 = C(2) * term;
		part3 = C(3) * costi * cos2tj;
		C part4 = costi + fmadd(part2, costj, part3);
		C mpTi = part1 * part4;
		part3 = C(3) * costj * cos2ti;
		part4 = costj + fmadd(part2, costi, part3);
		C mpTj = part1 * part4;
		C pG = qfac * C(4) * term;
		C fac = fnmadd((mpTi * costi + mpTj * costj), invdr, mpR);
		f[0] = sp(fac, mpTi, mpTj, dx, ei[0], ej[0]);
		f[1] = sp(fac, mpTi, mpTj, dy, ei[1], ej[1]);
		f[2] = sp(fac, mpTi, mpTj, dz, ei[2], ej[2]);
		for (int d = 0; d < 3; d++) V[d] = (double) ((a.mr[d] - b.mr[d]) * f[d]);
		if (macro) { s.upotXpoles += (double) upot; s.virial += V[0] + V[1] + V[2]; }
		C cz = fmsub(ei[0], ej[1], ei[1] * ej[0]), cx = fmsub(ei[1], ej[2], ei[2] * ej[1]), cy = fmsub(ei[2], ej[0], ei[0] * ej[2]);
		C gx = pG * cx, gy = pG * cy, gz = pG * cz;
		M1[0] = (double) fmsub(mpTi, fmsub(ei[1], dz, ei[2] * dy), gx);
		M1[1] = (double) fmsub(mpTi, fmsub(ei[2], dx, ei[0] * dz), gy);
		M1[2] = (double) fmsub(mpTi, fmsub(ei[0], dy, ei[1] * dx), gz);
		M2[0] = (double) fmadd(mpTj, fmsub(ej[1], dz, ej[2] * dy), gx);
		M2[1] = (double) fmadd(mpTj, fmsub(ej[2], dx, ej[0] * dz), gy);
		M2[2] = (double) fmadd(mpTj, fmsub(ej[0], dy, ej[1] * dx), gz);
	}
};

struct Sim {
	int precision = 0, geom_mode = 0;
	int periodic[3] = {0, 0, 0};
	double cutoff = 0, cutoff_lj = 0, eps_rf = 0;
	Geometry g;
	std::vector<Component> comps;
	std::vector<Site> sites;
	uint32_t n_lj_centers = 0;
	std::vector<double> eps24, sig2, shift6;
	// molecule state (FullMolecule, molecules/FullMolecule.h:330-346)
	std::vector<uint64_t> id; std::vector<uint32_t> cid;
	std::vector<double> r, v, q, L, F, M, Vi;
	std::vector<uint64_t> cell;
	// halo copies of the last update: owner index + shifted position
	std::vector<long> halo_owner; std::vector<double> halo_r; std::vector<uint64_t> halo_cell;
	FlopCount flops{};
	// Domain's "explosion" heuristics (Domain.cpp:74-79, :259-314, :363-368): _universalSelectiveThermostatCounter / Warning / Error,
	// and whether the evaluation of Simulation::prepare_start (Simulation.cpp:991) has happened for the uploaded state
	int expl_counter = 0, expl_warning = 0, expl_error = 0;
	bool globals_initialised = false;
	uint64_t clipped = 0;      // molecules the selective thermostat rescaled so far (diagnostics)
	size_t n() const { return id.size(); }
};

static inline float next_toward_f(float r, float dirv) { return std::nexttoward(r, (long double) dirv); }

// DomainDecompBase::handleDomainLeavingParticles (parallel/DomainDecompBase.cpp:199-249): wrap with clamp.
// vcp_real_calc is float in the mixed build and double in the double build.
template <class C> static double wrap_coord(double x, double mn, double mx) {
	const double Lbox = mx - mn;
	if (x < mn) {           // crossed the lower boundary: positive shift
		x = x + Lbox;
		if (x >= mx) { C rr = (C) mx; x = (double) std::nexttoward(rr, (long double) (rr - C(1))); }
	} else if (x >= mx) {   // crossed the upper boundary: negative shift
		x = x - Lbox;
		if (x <= mn) x = mn;
	}
	return x;
}
// DomainDecompBase::populateHaloLayerWithCopies (parallel/DomainDecompBase.cpp:317-376): shifted copy with clamp.
template <class C> static double halo_coord(double x, double mn, double mx, int dir /* +1: copy goes above max */) {
	const double Lbox = mx - mn;
	if (dir > 0) {
		double y = x + Lbox;
		if (y < mx) { C rr = (C) mx; y = (double) std::nexttoward(rr, (long double) (rr + C(1))); }
		return y;
	} else {
		double y = x - Lbox;
		if (y >= mn) { C rr = (C) mn; y = (double) std::nexttoward(rr, (long double) (rr - C(1))); }
		return y;
	}
}

// LinkedCells::update + DomainDecompBase::exchangeMolecules (Simulation.cpp:1417-1455)
template <class C> static void update(Sim& s) {
	const size_t n = s.n();
	s.g.init(s.g.bmin, s.g.bmax, s.cutoff, s.geom_mode);
	s.cell.assign(n, 0);
	s.halo_owner.clear(); s.halo_r.clear(); s.halo_cell.clear();
	for (size_t i = 0; i < n; i++) {
		double* x = &s.r[3 * i];
		for (int d = 0; d < 3; d++) if (s.periodic[d]) x[d] = wrap_coord<C>(x[d], s.g.bmin[d], s.g.bmax[d]);
		int c[3];
		for (int d = 0; d < 3; d++) c[d] = s.g.cell1d(x[d], d);
		bool out = false;
		for (int d = 0; d < 3; d++) if (c[d] < 0 || c[d] >= s.g.cells[d]) out = true;
		s.cell[i] = out ? (uint64_t) -1 : (uint64_t) s.g.index(c[0], c[1], c[2]);
		// halo copies: per dim the set of shifts {0} u {+L if within rc of the lower face} u {-L if within rc of the upper face}
		int nsh[3]; int sh[3][3];
		for (int d = 0; d < 3; d++) {
			nsh[d] = 0; sh[d][nsh[d]++] = 0;
			if (s.periodic[d]) {
				if (x[d] >= s.g.bmin[d] && x[d] < s.g.bmin[d] + s.cutoff) sh[d][nsh[d]++] = +1;
				if (x[d] >= s.g.bmax[d] - s.cutoff && x[d] < s.g.bmax[d]) sh[d][nsh[d]++] = -1;
			}
		}
		for (int a = 0; a < nsh[0]; a++) for (int b = 0; b < nsh[1]; b++) for (int cc = 0; cc < nsh[2]; cc++) {
			int sd[3] = {sh[0][a], sh[1][b], sh[2][cc]};
			if (!sd[0] && !sd[1] && !sd[2]) continue;
			double y[3]; int hc[3];
			for (int d = 0; d < 3; d++) {
				y[d] = sd[d] ? halo_coord<C>(x[d], s.g.bmin[d], s.g.bmax[d], sd[d]) : x[d];
				hc[d] = s.g.cell1d(y[d], d);
			}
			s.halo_owner.push_back((long) i);
			for (int d = 0; d < 3; d++) s.halo_r.push_back(y[d]);
			s.halo_cell.push_back((uint64_t) s.g.index(hc[0], hc[1], hc[2]));
		}
	}
}

// FullParticleCell::buildSoACaches / FullMolecule::setupSoACache (FullParticleCell.cpp:141-191, FullMolecule.cpp:711-767)
template <class C> static void push_molecule(const Sim& s, Cell<C>& cell, long mol, const double* r, bool halo) {
	const Component& c = s.comps[s.cid[mol]];
	double qn[4];
	{   // normalizeQuaternion
		const double* q = &s.q[4 * mol];
		double m2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
		double sc = 1. / std::sqrt(m2);
		for (int k = 0; k < 4; k++) qn[k] = q[k] * sc;
	}
	SoaMol<C> sm;
	for (int d = 0; d < 3; d++) sm.mr[d] = (C) r[d];
	sm.mol = mol; sm.halo = halo;
	const uint32_t cnt[4] = {c.n_lj, c.n_charge, c.n_dipole, c.n_quadrupole};
	uint32_t so = c.site_offset;
	for (int t = 0; t < 4; t++) {
		sm.first[t] = (int) cell.sites[t].size(); sm.num[t] = (int) cnt[t];
		for (uint32_t j = 0; j < cnt[t]; j++, so++) {
			const Site& st = s.sites[so];
			SoaSite<C> ss{};
			double d3[3]; q_rotate(qn, st.r, d3);
			for (int d = 0; d < 3; d++) { ss.mr[d] = (C) r[d]; ss.r[d] = (C) (d3[d] + r[d]); }
			if (t >= 2) { double e3[3]; q_rotate(qn, st.e, e3); for (int d = 0; d < 3; d++) ss.e[d] = (C) e3[d]; }
			ss.p = (C) st.param;
			ss.ljid = (t == 0) ? c.lj_lookup_id + j : 0;
			cell.sites[t].push_back(ss);
		}
	}
	cell.mols.push_back(sm);
}

// One molecule pair: adapter/VectorizedCellProcessor.cpp:1006-2728 (_calculatePairs body for molecule i vs molecule j),
// sign conventions per SURVEY.md Appendix A.3.
template <class C> static void mol_pair(Sim& s, Cell<C>& c1, const SoaMol<C>& mi, Cell<C>& c2, const SoaMol<C>& mj,
                                        bool single_cell, bool macro, Sums& sums) {
	C dx = mi.mr[0] - mj.mr[0], dy = mi.mr[1] - mj.mr[1], dz = mi.mr[2] - mj.mr[2];
	C r2 = sp(dx, dy, dz, dx, dy, dz);
	s.flops.mol_dist += 1;
	if (single_cell && !(r2 != C(0))) return;  // SingleCellPolicy_::GetForceMask, SIMD_..Helpers.h:367-376
	const C rc2 = (C) (s.cutoff * s.cutoff), ljrc2 = (C) (s.cutoff_lj * s.cutoff_lj);
	const bool in_lj = r2 < ljrc2, in_rc = r2 < rc2;
	const C epsRFInvrc3 = (C) (2. * (s.eps_rf - 1.) / ((s.cutoff * s.cutoff * s.cutoff) * (2. * s.eps_rf + 1.)));
	C f[3]; double V[3], M1[3], M2[3];
	auto addf = [](SoaSite<C>& a, SoaSite<C>& b, const C* f, const double* V, double sign) {
		for (int d = 0; d < 3; d++) { a.f[d] += sign * (double) f[d]; b.f[d] -= sign * (double) f[d]; a.V[d] += V[d]; b.V[d] += V[d]; }
	};
	auto addM = [](SoaSite<C>& a, const double* M) { for (int d = 0; d < 3; d++) a.M[d] += M[d]; };
	auto S1 = [&](int t, int k) -> SoaSite<C>& { return c1.sites[t][mi.first[t] + k]; };
	auto S2 = [&](int t, int k) -> SoaSite<C>& { return c2.sites[t][mj.first[t] + k]; };
	if (in_lj) {
		for (int a = 0; a < mi.num[0]; a++) for (int b = 0; b < mj.num[0]; b++) {
			SoaSite<C>& A = S1(0, a); SoaSite<C>& B = S2(0, b);
			size_t k = (size_t) A.ljid * s.n_lj_centers + B.ljid;
			PK<C>::lj(A, B, (C) s.eps24[k], (C) s.sig2[k], (C) s.shift6[k], macro, sums, f, V);
			addf(A, B, f, V, 1.);
			s.flops.k[0] += 1; if (macro) s.flops.m[0] += 1;
		}
	}
	if (!in_rc) return;
	// charge(i) - charge(j)
	for (int a = 0; a < mi.num[1]; a++) for (int b = 0; b < mj.num[1]; b++) {
		PK<C>::charge(S1(1, a), S2(1, b), macro, sums, f, V); addf(S1(1, a), S2(1, b), f, V, 1.);
		s.flops.k[1] += 1; if (macro) s.flops.m[1] += 1;
	}
	// dipole(i) - charge(j): kernel(charge=j, dipole=i): f on charge j  (.cpp:1363-1397)
	for (int a = 0; a < mi.num[2]; a++) for (int b = 0; b < mj.num[1]; b++) {
		PK<C>::charge_dipole(S2(1, b), S1(2, a), macro, sums, f, V, M1); addf(S1(2, a), S2(1, b), f, V, -1.); addM(S1(2, a), M1);
		s.flops.k[2] += 1; if (macro) s.flops.m[2] += 1;
	}
	// quadrupole(i) - charge(j): kernel(charge=j, quadrupole=i) (.cpp:1525-1560)
	for (int a = 0; a < mi.num[3]; a++) for (int b = 0; b < mj.num[1]; b++) {
		PK<C>::charge_quadrupole(S2(1, b), S1(3, a), macro, sums, f, V, M1); addf(S1(3, a), S2(1, b), f, V, -1.); addM(S1(3, a), M1);
		s.flops.k[4] += 1; if (macro) s.flops.m[4] += 1;
	}
	// dipole(i) - dipole(j)
	for (int a = 0; a < mi.num[2]; a++) for (int b = 0; b < mj.num[2]; b++) {
		PK<C>::dipole(S1(2, a), S2(2, b), epsRFInvrc3, macro, sums, f, V, M1, M2); addf(S1(2, a), S2(2, b), f, V, 1.);
		addM(S1(2, a), M1); addM(S2(2, b), M2);
		s.flops.k[3] += 1; if (macro) s.flops.m[3] += 1;
	}
	// charge(i) - dipole(j)
	for (int a = 0; a < mi.num[1]; a++) for (int b = 0; b < mj.num[2]; b++) {
		PK<C>::charge_dipole(S1(1, a), S2(2, b), macro, sums, f, V, M2); addf(S1(1, a), S2(2, b), f, V, 1.); addM(S2(2, b), M2);
		s.flops.k[2] += 1; if (macro) s.flops.m[2] += 1;
	}
	// quadrupole(i) - dipole(j): kernel(dipole=j, quadrupole=i) (.cpp:2053-2100)
	for (int a = 0; a < mi.num[3]; a++) for (int b = 0; b < mj.num[2]; b++) {
		PK<C>::dipole_quadrupole(S2(2, b), S1(3, a), macro, sums, f, V, M2, M1); addf(S1(3, a), S2(2, b), f, V, -1.);
		addM(S1(3, a), M1); addM(S2(2, b), M2);
		s.flops.k[5] += 1; if (macro) s.flops.m[5] += 1;
	}
	// quadrupole(i) - quadrupole(j)
	for (int a = 0; a < mi.num[3]; a++) for (int b = 0; b < mj.num[3]; b++) {
		PK<C>::quadrupole(S1(3, a), S2(3, b), macro, sums, f, V, M1, M2); addf(S1(3, a), S2(3, b), f, V, 1.);
		addM(S1(3, a), M1); addM(S2(3, b), M2);
		s.flops.k[6] += 1; if (macro) s.flops.m[6] += 1;
	}
	// charge(i) - quadrupole(j)
	for (int a = 0; a < mi.num[1]; a++) for (int b = 0; b < mj.num[3]; b++) {
		PK<C>::charge_quadrupole(S1(1, a), S2(3, b), macro, sums, f, V, M2); addf(S1(1, a), S2(3, b), f, V, 1.); addM(S2(3, b), M2);
		s.flops.k[4] += 1; if (macro) s.flops.m[4] += 1;
	}
	// dipole(i) - quadrupole(j)
	for (int a = 0; a < mi.num[2]; a++) for (int b = 0; b < mj.num[3]; b++) {
		PK<C>::dipole_quadrupole(S1(2, a), S2(3, b), macro, sums, f, V, M1, M2); addf(S1(2, a), S2(3, b), f, V, 1.);
		addM(S1(2, a), M1); addM(S2(3, b), M2);
		s.flops.k[5] += 1; if (macro) s.flops.m[5] += 1;
	}
}

// LinkedCells::traverseCells with the C08 pair set (LinkedCellTraversals/C08BasedTraversals.h:48-99: every unordered
// pair of neighbouring cells exactly once, halo-halo pairs skipped), VectorizedCellProcessor::processCell /
// processCellPair rules (VectorizedCellProcessor.cpp:2732-2819) and FullMolecule::calcFM (FullMolecule.cpp:548-626).
template <class C> static void traverse(Sim& s, Macro* out) {
	const Geometry& g = s.g;
	std::vector<Cell<C>> cells((size_t) g.ncells());
	for (int iz = 0; iz < g.cells[2]; iz++) for (int iy = 0; iy < g.cells[1]; iy++) for (int ix = 0; ix < g.cells[0]; ix++) {
		Cell<C>& c = cells[(size_t) g.index(ix, iy, iz)];
		c.index = g.index(ix, iy, iz); c.halo = g.is_halo(ix, iy, iz);
	}
	const size_t n = s.n();
	for (size_t i = 0; i < n; i++) if (s.cell[i] != (uint64_t) -1) {
		Cell<C>& c = cells[s.cell[i]];
		push_molecule<C>(s, c, (long) i, &s.r[3 * i], c.halo);
	}
	for (size_t h = 0; h < s.halo_owner.size(); h++) push_molecule<C>(s, cells[s.halo_cell[h]], s.halo_owner[h], &s.halo_r[3 * h], true);

	Sums sums;
	memset(&s.flops, 0, sizeof(s.flops));
	for (int iz = 0; iz < g.cells[2]; iz++) for (int iy = 0; iy < g.cells[1]; iy++) for (int ix = 0; ix < g.cells[0]; ix++) {
		Cell<C>& c1 = cells[(size_t) g.index(ix, iy, iz)];
		// processCell: skip halo cells and cells with < 2 molecules (.cpp:2732-2742)
		if (!c1.halo && c1.mols.size() >= 2) {
			for (size_t a = 0; a < c1.mols.size(); a++) for (size_t b = a + 1; b < c1.mols.size(); b++)
				mol_pair<C>(s, c1, c1.mols[a], c1, c1.mols[b], true, true, sums);
		}
		// 13 forward neighbours
		for (int dz = 0; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
			if (dz == 0 && (dy < 0 || (dy == 0 && dx <= 0))) continue;
			int jx = ix + dx, jy = iy + dy, jz = iz + dz;
			if (jx < 0 || jy < 0 || jz < 0 || jx >= g.cells[0] || jy >= g.cells[1] || jz >= g.cells[2]) continue;
			Cell<C>& c2 = cells[(size_t) g.index(jx, jy, jz)];
			if (c1.mols.empty() || c2.mols.empty()) continue;
			if (c1.halo && c2.halo) continue;
			// macroscopic values iff neither is halo, or (exactly one halo) index(c1) < index(c2) in call order; the reference
			// passes the non-halo cell first (C08BasedTraversals.h:83-97) and tests c1.index < c2.index (.cpp:2795-2796)
			Cell<C>* p1 = &c1; Cell<C>* p2 = &c2;
			if (p1->halo) std::swap(p1, p2);   // non-halo first
			bool macro = (!p1->halo && !p2->halo) || (p1->index < p2->index);
			// soa1 = the smaller cell (.cpp:2763)
			Cell<C>* a = p1; Cell<C>* b = p2;
			if (a->mols.size() > b->mols.size()) std::swap(a, b);
			for (auto& mi : a->mols) for (auto& mj : b->mols) mol_pair<C>(s, *a, mi, *b, mj, false, macro, sums);
		}
	}
	// calcFM for every non-halo molecule
	std::fill(s.F.begin(), s.F.end(), 0.); std::fill(s.M.begin(), s.M.end(), 0.); std::fill(s.Vi.begin(), s.Vi.end(), 0.);
	for (auto& c : cells) {
		if (c.halo) continue;
		for (auto& m : c.mols) {
			const Component& comp = s.comps[s.cid[m.mol]];
			double qn[4]; const double* q = &s.q[4 * m.mol];
			double sc = 1. / std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
			for (int k = 0; k < 4; k++) qn[k] = q[k] * sc;
			double F[3] = {0, 0, 0}, M[3] = {0, 0, 0}, Vi[3] = {0, 0, 0}, tM[3] = {0, 0, 0};
			uint32_t so = comp.site_offset;
			for (int t = 0; t < 4; t++) for (int k = 0; k < m.num[t]; k++, so++) {
				const SoaSite<C>& ss = c.sites[t][m.first[t] + k];
				double d3[3]; q_rotate(qn, s.sites[so].r, d3);
				for (int d = 0; d < 3; d++) { F[d] += ss.f[d]; Vi[d] += ss.V[d]; }
				M[0] += d3[1] * ss.f[2] - d3[2] * ss.f[1];
				M[1] += d3[2] * ss.f[0] - d3[0] * ss.f[2];
				M[2] += d3[0] * ss.f[1] - d3[1] * ss.f[0];
				if (t >= 2) for (int d = 0; d < 3; d++) tM[d] += ss.M[d];
			}
			for (int d = 0; d < 3; d++) { s.F[3 * m.mol + d] = F[d]; s.M[3 * m.mol + d] = M[d] + tM[d]; s.Vi[3 * m.mol + d] = 0.5 * Vi[d]; }
		}
	}
	// endTraversal (.cpp:122-155); myRF sign flip at .cpp:2727-2728
	Macro mac;
	mac.upot6lj = sums.upot6lj; mac.upot_xpoles = sums.upotXpoles; mac.virial_raw = sums.virial; mac.my_rf = -sums.myRF;
	mac.virial = mac.virial_raw + 3.0 * mac.my_rf;
	mac.upot = mac.upot6lj / 6.0 + mac.upot_xpoles + mac.my_rf;
	// FlopCounter total (adapter/FlopCounter.cpp:20-74, FlopCounter.h:263-270)
	static const double kk[7] = {15, 10, 34, 101, 52, 121, 131}, mm[7] = {4, 2, 3, 5, 2, 2, 2}, ssum[7] = {12, 12, 15, 18, 15, 18, 18}, smac[7] = {2, 2, 2, 3, 2, 2, 2};
	s.flops.flops = 8 * s.flops.mol_dist;
	for (int t = 0; t < 7; t++) s.flops.flops += (8 + kk[t] + ssum[t]) * s.flops.k[t] + (mm[t] + smac[t]) * s.flops.m[t];
	if (out) *out = mac;
}

// FullMolecule::upd_preF (molecules/FullMolecule.cpp:331-361)
static void leapfrog_pre(Sim& s, double dt) {
	const double dth = .5 * dt;
	for (size_t i = 0; i < s.n(); i++) {
		const Component& c = s.comps[s.cid[i]];
		const double dtInv2m = dth / c.mass;
		double* v = &s.v[3 * i]; double* r = &s.r[3 * i]; double* q = &s.q[4 * i]; double* L = &s.L[3 * i];
		const double* F = &s.F[3 * i]; const double* M = &s.M[3 * i];
		for (int d = 0; d < 3; d++) { v[d] += dtInv2m * F[d]; r[d] += dt * v[d]; }
		double invI[3];
		for (int d = 0; d < 3; d++) invI[d] = (c.inertia[d] != 0.) ? 1. / c.inertia[d] : 0.;
		double w[3]; q_rotateinv(q, L, w);
		for (int d = 0; d < 3; d++) w[d] *= invI[d];
		double qh[4]; q_differentiate(q, w, qh);
		for (int k = 0; k < 4; k++) qh[k] = qh[k] * dth + q[k];
		double qc = 1. / std::sqrt(qh[0] * qh[0] + qh[1] * qh[1] + qh[2] * qh[2] + qh[3] * qh[3]);
		for (int k = 0; k < 4; k++) qh[k] *= qc;
		for (int d = 0; d < 3; d++) L[d] += dth * M[d];
		q_rotateinv(qh, L, w);
		for (int d = 0; d < 3; d++) w[d] *= invI[d];
		double qi[4]; q_differentiate(qh, w, qi);
		for (int k = 0; k < 4; k++) q[k] += qi[k] * dt;
		qc = 1. / std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
		for (int k = 0; k < 4; k++) q[k] *= qc;
	}
}
// FullMolecule::upd_postF (molecules/FullMolecule.cpp:363-386) + Leapfrog::transition2to3 sums (Leapfrog.cpp:113-143)
static void leapfrog_post(Sim& s, double dt, Kinetic* out) {
	const double dth = .5 * dt;
	Kinetic k{0, 0, 0, 0};
	for (size_t i = 0; i < s.n(); i++) {
		const Component& c = s.comps[s.cid[i]];
		const double dtInv2m = dth / c.mass;
		double* v = &s.v[3 * i]; double* L = &s.L[3 * i];
		double v2 = 0;
		for (int d = 0; d < 3; d++) { v[d] += dtInv2m * s.F[3 * i + d]; v2 += v[d] * v[d]; L[d] += dth * s.M[3 * i + d]; }
		k.summv2 += c.mass * v2;
		double w[3]; q_rotateinv(&s.q[4 * i], L, w);
		double Iw2 = 0;
		for (int d = 0; d < 3; d++) { double invI = (c.inertia[d] != 0.) ? 1. / c.inertia[d] : 0.; w[d] *= invI; Iw2 += c.inertia[d] * w[d] * w[d]; }
		k.sum_iw2 += Iw2;
		k.n++; k.rot_dof += c.rot_dof;
	}
	if (out) *out = k;
}

}  // namespace

// =========================================================================================================
// C API (ctypes): mirrors include/ls1b200.h one to one so that the parity tests drive both the same way.
// =========================================================================================================
extern "C" {

struct orc_config { double box_min[3], box_max[3], cutoff, cutoff_lj, epsilon_rf; int32_t precision, geometry, periodic[3], device, reserved[4]; };

void* orc_create(const orc_config* c) {
	Sim* s = new Sim();
	s->precision = c->precision; s->geom_mode = c->geometry;
	for (int d = 0; d < 3; d++) { s->periodic[d] = c->periodic[d]; s->g.bmin[d] = c->box_min[d]; s->g.bmax[d] = c->box_max[d]; }
	s->cutoff = c->cutoff; s->cutoff_lj = c->cutoff_lj; s->eps_rf = c->epsilon_rf;
	s->g.init(c->box_min, c->box_max, c->cutoff, c->geometry);
	return s;
}
void orc_destroy(void* h) { delete (Sim*) h; }

int orc_set_components(void* h, uint32_t nc, const Component* comps, uint32_t ns, const Site* sites,
                       uint32_t nlj, const double* eps24, const double* sig2, const double* shift6) {
	Sim& s = *(Sim*) h;
	s.comps.assign(comps, comps + nc); s.sites.assign(sites, sites + ns); s.n_lj_centers = nlj;
	s.eps24.assign(eps24, eps24 + (size_t) nlj * nlj); s.sig2.assign(sig2, sig2 + (size_t) nlj * nlj); s.shift6.assign(shift6, shift6 + (size_t) nlj * nlj);
	return 0;
}
int orc_upload(void* h, uint64_t n, const uint64_t* id, const uint32_t* cid, const double* r, const double* v, const double* q, const double* L) {
	Sim& s = *(Sim*) h;
	s.id.assign(id, id + n); s.cid.assign(cid, cid + n);
	s.r.assign(r, r + 3 * n); s.v.assign(v, v + 3 * n);
	s.q.assign(4 * n, 0.); s.L.assign(3 * n, 0.);
	for (uint64_t i = 0; i < n; i++) { if (q) for (int k = 0; k < 4; k++) s.q[4 * i + k] = q[4 * i + k]; else s.q[4 * i] = 1.; }
	if (L) s.L.assign(L, L + 3 * n);
	s.F.assign(3 * n, 0.); s.M.assign(3 * n, 0.); s.Vi.assign(3 * n, 0.); s.cell.assign(n, 0);
	s.expl_counter = s.expl_warning = s.expl_error = 0; s.globals_initialised = false; s.clipped = 0;
	return 0;
}
uint64_t orc_clipped(void* h) { return ((Sim*) h)->clipped; }
uint64_t orc_num_molecules(void* h) { return ((Sim*) h)->n(); }
int orc_geometry(void* h, int32_t cpd[3], double clen[3], double hmin[3]) {
	Sim& s = *(Sim*) h;
	for (int d = 0; d < 3; d++) { cpd[d] = s.g.cells[d]; clen[d] = s.g.clen[d]; hmin[d] = s.g.hmin[d]; }
	return 0;
}
int orc_update(void* h) { Sim& s = *(Sim*) h; if (s.precision == 0) update<float>(s); else update<double>(s); return 0; }
int orc_traverse_cells(void* h, Macro* out) { Sim& s = *(Sim*) h; if (s.precision == 0) traverse<float>(s, out); else traverse<double>(s, out); return 0; }
int orc_leapfrog_pre(void* h, double dt) { leapfrog_pre(*(Sim*) h, dt); return 0; }
int orc_leapfrog_post(void* h, double dt, Kinetic* out) { leapfrog_post(*(Sim*) h, dt, out); return 0; }
// VelocityScalingThermostat::apply (thermostats/VelocityScalingThermostat.cpp:68-89)
int orc_scale_velocities(void* h, double bt, double br) {
	Sim& s = *(Sim*) h;
	for (auto& x : s.v) x *= bt;
	for (auto& x : s.L) x *= br;
	return 0;
}
// field ids as enum ls1b200_field
int orc_download(void* h, int field, void* dst, uint64_t n) {
	Sim& s = *(Sim*) h;
	if (n != s.n()) return 1;
	switch (field) {
		case 0: memcpy(dst, s.id.data(), 8 * n); break;
		case 1: memcpy(dst, s.cid.data(), 4 * n); break;
		case 2: memcpy(dst, s.r.data(), 24 * n); break;
		case 3: memcpy(dst, s.v.data(), 24 * n); break;
		case 4: memcpy(dst, s.q.data(), 32 * n); break;
		case 5: memcpy(dst, s.L.data(), 24 * n); break;
		case 6: memcpy(dst, s.F.data(), 24 * n); break;
		case 7: memcpy(dst, s.M.data(), 24 * n); break;
		case 8: memcpy(dst, s.Vi.data(), 24 * n); break;
		case 9: memcpy(dst, s.cell.data(), 8 * n); break;
		default: return 2;
	}
	return 0;
}
// halo copies of the last update: returns count; fills owner index, shifted position, cell index if non-null
uint64_t orc_halo(void* h, int64_t* owner, double* r, uint64_t* cell) {
	Sim& s = *(Sim*) h;
	size_t nh = s.halo_owner.size();
	if (owner) for (size_t i = 0; i < nh; i++) owner[i] = s.halo_owner[i];
	if (r) memcpy(r, s.halo_r.data(), 24 * nh);
	if (cell) memcpy(cell, s.halo_cell.data(), 8 * nh);
	return nh;
}
int orc_count_flops(void* h, FlopCount* out) { *out = ((Sim*) h)->flops; return 0; }

// Domain::calculateGlobalValues for one global thermostat (Domain.cpp:232-314, :363-368): temperature, the velocity-scaling factors,
// and the heuristic handling of an "explosion" -- a factor below MIN_BETA = 0.9 while no error is pending: every molecule whose
// translational (rotational) kinetic energy exceeds 10 T (3 rotDOF T) is scaled down to that limit together with its force (torque),
// beta_trans becomes 1, beta_rot its 0.0091-th power; three counters decide for how long this may go on (explosion still there after
// 40 steps or back within 4000: warning; a second time: error, after which the plain factors apply again). Returns bt, br.
static void global_values(Sim& s, const Kinetic& kin, uint64_t N, double Ttarget, double& bt, double& br) {
	const double sumIw2 = kin.rot_dof > 0 ? kin.sum_iw2 : 0.0;
	bt = 1.; br = 1.;
	if (Ttarget > 0 && N > 0) {
		bt = std::pow(3.0 * N * Ttarget / kin.summv2, 0.4);
		br = (sumIw2 == 0.0) ? 1.0 : std::pow(kin.rot_dof * Ttarget / sumIw2, 0.4);
	}
	if ((bt < 0.9 || br < 0.9) && s.expl_error <= 0) {
		const double limit_energy = 10.0 * Ttarget;
		for (size_t i = 0; i < s.n(); i++) {
			const Component& c = s.comps[s.cid[i]];
			double* v = &s.v[3 * i];
			const double Utrans = 0.5 * c.mass * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
			bool hit = false;
			if (Utrans > limit_energy) {
				const double vcorr = std::sqrt(limit_energy / Utrans);
				for (int d = 0; d < 3; d++) { v[d] *= vcorr; s.F[3 * i + d] *= vcorr; }
				hit = true;
			}
			if (c.rot_dof > 0) {
				const double limit_rot_energy = 3.0 * c.rot_dof * Ttarget;
				double w[3]; q_rotateinv(&s.q[4 * i], &s.L[3 * i], w);
				double Iw2 = 0;
				for (int d = 0; d < 3; d++) { double invI = (c.inertia[d] != 0.) ? 1. / c.inertia[d] : 0.; w[d] *= invI; Iw2 += c.inertia[d] * w[d] * w[d]; }
				const double Urot = 0.5 * Iw2;
				if (Urot > limit_rot_energy) {
					const double Dcorr = std::sqrt(limit_rot_energy / Urot);
					for (int d = 0; d < 3; d++) { s.L[3 * i + d] *= Dcorr; s.M[3 * i + d] *= Dcorr; }
					hit = true;
				}
			}
			if (hit) s.clipped++;
		}
		const int explosionReappearanceLimit = 4000, explosionVanishGracePeriod = 40;
		const int stepsSinceLastExplosion = explosionReappearanceLimit - s.expl_counter;
		if (stepsSinceLastExplosion >= explosionVanishGracePeriod) {
			if (s.expl_warning > 0) s.expl_error = s.expl_warning;
			if (s.expl_counter > 0) s.expl_warning = s.expl_counter;
			s.expl_counter = explosionReappearanceLimit;
		}
		bt = 1.0;
		br = std::pow(br, 0.0091);
	}
	if (s.expl_counter > 0) s.expl_counter--;
	if (s.expl_warning > 0) s.expl_warning--;
	if (s.expl_error > 0) s.expl_error--;
}

// Simulation::simulate hot loop (Simulation.cpp:1076-1261) + Domain::calculateGlobalValues (Domain.cpp:157-314)
// + VelocityScalingThermostat::apply. Forces must be valid on entry. The first call after an upload evaluates the global values
// once on the state as it is, like Simulation::prepare_start (Simulation.cpp:986-992): nothing is scaled by the factors there, but
// the explosion heuristics see the initial state (and start their counters).
int orc_simulate(void* h, uint64_t nsteps, double dt, double Ttarget, uint64_t globalN, double globalV, StepRecord* rec) {
	Sim& s = *(Sim*) h;
	if (!s.globals_initialised && nsteps > 0) {
		Kinetic kin;
		leapfrog_post(s, 0.0, &kin);      // the sums of the state as it is (a kick of length 0)
		double bt, br;
		global_values(s, kin, globalN ? globalN : kin.n, Ttarget, bt, br);
		s.globals_initialised = true;
	}
	for (uint64_t st = 0; st < nsteps; st++) {
		leapfrog_pre(s, dt);
		Macro mac; Kinetic kin;
		if (s.precision == 0) { update<float>(s); traverse<float>(s, &mac); } else { update<double>(s); traverse<double>(s, &mac); }
		leapfrog_post(s, dt, &kin);
		uint64_t N = globalN ? globalN : kin.n;
		double V = globalV;
		if (V <= 0) V = (s.g.bmax[0] - s.g.bmin[0]) * (s.g.bmax[1] - s.g.bmin[1]) * (s.g.bmax[2] - s.g.bmin[2]);
		double sumIw2 = kin.rot_dof > 0 ? kin.sum_iw2 : 0.0;
		double T = (kin.summv2 + sumIw2) / (double) (3 * N + kin.rot_dof);
		double bt = 1., br = 1.;
		global_values(s, kin, N, Ttarget, bt, br);
		double rho = N / V;
		double p = T * rho + rho * (mac.virial / N) / 3.;
		if (rec) rec[st] = StepRecord{mac.upot, mac.virial, kin.summv2, sumIw2, T, p, bt, br};
		if (Ttarget > 0) orc_scale_velocities(h, bt, br);
	}
	return 0;
}

// Comp2Param::initialize, LJ part (molecules/Comp2Param.cpp:11-111), with Component::calculateLJshift
// (molecules/Component.cpp:118-122). eps/sigma/shift flag per LJ centre; xi/eta per component pair (row-major nc*nc,
// only i<j used); outputs C*C tables laid out like VectorizedCellProcessor.cpp:46-81.
int orc_lj_tables(uint32_t nc, const uint32_t* nlj, const double* eps, const double* sigma, const int32_t* do_shift, const double* rc_site,
                  const double* xi, const double* eta, double rc_lj, double* eps24, double* sig2, double* shift6) {
	std::vector<uint32_t> start(nc); uint32_t C = 0;
	for (uint32_t i = 0; i < nc; i++) { start[i] = C; C += nlj[i]; }
	auto shift_of = [&](double e, double s, double rcs) { double s2 = s * s / (rcs * rcs); double s6 = s2 * s2 * s2; return 24.0 * e * (s6 - s6 * s6); };
	for (uint32_t ci = 0; ci < nc; ci++) {
		for (uint32_t a = 0; a < nlj[ci]; a++) for (uint32_t b = 0; b < nlj[ci]; b++) {
			uint32_t ia = start[ci] + a, ib = start[ci] + b;
			double e24 = 24. * std::sqrt(eps[ia] * eps[ib]);
			double s2 = .5 * (sigma[ia] + sigma[ib]); s2 *= s2;
			eps24[(size_t) ia * C + ib] = e24; sig2[(size_t) ia * C + ib] = s2;
			shift6[(size_t) ia * C + ib] = do_shift[ia] ? shift_of(eps[ia], sigma[ia], rc_site[ia]) : 0.0;
		}
		for (uint32_t cj = ci + 1; cj < nc; cj++) {
			double x = xi[ci * nc + cj], et = eta[ci * nc + cj];
			for (uint32_t a = 0; a < nlj[ci]; a++) for (uint32_t b = 0; b < nlj[cj]; b++) {
				uint32_t ia = start[ci] + a, jb = start[cj] + b;
				double e24 = 24. * (x * std::sqrt(eps[ia] * eps[jb]));
				double s2 = 0.5 * (et * (sigma[ia] + sigma[jb])); s2 *= s2;
				double sp2 = s2 / (rc_lj * rc_lj), sp6 = sp2 * sp2 * sp2;
				double sh = e24 * (sp6 - sp6 * sp6);
				eps24[(size_t) ia * C + jb] = e24; sig2[(size_t) ia * C + jb] = s2; shift6[(size_t) ia * C + jb] = sh;
				eps24[(size_t) jb * C + ia] = e24; sig2[(size_t) jb * C + ia] = s2; shift6[(size_t) jb * C + ia] = sh;
			}
		}
	}
	return 0;
}

}  // extern "C"
