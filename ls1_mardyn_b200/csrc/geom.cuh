// geom.cuh -- device-side LinkedCells geometry: bit-exact cell assignment and periodic wrap / halo-copy coordinates.
// Reference: particleContainer/CellBorderAndFlagManager.h:134-152 (getBorder), particleContainer/LinkedCells.cpp:870-923
// (guess + correction), molecules/MoleculeInterface.h:299-309 (inBox: min <= r < max),
// parallel/DomainDecompBase.cpp:199-249 (wrap), :317-376 (halo copies).
#pragma once
#include "common.cuh"

namespace ls1b200 {

// Border i of dimension d. The reference binary evaluates the generic case as ONE fused multiply-add (g++ -mfma contracts
// index*_cellLength + _haloBoundingBoxMin); tests/golden/periodic_border_fma.npz pins that.
template <int D>
__device__ __forceinline__ double cell_border(const Geom& g, int i) {
	if (i == 0) return g.hmin[D];
	if (i == 1) return g.bmin[D];
	if (i == g.cells[D] - 1) return g.bmax[D];
	if (i == g.cells[D]) return g.hmax[D];
	return fma((double) i, g.clen[D], g.hmin[D]);
}

// The unique i with border(i) <= x < border(i+1); -1 / cells[D] if x lies outside the halo box.
template <int D>
__device__ __forceinline__ int cell1d(const Geom& g, double x) {
	int c = (int) floor(__dmul_rn(__dsub_rn(x, g.bmin[D]), g.crecip[D])) + 1;
	c = min(max(c, 0), g.cells[D] - 1);
	while (c > 0 && x < cell_border<D>(g, c)) c--;
	while (c < g.cells[D] - 1 && x >= cell_border<D>(g, c + 1)) c++;
	if (x < g.hmin[D]) return -1;
	if (x >= g.hmax[D]) return g.cells[D];
	return c;
}
// The same with the borders read from a table (`tab[i]` == cell_border<D>(g, i), filled once by k_borders): the binning
// kernel is bound by FP64 conversions (int -> double per border evaluation), not by HBM -- profiles/r01u_hbm_kernels.md.
template <int D>
__device__ __forceinline__ int cell1d_tab(const Geom& g, const double* __restrict__ tab, double x) {
	int c = __double2int_rd(__dmul_rn(__dsub_rn(x, g.bmin[D]), g.crecip[D])) + 1;
	c = min(max(c, 0), g.cells[D] - 1);
	while (c > 0 && x < __ldg(tab + c)) c--;
	while (c < g.cells[D] - 1 && x >= __ldg(tab + c + 1)) c++;
	if (x < g.hmin[D]) return -1;
	if (x >= g.hmax[D]) return g.cells[D];
	return c;
}

// Fine x-bin of a coordinate inside ITS reference cell (lower border `border` from the same table the cell assignment uses):
// the counting sort orders the molecules of a cell by this key, so a row of cells is (coarsely) sorted by x and the force
// kernels can trim their candidate ranges to [x - rc, x + rc]. Not part of the reference's semantics -- the reference cell
// (and everything reported as CELL) is unchanged; the SAME FP64 expression is evaluated by the sort and by the force kernel,
// which makes the trim rigorous (floating-point operations are monotone).
__device__ __forceinline__ int fine_sub(double x, double border, double finv, int S) {
	const int s = __double2int_rd(__dmul_rn(__dsub_rn(x, border), finv));
	return min(max(s, 0), S - 1);
}

__device__ __forceinline__ uint32_t cell_index(const Geom& g, int cx, int cy, int cz) {
	return (uint32_t) ((cz * g.cells[1] + cy) * g.cells[0] + cx);
}

__device__ __forceinline__ bool is_halo_cell(const Geom& g, uint32_t c) {
	const int nx = g.cells[0], ny = g.cells[1], nz = g.cells[2];
	const int cx = c % nx, cy = (c / nx) % ny, cz = c / (nx * ny);
	return cx == 0 || cy == 0 || cz == 0 || cx == nx - 1 || cy == ny - 1 || cz == nz - 1;
}

// DomainDecompBase::handleDomainLeavingParticles: wrap x into [bmin, bmax) with the reference's clamps.
template <int D>
__device__ __forceinline__ double wrap_coord(const Geom& g, double x) {
	if (x < g.bmin[D]) {
		x = __dadd_rn(x, g.lbox[D]);
		if (x >= g.bmax[D]) x = g.wrap_hi[D];
	} else if (x >= g.bmax[D]) {
		x = __dsub_rn(x, g.lbox[D]);
		if (x <= g.bmin[D]) x = g.bmin[D];
	}
	return x;
}

// DomainDecompBase::populateHaloLayerWithCopies: coordinate of the copy shifted by +L (dir 1) or -L (dir 2).
template <int D>
__device__ __forceinline__ double halo_coord(const Geom& g, double x, int dir) {
	if (dir == 1) {
		double y = __dadd_rn(x, g.lbox[D]);
		if (y < g.bmax[D]) y = g.halo_hi[D];
		return y;
	} else if (dir == 2) {
		double y = __dsub_rn(x, g.lbox[D]);
		if (y >= g.bmin[D]) y = g.halo_lo[D];
		return y;
	}
	return x;
}

// which periodic copies does a molecule at x need in dimension D: bit0 -> +L copy (near the lower face), bit1 -> -L copy
template <int D>
__device__ __forceinline__ uint32_t shift_mask(const Geom& g, double x) {
	uint32_t m = 0;
	if (g.periodic[D]) {
		if (x >= g.bmin[D] && x < g.lo_band[D]) m |= 1u;
		if (x >= g.hi_band[D] && x < g.bmax[D]) m |= 2u;
	}
	return m;
}

}  // namespace ls1b200
