// refdump -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A small driver of OUR OWN that links against the unmodified ls1-MarDyn objects built by
// oracle/build_ref.sh and calls the reference's public API the same way its own unit-test fixture does
// (reference: src/particleContainer/tests/ParticleContainerFactory.cpp:52-102 for the set-up sequence,
//  src/particleContainer/adapter/tests/VectorizedCellProcessorTest.cpp:136-214 for the force sequence).
// It dumps what the parity tests need as golden vectors:
//   * per molecule: id, cid, cell index, halo flag, r, v, q, D(=L), F, M, Vi   (after traverseCells + calcFM)
//   * local Upot and Virial as written into Domain by the cell processor's endTraversal()
//   * the LJ pair-parameter streams (24*eps, sigma^2, 6*shift) from Comp2Param
//   * the LinkedCells geometry (cells per dimension, cell length, halo bounding box)
//
// usage: refdump <file.inp> <cutoff> <out.bin> <vcp|legacy> <open|periodic> <ctor|xml>
//   open     : no halo population (== the reference unit-test fixture)
//   periodic : DomainDecompBase::exchangeMolecules() is called first (wrap + periodic halo copies),
//              i.e. what Simulation::updateParticleContainerAndDecomposition does in a sequential build
//   ctor     : LinkedCells(bBoxMin,bBoxMax,cutoff) geometry (pure-double cell length, LinkedCells.cpp:47-58)
//   xml      : default-constructed LinkedCells + rebuild() (float cast of rc, LinkedCells.cpp:148-166)
//
// Output: little-endian binary, see tests/golden/make_golden.py for the reader.

#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "Simulation.h"
#include "Domain.h"
#include "ensemble/EnsembleBase.h"
#include "io/ASCIIReader.h"
#include "molecules/Molecule.h"
#include "molecules/Comp2Param.h"
#include "parallel/DomainDecompBase.h"
#include "particleContainer/LinkedCells.h"
#include "particleContainer/adapter/LegacyCellProcessor.h"
#include "particleContainer/adapter/ParticlePairs2PotForceAdapter.h"
#include "particleContainer/adapter/VectorizedCellProcessor.h"
#include "utils/Logger.h"

static void wr(FILE* f, const void* p, size_t n) { if (fwrite(p, 1, n, f) != n) { perror("fwrite"); exit(3); } }
static void wr_u64(FILE* f, uint64_t v) { wr(f, &v, 8); }
static void wr_f64(FILE* f, double v) { wr(f, &v, 8); }

int main(int argc, char** argv) {
	if (argc < 7) {
		fprintf(stderr, "usage: %s <file.inp> <cutoff> <out.bin> <vcp|legacy> <open|periodic> <ctor|xml>\n", argv[0]);
		return 2;
	}
	const std::string inp = argv[1];
	const double cutoff = atof(argv[2]);
	const std::string out = argv[3];
	const bool legacy = std::string(argv[4]) == "legacy";
	const bool periodic = std::string(argv[5]) == "periodic";
	const bool xmlgeom = std::string(argv[6]) == "xml";

	Log::global_log = std::make_unique<Log::Logger>(Log::Error);
	Simulation* sim = new Simulation();  // assigns global_simulation
	(void) sim;
	Domain* domain = global_simulation->getDomain();
	DomainDecompBase* dd = &global_simulation->domainDecomposition();
	global_simulation->setcutoffRadius(cutoff);
	global_simulation->setLJCutoff(cutoff);

	ASCIIReader reader;
	reader.setPhaseSpaceHeaderFile(inp);
	reader.setPhaseSpaceFile(inp);
	reader.readPhaseSpaceHeader(domain, 1.0);
	double bmin[3], bmax[3];
	for (int d = 0; d < 3; d++) { bmin[d] = dd->getBoundingBoxMin(d, domain); bmax[d] = dd->getBoundingBoxMax(d, domain); }

	LinkedCells* lc;
	if (xmlgeom) {
		lc = new LinkedCells();
		lc->setCutoff(cutoff);
		lc->rebuild(bmin, bmax);
	} else {
		lc = new LinkedCells(bmin, bmax, cutoff);
	}
	reader.readPhaseSpace(lc, domain, dd);
	lc->deleteOuterParticles();
	lc->update();
	if (periodic) {
		dd->exchangeMolecules(lc, domain);
	}
	lc->updateMoleculeCaches();
	domain->initParameterStreams(cutoff, cutoff);

	if (legacy) {
		ParticlePairs2PotForceAdapter adapter(*domain);
		LegacyCellProcessor cp(cutoff, cutoff, &adapter);
		lc->traverseCells(cp);
	} else {
		VectorizedCellProcessor cp(*domain, cutoff, cutoff);
		lc->traverseCells(cp);
	}
	for (auto m = lc->iterator(ParticleIterator::ALL_CELLS); m.isValid(); ++m) m->calcFM();

	FILE* f = fopen(out.c_str(), "wb");
	if (!f) { perror("fopen"); return 3; }
	wr(f, "LS1REFD1", 8);
	// geometry
	for (int d = 0; d < 3; d++) wr_f64(f, bmin[d]);
	for (int d = 0; d < 3; d++) wr_f64(f, bmax[d]);
	for (int d = 0; d < 3; d++) wr_f64(f, lc->getHaloWidthNumCells() * 1.0);
	double clen[3];
	{
		double* cl = lc->getCellLength();
		for (int d = 0; d < 3; d++) { clen[d] = cl[d]; wr_f64(f, clen[d]); }
	}
	uint64_t cpd[3];
	for (int d = 0; d < 3; d++) { cpd[d] = (uint64_t) (lc->getBoxWidthInNumCells()[d] + 2 * lc->getHaloWidthNumCells()); wr_u64(f, cpd[d]); }  // incl. halo
	wr_f64(f, cutoff);
	wr_f64(f, domain->getepsilonRF());
	wr_f64(f, domain->getLocalUpot());
	wr_f64(f, domain->getLocalVirial());

	// LJ parameter streams, laid out like VectorizedCellProcessor's tables (VectorizedCellProcessor.cpp:46-81)
	const auto& comps = *(global_simulation->getEnsemble()->getComponents());
	size_t centers = 0;
	for (auto& c : comps) centers += c.numLJcenters();
	wr_u64(f, comps.size());
	wr_u64(f, centers);
	std::vector<size_t> start(comps.size());
	{ size_t s = 0; for (size_t i = 0; i < comps.size(); i++) { start[i] = s; s += comps[i].numLJcenters(); } }
	std::vector<double> eps24(centers * centers), sig2(centers * centers), shift6(centers * centers);
	for (size_t ci = 0; ci < comps.size(); ci++)
		for (size_t cj = 0; cj < comps.size(); cj++) {
			ParaStrm& p = domain->getComp2Params()(comps[ci].ID(), comps[cj].ID());
			p.reset_read();
			for (size_t a = 0; a < comps[ci].numLJcenters(); a++)
				for (size_t b = 0; b < comps[cj].numLJcenters(); b++) {
					double e, s, sh;
					p >> e; p >> s; p >> sh;
					size_t k = (start[ci] + a) * centers + (start[cj] + b);
					eps24[k] = e; sig2[k] = s; shift6[k] = sh;
				}
		}
	wr(f, eps24.data(), 8 * eps24.size());
	wr(f, sig2.data(), 8 * sig2.size());
	wr(f, shift6.data(), 8 * shift6.size());
	// per component: mass, I[3], rotDOF, site counts
	for (auto& c : comps) {
		wr_f64(f, c.m()); wr_f64(f, c.I11()); wr_f64(f, c.I22()); wr_f64(f, c.I33());
		wr_u64(f, c.getRotationalDegreesOfFreedom());
		wr_u64(f, c.numLJcenters()); wr_u64(f, c.numCharges()); wr_u64(f, c.numDipoles()); wr_u64(f, c.numQuadrupoles());
	}

	// molecules
	uint64_t n = 0;
	for (auto m = lc->iterator(ParticleIterator::ALL_CELLS); m.isValid(); ++m) n++;
	wr_u64(f, n);
	for (auto m = lc->iterator(ParticleIterator::ALL_CELLS); m.isValid(); ++m) {
		const uint64_t cell = m.getCellIndex();
		// recompute halo flag from the 3D index
		uint64_t nx = cpd[0], ny = cpd[1], nz = cpd[2];
		uint64_t ix = cell % nx, iy = (cell / nx) % ny, iz = cell / (nx * ny);
		uint64_t halo = (ix == 0 || iy == 0 || iz == 0 || ix == nx - 1 || iy == ny - 1 || iz == nz - 1) ? 1 : 0;
		wr_u64(f, m->getID());
		wr_u64(f, m->componentid());
		wr_u64(f, cell);
		wr_u64(f, halo);
		for (int d = 0; d < 3; d++) wr_f64(f, m->r(d));
		for (int d = 0; d < 3; d++) wr_f64(f, m->v(d));
		wr_f64(f, m->q().qw()); wr_f64(f, m->q().qx()); wr_f64(f, m->q().qy()); wr_f64(f, m->q().qz());
		for (int d = 0; d < 3; d++) wr_f64(f, m->D(d));
		for (int d = 0; d < 3; d++) wr_f64(f, m->F(d));
		for (int d = 0; d < 3; d++) wr_f64(f, m->M(d));
		for (int d = 0; d < 3; d++) wr_f64(f, m->Vi(d));
	}
	fclose(f);
	return 0;
}
