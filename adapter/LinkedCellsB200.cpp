#include "particleContainer/LinkedCellsB200.h"

#include <algorithm>
#include <unordered_map>
#include <vector>

#include "Domain.h"
#include "Simulation.h"
#include "ensemble/EnsembleBase.h"
#include "molecules/Comp2Param.h"
#include "molecules/Component.h"
#include "particleContainer/adapter/LegacyCellProcessor.h"
#include "particleContainer/adapter/VectorizedCellProcessor.h"
#include "utils/Logger.h"
#include "utils/mardyn_assert.h"
#include "utils/xmlfileUnits.h"

#include <cstdlib>

#include "ls1b200.h"

#define B200_CHECK(call)                                                             \
	do {                                                                               \
		if ((call) != 0) {                                                               \
			std::ostringstream error_message;                                              \
			error_message << "LinkedCellsB200: " #call ": " << ls1b200_last_error() << std::endl; \
			MARDYN_EXIT(error_message.str());                                              \
		}                                                                                \
	} while (0)

LinkedCellsB200::~LinkedCellsB200() {
	if (_ctx) ls1b200_destroy(_ctx);
}

void LinkedCellsB200::readXML(XMLfileUnits& xmlconfig) {
	LinkedCells::readXML(xmlconfig);
	const int cic = xmlconfig.getNodeValue_int("cellsInCutoffRadius", 1);      // private in LinkedCells: read it again (LinkedCells.cpp:125)
	if (cic < 1) {
		std::ostringstream error_message;
		error_message << "LinkedCellsB200: cellsInCutoffRadius = " << cic << " is not valid" << std::endl;
		MARDYN_EXIT(error_message.str());
	}
	// cellsInCutoffRadius > 1 refines the HOST cell grid (the inherited mirror, its iterators and plugins see the cells the input asked for).
	// The device keeps one cell per cutoff radius: which pairs interact is decided by the cutoff test alone, so forces, energies and the
	// virial are those of the reference with any cell size (tests/test_gpu_dropin.py::test_cells_in_cutoff_radius_2, against the stock binary).
	if (cic != 1) Log::global_log->info() << "LinkedCellsB200: cellsInCutoffRadius = " << cic << " applies to the host mirror; the device traversal uses cells of one cutoff radius" << std::endl;
	int f = 0;
	xmlconfig.getNodeValue("hostSyncFrequency", f);
	_hostSyncFrequency = f > 0 ? (unsigned long) f : 0ul;
	xmlconfig.getNodeValue("device", _device);
	xmlconfig.getNodeValue("syncAtLastStep", _syncAtLastStep);
	Log::global_log->info() << "LinkedCellsB200: host mirror refreshed every " << _hostSyncFrequency << " steps (0: only at the last step)" << std::endl;
}

// Everything VectorizedCellProcessor's constructor derives from Domain / Ensemble (VectorizedCellProcessor.cpp:19-81) goes
// through ls1b200_set_components: the component list and the (24 eps, sigma^2, 6 shift) tables of Comp2Param.
void LinkedCellsB200::createContext() {
	Domain* domain = global_simulation->getDomain();
	ls1b200_config cfg{};
	for (int d = 0; d < 3; d++) {
		cfg.box_min[d] = getBoundingBoxMin(d);
		cfg.box_max[d] = getBoundingBoxMax(d);
		cfg.periodic[d] = 1;          // sequential DomainDecompBase: every dimension wraps onto itself
	}
	cfg.cutoff = global_simulation->getcutoffRadius();
	cfg.cutoff_lj = global_simulation->getLJCutoff();
	cfg.epsilon_rf = domain->getepsilonRF();
#if defined(MARDYN_DPDP)
	cfg.precision = LS1B200_DOUBLE;
#else
	cfg.precision = LS1B200_MIXED;    // MARDYN_SPDP
#endif
	cfg.geometry = LS1B200_GEOM_XML;
	// device ordinal: <device> of the datastructure, else LS1B200_DEVICE (one process per GPU: the launcher exports it per rank), else 0.
	// This build of ls1 is sequential (DomainDecompBase: every dimension wraps onto itself, periodic = 1 above); under
	// DomainDecomposition each rank would pass its block and its periodic flags and call ls1b200_decomp_attach.
	cfg.device = _device >= 0 ? _device : (getenv("LS1B200_DEVICE") ? atoi(getenv("LS1B200_DEVICE")) : 0);
	B200_CHECK(ls1b200_create(&cfg, &_ctx));

	std::vector<Component>& components = *(global_simulation->getEnsemble()->getComponents());
	std::vector<ls1b200_component> comps(components.size());
	std::vector<ls1b200_site> sites;
	std::vector<size_t> ljStart(components.size());
	size_t centers = 0;
	for (size_t i = 0; i < components.size(); ++i) {
		const Component& c = components[i];
		mardyn_assert(c.ID() == i);
		ls1b200_component& o = comps[i];
		o = ls1b200_component{};
		o.n_lj = c.numLJcenters(); o.n_charge = c.numCharges(); o.n_dipole = c.numDipoles(); o.n_quadrupole = c.numQuadrupoles();
		o.site_offset = sites.size(); o.lj_lookup_id = centers; o.rot_dof = c.getRotationalDegreesOfFreedom();
		o.mass = c.m(); o.inertia[0] = c.I11(); o.inertia[1] = c.I22(); o.inertia[2] = c.I33();
		ljStart[i] = centers; centers += c.numLJcenters();
		auto push = [&sites](const Site& s, const double* e, double param) {
			ls1b200_site t{};
			t.r[0] = s.rx(); t.r[1] = s.ry(); t.r[2] = s.rz();
			if (e) { t.e[0] = e[0]; t.e[1] = e[1]; t.e[2] = e[2]; }
			t.param = param;
			sites.push_back(t);
		};
		for (unsigned j = 0; j < c.numLJcenters(); ++j) push(c.ljcenter(j), nullptr, 0.);
		for (unsigned j = 0; j < c.numCharges(); ++j) push(c.charge(j), nullptr, c.charge(j).q());
		for (unsigned j = 0; j < c.numDipoles(); ++j) { auto e = c.dipole(j).e(); push(c.dipole(j), e.data(), c.dipole(j).absMy()); }
		for (unsigned j = 0; j < c.numQuadrupoles(); ++j) { auto e = c.quadrupole(j).e(); push(c.quadrupole(j), e.data(), c.quadrupole(j).absQ()); }
	}
	std::vector<double> eps24(centers * centers), sig2(centers * centers), shift6(centers * centers);
	for (size_t ci = 0; ci < components.size(); ++ci) for (size_t cj = 0; cj < components.size(); ++cj) {
		ParaStrm& p = domain->getComp2Params()(components[ci].ID(), components[cj].ID());
		p.reset_read();
		for (size_t a = 0; a < components[ci].numLJcenters(); ++a) for (size_t b = 0; b < components[cj].numLJcenters(); ++b) {
			double eps, sig, shift;
			p >> eps; p >> sig; p >> shift;
			const size_t k = (ljStart[ci] + a) * centers + ljStart[cj] + b;
			eps24[k] = eps; sig2[k] = sig; shift6[k] = shift;
		}
	}
	B200_CHECK(ls1b200_set_components(_ctx, comps.size(), comps.data(), sites.size(), sites.data(), centers, eps24.data(), sig2.data(), shift6.data()));
}

void LinkedCellsB200::uploadAll() {
	std::vector<uint64_t> id; std::vector<uint32_t> cid; std::vector<double> r, v, q, D;
	for (auto m = LinkedCells::iterator(ParticleIterator::ONLY_INNER_AND_BOUNDARY); m.isValid(); ++m) {
		id.push_back(m->getID()); cid.push_back(m->componentid());
		for (int d = 0; d < 3; d++) { r.push_back(m->r(d)); v.push_back(m->v(d)); D.push_back(m->D(d)); }
		const Quaternion& qq = m->q();
		q.push_back(qq.qw()); q.push_back(qq.qx()); q.push_back(qq.qy()); q.push_back(qq.qz());
	}
	_numMolecules = id.size();
	B200_CHECK(ls1b200_upload(_ctx, id.size(), id.data(), cid.data(), r.data(), v.data(), q.data(), D.data()));
}

void LinkedCellsB200::update() {
	// Before the first force calculation (input readers, Simulation::initialize) this is the stock host container: the
	// pair-parameter streams the device tables are built from do not exist yet (Domain::initParameterStreams runs in
	// Simulation::prepare_start).
	if (!_deviceMode) { LinkedCells::update(); return; }
	B200_CHECK(ls1b200_update(_ctx));          // == update() + exchangeMolecules() + updateMoleculeCaches()
}

void LinkedCellsB200::traverseCells(CellProcessor& cellProcessor) {
	const bool forceProcessor = dynamic_cast<VectorizedCellProcessor*>(&cellProcessor) || dynamic_cast<LegacyCellProcessor*>(&cellProcessor);
	if (!_deviceMode && forceProcessor) {
		// first force calculation (Simulation::prepare_start, Simulation.cpp:908-912): the molecules move to the device
		LinkedCells::deleteOuterParticles();   // the host halo copies made by exchangeMolecules are not owned molecules
		if (!_ctx) createContext();
		uploadAll();
		_deviceMode = true;
		B200_CHECK(ls1b200_update(_ctx));
		Log::global_log->info() << "LinkedCellsB200: " << _numMolecules << " molecules now live on the device" << std::endl;
	}
	if (_deviceMode && forceProcessor) {
		ls1b200_macro mac{};
		B200_CHECK(ls1b200_traverse_cells(_ctx, &mac));      // forces, torques, per-molecule virial stay on the device
		Domain* domain = global_simulation->getDomain();
		// The per-molecule part of the long-range correction: Homogeneous::calculateLongRange adds VirialCorr / 3N to every molecule's
		// Vi right after the traversal (longRange/Homogeneous.cpp:143-155, Simulation.cpp:1175) -- a host loop that iterates nothing in
		// device mode (and the view cannot open before Simulation::updateForces has run: calcFM would add the mirror's stale site
		// forces). syncToHost() adds the term instead. Domain keeps VirialCorr private, but publishes _globalVirial / N
		// (getAverageGlobalVirial, Domain.cpp:148, :187: _globalVirial = Virial + _VirialCorr): at this point it still holds the
		// values Domain::calculateGlobalValues computed from the PREVIOUS traversal's virial. VirialCorr depends on N, V and the
		// components only. (Sequential build: the local virial is the global one.)
		if (_havePrevVirial) _virialCorr = domain->getAverageGlobalVirial() * (double) domain->getglobalNumMolecules(false) - _prevLocalVirial;
		_prevLocalVirial = mac.virial; _havePrevVirial = true;
		domain->setLocalVirial(mac.virial);                  // as VectorizedCellProcessor::endTraversal (.cpp:153-154)
		domain->setLocalUpot(mac.upot);
		return;
	}
	// any other processor (RDFCellProcessor, FlopCounter, ...) works on the host mirror
	if (_deviceMode) syncToHost();
	const bool hv = _hostView; _hostView = true;
	LinkedCells::traverseCells(cellProcessor);
	_hostView = hv;
}

unsigned long LinkedCellsB200::getNumberOfParticles(ParticleIterator::Type t) {
	if (_deviceMode) return ls1b200_num_molecules(_ctx);
	return LinkedCells::getNumberOfParticles(t);
}

ParticleIterator LinkedCellsB200::iterator(ParticleIterator::Type t) {
	if (_deviceMode && !_hostView) return ParticleIterator();      // invalid iterator: the loop body never runs
	return LinkedCells::iterator(t);
}

RegionParticleIterator LinkedCellsB200::regionIterator(const double startRegion[3], const double endRegion[3], ParticleIterator::Type type) {
	if (_deviceMode && !_hostView) return RegionParticleIterator();
	return LinkedCells::regionIterator(startRegion, endRegion, type);
}

void LinkedCellsB200::syncToHost() {
	if (!_deviceMode) return;
	const uint64_t n = ls1b200_num_molecules(_ctx);
	std::vector<uint64_t> id(n);
	std::vector<double> r(3 * n), v(3 * n), q(4 * n), D(3 * n), F(3 * n), M(3 * n), Vi(3 * n);
	B200_CHECK(ls1b200_download(_ctx, LS1B200_F_ID, id.data(), n));
	B200_CHECK(ls1b200_download(_ctx, LS1B200_F_R, r.data(), n)); B200_CHECK(ls1b200_download(_ctx, LS1B200_F_V, v.data(), n));
	B200_CHECK(ls1b200_download(_ctx, LS1B200_F_Q, q.data(), n)); B200_CHECK(ls1b200_download(_ctx, LS1B200_F_L, D.data(), n));
	B200_CHECK(ls1b200_download(_ctx, LS1B200_F_F, F.data(), n)); B200_CHECK(ls1b200_download(_ctx, LS1B200_F_M, M.data(), n));
	B200_CHECK(ls1b200_download(_ctx, LS1B200_F_VI, Vi.data(), n));
	// id -> row of the downloaded arrays: a flat table when the ids are dense (generators, checkpoints), a hash map otherwise
	uint64_t maxId = 0;
	for (size_t i = 0; i < n; i++) maxId = std::max(maxId, id[i]);
	const bool dense = maxId < 8 * n + 1024;
	std::vector<uint32_t> flat;
	std::unordered_map<uint64_t, size_t> at;
	if (dense) {
		flat.assign(maxId + 1, 0xFFFFFFFFu);
		for (size_t i = 0; i < n; i++) flat[id[i]] = (uint32_t) i;
	} else {
		at.reserve(n);
		for (size_t i = 0; i < n; i++) at[id[i]] = i;
	}
	const unsigned long nGlobal = global_simulation->getDomain()->getglobalNumMolecules(false);
	const double viAdd = nGlobal > 0 ? _virialCorr / (3. * (double) nGlobal) : 0.;
	#if defined(_OPENMP)
	#pragma omp parallel
	#endif
	for (auto m = LinkedCells::iterator(ParticleIterator::ALL_CELLS); m.isValid(); ++m) {
		size_t i;
		if (dense) {
			const uint64_t mid = m->getID();
			if (mid > maxId || flat[mid] == 0xFFFFFFFFu) continue;
			i = flat[mid];
		} else {
			auto it = at.find(m->getID());
			if (it == at.end()) continue;
			i = it->second;
		}
		for (int d = 0; d < 3; d++) { m->setr(d, r[3 * i + d]); m->setv(d, v[3 * i + d]); m->setD(d, D[3 * i + d]); }
		m->setq(Quaternion(q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]));
		// pair virial from the device + what Homogeneous::calculateLongRange would have added to this molecule (see traverseCells)
		const double vi[3] = {Vi[3 * i] + viAdd, Vi[3 * i + 1] + viAdd, Vi[3 * i + 2] + viAdd};
		m->setF(&F[3 * i]); m->setM(&M[3 * i]); m->setVi(const_cast<double*>(vi));
	}
	LinkedCells::update();                   // re-sort the host mirror
}

void LinkedCellsB200::syncToDevice() {
	if (!_ctx) return;
	uploadAll();
	B200_CHECK(ls1b200_update(_ctx));
}
