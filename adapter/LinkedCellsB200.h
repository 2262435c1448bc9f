// LinkedCellsB200 -- the ls1-MarDyn side of the drop-in: a ParticleContainer that keeps the molecules on a B200 through
// libls1b200.so (include/ls1b200.h). Compiled AGAINST ls1's headers (this file is added to ls1's src/ by
// adapter/build_adapter.sh; it is not part of libls1b200.so).
//
// It derives from the reference's own LinkedCells, which stays the host mirror: input readers / generators fill it through
// addParticle(s) / initCubicGrid exactly as before (Simulation::prepare_start), and plugins that iterate molecules see it
// after a lazy device -> host sync. From the first force traversal on the device copy is authoritative ("device mode"): the
// per-step host loops of the stock time step find an empty iterator range and cost nothing.
// Precedent in the reference for a container that ignores the CellProcessor's kernels and writes Upot / Virial itself:
// AutoPasContainer::traverseCells (particleContainer/AutoPasContainer.cpp:554-606).
#pragma once

#include "particleContainer/LinkedCells.h"

struct ls1b200_ctx;

class LinkedCellsB200 : public LinkedCells {
public:
	LinkedCellsB200() : LinkedCells() {}
	~LinkedCellsB200() override;

	// <datastructure type="LinkedCellsB200"> takes the same children as LinkedCells plus
	//   <hostSyncFrequency>INT</hostSyncFrequency>   every that many steps (and always at the last step, for the final checkpoint) the
	//                                                host mirror is refreshed after the forces: output plugins, checkpoint writers,
	//                                                RDF etc. then see the molecules of that step (0 = only at the last step)
	//   <syncAtLastStep>BOOL</syncAtLastStep>        default true; false skips the refresh at the last step (runs with --final-checkpoint=0
	//                                                and no molecule-reading plugin at the end: the refresh of 16 M molecules costs seconds)
	//   <device>INT</device>                         CUDA device ordinal (default: environment LS1B200_DEVICE, else 0)
	void readXML(XMLfileUnits& xmlconfig) override;
	bool wantsHostSync(unsigned long step, unsigned long lastStep) const { return (_syncAtLastStep && step >= lastStep) || (_hostSyncFrequency > 0 && step % _hostSyncFrequency == 0); }
	void update() override;                                  // -> ls1b200_update once the molecules live on the device
	void traverseCells(CellProcessor& cellProcessor) override;   // -> ls1b200_traverse_cells, Domain::setLocalUpot/Virial (first call: upload)
	void updateMoleculeCaches() override { if (!_deviceMode) LinkedCells::updateMoleculeCaches(); }
	void updateInnerMoleculeCaches() override { if (!_deviceMode) LinkedCells::updateInnerMoleculeCaches(); }
	void updateBoundaryAndHaloMoleculeCaches() override { if (!_deviceMode) LinkedCells::updateBoundaryAndHaloMoleculeCaches(); }
	void deleteOuterParticles() override { if (!_deviceMode) LinkedCells::deleteOuterParticles(); }   // halo is transient device data
	unsigned long getNumberOfParticles(ParticleIterator::Type t = ParticleIterator::ALL_CELLS) override;
	bool requiresForceExchange() const override { return false; }   // full-shell on the device, like plain C08 (C08CellPairTraversal.h:35)

	// Until LeapfrogB200 starts the first time step the host mirror still equals the device state and stays visible
	// (Domain::calculateVelocitySums etc. in Simulation::prepare_start, Simulation.cpp:986-992, need it).
	// In device mode the host loops of the time step (Leapfrog, exchangeMolecules, updateForces, thermostat) iterate nothing;
	// a caller that really wants the molecules (writers, checkpoints, RDF) calls syncToHost() first -- or sets hostView(true).
	ParticleIterator iterator(ParticleIterator::Type t) override;
	RegionParticleIterator regionIterator(const double startRegion[3], const double endRegion[3], ParticleIterator::Type type) override;

	void syncToHost();            // device -> host Molecule objects (r, v, q, D, F, M, Vi), then the host cell structure
	void syncToDevice();          // host -> device after a plugin changed molecules
	void hostView(bool on) { _hostView = on; }
	ls1b200_ctx* context() { return _ctx; }
	bool deviceMode() const { return _deviceMode; }

private:
	void createContext();
	void uploadAll();

	ls1b200_ctx* _ctx = nullptr;
	bool _deviceMode = false;
	bool _hostView = true;        // until the first time step: the init-time loops of prepare_start read the host mirror
	unsigned long _numMolecules = 0;
	// Homogeneous long-range correction, per-molecule part (see traverseCells): Domain's VirialCorr reconstructed from its public values
	double _prevLocalVirial = 0.;
	bool _havePrevVirial = false;
	double _virialCorr = 0.;
	unsigned long _hostSyncFrequency = 0;
	bool _syncAtLastStep = true;
	int _device = -1;
};
