// LeapfrogB200 -- <integrator type="LeapfrogB200">: the Leapfrog integrator (integrators/Leapfrog.cpp) on the device state
// of a LinkedCellsB200 container. Falls back to nothing: with any other container it aborts (MARDYN_EXIT).
#pragma once

#include "integrators/Integrator.h"

class LeapfrogB200 : public Integrator {
public:
	LeapfrogB200() {}
	void readXML(XMLfileUnits& xmlconfig) override;
	void init() override {}
	void eventNewTimestep(ParticleContainer* moleculeContainer, Domain* domain) override;       // transition1to2: upd_preF
	void eventForcesCalculated(ParticleContainer* moleculeContainer, Domain* domain) override;  // transition2to3: upd_postF + sums
private:
	bool _thermostatPending = false;   // a step has completed: its velocity-scaling factors are applied before the next kick
	double _lastSummv2 = 0., _lastSumIw2 = 0.;   // the kinetic sums of that step, as handed to Domain (eventForcesCalculated)
	unsigned long _lastRotDOF = 0;
};
