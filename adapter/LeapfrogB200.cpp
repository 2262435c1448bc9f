#include "integrators/LeapfrogB200.h"

#include "Domain.h"
#include "Simulation.h"
#include "particleContainer/LinkedCellsB200.h"
#include "utils/Logger.h"
#include "utils/mardyn_assert.h"
#include "utils/xmlfileUnits.h"

#include <cmath>

#include "ls1b200.h"

static LinkedCellsB200* deviceContainer(ParticleContainer* pc) {
	LinkedCellsB200* c = dynamic_cast<LinkedCellsB200*>(pc);
	if (!c) {
		std::ostringstream error_message;
		error_message << "LeapfrogB200 needs <datastructure type=\"LinkedCellsB200\">" << std::endl;
		MARDYN_EXIT(error_message.str());
	}
	return c;
}

void LeapfrogB200::readXML(XMLfileUnits& xmlconfig) {
	_timestepLength = 0;
	xmlconfig.getNodeValueReduced("timestep", _timestepLength);      // as Leapfrog::readXML (Leapfrog.cpp:23-28)
	Log::global_log->info() << "Timestep: " << _timestepLength << std::endl;
	mardyn_assert(_timestepLength > 0);
}

void LeapfrogB200::eventNewTimestep(ParticleContainer* moleculeContainer, Domain* domain) {
	LinkedCellsB200* c = deviceContainer(moleculeContainer);
	if (!c->deviceMode()) return;
	c->hostView(false);                     // from here on the device state is authoritative; host loops of the step iterate nothing
	// What the device path does NOT implement must not be skipped silently: with component-wise thermostats the host loop that would
	// scale the velocities is vacuous in device mode (Domain.cpp:195-215, VelocityScalingThermostat.cpp:68-89), and the
	// TemperatureControl plugin edits host molecules the device never sees.
	if (domain->severalThermostats()) {
		std::ostringstream error_message;
		error_message << "LeapfrogB200: component-wise thermostats are not supported on the device (single global velocity scaling or NVE only)" << std::endl;
		MARDYN_EXIT(error_message.str());
	}
	if (global_simulation->getPlugin("TemperatureControl") != nullptr) {
		std::ostringstream error_message;
		error_message << "LeapfrogB200: the TemperatureControl plugin modifies host molecules every step and is not supported with LinkedCellsB200" << std::endl;
		MARDYN_EXIT(error_message.str());
	}
	// VelocityScalingThermostat::apply of the previous step found an empty iterator range (device mode); its factors are
	// Domain's global betas (Domain.cpp:248-252) and are applied here, before the kick, which is the same arithmetic
	// (Simulation.cpp:1235-1261; single global thermostat -- the component-wise variant needs a per-component factor table).
	if (_thermostatPending && !domain->NVE()) {
		// Domain's "explosion" heuristics (Domain.cpp:259-314): when a plain factor falls below MIN_BETA and no error is pending,
		// calculateGlobalValues caps the kinetic energy of single molecules in a host loop -- vacuous in device mode -- and then reports
		// beta_trans = 1 and beta_rot^0.0091. The decision stays Domain's; it is recognised here by exactly that signature (the plain
		// factors follow from the sums this integrator handed to Domain), and the loop runs on the device before the factors are applied.
		const double Ti = global_simulation->Tfactor(global_simulation->getSimulationStep() - 1) * domain->getTargetTemperature(0);
		const double N = (double) domain->getglobalNumMolecules(false);
		if (Ti > 0. && N > 0. && _lastSummv2 > 0.) {
			const double btPlain = std::pow(3.0 * N * Ti / _lastSummv2, 0.4);
			const double brPlain = (_lastSumIw2 == 0.0) ? 1.0 : std::pow((double) _lastRotDOF * Ti / _lastSumIw2, 0.4);
			if ((btPlain < MIN_BETA || brPlain < MIN_BETA) && domain->getGlobalBetaTrans() == 1.0) {
				if (ls1b200_selective_thermostat(c->context(), Ti) != 0) MARDYN_EXIT(ls1b200_last_error());
			}
		}
		if (ls1b200_scale_velocities(c->context(), domain->getGlobalBetaTrans(), domain->getGlobalBetaRot()) != 0) MARDYN_EXIT(ls1b200_last_error());
	}
	if (ls1b200_leapfrog_pre(c->context(), _timestepLength) != 0) MARDYN_EXIT(ls1b200_last_error());
}

void LeapfrogB200::eventForcesCalculated(ParticleContainer* moleculeContainer, Domain* domain) {
	LinkedCellsB200* c = deviceContainer(moleculeContainer);
	if (!c->deviceMode()) return;
	ls1b200_kinetic kin{};
	if (ls1b200_leapfrog_post(c->context(), _timestepLength, &kin) != 0) MARDYN_EXIT(ls1b200_last_error());
	domain->setLocalSummv2(kin.summv2, 0);                        // Leapfrog.cpp:139-143
	domain->setLocalSumIw2(kin.sum_iw2, 0);
	domain->setLocalNrotDOF(0, kin.n_molecules, kin.rot_dof);
	_lastSummv2 = kin.summv2; _lastSumIw2 = kin.rot_dof > 0 ? kin.sum_iw2 : 0.; _lastRotDOF = kin.rot_dof;      // sequential build: local sums are global
	_thermostatPending = true;
	// Host coherence: on the steps of the container's sync schedule -- and at the last step, for the final
	// checkpoint -- the host mirror is refreshed HERE and the host view opens for the rest of the step. What follows in the stock
	// loop then acts on the mirror exactly as in the reference: the thermostat scales its velocities (the device applies the same
	// factors at the next eventNewTimestep), Homogeneous adds the per-molecule virial correction, output plugins
	// (CheckpointWriter, MmpldWriter, RDF, ...) iterate real molecules. eventNewTimestep closes the view again.
	const unsigned long step = global_simulation->getSimulationStep();
	const unsigned long last = global_simulation->getNumInitTimesteps() + global_simulation->getNumTimesteps();
	if (c->wantsHostSync(step, last)) { c->syncToHost(); c->hostView(true); }
}
