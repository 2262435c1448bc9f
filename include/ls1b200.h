/* ls1b200.h -- C ABI of libls1b200.so: the B200-native short-range force path of ls1-MarDyn.
 *
 * This is the drop-in boundary. Every entry point below replaces one host-side call of the reference's
 * time-step loop (Simulation::simulateOneTimestep, /root/reference/src/Simulation.cpp:1076-1300); the
 * reference-side binding (a ParticleContainer / Integrator adapter compiled against ls1's headers) is shown in
 * INTEGRATION.md. All "file:line" citations are relative to the reference tree (/root/reference).
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success, non-zero on error; the message is available
 *     through ls1b200_last_error() (the reference aborts via MARDYN_EXIT -- utils/mardyn_assert.h -- the adapter
 *     turns a non-zero return into exactly that);
 *   - never throws, never falls back to a CPU path: without a CUDA device ls1b200_create() fails;
 *   - one caller thread per context (the reference calls these phases from the serial master context of
 *     simulateOneTimestep, Simulation.cpp:1104-1300); a context owns all of its device memory;
 *   - host pointers are borrowed for the duration of a call only;
 *   - molecule state is FP64 (as FullMolecule, molecules/FullMolecule.h:341-346), pair arithmetic is FP32 with
 *     FP64 accumulation (LS1B200_MIXED == -DMARDYN_SPDP) or FP64 (LS1B200_DOUBLE == -DMARDYN_DPDP),
 *     vectorization/SIMD_TYPES.h:108-110.
 */
#ifndef LS1B200_H
#define LS1B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LS1B200_ABI_VERSION 2

typedef struct ls1b200_ctx ls1b200_ctx;

enum { LS1B200_MIXED = 0, LS1B200_DOUBLE = 1 };
/* How LinkedCells derives its cell length: LS1B200_GEOM_XML follows LinkedCells::rebuild (float cast of
 * rc/cellsInCutoff, particleContainer/LinkedCells.cpp:148-166), LS1B200_GEOM_CTOR the (bBoxMin,bBoxMax,cutoff)
 * constructor used by the reference's unit-test fixture (LinkedCells.cpp:47-58). */
enum { LS1B200_GEOM_XML = 0, LS1B200_GEOM_CTOR = 1 };

typedef struct ls1b200_config {
	double box_min[3];      /* bounding box of THIS rank's domain (LinkedCells::_boundingBoxMin/Max)            */
	double box_max[3];
	double cutoff;          /* rc   (CellProcessor::_cutoffRadiusSquare = rc^2, adapter/CellProcessor.h:31)   */
	double cutoff_lj;       /* rcLJ (CellProcessor::_LJCutoffRadiusSquare, :32); Simulation forces rc>=rcLJ  */
	double epsilon_rf;      /* Domain::getepsilonRF(); enters dipole-dipole only (VectorizedCellProcessor.cpp:22) */
	int32_t precision;      /* LS1B200_MIXED | LS1B200_DOUBLE                                                   */
	int32_t geometry;       /* LS1B200_GEOM_XML | LS1B200_GEOM_CTOR                                             */
	int32_t periodic[3];    /* 1: this rank wraps onto itself in that dimension (sequential DomainDecompBase,
	                           parallel/DomainDecompBase.cpp:80-114); 0: no local periodic images are made --
	                           open boundary (the unit-test fixture) or halo supplied by a neighbour rank via
	                           ls1b200_halo_import                                                               */
	int32_t device;         /* CUDA device ordinal                                                              */
	int32_t reserved[4];
} ls1b200_config;

/* One Component (molecules/Component.h). Sites are listed LJ | charge | dipole | quadrupole, the order of
 * CellDataSoA / ConcatenatedSites (utils/ConcatenatedSites.h:55-87). */
typedef struct ls1b200_component {
	uint32_t n_lj, n_charge, n_dipole, n_quadrupole;
	uint32_t site_offset;   /* first row of this component in the site table                                    */
	uint32_t lj_lookup_id;  /* Component::getLookUpID(): first row/column of this component in the LJ tables    */
	uint32_t rot_dof;       /* Component::getRotationalDegreesOfFreedom()                                       */
	uint32_t reserved;
	double mass;            /* Component::m()                                                                   */
	double inertia[3];      /* Component::I11(), I22(), I33() (principal axes)                                  */
} ls1b200_component;

/* One row of the site table: body-frame position, body-frame orientation (dipole / quadrupole; 0 otherwise) and
 * the scalar the SoA carries for that site (charge q, |mu|, |Q|; unused for LJ), molecules/Site.h. */
typedef struct ls1b200_site {
	double r[3];
	double e[3];
	double param;
} ls1b200_site;

/* Macroscopic values of one traversal: what VectorizedCellProcessor::endTraversal hands to
 * Domain::setLocalUpot / setLocalVirial (VectorizedCellProcessor.cpp:122-155). */
typedef struct ls1b200_macro {
	double upot;            /* U6_LJ/6 + U_xpoles + myRF                                                        */
	double virial;          /* Virial + 3*myRF                                                                  */
	double upot6lj, upot_xpoles, virial_raw, my_rf;   /* the four raw accumulators, for tests                  */
} ls1b200_macro;

/* What Leapfrog::transition2to3 hands to Domain::setLocalSummv2 / setLocalSumIw2 / setLocalNrotDOF
 * (integrators/Leapfrog.cpp:139-143). */
typedef struct ls1b200_kinetic {
	double summv2;
	double sum_iw2;
	uint64_t n_molecules;
	uint64_t rot_dof;
} ls1b200_kinetic;

/* Per-step record of ls1b200_simulate: the numbers of the reference's "Simstep" log line and ResultWriter
 * (Simulation.cpp:1383-1386, io/ResultWriter.cpp:81). */
typedef struct ls1b200_step_record {
	double upot;        /* Domain::getGlobalUpot() (no long-range correction)                                   */
	double virial;      /* Domain::_globalVirial   (no long-range correction)                                   */
	double summv2, sum_iw2;
	double temperature; /* Domain::getGlobalCurrentTemperature(), Domain.cpp:240-241                            */
	double pressure;    /* Domain::getGlobalPressure(), Domain.cpp:142-146                                      */
	double beta_trans, beta_rot;   /* Domain.cpp:248-252                                                        */
} ls1b200_step_record;

enum ls1b200_field {
	LS1B200_F_ID = 0,   /* uint64 [n]                                                                           */
	LS1B200_F_CID,      /* uint32 [n]                                                                           */
	LS1B200_F_R,        /* double [n][3]                                                                        */
	LS1B200_F_V,        /* double [n][3]                                                                        */
	LS1B200_F_Q,        /* double [n][4]  (w,x,y,z)                                                             */
	LS1B200_F_L,        /* double [n][3]  angular momentum D                                                    */
	LS1B200_F_F,        /* double [n][3]  FullMolecule::_F after calcFM                                         */
	LS1B200_F_M,        /* double [n][3]  FullMolecule::_M after calcFM                                         */
	LS1B200_F_VI,       /* double [n][3]  FullMolecule::_Vi after calcFM                                        */
	LS1B200_F_CELL,     /* uint64 [n]     LinkedCells cell index (incl. halo layer) of the molecule             */
	LS1B200_F_COUNT
};

const char* ls1b200_last_error(void);
int ls1b200_abi_version(void);
/* Number of CUDA kernel launches issued by this library since load (bench.py's gpu_launches). */
uint64_t ls1b200_kernel_launch_count(void);

/* The CUDA stream (cudaStream_t) every kernel of this context is launched on: lets the caller record its own CUDA
 * events around calls (bench.py) or order its own work (NCCL halo exchange) against the library's. */
void* ls1b200_stream(ls1b200_ctx* ctx);

/* new LinkedCells + VectorizedCellProcessor + Leapfrog on one device
 * (Simulation.cpp:484, :854, :181). */
int ls1b200_create(const ls1b200_config* cfg, ls1b200_ctx** out);
int ls1b200_destroy(ls1b200_ctx* ctx);

/* Component list + the LJ pair tables VectorizedCellProcessor builds from Comp2Param
 * (VectorizedCellProcessor.cpp:39-81): eps24[i*C+j] = 24*eps_ij, sig2 = sigma_ij^2, shift6 = 6*shift_ij,
 * C = total number of LJ centres over all components. Values are FP64 and are cast to the calc type inside. */
int ls1b200_set_components(ls1b200_ctx* ctx, uint32_t n_components, const ls1b200_component* comps,
                           uint32_t n_sites, const ls1b200_site* sites,
                           uint32_t n_lj_centers, const double* eps24, const double* sig2, const double* shift6);

/* ParticleContainer::addParticles (ParticleContainer.h:134): replaces the molecule set. cid is 0-based.
 * q may be NULL (identity), L may be NULL (zero). Molecules outside [box_min, box_max) are kept and treated
 * exactly like the reference treats them at the next update (wrapped if periodic). */
int ls1b200_upload(ls1b200_ctx* ctx, uint64_t n, const uint64_t* id, const uint32_t* cid,
                   const double* r, const double* v, const double* q, const double* L);
uint64_t ls1b200_num_molecules(const ls1b200_ctx* ctx);
int ls1b200_download(ls1b200_ctx* ctx, int field, void* dst, uint64_t n);
/* Asynchronous variant for per-step consumers (the host-coherence layer of plugins / writers, FullMolecule fields as in
 * io/CheckpointWriter): _begin enqueues the gather behind the kernels already submitted and starts the copy to `dst`
 * (pinned memory for real overlap) on a side stream, so that e.g. the positions of this step travel to the host while the
 * force traversal runs; `dst` is borrowed until _end returns. One download in flight per context; R, V, Q, L, F, M, VI;
 * not on decomposed contexts. */
int ls1b200_download_begin(ls1b200_ctx* ctx, int field, void* dst, uint64_t n);
int ls1b200_download_end(ls1b200_ctx* ctx);
/* Overwrite one per-molecule field (R, V, Q, L or F, M) from host memory: the lazy host->device sync the adapter
 * needs when a plugin modified molecules on the host. */
int ls1b200_set_field(ls1b200_ctx* ctx, int field, const void* src, uint64_t n);

/* LinkedCells geometry as the reference computes it (LinkedCells.cpp:47-58 / :148-166):
 * cells per dimension incl. the halo layer, cell length, halo bounding box min. */
int ls1b200_geometry(const ls1b200_ctx* ctx, int32_t cells_per_dim[3], double cell_length[3], double halo_min[3]);

/* == LinkedCells::update() + DomainDecompBase::exchangeMolecules() + LinkedCells::updateMoleculeCaches()
 * (Simulation.cpp:1417-1455): wraps leaving molecules (periodic dims), assigns every molecule to its cell
 * bit-exactly like CellBorderAndFlagManager::getBorder (CellBorderAndFlagManager.h:134-152), creates the
 * periodic halo copies (DomainDecompBase.cpp:317-376) and rebuilds the cell-sorted SoA. */
int ls1b200_update(ls1b200_ctx* ctx);

/* == ParticleContainer::traverseCells(VectorizedCellProcessor&) + Simulation::updateForces (calcFM)
 * (Simulation.cpp:1161, :1178; LinkedCells.cpp:565-575; FullMolecule.cpp:548-626). Writes per-molecule F, M, Vi
 * on the device; *out receives the macroscopic sums (may be NULL). */
int ls1b200_traverse_cells(ls1b200_ctx* ctx, ls1b200_macro* out);

/* == Leapfrog::eventNewTimestep -> transition1to2 -> upd_preF(dt) (Leapfrog.cpp:45-61, FullMolecule.cpp:331-361). */
int ls1b200_leapfrog_pre(ls1b200_ctx* ctx, double dt);
/* == Leapfrog::eventForcesCalculated -> transition2to3 -> upd_postF(dt/2) + sums (Leapfrog.cpp:63-147). */
int ls1b200_leapfrog_post(ls1b200_ctx* ctx, double dt, ls1b200_kinetic* out);
/* == VelocityScalingThermostat::apply, global variant (thermostats/VelocityScalingThermostat.cpp:68-89). */
int ls1b200_scale_velocities(ls1b200_ctx* ctx, double beta_trans, double beta_rot);
/* == the per-molecule loop of Domain's "explosion" heuristics (Domain.cpp:266-296; a thermostat factor below MIN_BETA = 0.9): every
 * molecule whose translational kinetic energy exceeds KINLIMIT_PER_T * T = 10 T is scaled down to that limit together with its
 * force, every molecule with more than 3 rotDOF T of rotational energy likewise with its angular momentum and torque. For
 * phase-by-phase drivers whose host Domain takes the decision (the adapter); ls1b200_simulate evaluates the whole heuristics
 * (decision, counters, beta_trans = 1, beta_rot^0.0091) on the device. */
int ls1b200_selective_thermostat(ls1b200_ctx* ctx, double target_temperature);

/* The whole hot path of Simulation::simulate for n_steps steps with the NVT velocity-scaling thermostat
 * (target_temperature > 0) or NVE (<= 0), without any host round trip inside the loop:
 *   leapfrog_pre -> update -> traverse_cells -> leapfrog_post -> Domain::calculateGlobalValues -> thermostat.
 * Forces must be valid on entry (call ls1b200_update + ls1b200_traverse_cells once first, as
 * Simulation::prepare_start does, Simulation.cpp:896-928). records (may be NULL) receives n_steps entries.
 * global_n / global_volume: N and V of the whole system (for rho in p); pass 0 to use this context's own.
 * Domain::calculateGlobalValues is evaluated in full (Domain.cpp:232-314, :363-368), i.e. with its heuristic handling of an
 * "explosion": while a velocity-scaling factor is below MIN_BETA = 0.9 and no error is pending, molecules above 10 T of
 * translational (3 rotDOF T of rotational) kinetic energy are scaled down to that limit with their force (torque), beta_trans is
 * reported and applied as 1 and beta_rot as its 0.0091-th power, and the warning / error counters run (a persistent or returning
 * explosion ends the special handling). The first call after an upload evaluates the global values once on the state as it is,
 * like Simulation::prepare_start (Simulation.cpp:986-992): no factor is applied there, but the heuristics start. The reference's
 * own CubicGridGenerator starts every rotating model inside these heuristics (all thermal energy in the translation).
 * (The reduced-memory loop, ls1b200_rmm_simulate, applies the plain factors only.) */
int ls1b200_simulate(ls1b200_ctx* ctx, uint64_t n_steps, double dt, double target_temperature,
                     uint64_t global_n, double global_volume, ls1b200_step_record* records);

/* ---- multi-GPU: 3D block decomposition (replaces parallel/DomainDecomposition.cpp, DomainDecompMPIBase.cpp and
 * CommunicationPartner.cpp; one process and one context per GPU) ----
 * The box is cut into grid[0] x grid[1] x grid[2] equal blocks (DomainDecomposition::getBoundingBoxMin/Max,
 * DomainDecomposition.cpp:130-139); rank = (coords[0]*grid[1] + coords[1])*grid[2] + coords[2] (MPI_Cart order).
 * The context of a rank is created with ITS block as box_min/box_max and periodic[d] = (grid[d] == 1): dimensions that
 * are not cut wrap locally (DomainDecompBase), cut dimensions get their halo from the neighbour ranks.
 * Per step, between leapfrog_pre and update, leaving molecules and halo copies are exchanged dimension by dimension
 * (x, then y, then z; copies received in one stage are forwarded by the next, which covers edges and corners) --
 * == DomainDecompMPIBase::exchangeMoleculesMPI(LEAVING_AND_HALO_COPIES) with the indirect neighbour scheme. Positions that
 * cross the global periodic boundary are shifted by the global box length with the reference's clamps
 * (CommunicationPartner.cpp:526-563). Full-shell halo: no force exchange (C08CellPairTraversal.h:35). */
typedef struct ls1b200_decomp {
	int32_t rank, n_ranks;
	int32_t grid[3];
	int32_t coords[3];
	double global_min[3], global_max[3];   /* Domain: the whole periodic box                                          */
	uint8_t nccl_unique_id[128];           /* ncclUniqueId of ls1b200_nccl_unique_id(), identical on all ranks;
	                                          ignored when use_nccl == 0                                             */
	int32_t use_nccl;                      /* 1: the library owns an NCCL communicator and ls1b200_exchange /
	                                          ls1b200_simulate do the transport (NVLink); 0: the caller moves the
	                                          messages itself with ls1b200_halo_export / ls1b200_halo_import*         */
	int32_t reserved[3];
} ls1b200_decomp;

/* Message record == ParticleDataFull (parallel/ParticleDataFull.h:30-35: id, cid, r, v, q, D). */
typedef struct ls1b200_particle {
	uint64_t id;
	uint32_t cid;
	uint32_t flags;         /* bit0: 1 = halo copy (not owned by the receiver), 0 = leaving molecule (ownership moves) */
	double r[3], v[3], q[4], L[3];
} ls1b200_particle;

/* ncclGetUniqueId on the calling rank (rank 0 calls it and broadcasts the 128 bytes with whatever it has: MPI,
 * torch.distributed, a file). NCCL is loaded at run time (dlopen libnccl.so.2); fails loudly if it is absent. */
int ls1b200_nccl_unique_id(uint8_t out[128]);
/* Must be called after ls1b200_set_components and before ls1b200_upload; collective over all ranks when use_nccl == 1
 * (ncclCommInitRank). After it, ls1b200_upload takes only the molecules of this rank's block, ids travel with the
 * molecules, and ls1b200_download returns the molecules this rank owns NOW, in slot order (download LS1B200_F_ID
 * to match them; ls1b200_num_molecules gives the count). */
int ls1b200_decomp_attach(ls1b200_ctx* ctx, const ls1b200_decomp* dd);
/* Peer-memory transport (the GPUs of one NVLink / NVSwitch domain, one process per GPU): after ls1b200_upload every rank
 * publishes the CUDA IPC handle of its receive buffers (ls1b200_peer_handle), the caller gathers the n_ranks handles with
 * whatever it has (MPI_Allgather, torch.distributed.all_gather) and hands the table to ls1b200_peer_attach, which maps the
 * buffers of the (at most six) neighbours. From then on ls1b200_exchange / ls1b200_simulate do not call NCCL for the
 * molecule exchange: the packing kernel stores the records straight into the neighbour's receive buffer over NVLink and
 * publishes {sequence number, count} when the last record is out; the import kernel of the neighbour waits for that
 * sequence number (bounded wait: a vanished peer is reported as an error, not a hang). Replaces DomainDecompMPIBase::exchangeMoleculesMPI's MPI_Isend / Irecv
 * (parallel/CommunicationPartner.cpp:140-330) for the ranks of one box; a new ls1b200_upload invalidates the mapping. */
#define LS1B200_PEER_HANDLE_BYTES 96
int ls1b200_peer_handle(ls1b200_ctx* ctx, uint8_t out[LS1B200_PEER_HANDLE_BYTES]);
int ls1b200_peer_attach(ls1b200_ctx* ctx, const uint8_t* handles /* [n_ranks][LS1B200_PEER_HANDLE_BYTES] */, int32_t n_ranks);
/* Unmaps the neighbours' buffers (back to the NCCL transport, if any). Every rank calls it -- followed by a barrier of the
 * caller's -- BEFORE any rank destroys its context or uploads again: memory must not be freed while a peer has it mapped. */
int ls1b200_peer_detach(ls1b200_ctx* ctx);
/* One exchange (all stages) on the context's stream -- peer-memory transport if attached, else NCCL send / recv; collective.
 * The NCCL transport sends fixed-capacity messages; the capacity is agreed over all ranks (the maximum of what each rank
 * derives from its own upload: one ncclAllReduce inside ls1b200_upload, which is therefore collective on contexts with
 * use_nccl = 1), so blocks of different population are fine; the peer-memory transport honours the capacity of every receiver.
 * Failure handling: a message that does not fit, a slot overflow or a neighbour that does not publish its message within
 * LS1B200_PEER_TIMEOUT_S seconds (environment, default 60) marks the step as failed ON THE DEVICE: from then on the
 * integration and export kernels leave the state untouched, the flag travels to every rank with the allreduce of the
 * macroscopic sums, and ls1b200_simulate / the next synchronising call returns the error on all ranks.
 * ls1b200_simulate calls it every step; phase-by-phase drivers call it between ls1b200_leapfrog_pre and ls1b200_update
 * (and once before the first update). */
int ls1b200_exchange(ls1b200_ctx* ctx);
/* Sums the macroscopic values of the last traversal / leapfrog_post over all ranks (ncclAllReduce; replaces the
 * collective communication of Domain::calculateGlobalValues, Domain.cpp:176-182, :220-230); collective. */
int ls1b200_allreduce_macro(ls1b200_ctx* ctx, ls1b200_macro* macro, ls1b200_kinetic* kinetic);

/* Transport-agnostic stage (use_nccl == 0, or MPI inside ls1): packs on the device what has to go to the neighbour on
 * `side` (0 low, 1 high) of dimension `dim` -- leaving molecules (they stay behind as this rank's halo copy) and halo
 * copies, positions already shifted for the periodic wrap -- and returns the device buffer of ls1b200_particle records
 * and their count. Both sides of a dimension must be exported before anything of that stage is imported. The buffer
 * stays valid until the next export for the same side. */
int ls1b200_halo_export(ls1b200_ctx* ctx, int32_t dim, int32_t side, void** device_buf, uint64_t* count);
/* Same, with the records copied to host memory; if host_cap is too small the call fails and *count holds the need --
 * NB: the export has then already happened (leaving molecules were converted), so size the buffer generously. */
int ls1b200_halo_export_host(ls1b200_ctx* ctx, int32_t dim, int32_t side, ls1b200_particle* host_buf, uint64_t host_cap,
                             uint64_t* count);
/* Receive side: appends leaving molecules as owned ones and halo copies as halo molecules; they are binned by the next
 * ls1b200_update. device_buf is device memory of this context's device. */
int ls1b200_halo_import(ls1b200_ctx* ctx, const void* device_buf, uint64_t count);
/* Same, from host memory (tests, MPI without GPU-direct). */
int ls1b200_halo_import_host(ls1b200_ctx* ctx, const ls1b200_particle* host_buf, uint64_t count);

/* ---- measurement helpers (bench.py) ---- */
/* The reference's CubicGridGenerator on the device (io/CubicGridGeneratorInternal.cpp:62-123, ParticleCellBase::initCubicGrid,
 * particleContainer/ParticleCellBase.cpp:51-159): a body-centred cubic lattice of 2 * nx * ny * nz molecules of component 0 over
 * the whole simulation box [0, sim_box) -- two sub-lattices at 1/4 and 3/4 of the spacing, coordinates created in the calc type
 * with one fused multiply-add, as the reference binary evaluates them -- of which this context keeps those inside ITS box
 * (min <= r < max): every rank of a decomposition generates its own block, no host array of the system ever exists.
 * Velocities: |v| = sqrt((3 + rotDOF) T / m), direction from two uniform numbers (EqualVelocityAssigner.h:16-29);
 * q = (1, 0, 0, 0), L = 0; id = lattice index + 1. `flags`:
 *   LS1B200_GEN_REMOVE_MOMENTUM        IOHelpers::removeMomentum (io/IOHelpers.cpp:8-115): subtracts sum(m v) / sum(m) over all ranks,
 *                                      then scales the velocities back to `temperature` (collective on contexts with a communicator);
 *   LS1B200_GEN_REFERENCE_VELOCITIES   the reference's own velocity sequence: initCubicGrid seeds a fresh std::mt19937(0) in every
 *                                      LinkedCells cell, so the k-th molecule inserted into any cell gets the k-th direction of one
 *                                      sequence. With both flags the state is the one the unmodified reference binary generates from
 *                                      the same XML (tests/golden/generator_1clj_2662.npz: its step-0 checkpoint and 20-step trace);
 *                                      it depends on the cell grid, i.e. on the decomposition, exactly as the reference's does.
 *                                      Without this flag the uniform numbers come from a counter-based generator keyed by `seed` and
 *                                      the molecule's global lattice index (the state does not depend on the decomposition).
 * Replaces ls1b200_upload: afterwards the context is in the same state as after an upload of these molecules. */
#define LS1B200_GEN_REMOVE_MOMENTUM 1
#define LS1B200_GEN_REFERENCE_VELOCITIES 2
int ls1b200_generate_cubic_grid(ls1b200_ctx* ctx, const uint64_t mols_per_dim[3], const double sim_box[3], double temperature,
                                uint64_t seed, int32_t flags);

/* ---- reduced-memory mode (the reference's RMM build: ENABLE_REDUCED_MEMORY_MODE, PRECISION=SINGLE) -------------------------------------
 * The all-single-precision 1CLJ path: a molecule is 32 bytes (FP32 position, FP32 velocity, id: molecules/MoleculeRMM.h), the force
 * goes straight into the velocities (particleContainer/adapter/VCP1CLJRMM.cpp:238-369: v += f dt / m), positions advance with
 * r = fma(dt, v, r) (integrators/PositionCellProcessorRMM.h:43-58), sum v^2 is not mass weighted (VelocityCellProcessorRMM.h:119-127).
 * Needs a mixed-precision, single-component 1CLJ, non-decomposed context; replaces ls1b200_upload / _simulate on it.
 *   ls1b200_rmm_upload    FP64 host arrays are stored as FP32 (the reference's reader does the same into MoleculeRMM)
 *   ls1b200_rmm_start     Simulation::prepare_start: binning + halo, the initial force applied with dt / 2, global values (the thermostat
 *                         factor is reported, not applied); upot_corr / virial_corr = Domain's long-range corrections (Homogeneous.cpp)
 *   ls1b200_rmm_simulate  n_steps of: positions, binning + halo, forces into the velocities, sum v^2, global values, velocity scaling
 *   ls1b200_rmm_download  the owned molecules (id, r, v as FP64), order unspecified                                                   */
int ls1b200_rmm_upload(ls1b200_ctx* ctx, uint64_t n, const uint64_t* id, const double* r /* [n][3] */, const double* v /* [n][3] */);
int ls1b200_rmm_start(ls1b200_ctx* ctx, double dt, double target_temperature, double upot_corr, double virial_corr, ls1b200_step_record* record);
int ls1b200_rmm_simulate(ls1b200_ctx* ctx, uint64_t n_steps, double dt, double target_temperature, double upot_corr, double virial_corr,
                         ls1b200_step_record* records /* [n_steps] or NULL */);
int ls1b200_rmm_download(ls1b200_ctx* ctx, uint64_t* id, double* r, double* v, uint64_t n);
uint64_t ls1b200_rmm_state_bytes(const ls1b200_ctx* ctx);      /* 32 * number of molecules */

/* Device time (ms) spent in each phase since the last reset, measured with CUDA events on the library's stream:
 * [0] update (binning+halo+SoA), [1] force kernels, [2] leapfrog pre, [3] leapfrog post + reductions. */
int ls1b200_phase_times(ls1b200_ctx* ctx, double ms[4], uint64_t launches[4], int reset);
/* Decomposed contexts inside ls1b200_simulate: the exchange (export -> NVLink / NCCL -> import, three stages) runs on its own stream
 * beside the integration and binning of the slots it cannot touch (LS1B200_OVERLAP=0 puts it back in line, inside the update
 * phase). Device time (ms) of those overlapped exchanges since the last reset; it is NOT part of the four phases above. */
int ls1b200_exchange_overlapped_ms(ls1b200_ctx* ctx, double* ms, int reset);
/* Counts of the reference's FlopCounter for the current configuration (adapter/FlopCounter.cpp:20-74):
 * molecule-distance evaluations and in-range pairs per potential, Newton-3 counting (each pair once). */
typedef struct ls1b200_flopcount {
	double mol_dist;        /* candidate molecule pairs the C08 traversal evaluates                             */
	double lj, charge, charge_dipole, dipole, charge_quadrupole, dipole_quadrupole, quadrupole; /* site pairs   */
	double lj_macro, charge_macro, charge_dipole_macro, dipole_macro, charge_quadrupole_macro,
	       dipole_quadrupole_macro, quadrupole_macro;
	double flops;           /* the FlopCounter total                                                            */
} ls1b200_flopcount;
int ls1b200_count_flops(ls1b200_ctx* ctx, ls1b200_flopcount* out);
/* FP32 FMA micro-benchmark on this device: returns TFLOP/s (2 flops per FMA), the denominator of the
 * "% of FP32 pair-rate peak" metric (MEASURED_PEAKS.json has no FP32 figure). */
int ls1b200_measure_fp32_peak(int device, double* tflops);
/* The same with FP64 fused multiply-adds: the denominator for the full-FP64 variant (the reference's double build). */
int ls1b200_measure_fp64_peak(int device, double* tflops);
int ls1b200_measure_hbm_copy(int device, double* gbs);

#ifdef __cplusplus
}
#endif
#endif /* LS1B200_H */
