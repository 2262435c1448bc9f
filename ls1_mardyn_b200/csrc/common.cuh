// common.cuh -- shared declarations of libls1b200.so (sm_100a). Host side C++17, device side CUDA.
//
// Data layout in HBM (DESIGN.md "Data layout"):
//   * molecules live in SLOTS: one slot per owned molecule and per periodic/neighbour halo copy, sorted by LinkedCells
//     cell index (z-major, y, x fastest -- LinkedCells::cellIndexOf3DIndex, particleContainer/LinkedCells.cpp:925-927),
//     so that the three x-neighbour cells of any cell are ONE contiguous slot range;
//   * molecule state is FP64 structure-of-arrays in component planes (r[d*cap + slot], ...), double buffered: the
//     counting sort of every update() permutes it from `cur` into `nxt`;
//   * `mpos` is the calc-precision (float4 / double4) copy of the molecule centre + component id + halo flag that the pair
//     kernels read (== CellDataSoA::_mol_pos, adapter/CellDataSoA.h:41-74);
//   * multi-site components additionally get `spos`/`sori` (site position + parameter, site orientation), smax entries
//     per slot.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/ls1b200.h"

namespace ls1b200 {

// ---------------------------------------------------------------------------------------------------------------------
// errors (never throw across the C ABI)
// ---------------------------------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(const std::string& msg);
extern std::atomic<uint64_t> g_launches;

#define LS1_CUDA(call)                                                                                             \
	do {                                                                                                             \
		cudaError_t e__ = (call);                                                                                      \
		if (e__ != cudaSuccess)                                                                                        \
			return ::ls1b200::fail(std::string(#call) + ": " + cudaGetErrorString(e__) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
	} while (0)

#define LS1_TRY(call)          \
	do {                         \
		int rc__ = (call);         \
		if (rc__ != 0) return rc__; \
	} while (0)

// kernel launch with accounting (ls1b200_kernel_launch_count) and launch-error check
#define LS1_LAUNCH(ctx, phase, kernel, grid, block, smem, ...)                                     \
	do {                                                                                             \
		kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                               \
		::ls1b200::g_launches.fetch_add(1, std::memory_order_relaxed);                                 \
		(ctx)->phase_launches[(phase)]++;                                                              \
		LS1_CUDA(cudaGetLastError());                                                                  \
	} while (0)

enum { PH_UPDATE = 0, PH_FORCE = 1, PH_PRE = 2, PH_POST = 3, PH_COUNT = 4, PH_XCHG = 4, PH_INTERNAL = 5 };      // PH_XCHG: the exchange when it runs on its own stream beside the step (not part of ls1b200_phase_times' four)

// ---------------------------------------------------------------------------------------------------------------------
// geometry of LinkedCells, passed by value to kernels
// ---------------------------------------------------------------------------------------------------------------------
struct Geom {
	double bmin[3], bmax[3], hmin[3], hmax[3], clen[3], crecip[3], lbox[3];
	double wrap_hi[3];           // value that replaces x >= bmax after a +L wrap (DomainDecompBase.cpp:236-240)
	double halo_hi[3];           // value for a +L halo copy that landed < bmax   (DomainDecompBase.cpp:364-368)
	double halo_lo[3];           // value for a -L halo copy that landed >= bmin  (DomainDecompBase.cpp:358-362)
	double cutoff;
	double lo_band[3], hi_band[3];   // bmin + cutoff, bmax - cutoff: the layers that need periodic halo copies (DomainDecompBase.cpp:329)
	int cells[3];                // cells per dimension incl. the halo layer
	int periodic[3];
	uint32_t ncells;             // cells[0]*cells[1]*cells[2]; bin ncells is the "limbo" bin (outside the halo box)
	int pad;
};

constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;
constexpr uint64_t META_HALO = 1ull << 63;
// a halo copy that arrived (or was retained) during the halo exchange of THIS step: unlike a stale halo slot of the previous
// sort it takes part in the next binning (and is forwarded by the later stages of the exchange); always set with META_HALO
constexpr uint64_t META_FRESH = 1ull << 62;
__host__ __device__ inline bool meta_fresh(uint64_t m) { return (m & META_FRESH) != 0; }
__host__ __device__ inline uint32_t meta_orig(uint64_t m) { return (uint32_t) (m & 0xFFFFFFFFull); }
__host__ __device__ inline uint32_t meta_cid(uint64_t m) { return (uint32_t) ((m >> 32) & 0xFFFFull); }
__host__ __device__ inline bool meta_halo(uint64_t m) { return (m & META_HALO) != 0; }
__host__ __device__ inline uint64_t make_meta(uint32_t orig, uint32_t cid, bool halo) {
	return (uint64_t) orig | ((uint64_t) (cid & 0xFFFFu) << 32) | (halo ? META_HALO : 0ull);
}

// device view of one component (molecules/Component.h); site rows follow in the site table
struct CompDev {
	uint32_t n[4];               // LJ, charge, dipole, quadrupole
	uint32_t site_offset, lj_lookup_id, rot_dof, nsites;
	double mass, inv_mass;
	double I[3], invI[3];        // invI = 0 where I == 0 (FullMolecule ctor, molecules/FullMolecule.cpp:33-40)
};
struct SiteDev {
	double r[3], e[3], param;
	uint32_t type;               // 0 LJ, 1 charge, 2 dipole, 3 quadrupole
	uint32_t ljid;               // lj_lookup_id + j for LJ centres
};

// one double-buffered half of the molecule state
struct State {
	double* r = nullptr;         // [3][cap]
	double* v = nullptr;         // [3][cap]
	double* q = nullptr;         // [4][cap]  (only if the ensemble rotates)
	double* L = nullptr;         // [3][cap]
	uint64_t* meta = nullptr;    // [cap] orig index | cid << 32 | fresh << 62 | halo << 63
	uint64_t* id = nullptr;      // [cap] molecule id, travels with the molecule (decomposed mode only; otherwise ids stay in d_id)
};

// Which slots a kernel of the overlapped multi-GPU step works on. A slot is BOUNDARY if its cell at the last sort lies within four
// layers (cells 0 .. 3, n-4 .. n-1) of a face the block shares with another rank, or in limbo, or if it was appended after the
// sort (an import): exactly the slots k_export reads or writes and k_import appends. Everything else is INNER: nothing the exchange
// does can touch it, so its integration and its histogram run on the main stream while the exchange runs beside them.
struct Region {
	const uint8_t* flag;          // per slot, written by the sort (k_gather): 1 = boundary cell (see above), 0 = inner
	int mode;                     // 0: all slots, 1: boundary slots only, 2: inner slots only, 3: only the slots appended behind the sorted ones (needs no flag)
};
__device__ __forceinline__ bool region_selected(const Region& rg, uint32_t slot, uint32_t n_sorted) {
	if (rg.mode == 0) return true;
	if (rg.mode == 3) return slot >= n_sorted;
	const bool boundary = slot >= n_sorted || rg.flag[slot] != 0;      // one byte per slot: the kernels the step is split into stay bandwidth bound
	return boundary == (rg.mode == 1);
}
// the definition of a boundary cell (evaluated once per slot and sort)
__device__ __forceinline__ bool cell_is_boundary(uint32_t c, uint32_t ncells, int nx, int ny, int nz, const int* cut) {
	if (c >= ncells) return true;
	const int cx = (int) (c % (uint32_t) nx), cy = (int) ((c / (uint32_t) nx) % (uint32_t) ny), cz = (int) (c / (uint32_t) (nx * ny));
	return (cut[0] && (cx <= 3 || cx + 4 >= nx)) || (cut[1] && (cy <= 3 || cy + 4 >= ny)) || (cut[2] && (cz <= 3 || cz + 4 >= nz));
}

// scalars living on the device between kernels of one step
struct DevScalars {
	double upot6lj, upot_xpoles, virial, my_rf;   // force traversal
	double summv2, sum_iw2;                       // leapfrog post
	double fault;                                 // != 0: some rank reported overflow (summed over the ranks with the six doubles above)
	unsigned long long n_mol, rot_dof;
	double beta_trans, beta_rot;                  // thermostat factors applied by the next leapfrog_pre
	uint32_t n_slots2[2];                         // slots in use (owned + halo), ping-pong: [ctx->parity] is current
	uint32_t overflow;                            // 1: an update needed more than cap slots, 2: a halo message buffer was too small, 4: a peer timed out, 5: another rank failed.
	                                              // Once set, the integration kernels leave the state untouched (the step loop is frozen) until the host reports it
	uint32_t n_owned_dev;                         // owned molecules counted by the last count_owned() (decomposed mode)
	uint32_t n_sorted;                            // slots in cell order (count of the last update); imports of the exchange follow behind
	uint32_t n_pre;                               // fused step: slots binned by k_pre_key (the imports of this step's exchange are appended behind them)
	// Domain's "explosion" heuristics (Domain.cpp:74-79, :259-314, :363-368), evaluated by k_globals: while a thermostat factor is below
	// MIN_BETA = 0.9 and no error is pending, every molecule whose translational (rotational) kinetic energy exceeds 10 T (3 rotDOF T) is
	// scaled down to that limit together with its force (torque). clip_limit > 0: that scaling is pending; like the thermostat factors
	// it is applied on load by the next kernel that reads the velocities (k_pre_key, k_leapfrog_pre*, k_scale). The three counters are
	// _universalSelectiveThermostatCounter / Warning / Error.
	double clip_limit, clip_ti;
	int32_t expl_counter, expl_warning, expl_error, expl_pad;
};

// the pending selective-thermostat scaling of one molecule (Domain.cpp:268-296): factor for v and F ...
__device__ __forceinline__ double clip_factor_trans(double limit, double mass, double vx, double vy, double vz) {
	const double utrans = 0.5 * mass * (vx * vx + vy * vy + vz * vz);      // FullMolecule::U_trans
	return utrans > limit ? sqrt(limit / utrans) : 1.0;
}

}  // namespace ls1b200

// the opaque context of include/ls1b200.h
struct ls1b200_ctx {
	ls1b200_config cfg{};
	ls1b200::Geom geom{};
	int device = 0;
	cudaStream_t stream = nullptr;
	bool mixed = true;

	// components
	std::vector<ls1b200::CompDev> comps;
	std::vector<ls1b200::SiteDev> sites;
	ls1b200::CompDev* d_comps = nullptr;
	ls1b200::SiteDev* d_sites = nullptr;
	uint32_t n_lj = 0;
	std::vector<double> h_eps24, h_sig2, h_shift6;
	double* d_eps24 = nullptr; double* d_sig2 = nullptr; double* d_shift6 = nullptr;      // FP64 tables [n_lj*n_lj]
	float* d_eps24f = nullptr; float* d_sig2f = nullptr; float* d_shift6f = nullptr;      // cast to calc (mixed)
	uint32_t smax = 1;            // max sites per molecule
	bool all_1clj = false;        // every component: exactly one LJ centre at the origin and nothing else
	bool rotates = false;         // some component has rotational degrees of freedom
	bool components_set = false;

	// molecules
	uint64_t n_owned = 0;         // molecules uploaded
	uint64_t rot_dof_total = 0;   // sum of Component::getRotationalDegreesOfFreedom() over the owned molecules
	uint64_t cap = 0;             // slot capacity of every per-slot array
	uint32_t n_slots = 0;         // host copy (valid after a synchronising call)
	uint32_t parity = 0;          // which DevScalars::n_slots2 entry is current
	uint64_t* d_id = nullptr;     // [n_owned] molecule id, upload order (never permuted)
	ls1b200::State cur, nxt;
	double* F = nullptr; double* M = nullptr; double* Vi = nullptr;   // [3][cap] each, slot order
	bool sorted = false;          // update() has run since the last upload
	bool ever_sorted = false;     // some update() has run since the last upload (cell_of2 / DevScalars::n_sorted describe it)
	bool globals_initialised = false;      // ls1b200_simulate has evaluated the global values of the uploaded state once (Simulation::prepare_start)
	bool forces_valid = false;

	// cells
	uint32_t* cell_count = nullptr;   // [ncells+2]
	uint32_t* cell_start = nullptr;   // [ncells+2] exclusive scan (reference cells)
	uint32_t* cell_start_fine = nullptr;   // [ncells*fine_s+2] exclusive scan over the fine x-bins (fine_s > 1)
	uint32_t fine_s = 1;              // fine x-bins per reference cell: the in-cell order of the slots is by x sub-bin
	double fine_inv = 0.;             // fine_s / cell length in x
	uint32_t* cursor = nullptr;       // [ncells+2] next free slot of the halo-copy part of every bin (k_scatter)
	uint32_t* cell_count_halo = nullptr;   // [ncells+2] periodic halo copies per bin (cell_count holds the non-copies, whose rank k_key already knows)
	uint32_t* rank = nullptr;         // [cap] provisional rank of a slot's molecule inside its bin (returned by the histogram atomic)
	double* d_border = nullptr;       // [3][border_stride] cell borders (bit-identical to CellBorderAndFlagManager::getBorder)
	uint32_t border_stride = 0;
	uint32_t* scan_tmp = nullptr;
	uint32_t* cell_start_alt = nullptr;      // overlapped multi-GPU step: cell_start is double buffered (the exchange reads the previous order's while the sort writes the next)
	uint32_t* cell_count_b = nullptr; uint32_t* cell_start_b = nullptr; uint32_t* cursor_b = nullptr; uint32_t* scan_tmp_b = nullptr;      // bins of the imports (overlapped multi-GPU step; one allocation, update.cu)
	uint32_t* key = nullptr;          // [cap] cell | shiftmask << 26 per source slot
	uint32_t* entry = nullptr;        // [cap] src << 5 | shift code, per provisional slot
	uint32_t* cell_of = nullptr;      // [cap] cell of provisional / final slot
	uint32_t* cell_of2 = nullptr;     // [cap]
	uint8_t* region_flag = nullptr;   // [cap] decomposed contexts: 1 = slot in a boundary cell of a cut dimension (Region)

	// calc-precision copies
	void* mpos = nullptr;             // float4 / double4 [cap]
	void* spos = nullptr;             // [cap*smax] float4 / double4: site position + parameter
	void* sori = nullptr;             // [cap*smax] float4 / double4: site orientation
	double* sf = nullptr;             // [cap*smax][9]: per-site force, virial, torque (general path)

	// chunk list of the tiled 1CLJ kernel (force_1clj_tiled.cu)
	void* chunks = nullptr; uint32_t* chunk_row_count = nullptr; uint32_t* chunk_row_start = nullptr; uint32_t* work_counter = nullptr;
	void* chunks_b = nullptr; uint32_t* chunk_row_count_b = nullptr; uint32_t* chunk_row_start_b = nullptr; uint32_t* work_counter_b = nullptr;      // boundary chunks of the overlapped multi-GPU step (tile.cuh); chunks_b_nrows: rows of that list
	uint32_t chunks_b_nrows = 0; bool chunks_b_ready = false;
	int sm_count = 148;
	// reduced-memory mode (rmm.cu): FP32 planes r[3][cap], v[3][cap], id[cap] (bit 63: halo copy), double buffered
	float* rmm_r[2] = {nullptr, nullptr}; float* rmm_v[2] = {nullptr, nullptr}; uint64_t* rmm_id[2] = {nullptr, nullptr};
	int rmm_cur = 0; bool rmm = false;
	// overlapped multi-GPU step (ls1b200_simulate on decomposed contexts): the exchange runs on its own stream beside the inner slots' work
	cudaStream_t stream_x = nullptr; cudaEvent_t ev_xa = nullptr, ev_xb = nullptr; int overlap = 0; int xchg_phase = 0; int split_test = 0;      // overlap: LS1B200_OVERLAP=1 (measured neutral to slower, profiles/r02_multi_gpu.md: off by default)
	void* chunks2 = nullptr; size_t chunks2_bytes = 0;      // molecule-granular chunk list of force_1clj_v3.cu
	int pf_check = 0;                 // LS1B200_PF_CHECK=1: force_1clj_pf verifies that every window translation is exact (overflow code 6)
	int lj_kernel = 0;                // LS1B200_LJ_KERNEL: 0 tiled (default; shared-memory windows, TMA), 1 "global" (force_1clj.cu), 3 "mma" / 4 "tc": tensor-core prefilter experiments (force_1clj_mma.cu, force_1clj_tc.cu; measured slower, profiles/r01_tensor_prefilter.md), 5 "v3": molecule-granular chunks + x-trimmed ranges over fine bins (force_1clj_v3.cu; measured slower, profiles/r02_force_kernel_experiments.md)
	int multi_pd2 = 0;                // LS1B200_MULTI_PD2=1: prefetch the site records of two neighbours ahead instead of one
	int multi_stage = 0;              // LS1B200_MULTI_STAGE=0: do not stage neighbour site records in shared memory (force_multi_tiled.cu)
	int multi_max_ctas = 0;           // LS1B200_MULTI_MAXCTAS: resident CTAs per SM of the multi-site kernel (0: as many as fit)
	int multi_kernel = 0;             // 0: tiled thread-per-molecule (force_multi_tiled.cu), 1: v1 thread-per-site (force_general.cu)

	// reductions
	double* partials = nullptr;       // [n_partial_blocks][8]
	double* partials2 = nullptr;      // [1024][8] second reduction level
	// chunk list of the tiled 1CLJ kernel built beside the sort's scatter / gather (it needs cell_start only): update.cu forks, force_1clj_tiled.cu joins
	cudaStream_t stream_c = nullptr; cudaEvent_t ev_c0 = nullptr, ev_c1 = nullptr; bool chunks_ready = false; uint32_t chunks_nrows = 0; int chunks_early = 1;      // LS1B200_CHUNKS_EARLY=0: inside the force phase
	double* partials_kin = nullptr;   // [n_partial_blocks][4] kinetic partial sums of the fused step (force_1clj_tiled.cu)
	double kin_dth = 0.;              // != 0: the next force traversal also delivers sum m (v + kin_dth / m F)^2 (fused step of ls1b200_simulate)
	int tile_max_ctas = 0, tile_pass1_ctas = 0;      // LS1B200_TILE_MAXCTAS / LS1B200_TILE_PASS1_CTAS: resident CTAs per SM of the tiled 1CLJ kernel (0: what fits)
	int force_pass = 0;               // overlapped multi-GPU step: which chunks the next force traversal visits (force_1clj_tiled.cu)
	int overlap2 = 0;                 // LS1B200_OVERLAP2=1: the exchange on its own stream beside the sort and the inner chunks (measured slower, profiles/r02_multi_gpu.md: off by default)
	cudaEvent_t ev_m = nullptr;       // overlapped step: the main sort has finished (the imports' sort may start)
	// lazy permutation (fused step): the sort wrote perm[] instead of moving r and v; `cur` holds meta / id of the new order, `nxt` r and v of the previous one
	uint32_t* perm = nullptr; bool state_lazy = false; bool lazy_next = false; int lazy = 1;      // LS1B200_LAZY=0: every sort moves the whole state
	bool keys_valid = false;          // ls1b200_leapfrog_pre binned the slots it integrated (k_pre_key): the next ls1b200_update only bins what was appended since
	int fused = 1;                    // LS1B200_FUSED=0: ls1b200_simulate runs the phases as separate kernels
	uint32_t partial_blocks = 0;
	ls1b200::DevScalars* d_scal = nullptr;
	ls1b200::DevScalars* h_scal = nullptr;   // pinned
	ls1b200_step_record* d_records = nullptr; uint64_t records_cap = 0;

	// staging buffer of download / set_field (grow-only: no cudaMalloc / cudaFree per call on the host-coherent path)
	void* stage = nullptr; size_t stage_bytes = 0;
	// ls1b200_download_begin / _end: one download in flight on a side stream while the main stream keeps computing
	cudaStream_t stream2 = nullptr; cudaEvent_t ev_dl = nullptr; void* stage2 = nullptr; size_t stage2_bytes = 0; bool dl_pending = false;

	// timing
	std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;
	struct Span { int phase; cudaEvent_t a, b; };
	std::vector<Span> spans;
	double phase_ms[ls1b200::PH_INTERNAL] = {0, 0, 0, 0, 0};
	uint64_t phase_launches[ls1b200::PH_INTERNAL] = {0, 0, 0, 0, 0};
	bool timing = true;

	// ---- 3D block decomposition (halo.cu; replaces parallel/DomainDecomposition.cpp + CommunicationPartner.cpp) ----
	struct Decomp {
		bool active = false;
		int rank = 0, n_ranks = 1;
		int grid[3] = {1, 1, 1}, coords[3] = {0, 0, 0};
		double gmin[3] = {0, 0, 0}, gmax[3] = {0, 0, 0};
		double wrap_hi[3], halo_hi[3], halo_lo[3];      // clamps against the GLOBAL box (CommunicationPartner.cpp:526-563)
		int nbr[3][2];                                   // neighbour rank per dimension, [0] low side, [1] high side
		void* comm = nullptr;                            // ncclComm_t
		uint64_t rot_dof_global = 0; bool rot_dof_valid = false;
	} dd;
	unsigned char* msg_out[2] = {nullptr, nullptr};    // [low, high] message: 16-byte header (count) + records
	unsigned char* msg_in[2] = {nullptr, nullptr};
	uint64_t msg_cap = 0;                               // records per message
	ls1b200_particle* import_stage = nullptr; uint64_t import_stage_cap = 0;   // host-staged imports (ls1b200_halo_import_host)
	// peer-memory transport (NVLink stores straight into the neighbour's receive buffers; CUDA IPC): [step parity][stage][side]
	unsigned char* pin_base = nullptr; size_t pin_stride = 0;      // my 12 receive buffers, one allocation
	unsigned char* peer_base[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};      // the neighbours' pin_base, mapped
	size_t peer_stride[3][2] = {{0, 0}, {0, 0}, {0, 0}}; uint64_t peer_cap[3][2] = {{0, 0}, {0, 0}, {0, 0}};      // the neighbours' buffer stride / record capacity
	void* peer_opened[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; int peer_opened_rank[6] = {-1, -1, -1, -1, -1, -1};
	bool peer_ok = false;
	uint32_t xseq = 0;                                  // exchange calls so far (the sequence number published with every message)
	uint32_t* xcnt = nullptr;                           // device: {count low, count high, pad, pad}
	cudaEvent_t xprof_ev[4] = {nullptr, nullptr, nullptr, nullptr}; double xprof_acc[3] = {0., 0., 0.}; int xprof_calls = 0;      // LS1B200_EXCHANGE_PROFILE
	uint64_t* ekey = nullptr;                           // [cap] (id << 5 | shift code) per provisional slot: content-based in-cell order
	uint64_t n_owned_local = 0;                         // owned molecules of this rank (host copy, decomposed mode)
	uint64_t rot_dof_local = 0;                         // their rotational degrees of freedom (count_owned)
};

namespace ls1b200 {
// implemented across the .cu files
int ctx_alloc_molecules(ls1b200_ctx* c, uint64_t n_owned, uint64_t cap);
void ctx_free_molecules(ls1b200_ctx* c);
void rmm_release(ls1b200_ctx* c);          // rmm.cu
void ctx_free_molecules(ls1b200_ctx* c);
int span_begin(ls1b200_ctx* c, int phase);
int span_end(ls1b200_ctx* c);
int do_update(ls1b200_ctx* c, int part = 0);      // part 0: everything; 1: histogram of the INNER slots only (overlapped step); 2: the remaining slots + scan, scatter, gather; 3: after do_pre_key (imports + scan, scatter, gather)
int build_chunks_tiled_1clj(ls1b200_ctx* c, uint32_t* nrows, int mode = 0, int set = 0);      // force_1clj_tiled.cu: the chunk list of the production 1CLJ kernel (launches on c->stream)
int do_update_imports(ls1b200_ctx* c);      // overlapped multi-GPU step: sorts the imports behind the slots of do_update(c, 4) (update.cu)
int do_pre_key(ls1b200_ctx* c, double dt, bool use_beta, bool post_kick);      // fused step: [post kick +] thermostat + pre kick + drift + binning (update.cu)
int do_forces(ls1b200_ctx* c);
int do_leapfrog_pre(ls1b200_ctx* c, double dt, bool use_beta, int region = 0);      // region: see Region::mode
int do_leapfrog_post(ls1b200_ctx* c, double dt);
int do_globals(ls1b200_ctx* c, double target_T, uint64_t global_n, double global_volume, ls1b200_step_record* d_rec, bool init = false);
int force_general(ls1b200_ctx* c);
int force_1clj(ls1b200_ctx* c);
int force_1clj_tiled(ls1b200_ctx* c);
int force_1clj_v3(ls1b200_ctx* c);          // round-2 experiment (force_1clj_v3.cu)
int force_1clj_pf(ls1b200_ctx* c);          // prefilter kernel, single-component FP32 (force_1clj_pf.cu)
int force_1clj_pf_rmm(ls1b200_ctx* c, float* v, float dt_over_m);
int force_1clj_mma(ls1b200_ctx* c);
int force_1clj_tc(ls1b200_ctx* c);
int force_multi_tiled(ls1b200_ctx* c);
int reduce_force_partials(ls1b200_ctx* c, uint32_t nblocks);
int do_exchange(ls1b200_ctx* c);                    // NCCL halo + leaving-molecule exchange of one step (no-op if not decomposed)
int do_allreduce_scalars(ls1b200_ctx* c);           // sums the 6 leading doubles of DevScalars over the ranks
int count_owned(ls1b200_ctx* c, bool want_index);
const char* overflow_message(uint32_t code);        // DevScalars::overflow -> text   // decomposed mode: n_owned_local (+ entry[slot] = index among owned slots)
int decomp_alloc(ls1b200_ctx* c);
void decomp_free(ls1b200_ctx* c);
}  // namespace ls1b200
