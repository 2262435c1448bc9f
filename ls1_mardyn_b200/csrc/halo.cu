// halo.cu -- multi-GPU 3D block decomposition: leaving-molecule + halo-copy exchange and the global reductions.
//
// Replaces (one process and one context per GPU): parallel/DomainDecomposition.cpp (:130-139 block bounds),
// DomainDecompMPIBase::exchangeMoleculesMPI(LEAVING_AND_HALO_COPIES) with the indirect (x, y, z staged) neighbour scheme
// (parallel/NeighbourCommunicationScheme.cpp), CommunicationPartner::collectMoleculesInRegion (CommunicationPartner.cpp:
// 435-575: which molecules, the periodic shift and its clamps) and the collective of Domain::calculateGlobalValues
// (Domain.cpp:176-182, :220-230).
//
// Per stage (dimension d with grid[d] > 1), all on the context's stream:
//   k_export   one pass over the slots: owned molecules outside [bmin, bmax) in d are packed as LEAVING for that neighbour
//              and stay behind as this rank's halo copy (META_FRESH); owned molecules and fresh halo copies within rc of
//              a face are packed as HALO copies. Records go straight into the two outgoing message buffers
//              (16-byte header holding the count + fixed-capacity record array), positions already shifted.
//   ncclGroup  { send low, send high, recv from high, recv from low }   -- NVLink / NVSwitch peer copies
//   k_import   appends the records of both incoming messages to the slot arrays (owned or fresh halo)
// The following ls1b200_update bins owned molecules and fresh halo copies alike and makes the local periodic copies of
// the dimensions that are not cut. In-cell order is by (molecule id, shift code), i.e. independent of message order.
// NCCL is bound at run time (dlopen), so the library loads -- and the single-GPU path runs -- without it.
#include <dlfcn.h>
#include <math.h>
#include <string.h>

#include <algorithm>

#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"

namespace ls1b200 {

// ---------------------------------------------------------------------------------------------------------------------
// NCCL, bound at run time
// ---------------------------------------------------------------------------------------------------------------------
namespace {
typedef struct { char internal[128]; } nccl_uid;
struct Nccl {
	void* h = nullptr;
	int (*GetUniqueId)(nccl_uid*) = nullptr;
	int (*CommInitRank)(void**, int, nccl_uid, int) = nullptr;
	int (*CommDestroy)(void*) = nullptr;
	int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
	int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
	int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
};
Nccl g_nccl;
constexpr int NCCL_UINT8 = 1, NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2;   // ncclDataType_t / ncclRedOp_t values of nccl.h (stable since 2.0)

int nccl_load() {
	if (g_nccl.h) return 0;
	void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	if (!h) return fail(std::string("ls1b200: cannot load NCCL (libnccl.so.2): ") + dlerror());
	Nccl n; n.h = h;
	#define SYM(field, name) *(void**) (&n.field) = dlsym(h, name); if (!n.field) return fail("ls1b200: NCCL symbol " name " not found")
	SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
	SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv"); SYM(AllReduce, "ncclAllReduce");
	SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
	#undef SYM
	g_nccl = n;
	return 0;
}
#define LS1_NCCL(call)                                                                                               \
	do {                                                                                                               \
		int e__ = (call);                                                                                                \
		if (e__ != 0) return ::ls1b200::fail(std::string(#call) + ": " + g_nccl.GetErrorString(e__));                    \
	} while (0)
}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// message records
// ---------------------------------------------------------------------------------------------------------------------
constexpr size_t MSG_HEADER = 16;
struct RecCompact { uint64_t id; uint32_t cid; uint32_t flags; double r[3], v[3]; };      // non-rotating ensembles over NCCL: 64 B
static_assert(sizeof(RecCompact) == 64 && sizeof(ls1b200_particle) == 120, "record layout");

struct XArgs {
	int dim;
	double bmin, bmax, rc;                 // this rank's block in dimension dim, halo width (== cutoff, DomainDecompBase.cpp:329)
	double glen, gmin, gmax;               // global box in dimension dim
	double wrap_hi, halo_hi, halo_lo;      // clamps against the global box
	int lo_wrap, hi_wrap;                  // the low / high neighbour lies across the global periodic boundary
	uint32_t cap, parity, msg_cap;
	int rotates;
	// slots [0, n_sorted) are in the cell order of the last update: only cells within `layers` of a face of dimension dim can hold
	// a molecule that leaves or needs a halo copy (a molecule moves less than a cell per step); cell_of == nullptr: scan everything
	const uint32_t* cell_of;
	uint32_t cstride, ncd, ncells;         // cell coordinate in dim = (cell / cstride) % ncd
	// the same filter as slot RANGES (slots are cell sorted): k_export walks only the rows / row ends within four cell layers of the
	// two faces of dimension dim, the limbo bin and the slots appended after the sort -- work proportional to the boundary, not to N
	const uint32_t* cell_start;
	int nx, ny, nz;
};

// writes one record (local message buffer or, peer transport, the neighbour's receive buffer over NVLink)
template <bool FULL>
__device__ __forceinline__ void pack(unsigned char* rec, uint32_t k, const State& st, size_t cap, uint32_t s, uint64_t meta,
                                     int dim, double xd, uint32_t flags, int rotates) {
	const double r[3] = {dim == 0 ? xd : st.r[s], dim == 1 ? xd : st.r[cap + s], dim == 2 ? xd : st.r[2 * cap + s]};
	if (FULL) {
		ls1b200_particle* p = reinterpret_cast<ls1b200_particle*>(rec) + k;
		p->id = st.id[s]; p->cid = meta_cid(meta); p->flags = flags;
		for (int d = 0; d < 3; d++) { p->r[d] = r[d]; p->v[d] = st.v[d * cap + s]; }
		if (rotates) {
			for (int d = 0; d < 4; d++) p->q[d] = st.q[d * cap + s];
			for (int d = 0; d < 3; d++) p->L[d] = st.L[d * cap + s];
		} else {
			p->q[0] = 1.; p->q[1] = p->q[2] = p->q[3] = 0.; p->L[0] = p->L[1] = p->L[2] = 0.;
		}
	} else {
		// 64 bytes as four 16-byte stores
		const double v0 = st.v[s], v1 = st.v[cap + s], v2 = st.v[2 * cap + s];
		uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<RecCompact*>(rec) + k);
		const uint64_t id = st.id[s];
		p[0] = make_uint4((uint32_t) id, (uint32_t) (id >> 32), meta_cid(meta), flags);
		double2* q = reinterpret_cast<double2*>(p + 1);
		q[0] = make_double2(r[0], r[1]); q[1] = make_double2(r[2], v0); q[2] = make_double2(v1, v2);
	}
}

struct XOut {
	unsigned char* rec[2];        // record arrays of the low / high message (local, or mapped peer memory)
	uint32_t* cnt[2];             // record counters (always local)
	unsigned long long* hdr[2];   // peer transport: the receive buffers' headers (k_publish)
	uint32_t seq;
	uint32_t cap[2];              // record capacity of the two targets
};

// which == 3: both sides (the transports inside the library), 1: low side only, 2: high side only (the staged C ABI exports one
// side per call; a leaving molecule is converted by the call for ITS side, so low + high together equal one which == 3 pass)
template <bool FULL>
__global__ void __launch_bounds__(256) k_export(XArgs a, State st, XOut o, DevScalars* scal, int which) {
	// persistent blocks striding over the slots: 68 000 blocks of a few hundred nanoseconds each cost more than the scan itself
	const size_t cap = a.cap;
	const uint32_t n_slots = scal->overflow ? 0u : min(scal->n_slots2[a.parity], a.cap);      // a failed step exports nothing (the step loop is frozen)
	constexpr int REC = FULL ? (int) sizeof(ls1b200_particle) : (int) sizeof(RecCompact);
	__shared__ __align__(16) unsigned char stage[8][32 * REC];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	// Work units of one warp: a cell range of one x-row (d = 0: the four cells at either end of every row; d = 1, 2: the whole rows of
	// the four layers at either face), or 32 slots of the tail (limbo bin + slots appended after the sort). Without a valid sort
	// (first exchange after an upload) the tail is everything.
	const bool ranged = a.cell_of != nullptr;
	const int lay = 4;
	const int nb0 = min(2 * lay, a.nx), nb1 = min(2 * lay, a.ny), nb2 = min(2 * lay, a.nz);
	const uint32_t ntask = !ranged ? 0u : a.dim == 0 ? (uint32_t) (a.ny * a.nz) * (a.nx > 2 * lay ? 2u : 1u) : a.dim == 1 ? (uint32_t) (a.nz * nb1) : (uint32_t) (nb2 * a.ny);
	// (cell_start[ncells], the start of the limbo bin, never exceeds the number of sorted slots; DevScalars::n_sorted is not read here: in the
	//  overlapped step the main stream's sort rewrites it while this kernel still walks the previous order, whose cell_start it was given)
	const uint32_t tail0 = ranged ? a.cell_start[a.ncells] : 0u;
	const uint32_t ntail = n_slots > tail0 ? (n_slots - tail0 + 31u) / 32u : 0u;
	(void) nb0;
	for (uint32_t task = blockIdx.x * 8u + (uint32_t) wid; task < ntask + ntail; task += gridDim.x * 8u) {
	uint32_t rb, re;
	if (task >= ntask) { rb = tail0 + (task - ntask) * 32u; re = min(rb + 32u, n_slots); }
	else if (a.dim == 0) {
		const bool two = a.nx > 2 * lay;
		const uint32_t row = two ? task >> 1 : task;
		const int c0 = !two ? 0 : (task & 1u) ? a.nx - lay : 0, c1 = !two ? a.nx : (task & 1u) ? a.nx : lay;
		rb = a.cell_start[row * (uint32_t) a.nx + c0]; re = a.cell_start[row * (uint32_t) a.nx + c1];
	} else if (a.dim == 1) {
		const int cz = (int) (task / (uint32_t) nb1), k = (int) (task % (uint32_t) nb1);
		const int cy = a.ny <= 2 * lay ? k : (k < lay ? k : a.ny - 2 * lay + k);
		const uint32_t row = (uint32_t) (cz * a.ny + cy);
		rb = a.cell_start[row * (uint32_t) a.nx]; re = a.cell_start[(row + 1u) * (uint32_t) a.nx];
	} else {
		const int k = (int) (task / (uint32_t) a.ny), cy = (int) (task % (uint32_t) a.ny);
		const int cz = a.nz <= 2 * lay ? k : (k < lay ? k : a.nz - 2 * lay + k);
		const uint32_t row = (uint32_t) (cz * a.ny + cy);
		rb = a.cell_start[row * (uint32_t) a.nx]; re = a.cell_start[(row + 1u) * (uint32_t) a.nx];
	}
	re = min(re, n_slots);
	for (uint32_t s0 = rb; s0 < re; s0 += 32u) {
	const uint32_t s = s0 + (uint32_t) lane;
	bool want[2] = {false, false}; double xn[2] = {0., 0.}; uint32_t fl[2] = {0u, 0u};
	uint64_t meta = 0;
	const bool candidate = s < re;
	if (candidate) {
		meta = st.meta[s];
		const bool halo = meta_halo(meta);
		if (!(halo && !meta_fresh(meta))) {                       // not a stale halo slot of the previous sort
			const double x = st.r[(size_t) a.dim * cap + s];
			if (!halo && x < a.bmin) {
				if (which & 1) {
					// leaving through the low face: LEAVING shift + clamp (CommunicationPartner.cpp:550-562) ...
					double xl = x;
					if (a.lo_wrap) { xl = __dadd_rn(x, a.glen); if (xl >= a.gmax) xl = a.wrap_hi; }
					want[0] = true; xn[0] = xl; fl[0] = 0u;
					// ... and it stays behind as my halo copy of the molecule its new owner now holds (HALO shift + clamp, :533-547)
					double xh = xl;
					if (a.lo_wrap) { xh = __dsub_rn(xl, a.glen); if (xh >= a.gmin) xh = a.halo_lo; }
					st.r[(size_t) a.dim * cap + s] = xh;
					st.meta[s] = meta | META_HALO | META_FRESH;
				}
			} else if (!halo && x >= a.bmax) {
				if (which & 2) {
					double xl = x;
					if (a.hi_wrap) { xl = __dsub_rn(x, a.glen); if (xl < a.gmin) xl = a.gmin; }
					want[1] = true; xn[1] = xl; fl[1] = 0u;
					double xh = xl;
					if (a.hi_wrap) { xh = __dadd_rn(xl, a.glen); if (xh < a.gmax) xh = a.halo_hi; }
					st.r[(size_t) a.dim * cap + s] = xh;
					st.meta[s] = meta | META_HALO | META_FRESH;
				}
			} else {
				// halo copies: an owned molecule inside the boundary layer, or a fresh halo copy of an earlier stage (forwarded to the
				// edge / corner neighbours; it may lie outside [bmin, bmax) in this dimension if its owner is about to hand it over)
				if ((which & 1) && x < __dadd_rn(a.bmin, a.rc)) {
					double xl = x;
					if (a.lo_wrap) { xl = __dadd_rn(x, a.glen); if (x >= a.bmin && xl < a.gmax) xl = a.halo_hi; }
					want[0] = true; xn[0] = xl; fl[0] = 1u;
				}
				if ((which & 2) && x >= __dsub_rn(a.bmax, a.rc)) {
					double xl = x;
					if (a.hi_wrap) { xl = __dsub_rn(x, a.glen); if (x < a.bmax && xl >= a.gmin) xl = a.halo_lo; }
					want[1] = true; xn[1] = xl; fl[1] = 1u;
				}
			}
		}
	}
	// One counter update per warp and side (the boundary layer is a few hundred thousand records on ONE counter). The records of
	// a warp are consecutive in the message: they are staged in shared memory and leave as whole 16-byte (8-byte for the 120-byte
	// records) lanes-contiguous stores -- over NVLink a 64-byte record written as four scattered stores costs ten times as much.
	#pragma unroll
	for (int side = 0; side < 2; side++) {
		const uint32_t b = __ballot_sync(0xffffffffu, want[side]);
		if (b == 0u) continue;
		const int leader = __ffs(b) - 1;
		const uint32_t cnt = (uint32_t) __popc(b);
		uint32_t base = 0;
		if (lane == leader) base = atomicAdd(o.cnt[side], cnt);
		base = __shfl_sync(0xffffffffu, base, leader);
		if (want[side]) pack<FULL>(stage[wid], (uint32_t) __popc(b & ((1u << lane) - 1u)), st, cap, s, meta, a.dim, xn[side], fl[side], a.rotates);
		__syncwarp();
		uint32_t n_copy = cnt;
		if (base + cnt > o.cap[side]) { n_copy = base < o.cap[side] ? o.cap[side] - base : 0u; if (lane == 0) scal->overflow = 2; }
		if (FULL) {
			const unsigned long long* src = reinterpret_cast<const unsigned long long*>(stage[wid]);
			unsigned long long* dst = reinterpret_cast<unsigned long long*>(o.rec[side] + (size_t) base * REC);
			for (uint32_t i = lane; i < n_copy * (REC / 8); i += 32) dst[i] = src[i];
		} else {
			const uint4* src = reinterpret_cast<const uint4*>(stage[wid]);
			uint4* dst = reinterpret_cast<uint4*>(o.rec[side] + (size_t) base * REC);
			for (uint32_t i = lane; i < n_copy * (REC / 16); i += 32) dst[i] = src[i];
		}
		__syncwarp();
	}
	}
	}
}

// peer transport: publishes {sequence number, count} in the receivers' headers. A kernel of its own behind k_export: the
// kernel boundary orders every record store before it, where a system-scope fence per block of k_export cost 0.8 ms.
__global__ void k_publish(XOut o) {
	if (blockIdx.x || threadIdx.x > 1) return;
	const bool hi = threadIdx.x == 1;      // (no dynamic indexing of the parameter struct: it would be copied to local memory)
	__threadfence_system();
	const uint32_t n = min(*reinterpret_cast<volatile uint32_t*>(hi ? o.cnt[1] : o.cnt[0]), hi ? o.cap[1] : o.cap[0]);
	*reinterpret_cast<volatile unsigned long long*>(hi ? o.hdr[1] : o.hdr[0]) = ((unsigned long long) o.seq << 32) | n;
	__threadfence_system();
}

// ticks of clock64() a receiver waits for its neighbour (set from LS1B200_PEER_TIMEOUT_S at ls1b200_peer_attach)
__device__ long long g_peer_timeout_ticks = 120000000000ll;

// peer transport: wait until the neighbour has published sequence number seq in this receive buffer (every thread of k_import
// and k_import_bump calls it; bounded: a peer that died must not hang the GPU -- overflow = 4 reports it)
__device__ __forceinline__ uint32_t msg_count(const unsigned char* msg, uint32_t seq, uint32_t msg_cap, DevScalars* scal) {
	const volatile unsigned long long* h = reinterpret_cast<const volatile unsigned long long*>(msg);
	unsigned long long v = *h;
	if (seq) {
		const long long t0 = clock64();
		while ((uint32_t) (v >> 32) != seq) {
			if (clock64() - t0 > g_peer_timeout_ticks) { scal->overflow = 4; return 0u; }      // LS1B200_PEER_TIMEOUT_S (default 60 s)
			__nanosleep(200);
			v = *h;
		}
		__threadfence_system();
	}
	return min((uint32_t) v, msg_cap);
}

// appends the records of one message behind the slots in use (+ the records of the message imported just before)
template <bool FULL>
__global__ void __launch_bounds__(256) k_import(const unsigned char* __restrict__ msg, const unsigned char* __restrict__ prev_msg, State st, uint32_t cap, uint32_t parity,
                                                 uint32_t msg_cap, int rotates, DevScalars* scal, uint32_t seq) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	__shared__ uint32_t s_n, s_prev;
	if (threadIdx.x == 0) { s_n = msg_count(msg, seq, msg_cap, scal); s_prev = prev_msg ? msg_count(prev_msg, seq, msg_cap, scal) : 0u; }
	__syncthreads();
	const uint32_t n = s_n;
	if (i >= n) return;
	const uint32_t base = scal->n_slots2[parity] + s_prev;
	const uint32_t t = base + i;
	if (t >= cap) { scal->overflow = 1; return; }
	const size_t c = cap;
	uint64_t id; uint32_t cid, flags; const double *r, *v;
	if (FULL) {
		const ls1b200_particle* p = reinterpret_cast<const ls1b200_particle*>(msg + MSG_HEADER) + i;
		id = p->id; cid = p->cid; flags = p->flags; r = p->r; v = p->v;
		if (rotates) {
			for (int d = 0; d < 4; d++) st.q[d * c + t] = p->q[d];
			for (int d = 0; d < 3; d++) st.L[d * c + t] = p->L[d];
		}
	} else {
		const RecCompact* p = reinterpret_cast<const RecCompact*>(msg + MSG_HEADER) + i;
		id = p->id; cid = p->cid; flags = p->flags; r = p->r; v = p->v;
	}
	for (int d = 0; d < 3; d++) { st.r[d * c + t] = r[d]; st.v[d * c + t] = v[d]; }
	st.id[t] = id;
	st.meta[t] = make_meta(0u, cid, false) | ((flags & 1u) ? (META_HALO | META_FRESH) : 0ull);
}
__global__ void k_import_bump(const unsigned char* __restrict__ m0, const unsigned char* __restrict__ m1, uint32_t cap, uint32_t parity, uint32_t msg_cap, DevScalars* scal, uint32_t seq) {
	if (threadIdx.x || blockIdx.x) return;
	uint32_t n = scal->n_slots2[parity] + msg_count(m0, seq, msg_cap, scal);
	if (m1) n += msg_count(m1, seq, msg_cap, scal);
	if (n > cap) { scal->overflow = 1; n = cap; }
	scal->n_slots2[parity] = n;
}

// ---- owned-molecule count / compaction index (decomposed mode: the set a rank owns changes every step) ----------------
__global__ void __launch_bounds__(256) k_flag_owned(const uint64_t* __restrict__ meta, uint32_t* __restrict__ flag, uint32_t cap, uint32_t parity, DevScalars* __restrict__ scal,
                                                     const CompDev* __restrict__ comps) {
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t rd = 0;
	if (s < cap) {
		const uint64_t m = meta[s];
		const bool own = s < min(scal->n_slots2[parity], cap) && !meta_halo(m);
		flag[s] = own ? 1u : 0u;
		if (own) rd = comps[meta_cid(m)].rot_dof;
	}
	// rotational degrees of freedom of the molecules owned NOW (Domain::setLocalNrotDOF of a phase-by-phase driver)
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) rd += __shfl_xor_sync(0xffffffffu, rd, o);
	if ((threadIdx.x & 31) == 0 && rd) atomicAdd(&scal->rot_dof, (unsigned long long) rd);
}
constexpr int CS_T = 256, CS_V = 8, CS_B = CS_T * CS_V;
__device__ __forceinline__ uint32_t cs_block_scan(uint32_t v, uint32_t* total, uint32_t* sh) {
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t inc = v;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
	if (lane == 31) sh[w] = inc;
	__syncthreads();
	if (w == 0) {
		uint32_t x = (lane < (blockDim.x >> 5)) ? sh[lane] : 0u, xi = x;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += t; }
		sh[lane] = xi - x;
		if (lane == 31) *total = xi;
	}
	__syncthreads();
	const uint32_t r = sh[w] + inc - v;
	__syncthreads();
	return r;
}
__global__ void __launch_bounds__(CS_T) k_cs_block(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t* __restrict__ bsum, uint32_t n) {
	__shared__ uint32_t sh[32]; __shared__ uint32_t tot;
	const uint32_t base = blockIdx.x * CS_B + threadIdx.x * CS_V;
	uint32_t v[CS_V], sum = 0;
	#pragma unroll
	for (int k = 0; k < CS_V; k++) { v[k] = (base + k < n) ? in[base + k] : 0u; sum += v[k]; }
	uint32_t off = cs_block_scan(sum, &tot, sh);
	#pragma unroll
	for (int k = 0; k < CS_V; k++) { if (base + k < n) out[base + k] = off; off += v[k]; }
	if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(CS_T) k_cs_top(uint32_t* __restrict__ bsum, uint32_t nb, DevScalars* scal) {
	__shared__ uint32_t sh[32]; __shared__ uint32_t tot;
	uint32_t carry = 0;
	for (uint32_t b0 = 0; b0 < nb; b0 += CS_T) {
		const uint32_t i = b0 + threadIdx.x;
		const uint32_t v = (i < nb) ? bsum[i] : 0u;
		const uint32_t off = cs_block_scan(v, &tot, sh);
		if (i < nb) bsum[i] = carry + off;
		carry += tot;
		__syncthreads();
	}
	if (threadIdx.x == 0) scal->n_owned_dev = carry;
}
__global__ void __launch_bounds__(CS_T) k_cs_add(uint32_t* __restrict__ out, const uint32_t* __restrict__ bsum, uint32_t n) {
	const uint32_t base = blockIdx.x * CS_B + threadIdx.x * CS_V;
	const uint32_t off = bsum[blockIdx.x];
	#pragma unroll
	for (int k = 0; k < CS_V; k++) if (base + k < n) out[base + k] += off;
}

int count_owned(ls1b200_ctx* c, bool want_index) {
	const uint32_t cap = (uint32_t) c->cap, g = (cap + 255) / 256, sb = (cap + CS_B - 1) / CS_B;
	uint32_t* bsum = reinterpret_cast<uint32_t*>(c->msg_out[0] + MSG_HEADER);      // scratch: the message buffers are idle here
	if ((uint64_t) sb * 4 > c->msg_cap * sizeof(RecCompact)) return fail("count_owned: scratch too small");
	LS1_CUDA(cudaMemsetAsync(&c->d_scal->rot_dof, 0, sizeof(unsigned long long), c->stream));
	LS1_LAUNCH(c, PH_UPDATE, k_flag_owned, g, 256, 0, c->cur.meta, c->key, cap, c->parity, c->d_scal, (const CompDev*) c->d_comps);
	LS1_LAUNCH(c, PH_UPDATE, k_cs_block, sb, CS_T, 0, c->key, c->entry, bsum, cap);
	LS1_LAUNCH(c, PH_UPDATE, k_cs_top, 1, CS_T, 0, bsum, sb, c->d_scal);
	if (want_index) LS1_LAUNCH(c, PH_UPDATE, k_cs_add, sb, CS_T, 0, c->entry, bsum, cap);
	LS1_CUDA(cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(DevScalars), cudaMemcpyDeviceToHost, c->stream));
	LS1_CUDA(cudaStreamSynchronize(c->stream));
	c->n_owned_local = c->h_scal->n_owned_dev;
	c->rot_dof_local = c->h_scal->rot_dof;
	return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
static XArgs stage_args(const ls1b200_ctx* c, int d) {
	const ls1b200_ctx::Decomp& dd = c->dd;
	XArgs a;
	a.dim = d; a.bmin = c->geom.bmin[d]; a.bmax = c->geom.bmax[d]; a.rc = c->geom.cutoff;
	a.gmin = dd.gmin[d]; a.gmax = dd.gmax[d]; a.glen = dd.gmax[d] - dd.gmin[d];
	a.wrap_hi = dd.wrap_hi[d]; a.halo_hi = dd.halo_hi[d]; a.halo_lo = dd.halo_lo[d];
	a.lo_wrap = dd.coords[d] == 0; a.hi_wrap = dd.coords[d] == dd.grid[d] - 1;
	a.cap = (uint32_t) c->cap; a.parity = c->parity; a.msg_cap = (uint32_t) c->msg_cap; a.rotates = c->rotates;
	// the cell filter needs the cell of every sorted slot: valid once an update has run (cell_of2 and n_sorted are of that update)
	a.cell_of = c->ever_sorted ? c->cell_of2 : nullptr;
	a.cstride = d == 0 ? 1u : d == 1 ? (uint32_t) c->geom.cells[0] : (uint32_t) (c->geom.cells[0] * c->geom.cells[1]);
	a.ncd = (uint32_t) c->geom.cells[d]; a.ncells = c->geom.ncells;
	a.cell_start = c->cell_start; a.nx = c->geom.cells[0]; a.ny = c->geom.cells[1]; a.nz = c->geom.cells[2];
	return a;
}

int decomp_alloc(ls1b200_ctx* c) {
	decomp_free(c);
	// records per message: the boundary layer of one face incl. forwarded copies, with head room
	double frac = 0.;
	for (int d = 0; d < 3; d++) if (c->dd.grid[d] > 1) frac = std::max(frac, std::min(1., c->geom.cutoff / c->geom.lbox[d]));
	double ring = 1.;
	for (int d = 0; d < 3; d++) ring *= 1. + 2. * std::min(1., c->geom.cutoff / c->geom.lbox[d]);
	c->msg_cap = (uint64_t) (1.6 * frac * ring * (double) std::max<uint64_t>(c->n_owned, 1)) + 8192;
	if (c->dd.comm) {
		// NCCL transport: Send and Recv of a pair must carry the same byte count, so all ranks agree on ONE capacity (the
		// largest): blocks of different population (balanced_cuts, uneven uploads) would otherwise size their messages differently
		double h = (double) c->msg_cap, *dcap = nullptr;
		LS1_CUDA(cudaMalloc((void**) &dcap, sizeof(double)));
		LS1_CUDA(cudaMemcpyAsync(dcap, &h, sizeof(double), cudaMemcpyHostToDevice, c->stream));
		LS1_NCCL(g_nccl.AllReduce(dcap, dcap, 1, NCCL_FLOAT64, NCCL_MAX, c->dd.comm, c->stream));
		LS1_CUDA(cudaMemcpyAsync(&h, dcap, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		LS1_CUDA(cudaStreamSynchronize(c->stream));
		cudaFree(dcap);
		c->msg_cap = (uint64_t) h;
	}
	const size_t bytes = MSG_HEADER + c->msg_cap * sizeof(ls1b200_particle);
	for (int k = 0; k < 2; k++) {
		LS1_CUDA(cudaMalloc((void**) &c->msg_out[k], bytes));
		LS1_CUDA(cudaMalloc((void**) &c->msg_in[k], bytes));
		LS1_CUDA(cudaMemsetAsync(c->msg_out[k], 0, MSG_HEADER, c->stream));
		LS1_CUDA(cudaMemsetAsync(c->msg_in[k], 0, MSG_HEADER, c->stream));
	}
	LS1_CUDA(cudaMalloc((void**) &c->ekey, sizeof(uint64_t) * c->cap));
	if (c->overlap || c->split_test) {      // the split step's per-slot region flag (k_gather writes it)
		LS1_CUDA(cudaMalloc((void**) &c->region_flag, c->cap));
		LS1_CUDA(cudaMemsetAsync(c->region_flag, 1, c->cap, c->stream));
	}
	// receive buffers of the peer-memory transport: [step parity][stage][side], one allocation (one IPC handle)
	c->pin_stride = (bytes + 255) & ~(size_t) 255;
	LS1_CUDA(cudaMalloc((void**) &c->pin_base, 12 * c->pin_stride));
	LS1_CUDA(cudaMemsetAsync(c->pin_base, 0, 12 * c->pin_stride, c->stream));
	LS1_CUDA(cudaMalloc((void**) &c->xcnt, 16));
	LS1_CUDA(cudaStreamSynchronize(c->stream));
	c->xseq = 0; c->peer_ok = false;
	return 0;
}
void decomp_free(ls1b200_ctx* c) {
	for (int k = 0; k < 2; k++) {
		if (c->msg_out[k]) cudaFree(c->msg_out[k]); c->msg_out[k] = nullptr;
		if (c->msg_in[k]) cudaFree(c->msg_in[k]); c->msg_in[k] = nullptr;
	}
	if (c->ekey) cudaFree(c->ekey); c->ekey = nullptr;
	if (c->region_flag) cudaFree(c->region_flag); c->region_flag = nullptr;
	for (int k = 0; k < 6; k++) { if (c->peer_opened[k]) cudaIpcCloseMemHandle(c->peer_opened[k]); c->peer_opened[k] = nullptr; c->peer_opened_rank[k] = -1; }
	for (int d = 0; d < 3; d++) c->peer_base[d][0] = c->peer_base[d][1] = nullptr;
	c->peer_ok = false;
	if (c->pin_base) cudaFree(c->pin_base); c->pin_base = nullptr;
	if (c->xcnt) cudaFree(c->xcnt); c->xcnt = nullptr;
	if (c->import_stage) cudaFree(c->import_stage); c->import_stage = nullptr; c->import_stage_cap = 0;
	c->msg_cap = 0;
}

// message buffers of this rank as the target of k_export (NCCL transport, staged C ABI): the count lives in the header
static XOut local_out(ls1b200_ctx* c) {
	XOut o;
	for (int k = 0; k < 2; k++) { o.rec[k] = c->msg_out[k] + MSG_HEADER; o.cnt[k] = reinterpret_cast<uint32_t*>(c->msg_out[k]); o.hdr[k] = nullptr; o.cap[k] = (uint32_t) c->msg_cap; }
	o.seq = 0;
	return o;
}

int do_exchange(ls1b200_ctx* c) {
	if (!c->dd.active) return 0;
	if (!c->dd.comm && !c->peer_ok) return fail("ls1b200_exchange: this context has no transport (use_nccl = 0 and no ls1b200_peer_attach): move the messages with ls1b200_halo_export / import");
	if (!c->msg_out[0]) return fail("ls1b200_exchange: call ls1b200_upload first");
	LS1_TRY(span_begin(c, c->xchg_phase));
	const bool full = c->rotates;
	const size_t rec = full ? sizeof(ls1b200_particle) : sizeof(RecCompact);
	const size_t bytes = MSG_HEADER + c->msg_cap * rec;
	const uint32_t g = std::min<uint32_t>((uint32_t) ((c->cap + 255) / 256), (uint32_t) c->sm_count * 8u), gi = (uint32_t) ((c->msg_cap + 255) / 256);
	// LS1B200_EXCHANGE_PROFILE=1: CUDA-event breakdown of the stages (export / transport / import), printed by rank 0 every 32 calls
	static const bool prof = getenv("LS1B200_EXCHANGE_PROFILE") != nullptr;
	cudaEvent_t* pe = c->xprof_ev; double* pacc = c->xprof_acc; int& pcalls = c->xprof_calls;
	if (prof && pcalls == 0 && !pe[0]) for (int k = 0; k < 4; k++) cudaEventCreate(&pe[k]);
	const bool peer = c->peer_ok;
	const uint32_t par = c->xseq & 1u;
	for (int d = 0; d < 3; d++) {
		if (c->dd.grid[d] == 1) continue;
		const XArgs a = stage_args(c, d);
		if (prof) cudaEventRecord(pe[0], c->stream);
		XOut o = local_out(c);
		const unsigned char* in[2] = {c->msg_in[0], c->msg_in[1]};
		uint32_t seq = 0;
		if (peer) {
			// ---- peer transport: k_export stores the records straight into the neighbours' receive buffers (NVLink), the last
			// block publishes {sequence number, count}; k_import waits for the sequence number in MY receive buffers ----
			seq = c->xseq * 4u + (uint32_t) d + 1u;
			LS1_CUDA(cudaMemsetAsync(c->xcnt, 0, 16, c->stream));
			const size_t slot_lo = ((size_t) par * 3 + d) * 2 + 0, slot_hi = ((size_t) par * 3 + d) * 2 + 1;
			// my low message is what the low neighbour receives "from its high side" (its slot 0), my high message goes to slot 1
			unsigned char* dst_lo = c->peer_base[d][0] + slot_lo * c->peer_stride[d][0];
			unsigned char* dst_hi = c->peer_base[d][1] + slot_hi * c->peer_stride[d][1];
			o.cap[0] = (uint32_t) c->peer_cap[d][0]; o.cap[1] = (uint32_t) c->peer_cap[d][1];
			o.rec[0] = dst_lo + MSG_HEADER; o.rec[1] = dst_hi + MSG_HEADER;
			o.cnt[0] = c->xcnt; o.cnt[1] = c->xcnt + 1;
			o.hdr[0] = reinterpret_cast<unsigned long long*>(dst_lo); o.hdr[1] = reinterpret_cast<unsigned long long*>(dst_hi);
			o.seq = seq;
			in[0] = c->pin_base + slot_lo * c->pin_stride; in[1] = c->pin_base + slot_hi * c->pin_stride;
		} else {
			LS1_CUDA(cudaMemsetAsync(c->msg_out[0], 0, MSG_HEADER, c->stream));
			LS1_CUDA(cudaMemsetAsync(c->msg_out[1], 0, MSG_HEADER, c->stream));
		}
		if (full) LS1_LAUNCH(c, c->xchg_phase, k_export<true>, g, 256, 0, a, c->cur, o, c->d_scal, 3);
		else LS1_LAUNCH(c, c->xchg_phase, k_export<false>, g, 256, 0, a, c->cur, o, c->d_scal, 3);
		if (peer) LS1_LAUNCH(c, PH_UPDATE, k_publish, 1, 32, 0, o);
		if (prof) cudaEventRecord(pe[1], c->stream);
		if (!peer) {
			// NCCL transport: fixed-size messages (the count travels in the header): no host round trip, the whole step stays asynchronous
			LS1_NCCL(g_nccl.GroupStart());
			LS1_NCCL(g_nccl.Send(c->msg_out[0], bytes, NCCL_UINT8, c->dd.nbr[d][0], c->dd.comm, c->stream));
			LS1_NCCL(g_nccl.Send(c->msg_out[1], bytes, NCCL_UINT8, c->dd.nbr[d][1], c->dd.comm, c->stream));
			LS1_NCCL(g_nccl.Recv(c->msg_in[0], bytes, NCCL_UINT8, c->dd.nbr[d][1], c->dd.comm, c->stream));      // the high neighbour's low message
			LS1_NCCL(g_nccl.Recv(c->msg_in[1], bytes, NCCL_UINT8, c->dd.nbr[d][0], c->dd.comm, c->stream));      // the low neighbour's high message
			LS1_NCCL(g_nccl.GroupEnd());
		}
		if (prof) cudaEventRecord(pe[2], c->stream);
		if (full) {
			LS1_LAUNCH(c, PH_UPDATE, k_import<true>, gi, 256, 0, in[0], (const unsigned char*) nullptr, c->cur, (uint32_t) c->cap, c->parity, (uint32_t) c->msg_cap, (int) c->rotates, c->d_scal, seq);
			LS1_LAUNCH(c, PH_UPDATE, k_import<true>, gi, 256, 0, in[1], in[0], c->cur, (uint32_t) c->cap, c->parity, (uint32_t) c->msg_cap, (int) c->rotates, c->d_scal, seq);
		} else {
			LS1_LAUNCH(c, PH_UPDATE, k_import<false>, gi, 256, 0, in[0], (const unsigned char*) nullptr, c->cur, (uint32_t) c->cap, c->parity, (uint32_t) c->msg_cap, (int) c->rotates, c->d_scal, seq);
			LS1_LAUNCH(c, PH_UPDATE, k_import<false>, gi, 256, 0, in[1], in[0], c->cur, (uint32_t) c->cap, c->parity, (uint32_t) c->msg_cap, (int) c->rotates, c->d_scal, seq);
		}
		LS1_LAUNCH(c, PH_UPDATE, k_import_bump, 1, 32, 0, in[0], in[1], (uint32_t) c->cap, c->parity, (uint32_t) c->msg_cap, c->d_scal, seq);
		if (prof) {
			cudaEventRecord(pe[3], c->stream); cudaEventSynchronize(pe[3]);
			++pcalls;
			if (pcalls > 16) for (int k = 0; k < 3; k++) { float ms = 0.f; cudaEventElapsedTime(&ms, pe[k], pe[k + 1]); pacc[k] += ms; }      // the first calls set up the connections
			if (pcalls > 16 && pcalls % 32 == 0 && c->dd.rank == 0)
				fprintf(stderr, "[ls1b200 exchange] %s transport, per stage: export %.3f ms, transport %.3f ms, import %.3f ms (message capacity %zu bytes)\n", peer ? "peer-memory" : "NCCL",
				        pacc[0] / (pcalls - 16), pacc[1] / (pcalls - 16), pacc[2] / (pcalls - 16), bytes);
		}
	}
	c->xseq++;
	LS1_TRY(span_end(c));
	c->sorted = false;
	return 0;
}

// sum of n doubles in device memory over the ranks of a decomposed context that has a communicator (no-op otherwise); collective
int allreduce_doubles(ls1b200_ctx* c, double* dev, int n) {
	if (!c->dd.active || !c->dd.comm) return 0;
	LS1_NCCL(g_nccl.AllReduce(dev, dev, (size_t) n, NCCL_FLOAT64, NCCL_SUM, c->dd.comm, c->stream));
	return 0;
}

int do_allreduce_scalars(ls1b200_ctx* c) {
	if (!c->dd.active || !c->dd.comm) return 0;
	// upot6lj, upot_xpoles, virial, my_rf, summv2, sum_iw2 + the fault flag: the seven leading doubles of DevScalars
	LS1_NCCL(g_nccl.AllReduce(c->d_scal, c->d_scal, 7, NCCL_FLOAT64, NCCL_SUM, c->dd.comm, c->stream));
	return 0;
}

}  // namespace ls1b200

using namespace ls1b200;

extern "C" {

int ls1b200_nccl_unique_id(uint8_t out[128]) {
	if (!out) return fail("ls1b200_nccl_unique_id: null argument");
	LS1_TRY(nccl_load());
	nccl_uid id;
	LS1_NCCL(g_nccl.GetUniqueId(&id));
	memcpy(out, id.internal, 128);
	return 0;
}

int ls1b200_decomp_attach(ls1b200_ctx* c, const ls1b200_decomp* dd) {
	if (!c || !dd) return fail("ls1b200_decomp_attach: null argument");
	if (!c->components_set) return fail("ls1b200_decomp_attach: call ls1b200_set_components first");
	if (c->cap) return fail("ls1b200_decomp_attach: must be called before ls1b200_upload");
	if (dd->n_ranks < 1 || dd->grid[0] * dd->grid[1] * dd->grid[2] != dd->n_ranks) return fail("ls1b200_decomp_attach: grid does not match the number of ranks");
	if (dd->rank < 0 || dd->rank >= dd->n_ranks) return fail("ls1b200_decomp_attach: rank out of range");
	LS1_CUDA(cudaSetDevice(c->device));
	ls1b200_ctx::Decomp& o = c->dd;
	o = ls1b200_ctx::Decomp();
	o.rank = dd->rank; o.n_ranks = dd->n_ranks;
	for (int d = 0; d < 3; d++) {
		o.grid[d] = dd->grid[d]; o.coords[d] = dd->coords[d];
		o.gmin[d] = dd->global_min[d]; o.gmax[d] = dd->global_max[d];
		if (o.coords[d] < 0 || o.coords[d] >= o.grid[d]) return fail("ls1b200_decomp_attach: coordinates outside the grid");
		if ((o.grid[d] == 1) != (c->geom.periodic[d] != 0) )
			return fail("ls1b200_decomp_attach: the context must be created with periodic[d] = (grid[d] == 1)");
		if (o.grid[d] > 1 && c->geom.lbox[d] < 2. * c->geom.cutoff) return fail("ls1b200_decomp_attach: a block must be at least two cutoff radii wide");
		// clamps of the periodic shift, in the calc type, against the global box (CommunicationPartner.cpp:533-562)
		if (c->mixed) {
			const float rmax = (float) o.gmax[d], rmin = (float) o.gmin[d];
			o.wrap_hi[d] = (double) nexttowardf(rmax, (long double) (rmax - 1.f));
			o.halo_hi[d] = (double) nexttowardf(rmax, (long double) (rmax + 1.f));
			o.halo_lo[d] = (double) nexttowardf(rmin, (long double) (rmin - 1.f));
		} else {
			const double rmax = o.gmax[d], rmin = o.gmin[d];
			o.wrap_hi[d] = nexttoward(rmax, (long double) (rmax - 1.));
			o.halo_hi[d] = nexttoward(rmax, (long double) (rmax + 1.));
			o.halo_lo[d] = nexttoward(rmin, (long double) (rmin - 1.));
		}
	}
	for (int d = 0; d < 3; d++) for (int side = 0; side < 2; side++) {
		int cc[3] = {o.coords[0], o.coords[1], o.coords[2]};
		cc[d] = (cc[d] + (side ? 1 : -1) + o.grid[d]) % o.grid[d];
		o.nbr[d][side] = (cc[0] * o.grid[1] + cc[1]) * o.grid[2] + cc[2];
	}
	{
		int cc = (o.coords[0] * o.grid[1] + o.coords[1]) * o.grid[2] + o.coords[2];
		if (cc != o.rank) return fail("ls1b200_decomp_attach: rank != (coords[0]*grid[1] + coords[1])*grid[2] + coords[2]");
	}
	if (dd->use_nccl) {
		LS1_TRY(nccl_load());
		nccl_uid id; memcpy(id.internal, dd->nccl_unique_id, 128);
		LS1_NCCL(g_nccl.CommInitRank(&o.comm, o.n_ranks, id, o.rank));
	}
	o.active = true;
	return 0;
}

int ls1b200_peer_handle(ls1b200_ctx* c, uint8_t out[LS1B200_PEER_HANDLE_BYTES]) {
	if (!c || !out) return fail("ls1b200_peer_handle: null argument");
	if (!c->dd.active || !c->pin_base) return fail("ls1b200_peer_handle: call ls1b200_decomp_attach and ls1b200_upload first");
	LS1_CUDA(cudaSetDevice(c->device));
	static_assert(sizeof(cudaIpcMemHandle_t) + 16 <= LS1B200_PEER_HANDLE_BYTES, "handle size");
	memset(out, 0, LS1B200_PEER_HANDLE_BYTES);
	cudaIpcMemHandle_t h;
	LS1_CUDA(cudaIpcGetMemHandle(&h, c->pin_base));
	memcpy(out, &h, sizeof(h));
	const uint64_t stride = c->pin_stride, cap = c->msg_cap;
	memcpy(out + sizeof(h), &stride, 8); memcpy(out + sizeof(h) + 8, &cap, 8);
	return 0;
}

int ls1b200_peer_attach(ls1b200_ctx* c, const uint8_t* handles, int32_t n_ranks) {
	if (!c || !handles) return fail("ls1b200_peer_attach: null argument");
	if (!c->dd.active || !c->pin_base) return fail("ls1b200_peer_attach: call ls1b200_decomp_attach and ls1b200_upload first");
	if (n_ranks != c->dd.n_ranks) return fail("ls1b200_peer_attach: one handle per rank of the decomposition is expected");
	LS1_CUDA(cudaSetDevice(c->device));
	// a second attach replaces the first: unmap what is mapped
	if (c->stream) LS1_CUDA(cudaStreamSynchronize(c->stream));
	for (int k = 0; k < 6; k++) { if (c->peer_opened[k]) cudaIpcCloseMemHandle(c->peer_opened[k]); c->peer_opened[k] = nullptr; c->peer_opened_rank[k] = -1; }
	c->peer_ok = false;
	int n_open = 0;
	for (int d = 0; d < 3; d++) for (int side = 0; side < 2; side++) {
		c->peer_base[d][side] = nullptr;
		if (c->dd.grid[d] == 1) continue;
		const int r = c->dd.nbr[d][side];
		const uint8_t* hp = handles + (size_t) r * LS1B200_PEER_HANDLE_BYTES;
		uint64_t stride = 0, cap = 0;
		memcpy(&stride, hp + sizeof(cudaIpcMemHandle_t), 8); memcpy(&cap, hp + sizeof(cudaIpcMemHandle_t) + 8, 8);
		// the sender addresses the receiver's slots and honours the receiver's capacity
		if (stride == 0 || cap == 0) return fail("ls1b200_peer_attach: empty handle (did every rank call ls1b200_peer_handle after its upload?)");
		c->peer_stride[d][side] = (size_t) stride; c->peer_cap[d][side] = cap;
		if (r == c->dd.rank) { c->peer_base[d][side] = c->pin_base; continue; }
		void* base = nullptr;
		for (int k = 0; k < n_open; k++) if (c->peer_opened_rank[k] == r) base = c->peer_opened[k];
		if (!base) {
			cudaIpcMemHandle_t h; memcpy(&h, hp, sizeof(h));
			LS1_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
			c->peer_opened[n_open] = base; c->peer_opened_rank[n_open] = r; n_open++;
		}
		c->peer_base[d][side] = (unsigned char*) base;
	}
	{
		// how long k_import waits for a neighbour's message: generous by default (a neighbour may be busy with I/O between two
		// ls1b200_simulate calls); LS1B200_PEER_TIMEOUT_S overrides. Ticks of clock64() at the SM clock.
		double sec = 60.;
		if (const char* e = getenv("LS1B200_PEER_TIMEOUT_S")) { const double v = atof(e); if (v > 0.) sec = v; }
		int khz = 1965000; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
		const long long ticks = (long long) (sec * 1e3 * (double) khz);
		LS1_CUDA(cudaMemcpyToSymbol(g_peer_timeout_ticks, &ticks, sizeof(ticks)));
	}
	c->peer_ok = true;
	return 0;
}

int ls1b200_peer_detach(ls1b200_ctx* c) {
	if (!c) return fail("ls1b200_peer_detach: null context");
	LS1_CUDA(cudaSetDevice(c->device));
	if (c->stream) LS1_CUDA(cudaStreamSynchronize(c->stream));
	for (int k = 0; k < 6; k++) { if (c->peer_opened[k]) cudaIpcCloseMemHandle(c->peer_opened[k]); c->peer_opened[k] = nullptr; c->peer_opened_rank[k] = -1; }
	for (int d = 0; d < 3; d++) c->peer_base[d][0] = c->peer_base[d][1] = nullptr;
	c->peer_ok = false;
	return 0;
}

int ls1b200_exchange(ls1b200_ctx* c) {
	if (!c) return fail("ls1b200_exchange: null context");
	LS1_CUDA(cudaSetDevice(c->device));
	return do_exchange(c);
}

int ls1b200_allreduce_macro(ls1b200_ctx* c, ls1b200_macro* m, ls1b200_kinetic* k) {
	if (!c) return fail("ls1b200_allreduce_macro: null context");
	LS1_CUDA(cudaSetDevice(c->device));
	LS1_TRY(do_allreduce_scalars(c));
	LS1_CUDA(cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(DevScalars), cudaMemcpyDeviceToHost, c->stream));
	LS1_CUDA(cudaStreamSynchronize(c->stream));
	const DevScalars& s = *c->h_scal;
	if (m) {
		m->upot6lj = s.upot6lj; m->upot_xpoles = s.upot_xpoles; m->virial_raw = s.virial; m->my_rf = s.my_rf;
		m->upot = s.upot6lj / 6.0 + s.upot_xpoles + s.my_rf; m->virial = s.virial + 3.0 * s.my_rf;
	}
	if (k) { k->summv2 = s.summv2; k->sum_iw2 = s.sum_iw2; }
	return 0;
}

int ls1b200_halo_export(ls1b200_ctx* c, int32_t dim, int32_t side, void** device_buf, uint64_t* count) {
	if (!c || !device_buf || !count) return fail("ls1b200_halo_export: null argument");
	if (!c->dd.active) return fail("ls1b200_halo_export: no decomposition attached");
	if (dim < 0 || dim > 2 || side < 0 || side > 1 || c->dd.grid[dim] == 1) return fail("ls1b200_halo_export: this dimension is not cut / bad side");
	if (!c->msg_out[0]) return fail("ls1b200_halo_export: call ls1b200_upload first");
	LS1_CUDA(cudaSetDevice(c->device));
	const XArgs a = stage_args(c, dim);
	LS1_CUDA(cudaMemsetAsync(c->msg_out[side], 0, MSG_HEADER, c->stream));
	LS1_LAUNCH(c, c->xchg_phase, k_export<true>, std::min<uint32_t>((uint32_t) ((c->cap + 255) / 256), (uint32_t) c->sm_count * 8u), 256, 0, a, c->cur, local_out(c), c->d_scal, side ? 2 : 1);
	uint32_t n = 0;
	LS1_CUDA(cudaMemcpyAsync(&n, c->msg_out[side], 4, cudaMemcpyDeviceToHost, c->stream));
	LS1_CUDA(cudaStreamSynchronize(c->stream));
	if (n > c->msg_cap) return fail("ls1b200_halo_export: message buffer too small");
	*device_buf = c->msg_out[side] + MSG_HEADER; *count = n;
	c->sorted = false;
	return 0;
}

}  // extern "C" (helpers below are C++)

static int export_side(ls1b200_ctx* c, int32_t dim, int32_t side, void** device_buf, uint64_t* count) { return ls1b200_halo_export(c, dim, side, device_buf, count); }

extern "C" {

/* ls1b200_halo_export + copy of the records to host memory (tests, MPI without GPU-direct). */
int ls1b200_halo_export_host(ls1b200_ctx* c, int32_t dim, int32_t side, ls1b200_particle* host_buf, uint64_t host_cap, uint64_t* count) {
	if (!count) return fail("ls1b200_halo_export_host: null argument");
	void* dbuf = nullptr;
	LS1_TRY(export_side(c, dim, side, &dbuf, count));
	if (*count > host_cap) return fail("ls1b200_halo_export_host: host buffer too small (count returned)");
	if (*count) {
		if (!host_buf) return fail("ls1b200_halo_export_host: null buffer");
		LS1_CUDA(cudaMemcpyAsync(host_buf, dbuf, *count * sizeof(ls1b200_particle), cudaMemcpyDeviceToHost, c->stream));
		LS1_CUDA(cudaStreamSynchronize(c->stream));
	}
	return 0;
}

static int import_records(ls1b200_ctx* c, const void* buf, uint64_t count, cudaMemcpyKind kind) {
	if (!c->dd.active) return fail("ls1b200_halo_import: no decomposition attached");
	if (!c->msg_in[0]) return fail("ls1b200_halo_import: call ls1b200_upload first");
	if (count == 0) return 0;
	if (!buf) return fail("ls1b200_halo_import: null buffer");
	if (count > c->msg_cap) return fail("ls1b200_halo_import: more records than the message capacity");
	LS1_CUDA(cudaSetDevice(c->device));
	const uint32_t n = (uint32_t) count;
	LS1_CUDA(cudaMemcpyAsync(c->msg_in[0], &n, 4, cudaMemcpyHostToDevice, c->stream));
	LS1_CUDA(cudaMemcpyAsync(c->msg_in[0] + MSG_HEADER, buf, count * sizeof(ls1b200_particle), kind, c->stream));
	LS1_LAUNCH(c, PH_UPDATE, k_import<true>, (uint32_t) ((count + 255) / 256), 256, 0, c->msg_in[0], (const unsigned char*) nullptr, c->cur, (uint32_t) c->cap, c->parity,
	           (uint32_t) c->msg_cap, (int) c->rotates, c->d_scal, 0u);
	LS1_LAUNCH(c, PH_UPDATE, k_import_bump, 1, 32, 0, c->msg_in[0], (const unsigned char*) nullptr, (uint32_t) c->cap, c->parity, (uint32_t) c->msg_cap, c->d_scal, 0u);
	LS1_CUDA(cudaStreamSynchronize(c->stream));      // n and a pageable host buffer are read by the copies above
	c->sorted = false;
	return 0;
}

int ls1b200_halo_import(ls1b200_ctx* c, const void* device_buf, uint64_t count) {
	if (!c) return fail("ls1b200_halo_import: null context");
	return import_records(c, device_buf, count, cudaMemcpyDeviceToDevice);
}
int ls1b200_halo_import_host(ls1b200_ctx* c, const ls1b200_particle* host_buf, uint64_t count) {
	if (!c) return fail("ls1b200_halo_import_host: null context");
	return import_records(c, host_buf, count, cudaMemcpyHostToDevice);
}

}  // extern "C"
