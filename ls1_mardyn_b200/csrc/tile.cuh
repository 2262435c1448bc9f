// tile.cuh -- machinery shared by the tiled cell-pair kernels (force_1clj_tiled.cu, force_multi_tiled.cu): the chunk list
// (runs of cells of one x-row whose molecules become the threads of one CTA pass and whose candidates are 9 contiguous slot
// ranges), its device-side builder, and the shared-memory / TMA / mbarrier helpers (inline PTX).
#pragma once
#include "pairs.cuh"
#include "types.cuh"

namespace ls1b200 {

struct Chunk { uint32_t cell0; uint16_t ncells; uint16_t flags; };
constexpr uint16_t CHUNK_BIG = 1;
constexpr uint16_t CHUNK_BND = 2;   // overlapped multi-GPU step: a chunk whose windows may hold imports (second slot range per cell, force_1clj_tiled.cu)
constexpr int CH_MAXC = 24;        // max cells per chunk (size of the cell_start cache)

// ---------------------------------------------------------------------------------------------------------------------
// chunk list
// ---------------------------------------------------------------------------------------------------------------------
// mode (overlapped multi-GPU step): 0 every inner cell of the grid; 1 only the chunks no import can touch -- cells further than two
// layers from every face shared with another rank (cut[d]); 2 the others, with the imports' own cell_start (csb) counted in
struct ChunkGeo { int nx, ny, nz; int tpb, w, ri; int mode; int cut[3]; };

// Chunk list of one row, one WARP per row. The greedy rule "extend the chunk while it fits" is evaluated in parallel: every
// lane computes, for its own candidate start cell xa, the last cell e(xa) a chunk starting there may reach (all constraints are
// monotone in the end cell, so a short linear scan over <= CH_MAXC cells of shared memory does it); lane 0 then only jumps
// xa -> e(xa) + 1 along the row (about ten dependent reads). The cell_start entries of the 9 neighbour rows and the prefix
// sum of threads per cell are staged in shared memory with coalesced loads first. (A serial walk per row -- with global or
// with shared loads -- cost 0.23 - 0.54 ms per pass at 115^3 cells: a dependent chain of ~2000 loads per row.)
constexpr int CHUNK_ROWS_PER_BLOCK = 4;
template <bool WRITE>
__global__ void __launch_bounds__(32 * CHUNK_ROWS_PER_BLOCK) k_chunks(ChunkGeo g, const uint32_t* __restrict__ cs, const uint32_t* __restrict__ csb, uint32_t nrows, uint32_t* __restrict__ cnt,
                                                                      const uint32_t* __restrict__ start, Chunk* __restrict__ chunks) {
	extern __shared__ uint32_t sm_cs[];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const uint32_t r = blockIdx.x * CHUNK_ROWS_PER_BLOCK + w;
	if (r >= nrows) return;
	const int nx = g.nx, P = nx + 1;
	const int cy = 1 + (int) (r % (uint32_t) (g.ny - 2)), cz = 1 + (int) (r / (uint32_t) (g.ny - 2));
	// Which x-ranges of this row the pass chunks (mode 0: all inner cells). Rows within two layers of a y / z face shared with another
	// rank belong to the boundary pass as a whole. In the other rows the boundary pass takes, where x is cut, the FIRST and the LAST chunk
	// the greedy rule gives on the sorted slots alone (they hold the two cells next to either face and are as full as any other chunk:
	// chunks of two cells would cost a CTA pass each for a sixth of the molecules); both passes evaluate the same rule, so they agree.
	const bool row_bnd = g.mode != 0 && ((g.cut[1] && (cy <= 2 || cy >= g.ny - 3)) || (g.cut[2] && (cz <= 2 || cz >= g.nz - 3)));
	if ((g.mode == 1 && row_bnd) || (g.mode == 2 && !row_bnd && !g.cut[0])) { if (!WRITE && lane == 0) cnt[r] = 0; return; }
	const bool split_x = g.mode != 0 && !row_bnd && g.cut[0];
	const int NA = csb ? 22 : 13;
	uint32_t* rowcs = sm_cs + (size_t) w * NA * P;       // [9][P] cell_start of the neighbour rows
	uint32_t* thr = rowcs + 9 * P;                        // [P] exclusive prefix of threads per cell of the centre row
	uint32_t* reach = thr + P;                            // [P] e(xa), or xa - 1 if not even the single cell fits
	uint32_t* rowcsb = reach + P;                         // [9][P] the imports' cell_start (mode 2)
	uint32_t* thrA = csb ? rowcsb + 9 * P : reach + P;    // [P], [P]: thr / reach of the sorted slots alone, with room left for imports (first / last chunk of a row)
	uint32_t* reachA = thrA + P;
	for (int k = 0; k < 9; k++) {
		const uint32_t row = (uint32_t) ((cz + k / 3 - 1) * g.ny + (cy + k % 3 - 1)) * nx;
		for (int x = lane; x <= nx; x += 32) { rowcs[k * P + x] = cs[row + x]; if (csb) rowcsb[k * P + x] = csb[row + x]; }
	}
	__syncwarp();
	const uint32_t* c4 = rowcs + 4 * P;                   // centre row
	const int last_x = nx - 2;
	const uint32_t* cb4 = rowcsb + 4 * P;
	uint32_t carry = 0;
	for (int x0 = 0; x0 < nx; x0 += 32) {                 // thr[x] = sum_{x' < x} ceil(n(x') / RI)
		const int x = x0 + lane;
		const uint32_t nmol = (x < nx) ? (c4[x + 1] - c4[x]) + (csb ? cb4[x + 1] - cb4[x] : 0u) : 0u;
		const uint32_t v = (nmol + (uint32_t) g.ri - 1u) / (uint32_t) g.ri;
		uint32_t inc = v;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
		if (x <= nx) thr[x] = carry + inc - v;
		carry += __shfl_sync(0xffffffffu, inc, 31);
	}
	if (lane == 0) thr[nx] = carry;
	// room the first / last chunk of a row leaves for the imports of the boundary pass (the halo cell next to it holds imports only)
	constexpr uint32_t BND_W = 24, BND_T = 8;
	if (csb && split_x) {
		uint32_t carryA = 0;
		for (int x0 = 0; x0 < nx; x0 += 32) {
			const int x = x0 + lane;
			const uint32_t v = (x < nx) ? (c4[x + 1] - c4[x] + (uint32_t) g.ri - 1u) / (uint32_t) g.ri : 0u;
			uint32_t inc = v;
			#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
			if (x <= nx) thrA[x] = carryA + inc - v;
			carryA += __shfl_sync(0xffffffffu, inc, 31);
		}
		if (lane == 0) thrA[nx] = carryA;
	}
	__syncwarp();
	for (int xa = 1 + lane; xa <= last_x; xa += 32) {
		uint32_t lo[9], lob[9];
		#pragma unroll
		for (int k = 0; k < 9; k++) { lo[k] = rowcs[k * P + xa - 1]; lob[k] = csb ? rowcsb[k * P + xa - 1] : 0u; }
		int xb = xa - 1;
		for (int x = xa; x <= last_x && x - xa + 1 <= CH_MAXC; x++) {
			bool f = thr[x + 1] - thr[xa] <= (uint32_t) g.tpb;
			#pragma unroll
			for (int k = 0; k < 9; k++) f = f && ((rowcs[k * P + x + 2] - lo[k]) + (csb ? rowcsb[k * P + x + 2] - lob[k] : 0u) <= (uint32_t) g.w);
			if (!f) break;
			xb = x;
		}
		reach[xa] = (uint32_t) xb;
		if (split_x) {      // the same rule on the sorted slots alone, with room for the imports
			const uint32_t* tA = csb ? thrA : thr;
			int xbA = xa - 1;
			for (int x = xa; x <= last_x && x - xa + 1 <= CH_MAXC; x++) {
				bool f = tA[x + 1] - tA[xa] + BND_T <= (uint32_t) g.tpb;
				#pragma unroll
				for (int k = 0; k < 9; k++) f = f && (rowcs[k * P + x + 2] - lo[k] + BND_W <= (uint32_t) g.w);
				if (!f) break;
				xbA = x;
			}
			reachA[xa] = (uint32_t) xbA;
		}
	}
	__syncwarp();
	int seg_lo[2] = {1, 0}, seg_hi[2] = {last_x, -1}, nseg = 1;
	if (split_x) {
		const uint32_t* rA = reachA;
		const int e1 = min(max((int) rA[1], 2), last_x);      // end of the first chunk (at least the two cells next to the face)
		int s2 = last_x - 1;                                   // start of the last one (at least the last two cells)
		for (int x0 = e1 + 1; x0 < last_x - 1; x0 += 32) {
			const int xa = x0 + lane;
			const uint32_t hit = __ballot_sync(0xffffffffu, xa < last_x - 1 && (int) rA[xa] >= last_x);
			if (hit) { s2 = x0 + __ffs(hit) - 1; break; }
		}
		if (s2 <= e1) { if (g.mode == 1) nseg = 0; }          // a short row: boundary as a whole
		else if (g.mode == 1) { seg_lo[0] = e1 + 1; seg_hi[0] = s2 - 1; nseg = seg_lo[0] <= seg_hi[0] ? 1 : 0; }
		else { nseg = 2; seg_lo[0] = 1; seg_hi[0] = e1; seg_lo[1] = s2; seg_hi[1] = last_x; }
	}
	if (nseg == 0) { if (!WRITE && lane == 0) cnt[r] = 0; return; }
	if (lane == 0) {
		const uint32_t R = (uint32_t) (cz * g.ny + cy) * nx;
		Chunk* out = WRITE ? chunks + start[r] : nullptr;
		uint32_t n = 0;
		for (int sg = 0; sg < nseg; sg++) {
			int xa = seg_lo[sg];
			while (xa <= seg_hi[sg]) {
				int xb = min((int) reach[xa], seg_hi[sg]);
				const bool ok = xb >= xa;
				if (!ok) xb = xa;                               // a single cell denser than the tile: BIG chunk (slow path of the kernel)
				if ((c4[xb + 1] - c4[xa]) + (csb ? cb4[xb + 1] - cb4[xa] : 0u) > 0) {                  // chunks without molecules are dropped
					if (WRITE) { Chunk c; c.cell0 = R + xa; c.ncells = (uint16_t) (xb - xa + 1); c.flags = (uint16_t) ((ok ? 0 : CHUNK_BIG) | (g.mode == 2 ? CHUNK_BND : 0)); out[n] = c; }
					n++;
				}
				xa = xb + 1;
			}
		}
		if (!WRITE) cnt[r] = n;
	}
}
// single-block exclusive scan of the per-row counts (a few thousand rows); start[nrows] = total; resets the work counter
static __global__ void __launch_bounds__(1024) k_chunks_scan(const uint32_t* __restrict__ cnt, uint32_t* __restrict__ start, uint32_t nrows, uint32_t* __restrict__ work_counter) {
	__shared__ uint32_t sh[32]; __shared__ uint32_t carry_s;
	if (threadIdx.x == 0) { carry_s = 0; *work_counter = 0; }
	__syncthreads();
	for (uint32_t b0 = 0; b0 < nrows; b0 += blockDim.x) {
		const uint32_t i = b0 + threadIdx.x;
		const uint32_t v = i < nrows ? cnt[i] : 0u;
		const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
		uint32_t inc = v;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
		if (lane == 31) sh[w] = inc;
		__syncthreads();
		if (w == 0) {
			uint32_t x = sh[lane], xi = x;
			#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += t; }
			sh[lane] = xi - x;
		}
		__syncthreads();
		const uint32_t carry = carry_s;
		if (i < nrows) start[i] = carry + sh[w] + inc - v;
		__syncthreads();
		if (threadIdx.x == blockDim.x - 1) carry_s = carry + sh[w] + inc;
		__syncthreads();
	}
	if (threadIdx.x == 0) start[nrows] = carry_s;
}

// ---------------------------------------------------------------------------------------------------------------------
// shared-memory / TMA helpers (inline PTX; cp.async.bulk shows up as UBLKCP in SASS)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
	uint32_t ok;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
	} while (!ok);
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(addr), "h"((unsigned short) v) : "memory"); }
__device__ __forceinline__ uint32_t lds16(uint32_t addr) { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
template <class C> __device__ __forceinline__ typename Vec4<C>::type lds_v4(uint32_t addr);
template <> __device__ __forceinline__ float4 lds_v4<float>(uint32_t addr) {
	float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory"); return v;
}
template <> __device__ __forceinline__ double4 lds_v4<double>(uint32_t addr) {
	double4 v;
	asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
	asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.z), "=d"(v.w) : "r"(addr + 16) : "memory");
	return v;
}

template <class C> __device__ __forceinline__ C rcp_fast(C x);
template <> __device__ __forceinline__ float rcp_fast<float>(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
template <> __device__ __forceinline__ double rcp_fast<double>(double x) { return 1.0 / x; }


// builds the chunk list for (tpb threads, w records per window row, ri molecules per thread); n_chunks = row_start[nrows]
// set 1: the second chunk list of the overlapped multi-GPU step (boundary chunks; built on the exchange stream while the inner chunks of set 0 are traversed)
static inline int build_chunks(ls1b200_ctx* c, int tpb, int w, int ri, uint32_t* nrows_out, int mode = 0, const uint32_t* csb = nullptr, int set = 0) {
	const Geom& g = c->geom;
	const uint32_t nrows = (uint32_t) (g.cells[1] - 2) * (uint32_t) (g.cells[2] - 2);
	ChunkGeo cg; cg.nx = g.cells[0]; cg.ny = g.cells[1]; cg.nz = g.cells[2]; cg.tpb = tpb; cg.w = w; cg.ri = ri;
	cg.mode = mode; for (int d = 0; d < 3; d++) cg.cut[d] = c->dd.active && c->dd.grid[d] > 1;
	if (set == 1 && !c->chunks_b) {
		const size_t nr = (size_t) g.cells[1] * g.cells[2] + 8;
		LS1_CUDA(cudaMalloc(&c->chunks_b, ((size_t) g.ncells + 8) * 8));
		LS1_CUDA(cudaMalloc((void**) &c->chunk_row_count_b, sizeof(uint32_t) * (2 * nr + 4)));
		c->chunk_row_start_b = c->chunk_row_count_b + nr; c->work_counter_b = c->chunk_row_start_b + nr;
	}
	Chunk* chunks = (Chunk*) (set ? c->chunks_b : c->chunks);
	uint32_t* row_count = set ? c->chunk_row_count_b : c->chunk_row_count; uint32_t* row_start = set ? c->chunk_row_start_b : c->chunk_row_start;
	uint32_t* work_counter = set ? c->work_counter_b : c->work_counter;
	const uint32_t gr = (nrows + CHUNK_ROWS_PER_BLOCK - 1) / CHUNK_ROWS_PER_BLOCK;
	const size_t csm = (size_t) CHUNK_ROWS_PER_BLOCK * (csb ? 22 : 13) * (cg.nx + 1) * sizeof(uint32_t);
	if (csm > 200 * 1024) return fail("tiled force kernel: more than ~1100 cells per row are not supported by the chunk builder");
	if (csm > 48 * 1024) {
		LS1_CUDA(cudaFuncSetAttribute(k_chunks<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) csm));
		LS1_CUDA(cudaFuncSetAttribute(k_chunks<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) csm));
	}
	LS1_LAUNCH(c, PH_FORCE, k_chunks<false>, gr, 32 * CHUNK_ROWS_PER_BLOCK, csm, cg, c->cell_start, csb, nrows, row_count, (const uint32_t*) nullptr, (Chunk*) nullptr);
	LS1_LAUNCH(c, PH_FORCE, k_chunks_scan, 1, 1024, 0, row_count, row_start, nrows, work_counter);
	LS1_LAUNCH(c, PH_FORCE, k_chunks<true>, gr, 32 * CHUNK_ROWS_PER_BLOCK, csm, cg, c->cell_start, csb, nrows, (uint32_t*) nullptr, row_start, chunks);
	*nrows_out = nrows;
	return 0;
}

}  // namespace ls1b200
