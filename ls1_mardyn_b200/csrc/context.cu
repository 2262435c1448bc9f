// context.cu -- C ABI of include/ls1b200.h: context life cycle, components, upload / download, geometry, timing.
// Reference for the geometry: LinkedCells ctor / rebuild (particleContainer/LinkedCells.cpp:47-58, :148-166).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "common.cuh"

namespace ls1b200 {

static thread_local std::string t_error;
std::atomic<uint64_t> g_launches{0};
void set_error(const std::string& msg) { t_error = msg; }
int fail(const std::string& msg) { t_error = msg; return 1; }

int span_begin(ls1b200_ctx* c, int phase) {
	if (!c->timing) return 0;
	if (c->ev_used + 2 > c->ev_pool.size()) {
		if (c->ev_pool.size() >= 16384) return 0;   // stop recording rather than grow without bound
		for (int i = 0; i < 64; i++) { cudaEvent_t e; LS1_CUDA(cudaEventCreate(&e)); c->ev_pool.push_back(e); }
	}
	ls1b200_ctx::Span s; s.phase = phase; s.a = c->ev_pool[c->ev_used++]; s.b = c->ev_pool[c->ev_used++];
	LS1_CUDA(cudaEventRecord(s.a, c->stream));
	c->spans.push_back(s);
	return 0;
}
int span_end(ls1b200_ctx* c) {
	if (!c->timing || c->spans.empty()) return 0;
	LS1_CUDA(cudaEventRecord(c->spans.back().b, c->stream));
	return 0;
}
static int spans_collect(ls1b200_ctx* c) {
	LS1_CUDA(cudaStreamSynchronize(c->stream));
	for (auto& s : c->spans) {
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) c->phase_ms[s.phase] += ms;
	}
	c->spans.clear(); c->ev_used = 0;
	return 0;
}

template <class T> static int dev_alloc(T** p, size_t n) {
	*p = nullptr;
	if (n == 0) n = 1;
	LS1_CUDA(cudaMalloc((void**) p, n * sizeof(T)));
	return 0;
}
template <class T> static void dev_free(T*& p) { if (p) cudaFree(p); p = nullptr; }

static int alloc_state(State& s, uint64_t cap, bool rot, bool ids) {
	LS1_TRY(dev_alloc(&s.r, 3 * cap)); LS1_TRY(dev_alloc(&s.v, 3 * cap)); LS1_TRY(dev_alloc(&s.meta, cap));
	if (ids) LS1_TRY(dev_alloc(&s.id, cap));
	if (rot) { LS1_TRY(dev_alloc(&s.q, 4 * cap)); LS1_TRY(dev_alloc(&s.L, 3 * cap)); }
	return 0;
}
static void free_state(State& s) { dev_free(s.r); dev_free(s.v); dev_free(s.q); dev_free(s.L); dev_free(s.meta); dev_free(s.id); }

void ctx_free_molecules(ls1b200_ctx* c) {
	rmm_release(c);
	free_state(c->cur); free_state(c->nxt);
	dev_free(c->d_id); dev_free(c->F); dev_free(c->M); dev_free(c->Vi);
	dev_free(c->rank);
	dev_free(c->key); dev_free(c->entry); dev_free(c->cell_of); dev_free(c->cell_of2);
	if (c->mpos) cudaFree(c->mpos); c->mpos = nullptr;
	if (c->spos) cudaFree(c->spos); c->spos = nullptr;
	if (c->sori) cudaFree(c->sori); c->sori = nullptr;
	dev_free(c->sf); dev_free(c->partials); dev_free(c->partials_kin); dev_free(c->perm); c->state_lazy = false;
	decomp_free(c);
	c->cap = 0; c->n_owned = 0; c->n_slots = 0;
}

int ctx_alloc_molecules(ls1b200_ctx* c, uint64_t n_owned, uint64_t cap) {
	ctx_free_molecules(c);
	if (cap >= (1ull << 27)) return fail("ls1b200: more than 2^27 slots per device are not supported");
	c->cap = cap; c->n_owned = n_owned;
	const bool rot = c->rotates;
	LS1_TRY(alloc_state(c->cur, cap, rot, c->dd.active)); LS1_TRY(alloc_state(c->nxt, cap, rot, c->dd.active));
	LS1_TRY(dev_alloc(&c->d_id, n_owned));
	LS1_TRY(dev_alloc(&c->F, 3 * cap)); LS1_TRY(dev_alloc(&c->Vi, 3 * cap));
	if (rot) LS1_TRY(dev_alloc(&c->M, 3 * cap));
	LS1_TRY(dev_alloc(&c->rank, cap));
	LS1_TRY(dev_alloc(&c->key, cap)); LS1_TRY(dev_alloc(&c->entry, cap)); LS1_TRY(dev_alloc(&c->cell_of, cap)); LS1_TRY(dev_alloc(&c->cell_of2, cap));
	const size_t v4 = c->mixed ? 16 : 32;
	// + 64 padding records; mpos[cap] is the far-away sentinel the 1CLJ kernel pads odd neighbour lists with
	LS1_CUDA(cudaMalloc(&c->mpos, (cap + 64) * v4));
	{
		const double far_d[4] = {1e18, 1e18, 1e18, 0.}; const float far_f[4] = {1e18f, 1e18f, 1e18f, 0.f};
		for (int k = 0; k < 64; k++)
			LS1_CUDA(cudaMemcpyAsync((char*) c->mpos + (cap + k) * v4, c->mixed ? (const void*) far_f : (const void*) far_d, v4, cudaMemcpyHostToDevice, c->stream));
		LS1_CUDA(cudaStreamSynchronize(c->stream));
	}
	if (!c->all_1clj) {
		LS1_CUDA(cudaMalloc(&c->spos, std::max<size_t>(cap, 1) * c->smax * v4));
		LS1_CUDA(cudaMalloc(&c->sori, std::max<size_t>(cap, 1) * c->smax * v4));
		LS1_TRY(dev_alloc(&c->sf, cap * c->smax * 9));
	}
	c->partial_blocks = (uint32_t) (((cap + 63) / 64) * std::max(c->smax, 1u) + 64);
	c->partial_blocks = std::max<uint32_t>(c->partial_blocks, c->geom.ncells + 64);   // tiled kernels: one partial per chunk
	{	// force_1clj_v3.cu: molecule-granular chunks (at most one per inner cell, or per 32 slots, plus one per row)
		const uint64_t nrows = (uint64_t) std::max(c->geom.cells[1] - 2, 0) * (uint64_t) std::max(c->geom.cells[2] - 2, 0);
		c->partial_blocks = (uint32_t) std::max<uint64_t>(c->partial_blocks, (uint64_t) c->geom.ncells + cap / 32 + nrows + 64);
	}
	LS1_TRY(dev_alloc(&c->partials, (size_t) c->partial_blocks * 8));
	LS1_TRY(dev_alloc(&c->partials_kin, (size_t) c->partial_blocks * 4));
	if (c->dd.active) LS1_TRY(decomp_alloc(c));
	LS1_CUDA(cudaMemsetAsync(c->F, 0, sizeof(double) * 3 * cap, c->stream));
	LS1_CUDA(cudaMemsetAsync(c->Vi, 0, sizeof(double) * 3 * cap, c->stream));
	if (rot) LS1_CUDA(cudaMemsetAsync(c->M, 0, sizeof(double) * 3 * cap, c->stream));
	return 0;
}

// LinkedCells geometry, host side (LinkedCells.cpp:47-58 ctor path, :148-166 XML path)
static int init_geometry(ls1b200_ctx* c) {
	Geom& g = c->geom;
	const ls1b200_config& f = c->cfg;
	uint64_t ncells = 1;
	for (int d = 0; d < 3; d++) {
		g.bmin[d] = f.box_min[d]; g.bmax[d] = f.box_max[d];
		const double diff = g.bmax[d] - g.bmin[d];
		if (!(diff > 0.)) return fail("ls1b200_create: empty bounding box");
		int box_cells;
		if (f.geometry == LS1B200_GEOM_CTOR) {
			box_cells = (int) floor(diff / f.cutoff);                 // LinkedCells.cpp:47-58
			if (box_cells == 0) box_cells = 1;
		} else {
			const float rc = (float) (f.cutoff / 1);                  // cellsInCutoff == 1; float cast: LinkedCells.cpp:148-166
			box_cells = (int) floor(diff / rc);
		}
		if (box_cells < 1) return fail("ls1b200_create: bounding box smaller than the cutoff radius");
		g.clen[d] = diff / box_cells;
		g.crecip[d] = box_cells / diff;
		g.hmin[d] = g.bmin[d] - g.clen[d];
		g.hmax[d] = g.bmax[d] + g.clen[d];
		g.cells[d] = box_cells + 2;
		g.lbox[d] = diff;
		g.periodic[d] = f.periodic[d] ? 1 : 0;
		g.lo_band[d] = g.bmin[d] + f.cutoff; g.hi_band[d] = g.bmax[d] - f.cutoff;      // one rounding each, like __dadd_rn / __dsub_rn on the device
		// clamp values of DomainDecompBase.cpp:236-240, :358-368 -- nexttoward in vcp_real_calc (float: mixed, double: double build)
		if (c->mixed) {
			const float rmax = (float) g.bmax[d], rmin = (float) g.bmin[d];
			g.wrap_hi[d] = (double) nexttowardf(rmax, (long double) (rmax - 1.f));
			g.halo_hi[d] = (double) nexttowardf(rmax, (long double) (rmax + 1.f));
			g.halo_lo[d] = (double) nexttowardf(rmin, (long double) (rmin - 1.f));
		} else {
			const double rmax = g.bmax[d], rmin = g.bmin[d];
			g.wrap_hi[d] = nexttoward(rmax, (long double) (rmax - 1.));
			g.halo_hi[d] = nexttoward(rmax, (long double) (rmax + 1.));
			g.halo_lo[d] = nexttoward(rmin, (long double) (rmin - 1.));
		}
		ncells *= (uint64_t) g.cells[d];
	}
	g.cutoff = f.cutoff;
	if (ncells + 2 >= (1ull << 26)) return fail("ls1b200_create: more than 2^26 cells per device are not supported");
	g.ncells = (uint32_t) ncells;
	return 0;
}

__global__ void k_deinterleave(const double* __restrict__ src, double* __restrict__ dst, uint64_t n, uint64_t cap, int dim) {
	const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n * dim) return;
	const uint64_t m = i / dim; const int d = (int) (i % dim);
	dst[(uint64_t) d * cap + m] = src[i];
}
__global__ void k_fill_q(double* __restrict__ q, uint64_t n, uint64_t cap) {
	const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	q[i] = 1.; q[cap + i] = 0.; q[2 * cap + i] = 0.; q[3 * cap + i] = 0.;
}
__global__ void k_make_meta(const uint32_t* __restrict__ cid, uint64_t* __restrict__ meta, uint64_t n) {
	const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) meta[i] = make_meta((uint32_t) i, cid[i], false);
}

// out[orig][d] = plane[d][slot] for the owned molecule in every slot
__global__ void k_download_planes(const double* __restrict__ planes, const uint64_t* __restrict__ meta, double* __restrict__ out, uint32_t n_slots, uint64_t cap, int dim,
                                  const uint32_t* __restrict__ owned_index) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n_slots) return;
	const uint64_t m = meta[t];
	if (meta_halo(m)) return;
	const uint64_t o = owned_index ? owned_index[t] : meta_orig(m);      // decomposed mode: slot order of the molecules owned now
	for (int d = 0; d < dim; d++) out[o * dim + d] = planes[(uint64_t) d * cap + t];
}
__global__ void k_upload_planes(double* __restrict__ planes, const uint64_t* __restrict__ meta, const double* __restrict__ in, uint32_t n_slots, uint64_t cap, int dim,
                                const uint32_t* __restrict__ owned_index) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n_slots) return;
	const uint64_t m = meta[t];
	if (meta_halo(m)) return;
	const uint64_t o = owned_index ? owned_index[t] : meta_orig(m);
	for (int d = 0; d < dim; d++) planes[(uint64_t) d * cap + t] = in[o * dim + d];
}
__global__ void k_download_meta(const uint64_t* __restrict__ meta, const uint32_t* __restrict__ cell_of, uint32_t ncells, uint32_t* __restrict__ cid_out,
                                uint64_t* __restrict__ cell_out, uint32_t n_slots, const uint32_t* __restrict__ owned_index, const uint64_t* __restrict__ id_plane,
                                uint64_t* __restrict__ id_out) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n_slots) return;
	const uint64_t m = meta[t];
	if (meta_halo(m)) return;
	const uint64_t o = owned_index ? owned_index[t] : meta_orig(m);
	if (id_out) id_out[o] = id_plane[t];
	if (cid_out) cid_out[o] = meta_cid(m);
	if (cell_out) { const uint32_t c = cell_of[t]; cell_out[o] = (c >= ncells) ? ~0ull : (uint64_t) c; }
}

}  // namespace ls1b200

using namespace ls1b200;

extern "C" {

const char* ls1b200_last_error(void) { return t_error.c_str(); }
int ls1b200_abi_version(void) { return LS1B200_ABI_VERSION; }
uint64_t ls1b200_kernel_launch_count(void) { return g_launches.load(); }

int ls1b200_create(const ls1b200_config* cfg, ls1b200_ctx** out) {
	if (!cfg || !out) return fail("ls1b200_create: null argument");
	*out = nullptr;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
		return fail("ls1b200_create: no CUDA device (this library has no CPU path)");
	if (cfg->device < 0 || cfg->device >= ndev) return fail("ls1b200_create: invalid device ordinal");
	if (!(cfg->cutoff > 0.) || !(cfg->cutoff_lj > 0.)) return fail("ls1b200_create: cutoff radii must be positive");
	if (cfg->precision != LS1B200_MIXED && cfg->precision != LS1B200_DOUBLE) return fail("ls1b200_create: unknown precision");
	ls1b200_ctx* c = new (std::nothrow) ls1b200_ctx();
	if (!c) return fail("ls1b200_create: out of host memory");
	c->cfg = *cfg; c->device = cfg->device; c->mixed = cfg->precision == LS1B200_MIXED;
	if (c->cfg.cutoff < c->cfg.cutoff_lj) c->cfg.cutoff = c->cfg.cutoff_lj;   // Simulation.cpp:267: rc = max(rc, rcLJ)
	int rc = init_geometry(c);
	if (rc) { delete c; return rc; }
	auto cleanup = [&](int code) { ls1b200_destroy(c); return code; };
	if (cudaSetDevice(c->device) != cudaSuccess) return cleanup(fail("ls1b200_create: cudaSetDevice failed"));
	if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return cleanup(fail("ls1b200_create: cannot create stream"));
	// fine x-bins per reference cell (the sort key inside a cell; the force kernels trim their candidate ranges with it)
	c->fine_s = 1;      // default off: measured +0.08 ms in the sort and no gain in the force kernel (profiles/r02_force_kernel_experiments.md)
	if (const char* e = getenv("LS1B200_FINE_S")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) c->fine_s = (uint32_t) v; }
	if ((uint64_t) c->geom.ncells * c->fine_s + 2 >= (1ull << 28)) c->fine_s = 1;
	c->fine_inv = (double) c->fine_s * c->geom.crecip[0];
	const size_t nb = (size_t) c->geom.ncells * c->fine_s + 4;
	if (dev_alloc(&c->cell_count, 2 * nb) || dev_alloc(&c->cell_start, (size_t) c->geom.ncells + 4) || dev_alloc(&c->cursor, nb) || dev_alloc(&c->scan_tmp, nb / 2048 + 16)) return cleanup(1);
	if (c->fine_s > 1 && dev_alloc(&c->cell_start_fine, nb)) return cleanup(1);
	c->cell_count_halo = c->cell_count + nb;      // one allocation, one memset per update
	{
		// border table, evaluated exactly like cell_border<D> (one fused multiply-add; the first / last two are the stored constants)
		int maxc = std::max(c->geom.cells[0], std::max(c->geom.cells[1], c->geom.cells[2]));
		c->border_stride = (uint32_t) maxc + 2;
		std::vector<double> tab(3 * (size_t) c->border_stride, 0.);
		for (int d = 0; d < 3; d++) for (int i = 0; i <= c->geom.cells[d]; i++) {
			double b;
			if (i == 0) b = c->geom.hmin[d]; else if (i == 1) b = c->geom.bmin[d]; else if (i == c->geom.cells[d] - 1) b = c->geom.bmax[d];
			else if (i == c->geom.cells[d]) b = c->geom.hmax[d]; else b = fma((double) i, c->geom.clen[d], c->geom.hmin[d]);
			tab[(size_t) d * c->border_stride + i] = b;
		}
		if (dev_alloc(&c->d_border, tab.size())) return cleanup(1);
		if (cudaMemcpy(c->d_border, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice) != cudaSuccess) return cleanup(fail("ls1b200_create: memcpy failed"));
	}
	{
		const size_t nrows = (size_t) c->geom.cells[1] * c->geom.cells[2] + 8;
		if (cudaMalloc(&c->chunks, ((size_t) c->geom.ncells + 8) * 8) != cudaSuccess) return cleanup(fail("ls1b200_create: out of device memory"));
		if (dev_alloc(&c->chunk_row_count, nrows) || dev_alloc(&c->chunk_row_start, nrows) || dev_alloc(&c->work_counter, 4)) return cleanup(1);
		cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device);
		const char* k = getenv("LS1B200_LJ_KERNEL");
		c->lj_kernel = (k && std::string(k) == "global") ? 1 : (k && std::string(k) == "mma") ? 3 : (k && std::string(k) == "tc") ? 4 : (k && std::string(k) == "v3") ? 5 : (k && std::string(k) == "pf") ? 6 : 0;      // default: the tiled kernel (v3 measured slower, profiles/r02_force_kernel_experiments.md)
		if (const char* pc = getenv("LS1B200_PF_CHECK")) c->pf_check = atoi(pc);
		if (const char* fu = getenv("LS1B200_FUSED")) c->fused = atoi(fu);
		if (const char* lz = getenv("LS1B200_LAZY")) c->lazy = atoi(lz);
		if (const char* ce = getenv("LS1B200_CHUNKS_EARLY")) c->chunks_early = atoi(ce);
		if (const char* o2 = getenv("LS1B200_OVERLAP2")) c->overlap2 = atoi(o2);
		if (const char* tm = getenv("LS1B200_TILE_MAXCTAS")) c->tile_max_ctas = atoi(tm);
		if (const char* tp = getenv("LS1B200_TILE_PASS1_CTAS")) c->tile_pass1_ctas = atoi(tp);
		const char* m = getenv("LS1B200_MULTI_KERNEL");
		if (const char* mc = getenv("LS1B200_MULTI_MAXCTAS")) c->multi_max_ctas = atoi(mc);
		if (const char* ms = getenv("LS1B200_MULTI_STAGE")) c->multi_stage = atoi(ms);
		if (const char* mp = getenv("LS1B200_MULTI_PD2")) c->multi_pd2 = atoi(mp);
		if (const char* ov = getenv("LS1B200_OVERLAP")) c->overlap = atoi(ov);
		if (const char* st = getenv("LS1B200_SPLIT_TEST")) c->split_test = atoi(st);
		c->multi_kernel = (m && std::string(m) == "general") ? 1 : (m && std::string(m) == "generic") ? 2 : 0;   // 2: tiled, but never specialised
	}
	if (dev_alloc(&c->d_scal, 1) || dev_alloc(&c->partials2, 1024 * 8)) return cleanup(1);
	if (cudaMallocHost((void**) &c->h_scal, sizeof(DevScalars)) != cudaSuccess) return cleanup(fail("ls1b200_create: cudaMallocHost failed"));
	memset(c->h_scal, 0, sizeof(DevScalars));
	c->h_scal->beta_trans = 1.; c->h_scal->beta_rot = 1.;
	if (cudaMemcpy(c->d_scal, c->h_scal, sizeof(DevScalars), cudaMemcpyHostToDevice) != cudaSuccess) return cleanup(fail("ls1b200_create: memcpy failed"));
	*out = c;
	return 0;
}

int ls1b200_destroy(ls1b200_ctx* c) {
	if (!c) return 0;
	cudaSetDevice(c->device);
	if (c->stream) cudaStreamSynchronize(c->stream);
	ctx_free_molecules(c);
	dev_free(c->d_comps); dev_free(c->d_sites);
	dev_free(c->d_eps24); dev_free(c->d_sig2); dev_free(c->d_shift6); dev_free(c->d_eps24f); dev_free(c->d_sig2f); dev_free(c->d_shift6f);
	dev_free(c->cell_start_alt);
	dev_free(c->cell_count_b); c->cell_start_b = c->cursor_b = c->scan_tmp_b = nullptr;
	dev_free(c->cell_count); dev_free(c->cell_start); dev_free(c->cell_start_fine); dev_free(c->cursor); dev_free(c->scan_tmp); dev_free(c->d_border);
	dev_free(c->d_scal); dev_free(c->d_records); dev_free(c->partials2);
	if (c->chunks) cudaFree(c->chunks);
	if (c->chunks_b) cudaFree(c->chunks_b);
	dev_free(c->chunk_row_count_b);
	if (c->chunks2) cudaFree(c->chunks2);
	dev_free(c->chunk_row_count); dev_free(c->chunk_row_start); dev_free(c->work_counter);
	if (c->h_scal) cudaFreeHost(c->h_scal);
	for (auto e : c->ev_pool) cudaEventDestroy(e);
	for (auto e : c->xprof_ev) if (e) cudaEventDestroy(e);
	decomp_free(c);
	if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); }
	if (c->ev_m) cudaEventDestroy(c->ev_m);
	if (c->stream_c) { cudaStreamSynchronize(c->stream_c); cudaStreamDestroy(c->stream_c); if (c->ev_c0) cudaEventDestroy(c->ev_c0); if (c->ev_c1) cudaEventDestroy(c->ev_c1); }
	if (c->stream_x) { cudaStreamSynchronize(c->stream_x); cudaStreamDestroy(c->stream_x); if (c->ev_xa) cudaEventDestroy(c->ev_xa); if (c->ev_xb) cudaEventDestroy(c->ev_xb); }
	if (c->ev_dl) cudaEventDestroy(c->ev_dl);
	if (c->stage2) cudaFree(c->stage2);
	if (c->stage) cudaFree(c->stage);
	if (c->stream) cudaStreamDestroy(c->stream);
	delete c;
	return 0;
}

int ls1b200_set_components(ls1b200_ctx* c, uint32_t n_components, const ls1b200_component* comps, uint32_t n_sites, const ls1b200_site* sites,
                           uint32_t n_lj, const double* eps24, const double* sig2, const double* shift6) {
	if (!c || !comps || n_components == 0) return fail("ls1b200_set_components: null / empty argument");
	if (n_components > 0xFFFF) return fail("ls1b200_set_components: too many components");
	if (c->n_owned) return fail("ls1b200_set_components: must be called before ls1b200_upload");
	LS1_CUDA(cudaSetDevice(c->device));
	c->comps.clear(); c->sites.clear();
	c->smax = 1; c->all_1clj = true; bool any_inertia = false;
	for (uint32_t i = 0; i < n_components; i++) {
		const ls1b200_component& s = comps[i];
		CompDev d{};
		d.n[0] = s.n_lj; d.n[1] = s.n_charge; d.n[2] = s.n_dipole; d.n[3] = s.n_quadrupole;
		d.nsites = s.n_lj + s.n_charge + s.n_dipole + s.n_quadrupole;
		d.site_offset = s.site_offset; d.lj_lookup_id = s.lj_lookup_id; d.rot_dof = s.rot_dof;
		if (d.site_offset + d.nsites > n_sites) return fail("ls1b200_set_components: site table too short");
		// components made of massless sites only (dipoles / quadrupoles, as in the reference's Vectorization*.inp force fixtures)
		// have m = 0: forces are defined, integration is not (the reference divides by m as well)
		if (!(s.mass >= 0.)) return fail("ls1b200_set_components: negative component mass");
		d.mass = s.mass; d.inv_mass = s.mass > 0. ? 1. / s.mass : 0.;
		for (int k = 0; k < 3; k++) { d.I[k] = s.inertia[k]; d.invI[k] = (s.inertia[k] != 0.) ? 1. / s.inertia[k] : 0.; if (s.inertia[k] != 0.) any_inertia = true; }
		c->smax = std::max(c->smax, d.nsites);
		bool one_clj = d.n[0] == 1 && d.nsites == 1 && d.lj_lookup_id == i;
		if (one_clj) { const ls1b200_site& st = sites[d.site_offset]; one_clj = st.r[0] == 0. && st.r[1] == 0. && st.r[2] == 0.; }
		c->all_1clj = c->all_1clj && one_clj;
		c->comps.push_back(d);
	}
	c->sites.resize(n_sites);
	for (uint32_t i = 0; i < n_components; i++) {
		const CompDev& d = c->comps[i];
		uint32_t k = d.site_offset;
		for (int t = 0; t < 4; t++) for (uint32_t j = 0; j < d.n[t]; j++, k++) {
			SiteDev& o = c->sites[k];
			for (int x = 0; x < 3; x++) { o.r[x] = sites[k].r[x]; o.e[x] = sites[k].e[x]; }
			o.param = sites[k].param; o.type = (uint32_t) t; o.ljid = (t == 0) ? d.lj_lookup_id + j : 0u;
			if (t == 0 && o.ljid >= n_lj) return fail("ls1b200_set_components: LJ lookup id outside the pair tables");
		}
	}
	c->rotates = !c->all_1clj || any_inertia;
	c->n_lj = n_lj;
	const size_t nt = (size_t) n_lj * n_lj;
	if (nt && (!eps24 || !sig2 || !shift6)) return fail("ls1b200_set_components: LJ tables missing");
	c->h_eps24.assign(eps24, eps24 + nt); c->h_sig2.assign(sig2, sig2 + nt); c->h_shift6.assign(shift6, shift6 + nt);
	if (nt == 0) { c->h_eps24.assign(1, 0.); c->h_sig2.assign(1, 0.); c->h_shift6.assign(1, 0.); }
	dev_free(c->d_comps); dev_free(c->d_sites);
	dev_free(c->d_eps24); dev_free(c->d_sig2); dev_free(c->d_shift6); dev_free(c->d_eps24f); dev_free(c->d_sig2f); dev_free(c->d_shift6f);
	LS1_TRY(dev_alloc(&c->d_comps, c->comps.size())); LS1_TRY(dev_alloc(&c->d_sites, c->sites.size()));
	LS1_CUDA(cudaMemcpy(c->d_comps, c->comps.data(), sizeof(CompDev) * c->comps.size(), cudaMemcpyHostToDevice));
	if (!c->sites.empty()) LS1_CUDA(cudaMemcpy(c->d_sites, c->sites.data(), sizeof(SiteDev) * c->sites.size(), cudaMemcpyHostToDevice));
	const size_t na = c->h_eps24.size();
	std::vector<float> f24(na), fs2(na), fsh(na);
	for (size_t i = 0; i < na; i++) { f24[i] = (float) c->h_eps24[i]; fs2[i] = (float) c->h_sig2[i]; fsh[i] = (float) c->h_shift6[i]; }
	LS1_TRY(dev_alloc(&c->d_eps24, na)); LS1_TRY(dev_alloc(&c->d_sig2, na)); LS1_TRY(dev_alloc(&c->d_shift6, na));
	LS1_TRY(dev_alloc(&c->d_eps24f, na)); LS1_TRY(dev_alloc(&c->d_sig2f, na)); LS1_TRY(dev_alloc(&c->d_shift6f, na));
	LS1_CUDA(cudaMemcpy(c->d_eps24, c->h_eps24.data(), 8 * na, cudaMemcpyHostToDevice));
	LS1_CUDA(cudaMemcpy(c->d_sig2, c->h_sig2.data(), 8 * na, cudaMemcpyHostToDevice));
	LS1_CUDA(cudaMemcpy(c->d_shift6, c->h_shift6.data(), 8 * na, cudaMemcpyHostToDevice));
	LS1_CUDA(cudaMemcpy(c->d_eps24f, f24.data(), 4 * na, cudaMemcpyHostToDevice));
	LS1_CUDA(cudaMemcpy(c->d_sig2f, fs2.data(), 4 * na, cudaMemcpyHostToDevice));
	LS1_CUDA(cudaMemcpy(c->d_shift6f, fsh.data(), 4 * na, cudaMemcpyHostToDevice));
	c->components_set = true;
	return 0;
}

uint64_t ls1b200_num_molecules(const ls1b200_ctx* c) { return c ? (c->dd.active ? c->n_owned_local : c->n_owned) : 0; }
void* ls1b200_stream(ls1b200_ctx* c) { return c ? (void*) c->stream : nullptr; }

int ls1b200_geometry(const ls1b200_ctx* c, int32_t cells_per_dim[3], double cell_length[3], double halo_min[3]) {
	if (!c) return fail("ls1b200_geometry: null context");
	for (int d = 0; d < 3; d++) { cells_per_dim[d] = c->geom.cells[d]; cell_length[d] = c->geom.clen[d]; halo_min[d] = c->geom.hmin[d]; }
	return 0;
}

int ls1b200_upload(ls1b200_ctx* c, uint64_t n, const uint64_t* id, const uint32_t* cid, const double* r, const double* v, const double* q, const double* L) {
	if (!c) return fail("ls1b200_upload: null context");
	if (!c->components_set) return fail("ls1b200_upload: call ls1b200_set_components first");
	if (n && (!id || !cid || !r || !v)) return fail("ls1b200_upload: null array");
	LS1_CUDA(cudaSetDevice(c->device));
	for (uint64_t i = 0; i < n; i++) if (cid[i] >= c->comps.size()) return fail("ls1b200_upload: component id out of range");
	// slot capacity: owned molecules + periodic halo copies (boundary layer of width rc on every periodic face) with head room
	double fac = 1.;
	for (int d = 0; d < 3; d++) if (c->geom.periodic[d] || c->cfg.reserved[0] || c->dd.active) fac *= 1. + 2. * std::min(1., c->geom.cutoff / c->geom.lbox[d]);
	// decomposed mode: + 10 % for molecules migrating in
	uint64_t cap = n + (uint64_t) ((fac - 1.) * 1.5 * (double) n) + (c->dd.active ? n / 10 + 16384 : 0) + 4096;
	cap = (cap + 63) & ~63ull;      // even (the 16-byte streaming kernels handle two slots per thread); planes stay 512-byte aligned
	LS1_TRY(ctx_alloc_molecules(c, n, cap));
	c->rot_dof_total = 0;
	for (uint64_t i = 0; i < n; i++) c->rot_dof_total += c->comps[cid[i]].rot_dof;
	if (n) {
		double* stage = nullptr; uint32_t* cstage = nullptr;
		LS1_TRY(dev_alloc(&stage, 4 * n)); LS1_TRY(dev_alloc(&cstage, n));
		const uint32_t g3 = (uint32_t) ((3 * n + 255) / 256), g4 = (uint32_t) ((4 * n + 255) / 256), g1 = (uint32_t) ((n + 255) / 256);
		auto put = [&](const double* src, double* planes, int dim) -> int {
			LS1_CUDA(cudaMemcpyAsync(stage, src, sizeof(double) * dim * n, cudaMemcpyHostToDevice, c->stream));
			k_deinterleave<<<dim == 3 ? g3 : g4, 256, 0, c->stream>>>(stage, planes, n, cap, dim);
			g_launches++;
			LS1_CUDA(cudaGetLastError());
			return 0;
		};
		LS1_TRY(put(r, c->cur.r, 3)); LS1_TRY(put(v, c->cur.v, 3));
		if (c->rotates) {
			if (q) LS1_TRY(put(q, c->cur.q, 4)); else { k_fill_q<<<g1, 256, 0, c->stream>>>(c->cur.q, n, cap); g_launches++; }
			if (L) LS1_TRY(put(L, c->cur.L, 3)); else LS1_CUDA(cudaMemsetAsync(c->cur.L, 0, sizeof(double) * 3 * cap, c->stream));
		}
		LS1_CUDA(cudaMemcpyAsync(cstage, cid, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, c->stream));
		k_make_meta<<<g1, 256, 0, c->stream>>>(cstage, c->cur.meta, n); g_launches++;
		LS1_CUDA(cudaMemcpyAsync(c->d_id, id, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, c->stream));
		if (c->dd.active) LS1_CUDA(cudaMemcpyAsync(c->cur.id, id, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, c->stream));
		LS1_CUDA(cudaStreamSynchronize(c->stream));
		cudaFree(stage); cudaFree(cstage);
	}
	c->n_owned_local = n; c->rot_dof_local = c->rot_dof_total; c->dd.rot_dof_valid = false;
	c->parity = 0; c->n_slots = (uint32_t) n;
	c->h_scal->n_slots2[0] = (uint32_t) n; c->h_scal->n_slots2[1] = 0; c->h_scal->overflow = 0;
	c->h_scal->beta_trans = 1.; c->h_scal->beta_rot = 1.;
	c->h_scal->clip_limit = 0.; c->h_scal->expl_counter = c->h_scal->expl_warning = c->h_scal->expl_error = 0; c->globals_initialised = false;      // a new state: Domain's explosion heuristics start over
	LS1_CUDA(cudaMemcpy(c->d_scal, c->h_scal, sizeof(DevScalars), cudaMemcpyHostToDevice));
	c->sorted = false; c->ever_sorted = false; c->forces_valid = false; c->keys_valid = false;
	return 0;
}

static int stage_reserve(ls1b200_ctx* c, size_t bytes) {
	if (bytes <= c->stage_bytes) return 0;
	if (c->stage) { LS1_CUDA(cudaStreamSynchronize(c->stream)); cudaFree(c->stage); c->stage = nullptr; c->stage_bytes = 0; }
	LS1_CUDA(cudaMalloc(&c->stage, bytes));
	c->stage_bytes = bytes;
	return 0;
}

}  // extern "C"
namespace ls1b200 {
const char* overflow_message(uint32_t code) {
	switch (code) {
		case 2: return "a halo message buffer was too small (more boundary molecules than the message capacity)";
		case 4: return "a neighbour's message did not arrive within the time limit (peer-memory transport, LS1B200_PEER_TIMEOUT_S: did a rank fail or skip ls1b200_exchange?); the state was frozen at the failed step";
		case 5: return "another rank of the decomposition reported a failed step (overflow or time-out); the state was frozen at that step";
		case 6: return "force_1clj_pf (LS1B200_PF_CHECK): a window translation was not exact";
		default: return "slot capacity exceeded (more halo copies than reserved)";
	}
}
}  // namespace ls1b200
extern "C" {

// brings n_slots / overflow to the host (synchronises the stream)
static int sync_scalars(ls1b200_ctx* c, bool owned_index = false) {
	if (c->dd.active && c->cap) LS1_TRY(count_owned(c, owned_index));
	LS1_CUDA(cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(DevScalars), cudaMemcpyDeviceToHost, c->stream));
	LS1_CUDA(cudaStreamSynchronize(c->stream));
	c->n_slots = std::min<uint64_t>(c->h_scal->n_slots2[c->parity], c->cap);
	if (c->h_scal->overflow) return fail(std::string("ls1b200: ") + overflow_message(c->h_scal->overflow));
	return 0;
}

int ls1b200_update(ls1b200_ctx* c) {
	if (!c) return fail("ls1b200_update: null context");
	LS1_CUDA(cudaSetDevice(c->device));
	const bool kv = c->keys_valid;      // ls1b200_leapfrog_pre already binned the slots it integrated
	c->keys_valid = false;
	if (c->split_test && c->dd.active && c->region_flag && c->ever_sorted) {      // LS1B200_SPLIT_TEST: inner slots' histogram first, then the rest + the sort
		LS1_TRY(do_update(c, 1));
		LS1_TRY(do_update(c, 2));
	} else
	LS1_TRY(do_update(c, kv ? 3 : 0));
	return sync_scalars(c);
}

int ls1b200_traverse_cells(ls1b200_ctx* c, ls1b200_macro* out) {
	if (!c) return fail("ls1b200_traverse_cells: null context");
	if (!c->sorted) return fail("ls1b200_traverse_cells: call ls1b200_update first");
	LS1_CUDA(cudaSetDevice(c->device));
	LS1_TRY(do_forces(c));
	LS1_TRY(sync_scalars(c));
	if (out) {
		const DevScalars& s = *c->h_scal;
		out->upot6lj = s.upot6lj; out->upot_xpoles = s.upot_xpoles; out->virial_raw = s.virial; out->my_rf = s.my_rf;
		out->upot = s.upot6lj / 6.0 + s.upot_xpoles + s.my_rf;      // VectorizedCellProcessor.cpp:153-154
		out->virial = s.virial + 3.0 * s.my_rf;
	}
	return 0;
}

int ls1b200_download(ls1b200_ctx* c, int field, void* dst, uint64_t n) {
	if (!c || !dst) return fail("ls1b200_download: null argument");
	LS1_CUDA(cudaSetDevice(c->device));
	if (c->dd.active) {
		if (!c->cap) return fail("ls1b200_download: call ls1b200_upload first");
		LS1_TRY(sync_scalars(c, true));      // entry[slot] = index among the owned slots
		if (n != c->n_owned_local) return fail("ls1b200_download: n does not match ls1b200_num_molecules()");
	} else if (n != c->n_owned) return fail("ls1b200_download: n does not match the number of molecules");
	if (n == 0) return 0;
	if (!c->dd.active) LS1_TRY(sync_scalars(c));
	const uint32_t* oi = c->dd.active ? c->entry : nullptr;
	const uint32_t ns = c->n_slots, g = (ns + 255) / 256;
	const double* planes = nullptr; int dim = 3;
	switch (field) {
		case LS1B200_F_ID: if (c->dd.active) break; LS1_CUDA(cudaMemcpy(dst, c->d_id, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost)); return 0;
		case LS1B200_F_R: planes = c->cur.r; break;
		case LS1B200_F_V: planes = c->cur.v; break;
		case LS1B200_F_Q: planes = c->cur.q; dim = 4; break;
		case LS1B200_F_L: planes = c->cur.L; break;
		case LS1B200_F_F: planes = c->F; break;
		case LS1B200_F_M: planes = c->M; break;
		case LS1B200_F_VI: planes = c->Vi; break;
		case LS1B200_F_CID: case LS1B200_F_CELL: break;
		default: return fail("ls1b200_download: unknown field");
	}
	if (field == LS1B200_F_CID || field == LS1B200_F_CELL || field == LS1B200_F_ID) {
		if (field == LS1B200_F_CELL && !c->sorted) return fail("ls1b200_download: cell indices need ls1b200_update first");
		const size_t es = field == LS1B200_F_CID ? 4 : 8;
		LS1_TRY(stage_reserve(c, es * n));
		void* tmp = c->stage;
		k_download_meta<<<g, 256, 0, c->stream>>>(c->cur.meta, c->cell_of2, c->geom.ncells, field == LS1B200_F_CID ? (uint32_t*) tmp : nullptr,
		                                         field == LS1B200_F_CELL ? (uint64_t*) tmp : nullptr, ns, oi, c->cur.id, field == LS1B200_F_ID ? (uint64_t*) tmp : nullptr);
		g_launches++;
		LS1_CUDA(cudaMemcpyAsync(dst, tmp, es * n, cudaMemcpyDeviceToHost, c->stream));
		LS1_CUDA(cudaStreamSynchronize(c->stream));
		return 0;
	}
	if (!planes) {   // fields a non-rotating ensemble does not store: identity orientation, zero angular momentum / torque
		double* o = (double*) dst;
		for (uint64_t i = 0; i < n * dim; i++) o[i] = 0.;
		if (field == LS1B200_F_Q) for (uint64_t i = 0; i < n; i++) o[4 * i] = 1.;
		return 0;
	}
	LS1_TRY(stage_reserve(c, sizeof(double) * dim * n));
	double* tmp = (double*) c->stage;
	k_download_planes<<<g, 256, 0, c->stream>>>(planes, c->cur.meta, tmp, ns, c->cap, dim, oi);
	g_launches++;
	LS1_CUDA(cudaGetLastError());
	LS1_CUDA(cudaMemcpyAsync(dst, tmp, sizeof(double) * dim * n, cudaMemcpyDeviceToHost, c->stream));
	LS1_CUDA(cudaStreamSynchronize(c->stream));
	return 0;
}

int ls1b200_download_begin(ls1b200_ctx* c, int field, void* dst, uint64_t n) {
	if (!c || !dst) return fail("ls1b200_download_begin: null argument");
	if (c->dd.active) return fail("ls1b200_download_begin: not available on a decomposed context (use ls1b200_download)");
	if (c->dl_pending) return fail("ls1b200_download_begin: a download is already in flight (call ls1b200_download_end)");
	if (n != c->n_owned) return fail("ls1b200_download_begin: n does not match the number of molecules");
	LS1_CUDA(cudaSetDevice(c->device));
	const double* planes = nullptr; int dim = 3;
	switch (field) {
		case LS1B200_F_R: planes = c->cur.r; break;
		case LS1B200_F_V: planes = c->cur.v; break;
		case LS1B200_F_Q: planes = c->cur.q; dim = 4; break;
		case LS1B200_F_L: planes = c->cur.L; break;
		case LS1B200_F_F: planes = c->F; break;
		case LS1B200_F_M: planes = c->M; break;
		case LS1B200_F_VI: planes = c->Vi; break;
		default: return fail("ls1b200_download_begin: field must be one of R, V, Q, L, F, M, VI");
	}
	if (!planes) return fail("ls1b200_download_begin: this ensemble does not store the field (use ls1b200_download)");
	if (n == 0) return 0;
	LS1_TRY(sync_scalars(c));      // the slot count of the current sort
	if (!c->stream2) {
		LS1_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
		LS1_CUDA(cudaEventCreateWithFlags(&c->ev_dl, cudaEventDisableTiming));
	}
	const size_t bytes = sizeof(double) * dim * n;
	if (bytes > c->stage2_bytes) {
		if (c->stage2) { LS1_CUDA(cudaStreamSynchronize(c->stream2)); cudaFree(c->stage2); c->stage2 = nullptr; c->stage2_bytes = 0; }
		LS1_CUDA(cudaMalloc(&c->stage2, bytes));
		c->stage2_bytes = bytes;
	}
	const uint32_t ns = c->n_slots;
	// the gather into upload order runs on the main stream (ordered with the kernels that wrote the field), the copy to the host
	// on the side stream: it overlaps whatever the caller enqueues next
	k_download_planes<<<(ns + 255) / 256, 256, 0, c->stream>>>(planes, c->cur.meta, (double*) c->stage2, ns, c->cap, dim, (const uint32_t*) nullptr);
	g_launches++;
	LS1_CUDA(cudaGetLastError());
	LS1_CUDA(cudaEventRecord(c->ev_dl, c->stream));
	LS1_CUDA(cudaStreamWaitEvent(c->stream2, c->ev_dl, 0));
	LS1_CUDA(cudaMemcpyAsync(dst, c->stage2, bytes, cudaMemcpyDeviceToHost, c->stream2));
	c->dl_pending = true;
	return 0;
}

int ls1b200_download_end(ls1b200_ctx* c) {
	if (!c) return fail("ls1b200_download_end: null context");
	if (!c->dl_pending) return 0;
	LS1_CUDA(cudaSetDevice(c->device));
	LS1_CUDA(cudaStreamSynchronize(c->stream2));
	c->dl_pending = false;
	return 0;
}

int ls1b200_set_field(ls1b200_ctx* c, int field, const void* src, uint64_t n) {
	if (!c || !src) return fail("ls1b200_set_field: null argument");
	LS1_CUDA(cudaSetDevice(c->device));
	if (c->dd.active) {
		LS1_TRY(sync_scalars(c, true));
		if (n != c->n_owned_local) return fail("ls1b200_set_field: n does not match ls1b200_num_molecules()");
	} else if (n != c->n_owned) return fail("ls1b200_set_field: n does not match the number of molecules");
	if (n == 0) return 0;
	if (!c->dd.active) LS1_TRY(sync_scalars(c));
	const uint32_t* oi = c->dd.active ? c->entry : nullptr;
	double* planes = nullptr; int dim = 3;
	switch (field) {
		// new positions / orientations: the cell order, the halo copies and the calc-precision records (mpos, spos, sori) are stale --
		// ls1b200_traverse_cells must not run before the next ls1b200_update
		case LS1B200_F_R: planes = c->cur.r; c->forces_valid = false; c->ever_sorted = false; c->sorted = false; c->keys_valid = false; break;      // arbitrary new positions: the exchange scans every slot
		case LS1B200_F_V: planes = c->cur.v; break;
		case LS1B200_F_Q: planes = c->cur.q; dim = 4; if (c->cur.q) { c->forces_valid = false; c->sorted = false; c->keys_valid = false; } break;
		case LS1B200_F_L: planes = c->cur.L; break;
		case LS1B200_F_F: planes = c->F; break;
		case LS1B200_F_M: planes = c->M; break;
		default: return fail("ls1b200_set_field: field cannot be set");
	}
	if (!planes) return 0;   // non-rotating ensemble: nothing to store
	LS1_TRY(stage_reserve(c, sizeof(double) * dim * n));
	double* tmp = (double*) c->stage;
	LS1_CUDA(cudaMemcpyAsync(tmp, src, sizeof(double) * dim * n, cudaMemcpyHostToDevice, c->stream));
	const uint32_t ns = c->n_slots;
	k_upload_planes<<<(ns + 255) / 256, 256, 0, c->stream>>>(planes, c->cur.meta, tmp, ns, c->cap, dim, oi);
	g_launches++;
	LS1_CUDA(cudaGetLastError());
	LS1_CUDA(cudaStreamSynchronize(c->stream));      // src is borrowed for the duration of the call only
	return 0;
}

int ls1b200_phase_times(ls1b200_ctx* c, double ms[4], uint64_t launches[4], int reset) {
	if (!c) return fail("ls1b200_phase_times: null context");
	LS1_CUDA(cudaSetDevice(c->device));
	LS1_TRY(spans_collect(c));
	for (int i = 0; i < PH_COUNT; i++) { if (ms) ms[i] = c->phase_ms[i]; if (launches) launches[i] = c->phase_launches[i]; }
	if (reset) for (int i = 0; i < PH_COUNT; i++) { c->phase_ms[i] = 0.; c->phase_launches[i] = 0; }
	return 0;
}

int ls1b200_exchange_overlapped_ms(ls1b200_ctx* c, double* ms, int reset) {
	if (!c || !ms) return fail("ls1b200_exchange_overlapped_ms: null argument");
	LS1_CUDA(cudaSetDevice(c->device));
	LS1_TRY(spans_collect(c));
	*ms = c->phase_ms[PH_XCHG];
	if (reset) { c->phase_ms[PH_XCHG] = 0.; c->phase_launches[PH_XCHG] = 0; }
	return 0;
}

}  // extern "C"
