// halo.cu -- multi-GPU halo / leaving-molecule exchange buffers (see include/ls1b200.h). Placeholder until the
// decomposition row of SURVEY.md section 8(e) is built: the entry points exist and fail loudly.
#include "common.cuh"

using namespace ls1b200;

extern "C" {

int ls1b200_halo_export(ls1b200_ctx*, const int32_t*, const double*, void**, uint64_t*) { return fail("ls1b200_halo_export: not implemented yet"); }
int ls1b200_halo_import(ls1b200_ctx*, const void*, uint64_t) { return fail("ls1b200_halo_import: not implemented yet"); }
int ls1b200_halo_import_host(ls1b200_ctx*, const ls1b200_particle*, uint64_t) { return fail("ls1b200_halo_import_host: not implemented yet"); }

}
