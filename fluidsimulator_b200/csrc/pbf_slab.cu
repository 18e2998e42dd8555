// pbf_slab.cu — x-slab decomposition across GPUs.  Placeholder entry points
// until the halo-exchange path lands; they fail loudly rather than silently doing nothing.
#include <cstring>

#include "pbf_context.h"

extern "C" {

int pbf_comm_unique_id(void* id_bytes) {
  if (id_bytes) std::memset(id_bytes, 0, PBF_COMM_ID_BYTES);
  return PBF_E_COMM;
}

int pbf_comm_init(pbf_ctx* ctx, int, int, const void*) {
  if (ctx) ctx->error = "pbf_comm_init: slab exchange is not built yet";
  return PBF_E_COMM;
}

int pbf_slab_upload(pbf_ctx* ctx, size_t, const float*, const float*, const float*, const float*,
                    const float*, const float*) {
  if (ctx) ctx->error = "pbf_slab_upload: slab exchange is not built yet";
  return PBF_E_COMM;
}

size_t pbf_slab_owned(const pbf_ctx* ctx) { return ctx ? ctx->n : 0; }

int pbf_slab_download(pbf_ctx* ctx, int64_t*, float*, float*, float*, float*, float*, float*) {
  if (ctx) ctx->error = "pbf_slab_download: slab exchange is not built yet";
  return PBF_E_COMM;
}

}  // extern "C"
