// PLACEHOLDER (replaced by the tcgen05 implementation later in this round): the symbols exist so the
// library exports everything include/gcf.h declares; every call fails loudly.
#include "common.cuh"
using namespace gcf;
extern "C" size_t gcf_infonce_workspace_bytes(int64_t, int64_t, int32_t) { return 0; }
extern "C" int gcf_infonce_fwd(const float*, int64_t, int64_t, const float*, int64_t, int64_t, int32_t, int32_t, float,
                               const int64_t*, float*, float*, float*, void*, size_t, gcf_stream_t) {
  set_error("gcf_infonce_fwd: not implemented in this build");
  return GCF_EUNSUPPORTED;
}
extern "C" int gcf_infonce_bwd(const float*, int64_t, int64_t, const float*, int64_t, int64_t, int32_t, int32_t, float,
                               const int64_t*, const float*, const float*, const float*, const float*, const float*,
                               float*, int64_t, float*, int64_t, void*, size_t, gcf_stream_t) {
  set_error("gcf_infonce_bwd: not implemented in this build");
  return GCF_EUNSUPPORTED;
}
extern "C" size_t gcf_directau_workspace_bytes(int64_t, int32_t) { return 0; }
extern "C" int gcf_directau_fwd(const float*, int64_t, const float*, int64_t, int64_t, int32_t, float, float*, void*,
                                size_t, gcf_stream_t) {
  set_error("gcf_directau_fwd: not implemented in this build");
  return GCF_EUNSUPPORTED;
}
extern "C" int gcf_directau_bwd(const float*, int64_t, const float*, int64_t, int64_t, int32_t, float, const float*,
                                const float*, float*, int64_t, float*, int64_t, void*, size_t, gcf_stream_t) {
  set_error("gcf_directau_bwd: not implemented in this build");
  return GCF_EUNSUPPORTED;
}
