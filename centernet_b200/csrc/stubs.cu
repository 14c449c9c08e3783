// Entry points declared in include/centernet_b200.h whose kernels are not written yet.
// Each returns CNB_EUNSUPPORTED (never a CPU fallback); they are replaced one by one.
#include "common.cuh"

#define CNB_STUB(name) \
  cnb::set_error(name ": not implemented yet in this build"); \
  return CNB_EUNSUPPORTED

extern "C" {

size_t cnb_exct_workspace_bytes(int, int, int, int, int, int, int) { return 0; }
int cnb_exct_decode(const float *, const float *, const float *, const float *, const float *, const float *,
                    const float *, const float *, const float *, int, int, int, int, int, int, float, float, float,
                    int, int, float *, void *, size_t, void *) {
  CNB_STUB("cnb_exct_decode");
}
}  // extern "C"
