// Entry points declared in include/centernet_b200.h whose kernels are not written yet.
// Each returns CNB_EUNSUPPORTED (never a CPU fallback); they are replaced one by one.
#include "common.cuh"

#define CNB_STUB(name) \
  cnb::set_error(name ": not implemented yet in this build"); \
  return CNB_EUNSUPPORTED

extern "C" {

size_t cnb_multi_pose_workspace_bytes(int, int, int, int, int, int) { return 0; }
int cnb_multi_pose_decode(const float *, const float *, const float *, const float *, const float *, const float *,
                          int, int, int, int, int, int, float *, void *, size_t, void *) {
  CNB_STUB("cnb_multi_pose_decode");
}
int cnb_edge_aggregate(const float *, float *, int, int, int, int, float, int, void *) {
  CNB_STUB("cnb_edge_aggregate");
}
size_t cnb_exct_workspace_bytes(int, int, int, int, int, int, int) { return 0; }
int cnb_exct_decode(const float *, const float *, const float *, const float *, const float *, const float *,
                    const float *, const float *, const float *, int, int, int, int, int, int, float, float, float,
                    int, int, float *, void *, size_t, void *) {
  CNB_STUB("cnb_exct_decode");
}
size_t cnb_focal_workspace_bytes(long long) { return 0; }
int cnb_focal_loss(const float *, const float *, long long, int, float, float *, float *, void *, size_t, void *) {
  CNB_STUB("cnb_focal_loss");
}
int cnb_splat_gaussian(const int32_t *, const int32_t *, const int32_t *, const int32_t *, const uint8_t *, int, int,
                       int, int, int, float *, void *) {
  CNB_STUB("cnb_splat_gaussian");
}
int cnb_focal_splat_loss(const float *, const int32_t *, const int32_t *, const int32_t *, const int32_t *,
                         const uint8_t *, int, int, int, int, int, int, float, float *, float *, void *, size_t,
                         void *) {
  CNB_STUB("cnb_focal_splat_loss");
}
int cnb_reg_loss(const float *, const void *, const int64_t *, const float *, int, int, int, int, int, float, float *,
                 float *, void *) {
  CNB_STUB("cnb_reg_loss");
}
size_t cnb_dcnv2_workspace_bytes(int, int, int, int, int, int, int, int, int, int, int) { return 0; }
int cnb_dcnv2_forward(const float *, const float *, const float *, const float *, const float *, float *, int, int,
                      int, int, int, int, int, int, int, int, int, int, int, int, void *, size_t, void *) {
  CNB_STUB("cnb_dcnv2_forward");
}
int cnb_dcnv2_backward(const float *, const float *, const float *, const float *, const float *, float *, float *,
                       float *, float *, float *, int, int, int, int, int, int, int, int, int, int, int, int, int,
                       int, void *, size_t, void *) {
  CNB_STUB("cnb_dcnv2_backward");
}

}  // extern "C"
