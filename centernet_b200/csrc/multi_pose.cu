// multi_pose_decode (models/decode.py:497-571) on top of the fused NMS+top-K selection:
//   select(heat)  -> K centres per image         (decode.py:503-504)
//   select(hm_hp) -> K peaks per (image, joint)  (:528-533, _topk_channel on _nms(hm_hp))
//   k_pose_assemble: one CTA per (image, joint), one thread per detection; the joint's K
//   heat-map candidates live in shared memory; each thread does the gathers, the nearest-
//   candidate search and the reject tests of :534-566 and writes its slice of the 40-column row.
#include "select.cuh"

namespace cnb {

struct PoseArgs {
  // centre top-K (per image)
  const float *c_scores; const int64_t *c_inds; const int32_t *c_clses; const float *c_ys; const float *c_xs;
  // joint top-K (per image*joint), null when hm_hp is absent
  const float *j_scores; const int64_t *j_inds; const float *j_ys; const float *j_xs;
  const float *wh, *kps, *reg, *hp_offset;
  int B, J, H, W, K;
  float *dets;  // [B, K, 4 + 1 + 2J + 1]
};

__global__ void __launch_bounds__(1024) k_pose_assemble(const PoseArgs a) {
  extern __shared__ float sm[];  // [3][K]: candidate x, y, score of this (image, joint)
  const int b = blockIdx.x / a.J, j = blockIdx.x - b * a.J;
  const int K = a.K, D = 4 + 1 + 2 * a.J + 1;
  const long long HW = (long long)a.H * a.W;
  float *cx = sm, *cy = sm + K, *cs = sm + 2 * K;
  const bool has_hp = a.j_scores != nullptr;
  if (has_hp) {
    for (int m = threadIdx.x; m < K; m += blockDim.x) {
      const size_t o = ((size_t)b * a.J + j) * K + m;
      float s = a.j_scores[o], x = a.j_xs[o], y = a.j_ys[o];
      if (a.hp_offset) {  // :534-539
        const long long sp = a.j_inds[o];
        x += a.hp_offset[((long long)b * 2) * HW + sp];
        y += a.hp_offset[((long long)b * 2 + 1) * HW + sp];
      } else {  // :540-542
        x += 0.5f;
        y += 0.5f;
      }
      if (!(s > 0.1f)) {  // :544-547  mask = (hm_score > thresh)
        s = -1.0f;
        x = -10000.0f;
        y = -10000.0f;
      }
      cx[m] = x; cy[m] = y; cs[m] = s;
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const size_t o = (size_t)b * K + k;
    const long long sp = a.c_inds[o];
    const float xi = a.c_xs[o], yi = a.c_ys[o];
    // regressed joint location uses the INTEGER centre (:506-509)
    float kx = a.kps[((long long)b * 2 * a.J + 2 * j) * HW + sp] + xi;
    float ky = a.kps[((long long)b * 2 * a.J + 2 * j + 1) * HW + sp] + yi;
    float xs = xi, ys = yi;
    if (a.reg) {
      xs += a.reg[((long long)b * 2) * HW + sp];
      ys += a.reg[((long long)b * 2 + 1) * HW + sp];
    } else {
      xs += 0.5f;
      ys += 0.5f;
    }
    const float hw_ = a.wh[((long long)b * 2) * HW + sp] * 0.5f;
    const float hh_ = a.wh[((long long)b * 2 + 1) * HW + sp] * 0.5f;
    const float l = xs - hw_, t = ys - hh_, r = xs + hw_, bt = ys + hh_;
    float *d = a.dets + o * D;
    if (j == 0) {
      d[0] = l; d[1] = t; d[2] = r; d[3] = bt;
      d[4] = a.c_scores[o];
      d[D - 1] = (float)a.c_clses[o];
    }
    if (has_hp) {
      float best = 0.0f, bx = 0.0f, by = 0.0f, bs = 0.0f;
      for (int m = 0; m < K; ++m) {  // :548-551: first minimum of the L2 distance
        const float dx = kx - cx[m], dy = ky - cy[m];
        const float dist = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        if (m == 0 || dist < best) {
          best = dist; bx = cx[m]; by = cy[m]; bs = cs[m];
        }
      }
      const float lim = fmaxf(bt - t, r - l) * 0.3f;  // :558-565
      const bool reject = (bx < l) || (bx > r) || (by < t) || (by > bt) || (bs < 0.1f) || (best > lim);
      if (!reject) {
        kx = bx;
        ky = by;
      }
    }
    d[5 + 2 * j] = kx;
    d[5 + 2 * j + 1] = ky;
  }
}

static size_t raw_bytes_(long long n) { return align_up((size_t)n * 8, 256) + 4 * align_up((size_t)n * 4, 256); }

}  // namespace cnb

using namespace cnb;

extern "C" {

size_t cnb_multi_pose_workspace_bytes(int b, int c, int j, int h, int w, int k) {
  SelectPlan p1, p2;
  if (make_select_plan(nullptr, b, c, h, w, k, 1, &p1) != CNB_OK) return 0;
  size_t sel = select_workspace_bytes(p1);
  if (j > 0) {
    if (make_select_plan(nullptr, b * j, 1, h, w, k, 1, &p2) != CNB_OK) return 0;
    const size_t s2 = select_workspace_bytes(p2);
    sel = s2 > sel ? s2 : sel;
  }
  return sel + raw_bytes_((long long)b * k) + raw_bytes_((long long)b * j * k);
}

int cnb_multi_pose_decode(const float *heat, const float *wh, const float *kps, const float *reg, const float *hm_hp,
                          const float *hp_offset, int b, int c, int j, int h, int w, int k, float *dets,
                          void *workspace, size_t workspace_bytes, void *stream_) {
  CNB_REQUIRE(heat && wh && kps && dets && workspace, CNB_EINVAL, "cnb_multi_pose_decode: null pointer");
  CNB_REQUIRE(j > 0, CNB_EINVAL, "cnb_multi_pose_decode: j=%d", j);
  CNB_REQUIRE(hm_hp || !hp_offset, CNB_EINVAL, "cnb_multi_pose_decode: hp_offset without hm_hp");
  cudaStream_t stream = (cudaStream_t)stream_;
  SelectPlan p1, p2;
  int rc = make_select_plan(heat, b, c, h, w, k, 1, &p1);
  if (rc != CNB_OK) return rc;
  size_t sel = select_workspace_bytes(p1);
  if (hm_hp) {
    rc = make_select_plan(hm_hp, b * j, 1, h, w, k, 1, &p2);
    if (rc != CNB_OK) return rc;
    const size_t s2 = select_workspace_bytes(p2);
    sel = s2 > sel ? s2 : sel;
  }
  const size_t need = sel + raw_bytes_((long long)b * k) + raw_bytes_((long long)b * j * k);
  CNB_REQUIRE(workspace_bytes >= need, CNB_EWORKSPACE, "cnb_multi_pose_decode: workspace %zu < %zu",
              workspace_bytes, need);
  char *p = reinterpret_cast<char *>(workspace) + sel;
  const long long n1 = (long long)b * k, n2 = (long long)b * j * k;
  int64_t *c_inds = reinterpret_cast<int64_t *>(p); p += align_up((size_t)n1 * 8, 256);
  float *c_scores = reinterpret_cast<float *>(p); p += align_up((size_t)n1 * 4, 256);
  int32_t *c_clses = reinterpret_cast<int32_t *>(p); p += align_up((size_t)n1 * 4, 256);
  float *c_ys = reinterpret_cast<float *>(p); p += align_up((size_t)n1 * 4, 256);
  float *c_xs = reinterpret_cast<float *>(p); p += align_up((size_t)n1 * 4, 256);
  int64_t *j_inds = reinterpret_cast<int64_t *>(p); p += align_up((size_t)n2 * 8, 256);
  float *j_scores = reinterpret_cast<float *>(p); p += align_up((size_t)n2 * 4, 256);
  p += align_up((size_t)n2 * 4, 256);  // (unused class slot keeps the raw layout uniform)
  float *j_ys = reinterpret_cast<float *>(p); p += align_up((size_t)n2 * 4, 256);
  float *j_xs = reinterpret_cast<float *>(p);

  FinalizeOut o1 = {c_scores, c_inds, c_clses, c_ys, c_xs, nullptr, nullptr, 0, nullptr};
  rc = run_select(heat, p1, o1, workspace, stream);
  if (rc != CNB_OK) return rc;
  if (hm_hp) {
    FinalizeOut o2 = {j_scores, j_inds, nullptr, j_ys, j_xs, nullptr, nullptr, 0, nullptr};
    rc = run_select(hm_hp, p2, o2, workspace, stream);
    if (rc != CNB_OK) return rc;
  }
  PoseArgs a;
  a.c_scores = c_scores; a.c_inds = c_inds; a.c_clses = c_clses; a.c_ys = c_ys; a.c_xs = c_xs;
  a.j_scores = hm_hp ? j_scores : nullptr; a.j_inds = j_inds; a.j_ys = j_ys; a.j_xs = j_xs;
  a.wh = wh; a.kps = kps; a.reg = reg; a.hp_offset = hp_offset;
  a.B = b; a.J = j; a.H = h; a.W = w; a.K = k; a.dets = dets;
  const int threads = k < 1024 ? ((k + 31) / 32) * 32 : 1024;
  k_pose_assemble<<<b * j, threads, (size_t)3 * k * sizeof(float), stream>>>(a);
  CNB_CHECK_LAUNCH("cnb_multi_pose_decode assemble");
  count_launch();
  return CNB_OK;
}

}  // extern "C"
