// C-ABI entry points of the heat-map decode family (models/decode.py, models/utils.py).
#include "select.cuh"

namespace cnb {

// ------------------------------------------------------------------ A1: standalone _nms
// One thread per 4 consecutive pixels; neighbours come through L1/L2 so DRAM sees one
// read and one write of the map.  models/decode.py:9-15.
template <bool VEC>
__global__ void __launch_bounds__(256) k_nms(const float *__restrict__ heat, float *__restrict__ out,
                                             long long planes, int H, int W) {
  const int Wq = VEC ? W / 4 : W;
  const long long total = planes * H * Wq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xq = (int)(i % Wq);
    const long long t = i / Wq;
    const int y = (int)(t % H);
    const float *pl = heat + (t / H) * (long long)H * W;
    const float NI = CNB_NEG_INF;
    if (VEC) {
      const int x0 = xq * 4;
      float v[6] = {NI, NI, NI, NI, NI, NI};  // vertical max of columns x0-1 .. x0+4
      float4 ctr = make_float4(0, 0, 0, 0);
      for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        const float *row = pl + (long long)yy * W;
        const float4 q = __ldg(reinterpret_cast<const float4 *>(row + x0));
        if (dy == 0) ctr = q;
        v[1] = fmaxf(v[1], q.x); v[2] = fmaxf(v[2], q.y); v[3] = fmaxf(v[3], q.z); v[4] = fmaxf(v[4], q.w);
        if (x0 > 0) v[0] = fmaxf(v[0], __ldg(row + x0 - 1));
        if (x0 + 4 < W) v[5] = fmaxf(v[5], __ldg(row + x0 + 4));
      }
      float4 o;
      o.x = ctr.x * ((fmax3(v[0], v[1], v[2]) == ctr.x) ? 1.0f : 0.0f);
      o.y = ctr.y * ((fmax3(v[1], v[2], v[3]) == ctr.y) ? 1.0f : 0.0f);
      o.z = ctr.z * ((fmax3(v[2], v[3], v[4]) == ctr.z) ? 1.0f : 0.0f);
      o.w = ctr.w * ((fmax3(v[3], v[4], v[5]) == ctr.w) ? 1.0f : 0.0f);
      *reinterpret_cast<float4 *>(out + (t * W) + x0) = o;
    } else {
      const int x = xq;
      const float c = pl[(long long)y * W + x];
      float m = c;
      for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          const int xx = x + dx;
          if (xx < 0 || xx >= W) continue;
          m = fmaxf(m, __ldg(pl + (long long)yy * W + xx));
        }
      }
      out[t * W + x] = c * ((m == c) ? 1.0f : 0.0f);
    }
  }
}

// ------------------------------------------------------------------ A4: gather
// out[b, m, ch] = feat[b, ch, ind[b, m]]   (models/utils.py:12-26 without the transpose copy)
__global__ void __launch_bounds__(256) k_gather_feat(const float *__restrict__ feat,
                                                     const int64_t *__restrict__ ind, float *__restrict__ out,
                                                     int B, int C, long long HW, int M) {
  const long long total = (long long)B * M * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % C);
    const long long bm = i / C;
    const int b = (int)(bm / M);
    const long long sp = ind[bm];
    float v = 0.0f;
    if (sp >= 0 && sp < HW) v = __ldg(feat + ((long long)b * C + ch) * HW + sp);
    out[i] = v;
  }
}

// grad_feat[b, ch, ind[b, m]] += grad_out[b, m, ch]: the adjoint of the gather (autograd of `feat.gather(1, ind)`,
// models/utils.py:15).  An index may repeat inside an image, hence the atomic add; grad_feat arrives zero-filled.
__global__ void __launch_bounds__(256) k_gather_feat_bwd(const float *__restrict__ grad_out,
                                                         const int64_t *__restrict__ ind,
                                                         float *__restrict__ grad_feat, int B, int C, long long HW,
                                                         int M) {
  const long long total = (long long)B * M * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % C);
    const long long bm = i / C;
    const int b = (int)(bm / M);
    const long long sp = ind[bm];
    if (sp >= 0 && sp < HW) atomicAdd(grad_feat + ((long long)b * C + ch) * HW + sp, grad_out[i]);
  }
}

// ------------------------------------------------------------------ A9: ddd epilogue
// models/decode.py:432-460: [xs, ys, score, rot(8), depth, dim(3), (wh(2)), cls]
__global__ void __launch_bounds__(128) k_ddd_epilogue(const float *__restrict__ scores,
                                                      const int64_t *__restrict__ inds,
                                                      const int32_t *__restrict__ clses,
                                                      const float *__restrict__ ys, const float *__restrict__ xs,
                                                      const float *__restrict__ rot, const float *__restrict__ depth,
                                                      const float *__restrict__ dim, const float *__restrict__ wh,
                                                      const float *__restrict__ reg, int B, int K, long long HW,
                                                      float *__restrict__ dets) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * K) return;
  const int b = i / K;
  const long long sp = inds[i];
  const int D = wh ? 18 : 16;
  float *d = dets + (size_t)i * D;
  float x = xs[i], y = ys[i];
  if (reg) {
    x += reg[((long long)b * 2) * HW + sp];
    y += reg[((long long)b * 2 + 1) * HW + sp];
  } else {
    x += 0.5f;
    y += 0.5f;
  }
  d[0] = x; d[1] = y; d[2] = scores[i];
  for (int r = 0; r < 8; ++r) d[3 + r] = rot[((long long)b * 8 + r) * HW + sp];
  d[11] = depth[(long long)b * HW + sp];
  for (int r = 0; r < 3; ++r) d[12 + r] = dim[((long long)b * 3 + r) * HW + sp];
  int o = 15;
  if (wh) {
    d[15] = wh[((long long)b * 2) * HW + sp];
    d[16] = wh[((long long)b * 2 + 1) * HW + sp];
    o = 17;
  }
  d[o] = (float)clses[i];
}

struct RawTopk {
  float *scores;
  int64_t *inds;
  int32_t *clses;
  float *ys;
  float *xs;
};

static size_t raw_bytes(long long n) {  // n = n_img * K entries
  return align_up((size_t)n * 8, 256) + 4 * align_up((size_t)n * 4, 256);
}
static RawTopk carve_raw(char *p, long long n) {
  RawTopk r;
  r.inds = reinterpret_cast<int64_t *>(p); p += align_up((size_t)n * 8, 256);
  r.scores = reinterpret_cast<float *>(p); p += align_up((size_t)n * 4, 256);
  r.clses = reinterpret_cast<int32_t *>(p); p += align_up((size_t)n * 4, 256);
  r.ys = reinterpret_cast<float *>(p); p += align_up((size_t)n * 4, 256);
  r.xs = reinterpret_cast<float *>(p);
  return r;
}

}  // namespace cnb

using namespace cnb;

// out = 1 / (1 + exp(-x)), the expression of torch's CUDA sigmoid (4 elements per thread)
__global__ void __launch_bounds__(256) k_sigmoid(const float *__restrict__ x, float *__restrict__ out, long long n) {
  const long long i0 = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (i0 + j < n) out[i0 + j] = 1.0f / (1.0f + expf(-x[i0 + j]));
}

// Flip-test merge (detectors/ctdet.py:34-37, detectors/multi_pose.py:43-51): src [2n, c, h, w] holds the heads
// of n images followed by those of their horizontally flipped copies; dst [n, c, h, w] =
//   ( f(src[i, ch]) + sign[ch] * f(flip_w(src[n + i, perm[ch]])) ) / 2,     f = sigmoid (mode 1) or identity.
// perm / sign (nullable) express flip_lr (joint swap, models/utils.py:33-39) and flip_lr_off (joint swap + negated
// x offsets, :41-50); null = flip_tensor (:28-29).  One pass instead of the reference's sigmoid_, slices, host
// round trip through numpy, add and divide.
__global__ void __launch_bounds__(256) k_flip_merge(const float *__restrict__ src, float *__restrict__ dst, int n, int c,
                                                    int h, int w, int mode, const int *__restrict__ perm,
                                                    const float *__restrict__ sign) {
  const long long total = (long long)n * c * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    long long r = i / w;
    const int y = (int)(r % h);
    r /= h;
    const int ch = (int)(r % c), img = (int)(r / c);
    const int pc = perm ? perm[ch] : ch;
    float a = src[i];
    float b = src[(((long long)(n + img) * c + pc) * h + y) * w + (w - 1 - x)];
    if (mode == 1) {
      a = 1.0f / (1.0f + expf(-a));
      b = 1.0f / (1.0f + expf(-b));
    }
    if (sign) b = __fmul_rn(b, sign[ch]);
    dst[i] = __fadd_rn(a, b) / 2.0f;
  }
}

extern "C" {

int cnb_flip_merge(const float *src, float *dst, int n, int c, int h, int w, int apply_sigmoid, const int32_t *perm,
                   const float *sign, void *stream) {
  CNB_REQUIRE(src && dst, CNB_EINVAL, "cnb_flip_merge: null pointer");
  CNB_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, CNB_EINVAL, "cnb_flip_merge: non-positive dimension");
  const long long total = (long long)n * c * h * w;
  const int grid = (int)((total + 255) / 256 < 4736 ? (total + 255) / 256 : 4736);
  k_flip_merge<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, n, c, h, w, apply_sigmoid ? 1 : 0, perm, sign);
  CNB_CHECK_LAUNCH("cnb_flip_merge");
  count_launch();
  return CNB_OK;
}

int cnb_nms(const float *heat, float *out, int n, int c, int h, int w, void *stream) {
  CNB_REQUIRE(heat && out, CNB_EINVAL, "cnb_nms: null pointer");
  CNB_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, CNB_EINVAL, "cnb_nms: non-positive dimension");
  const long long planes = (long long)n * c;
  const bool vec = (w % 4 == 0) && (((uintptr_t)heat | (uintptr_t)out) & 15u) == 0;
  const long long work = planes * h * (vec ? w / 4 : w);
  const int grid = (int)((work + 255) / 256 < (long long)num_sms() * 16 ? (work + 255) / 256 : (long long)num_sms() * 16);
  if (vec)
    k_nms<true><<<grid, 256, 0, (cudaStream_t)stream>>>(heat, out, planes, h, w);
  else
    k_nms<false><<<grid, 256, 0, (cudaStream_t)stream>>>(heat, out, planes, h, w);
  CNB_CHECK_LAUNCH("cnb_nms");
  count_launch();
  return CNB_OK;
}

int cnb_gather_feat_backward(const float *grad_out, const int64_t *ind, float *grad_feat, int b, int c, int hw, int m,
                             void *stream) {
  CNB_REQUIRE(grad_out && ind && grad_feat, CNB_EINVAL, "cnb_gather_feat_backward: null pointer");
  CNB_REQUIRE(b > 0 && c > 0 && hw > 0 && m > 0, CNB_EINVAL, "cnb_gather_feat_backward: non-positive dimension");
  const long long total = (long long)b * m * c;
  const int grid = (int)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535);
  k_gather_feat_bwd<<<grid, 256, 0, (cudaStream_t)stream>>>(grad_out, ind, grad_feat, b, c, hw, m);
  CNB_CHECK_LAUNCH("cnb_gather_feat_backward");
  count_launch();
  return CNB_OK;
}

int cnb_gather_feat(const float *feat, const int64_t *ind, float *out, int b, int c, int hw, int m, void *stream) {
  CNB_REQUIRE(feat && ind && out, CNB_EINVAL, "cnb_gather_feat: null pointer");
  CNB_REQUIRE(b > 0 && c > 0 && hw > 0 && m > 0, CNB_EINVAL, "cnb_gather_feat: non-positive dimension");
  const long long total = (long long)b * m * c;
  const int grid = (int)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535);
  k_gather_feat<<<grid, 256, 0, (cudaStream_t)stream>>>(feat, ind, out, b, c, hw, m);
  CNB_CHECK_LAUNCH("cnb_gather_feat");
  count_launch();
  return CNB_OK;
}

size_t cnb_topk_workspace_bytes(int n_img, int c, int h, int w, int k) {
  SelectPlan pl;
  // worst case over alignment: the non-TMA plan has the same grid / slot count
  if (make_select_plan(nullptr, n_img, c, h, w, k, 1, &pl) != CNB_OK) return 0;
  return select_workspace_bytes(pl) + raw_bytes((long long)n_img * k);
}

int cnb_topk(const float *scores, int b, int c, int h, int w, int k, int fuse_nms, float *out_scores,
             int64_t *out_inds, int32_t *out_clses, float *out_ys, float *out_xs, void *workspace,
             size_t workspace_bytes, void *stream) {
  CNB_REQUIRE(scores && workspace, CNB_EINVAL, "cnb_topk: null pointer");
  SelectPlan pl;
  int rc = make_select_plan(scores, b, c, h, w, k, fuse_nms, &pl);
  if (rc != CNB_OK) return rc;
  CNB_REQUIRE(workspace_bytes >= select_workspace_bytes(pl), CNB_EWORKSPACE,
              "cnb_topk: workspace %zu < %zu", workspace_bytes, select_workspace_bytes(pl));
  FinalizeOut out = {out_scores, out_inds, out_clses, out_ys, out_xs, nullptr, nullptr, 0, nullptr};
  return run_select(scores, pl, out, workspace, (cudaStream_t)stream);
}

int cnb_topk_channel(const float *scores, int b, int c, int h, int w, int k, int fuse_nms, float *out_scores,
                     int64_t *out_inds, float *out_ys, float *out_xs, void *workspace, size_t workspace_bytes,
                     void *stream) {
  CNB_REQUIRE(scores && workspace, CNB_EINVAL, "cnb_topk_channel: null pointer");
  CNB_REQUIRE(b > 0 && c > 0, CNB_EINVAL, "cnb_topk_channel: non-positive dimension");
  SelectPlan pl;  // every (b, channel) plane is its own "image" with one class
  int rc = make_select_plan(scores, b * c, 1, h, w, k, fuse_nms, &pl);
  if (rc != CNB_OK) return rc;
  CNB_REQUIRE(workspace_bytes >= select_workspace_bytes(pl), CNB_EWORKSPACE,
              "cnb_topk_channel: workspace %zu < %zu", workspace_bytes, select_workspace_bytes(pl));
  FinalizeOut out = {out_scores, out_inds, nullptr, out_ys, out_xs, nullptr, nullptr, 0, nullptr};
  return run_select(scores, pl, out, workspace, (cudaStream_t)stream);
}

int cnb_ctdet_decode(const float *heat, const float *wh, const float *reg, int cat_spec_wh, int b, int c, int h,
                     int w, int k, float *dets, void *workspace, size_t workspace_bytes, void *stream) {
  CNB_REQUIRE(heat && wh && dets && workspace, CNB_EINVAL, "cnb_ctdet_decode: null pointer");
  SelectPlan pl;
  int rc = make_select_plan(heat, b, c, h, w, k, 1, &pl);
  if (rc != CNB_OK) return rc;
  CNB_REQUIRE(workspace_bytes >= select_workspace_bytes(pl), CNB_EWORKSPACE,
              "cnb_ctdet_decode: workspace %zu < %zu", workspace_bytes, select_workspace_bytes(pl));
  FinalizeOut out = {nullptr, nullptr, nullptr, nullptr, nullptr, wh, reg, cat_spec_wh, dets};
  return run_select(heat, pl, out, workspace, (cudaStream_t)stream);
}

/* N1: ctdet_decode on the head's raw logits, sigmoid fused (detectors/ctdet.py:31 +
 * models/decode.py:464-495).  Hot geometry: the selection kernel works in logit space and applies
 * the sigmoid only to the few candidates; other geometries materialise the heat map once in the
 * workspace and take the normal path. */
size_t cnb_ctdet_logits_workspace_bytes(const float *hm_logits, int b, int c, int h, int w, int k) {
  SelectPlan pl;
  if (make_select_plan(hm_logits, b, c, h, w, k, 1, &pl) != CNB_OK) return 0;
  size_t n = select_workspace_bytes(pl);
  if (!pl.hot) n = align_up(n, 256) + align_up((size_t)b * c * h * w * 4, 256);
  return n;
}

int cnb_ctdet_decode_logits(const float *hm_logits, const float *wh, const float *reg, int cat_spec_wh, int b, int c,
                            int h, int w, int k, float *dets, void *workspace, size_t workspace_bytes, void *stream) {
  CNB_REQUIRE(hm_logits && wh && dets && workspace, CNB_EINVAL, "cnb_ctdet_decode_logits: null pointer");
  SelectPlan pl;
  int rc = make_select_plan(hm_logits, b, c, h, w, k, 1, &pl);
  if (rc != CNB_OK) return rc;
  const size_t need = cnb_ctdet_logits_workspace_bytes(hm_logits, b, c, h, w, k);
  CNB_REQUIRE(workspace_bytes >= need, CNB_EWORKSPACE, "cnb_ctdet_decode_logits: workspace %zu < %zu", workspace_bytes,
              need);
  FinalizeOut out = {nullptr, nullptr, nullptr, nullptr, nullptr, wh, reg, cat_spec_wh, dets};
  if (pl.hot) {
    pl.logits = 1;
    return run_select(hm_logits, pl, out, workspace, (cudaStream_t)stream);
  }
  float *heat = reinterpret_cast<float *>(reinterpret_cast<char *>(workspace) + align_up(select_workspace_bytes(pl), 256));
  const long long n = (long long)b * c * h * w;
  k_sigmoid<<<(unsigned)((n + 1023) / 1024), 256, 0, (cudaStream_t)stream>>>(hm_logits, heat, n);
  CNB_CHECK_LAUNCH("cnb_ctdet_decode_logits sigmoid");
  count_launch();
  rc = make_select_plan(heat, b, c, h, w, k, 1, &pl);
  if (rc != CNB_OK) return rc;
  return run_select(heat, pl, out, workspace, (cudaStream_t)stream);
}

int cnb_sigmoid(const float *x, float *out, long long n, void *stream) {
  CNB_REQUIRE(x && out && n >= 0, CNB_EINVAL, "cnb_sigmoid: bad argument");
  if (n == 0) return CNB_OK;
  k_sigmoid<<<(unsigned)((n + 1023) / 1024), 256, 0, (cudaStream_t)stream>>>(x, out, n);
  CNB_CHECK_LAUNCH("cnb_sigmoid");
  count_launch();
  return CNB_OK;
}

int cnb_ddd_decode(const float *heat, const float *rot, const float *depth, const float *dim, const float *wh,
                   const float *reg, int b, int c, int h, int w, int k, float *dets, void *workspace,
                   size_t workspace_bytes, void *stream) {
  CNB_REQUIRE(heat && rot && depth && dim && dets && workspace, CNB_EINVAL, "cnb_ddd_decode: null pointer");
  SelectPlan pl;
  int rc = make_select_plan(heat, b, c, h, w, k, 1, &pl);
  if (rc != CNB_OK) return rc;
  const size_t sel = select_workspace_bytes(pl);
  CNB_REQUIRE(workspace_bytes >= sel + raw_bytes((long long)b * k), CNB_EWORKSPACE,
              "cnb_ddd_decode: workspace %zu < %zu", workspace_bytes, sel + raw_bytes((long long)b * k));
  RawTopk r = carve_raw(reinterpret_cast<char *>(workspace) + sel, (long long)b * k);
  FinalizeOut out = {r.scores, r.inds, r.clses, r.ys, r.xs, nullptr, nullptr, 0, nullptr};
  rc = run_select(heat, pl, out, workspace, (cudaStream_t)stream);
  if (rc != CNB_OK) return rc;
  k_ddd_epilogue<<<(b * k + 127) / 128, 128, 0, (cudaStream_t)stream>>>(r.scores, r.inds, r.clses, r.ys, r.xs, rot,
                                                                         depth, dim, wh, reg, b, k,
                                                                         (long long)h * w, dets);
  CNB_CHECK_LAUNCH("cnb_ddd_decode epilogue");
  count_launch();
  return CNB_OK;
}

}  // extern "C"
