// Deformable PSROI pooling, forward and backward (DCNv2 package surface: DCNv2Pooling / DCNPooling,
// SURVEY.md section 8f N4).  Semantics of
//   src/lib/models/networks/DCNv2/src/cuda/dcn_v2_psroi_pooling_cuda.cu:47-139 (forward), :141-255 (backward):
// every output bin (roi n, channel ctop, ph, pw) averages sample_per_part^2 bilinear samples of one
// position-sensitive input channel, the bin window optionally shifted by a learned (class, part) offset.
// No CenterNet network instantiates it, so this is a plain one-thread-per-bin kernel (the op is tiny:
// N * output_dim * P * P bins x <= 16 samples); the arithmetic follows the reference expression by
// expression, including its float/double mixing in the window geometry, so bins land on the same pixels.
#include "common.cuh"

namespace cnb {

struct PsroiArgs {
  const float *data, *rois, *trans;   // trans == nullptr when no_trans
  int B, C, H, W, N;
  int no_trans, output_dim, group_size, P, part_size, spp, num_classes, ch_each;
  float spatial_scale, trans_std;
};

struct PsroiBin {
  int n, ctop, ph, pw, batch, c, cls, part_h, part_w;
  float wstart, hstart, sub_w, sub_h, roi_w, roi_h;
};

// window geometry of one bin (:66-108): float arithmetic with the double literals of the source
__device__ __forceinline__ PsroiBin psroi_bin(const PsroiArgs &a, long long index) {
  PsroiBin b;
  const int P = a.P;
  b.pw = (int)(index % P);
  b.ph = (int)((index / P) % P);
  b.ctop = (int)((index / P / P) % a.output_dim);
  b.n = (int)(index / P / P / a.output_dim);
  const float *r = a.rois + (long long)b.n * 5;
  b.batch = (int)r[0];
  const float start_w = (float)((double)((float)round((double)r[1]) * a.spatial_scale) - 0.5);
  const float start_h = (float)((double)((float)round((double)r[2]) * a.spatial_scale) - 0.5);
  const float end_w = (float)((double)((float)(round((double)r[3]) + 1.) * a.spatial_scale) - 0.5);
  const float end_h = (float)((double)((float)(round((double)r[4]) + 1.) * a.spatial_scale) - 0.5);
  b.roi_w = (float)fmax((double)(end_w - start_w), 0.1);   // avoid 0
  b.roi_h = (float)fmax((double)(end_h - start_h), 0.1);
  const float bin_h = b.roi_h / (float)P, bin_w = b.roi_w / (float)P;
  b.sub_h = bin_h / (float)a.spp;
  b.sub_w = bin_w / (float)a.spp;
  b.part_h = (int)floorf((float)b.ph / (float)P * (float)a.part_size);
  b.part_w = (int)floorf((float)b.pw / (float)P * (float)a.part_size);
  b.cls = b.ctop / a.ch_each;
  float tx = 0.f, ty = 0.f;
  if (!a.no_trans) {
    const long long t0 = (((long long)b.n * a.num_classes + b.cls) * 2) * a.part_size;
    tx = a.trans[(t0 + b.part_h) * a.part_size + b.part_w] * a.trans_std;
    ty = a.trans[(t0 + a.part_size + b.part_h) * a.part_size + b.part_w] * a.trans_std;
  }
  b.wstart = __fadd_rn(__fadd_rn(__fmul_rn((float)b.pw, bin_w), start_w), __fmul_rn(tx, b.roi_w));
  b.hstart = __fadd_rn(__fadd_rn(__fmul_rn((float)b.ph, bin_h), start_h), __fmul_rn(ty, b.roi_h));
  int gw = (int)floorf((float)b.pw * (float)a.group_size / (float)P);
  int gh = (int)floorf((float)b.ph * (float)a.group_size / (float)P);
  gw = min(max(gw, 0), a.group_size - 1);
  gh = min(max(gh, 0), a.group_size - 1);
  b.c = (b.ctop * a.group_size + gh) * a.group_size + gw;
  return b;
}

// sample position (iw, ih) of a bin; false when it falls outside the map (:113-121)
__device__ __forceinline__ bool psroi_sample(const PsroiArgs &a, const PsroiBin &b, int iw, int ih, float *w, float *h) {
  float ww = __fadd_rn(b.wstart, __fmul_rn((float)iw, b.sub_w));
  float hh = __fadd_rn(b.hstart, __fmul_rn((float)ih, b.sub_h));
  if ((double)ww < -0.5 || (double)ww > (double)a.W - 0.5 || (double)hh < -0.5 || (double)hh > (double)a.H - 0.5) return false;
  *w = (float)fmin(fmax((double)ww, 0.), (double)a.W - 1.);
  *h = (float)fmin(fmax((double)hh, 0.), (double)a.H - 1.);
  return true;
}

__global__ void __launch_bounds__(256) k_psroi_forward(const PsroiArgs a, long long count, float *__restrict__ out,
                                                       float *__restrict__ top_count) {
  const long long index = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (index >= count) return;
  const PsroiBin b = psroi_bin(a, index);
  if (b.batch < 0 || b.batch >= a.B) {   // roi names an image that does not exist: empty bin instead of a wild read
    out[index] = 0.f;
    top_count[index] = 0.f;
    return;
  }
  const float *pl = a.data + ((long long)b.batch * a.C + b.c) * a.H * a.W;
  float sum = 0.f;
  int cnt = 0;
  for (int ih = 0; ih < a.spp; ++ih)
    for (int iw = 0; iw < a.spp; ++iw) {
      float w, h;
      if (!psroi_sample(a, b, iw, ih, &w, &h)) continue;
      const int x1 = (int)floorf(w), x2 = (int)ceilf(w), y1 = (int)floorf(h), y2 = (int)ceilf(h);   // :24-45
      const float dx = w - (float)x1, dy = h - (float)y1;
      const float v11 = pl[y1 * a.W + x1], v12 = pl[y2 * a.W + x1], v21 = pl[y1 * a.W + x2], v22 = pl[y2 * a.W + x2];
      const float val = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(1.f - dx, 1.f - dy), v11),
                                                       __fmul_rn(__fmul_rn(1.f - dx, dy), v12)),
                                            __fmul_rn(__fmul_rn(dx, 1.f - dy), v21)),
                                  __fmul_rn(__fmul_rn(dx, dy), v22));
      sum = __fadd_rn(sum, val);
      ++cnt;
    }
  out[index] = cnt == 0 ? 0.f : sum / (float)cnt;
  top_count[index] = (float)cnt;
}

__global__ void __launch_bounds__(256) k_psroi_backward(const PsroiArgs a, long long count,
                                                        const float *__restrict__ top_diff,
                                                        const float *__restrict__ top_count, float *grad_data,
                                                        float *grad_trans) {
  const long long index = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (index >= count) return;
  if (top_count[index] <= 0.f) return;
  const PsroiBin b = psroi_bin(a, index);
  if (b.batch < 0 || b.batch >= a.B) return;
  const float dv = top_diff[index] / top_count[index];
  const long long plane = ((long long)b.batch * a.C + b.c) * a.H * a.W;
  const float *pl = a.data + plane;
  float *gp = grad_data + plane;
  float gx = 0.f, gy = 0.f;
  for (int ih = 0; ih < a.spp; ++ih)
    for (int iw = 0; iw < a.spp; ++iw) {
      float w, h;
      if (!psroi_sample(a, b, iw, ih, &w, &h)) continue;
      const int x0 = (int)floorf(w), x1 = (int)ceilf(w), y0 = (int)floorf(h), y1 = (int)ceilf(h);
      const float dx = w - (float)x0, dy = h - (float)y0;
      atomicAdd(gp + y0 * a.W + x0, (1.f - dx) * (1.f - dy) * dv);   // :223-231
      atomicAdd(gp + y1 * a.W + x0, (1.f - dx) * dy * dv);
      atomicAdd(gp + y0 * a.W + x1, dx * (1.f - dy) * dv);
      atomicAdd(gp + y1 * a.W + x1, dx * dy * dv);
      if (a.no_trans) continue;
      const float u00 = pl[y0 * a.W + x0], u01 = pl[y1 * a.W + x0], u10 = pl[y0 * a.W + x1], u11 = pl[y1 * a.W + x1];
      gx += (u11 * dy + u10 * (1.f - dy) - u01 * dy - u00 * (1.f - dy)) * a.trans_std * dv * b.roi_w;   // :241-246
      gy += (u11 * dx + u01 * (1.f - dx) - u10 * dx - u00 * (1.f - dx)) * a.trans_std * dv * b.roi_h;
    }
  if (!a.no_trans) {  // one atomic per bin and direction (the reference issues one per sample)
    const long long t0 = (((long long)b.n * a.num_classes + b.cls) * 2) * a.part_size;
    atomicAdd(grad_trans + (t0 + b.part_h) * a.part_size + b.part_w, gx);
    atomicAdd(grad_trans + (t0 + a.part_size + b.part_h) * a.part_size + b.part_w, gy);
  }
}

static int fill_args(const char *who, PsroiArgs *a, const float *data, const float *rois, const float *trans, int b,
                     int c, int h, int w, int num_rois, int channels_trans, int no_trans, float spatial_scale,
                     int output_dim, int group_size, int pooled_size, int part_size, int sample_per_part,
                     float trans_std) {
  CNB_REQUIRE(data && rois, CNB_EINVAL, "%s: null pointer", who);
  CNB_REQUIRE(b > 0 && c > 0 && h > 0 && w > 0 && num_rois >= 0, CNB_EINVAL, "%s: bad shape", who);
  CNB_REQUIRE(output_dim > 0 && group_size > 0 && pooled_size > 0 && part_size > 0 && sample_per_part > 0, CNB_EINVAL,
              "%s: non-positive pooling parameter", who);
  CNB_REQUIRE(c >= output_dim * group_size * group_size, CNB_EINVAL,
              "%s: input has %d channels, output_dim * group_size^2 = %d needed", who, c,
              output_dim * group_size * group_size);
  CNB_REQUIRE(no_trans || (trans && channels_trans >= 2 && channels_trans % 2 == 0), CNB_EINVAL,
              "%s: offsets need an even, positive channel count (got %d)", who, channels_trans);
  a->data = data; a->rois = rois; a->trans = no_trans ? nullptr : trans;
  a->B = b; a->C = c; a->H = h; a->W = w; a->N = num_rois;
  a->no_trans = no_trans; a->output_dim = output_dim; a->group_size = group_size; a->P = pooled_size;
  a->part_size = part_size; a->spp = sample_per_part; a->spatial_scale = spatial_scale; a->trans_std = trans_std;
  a->num_classes = no_trans ? 1 : channels_trans / 2;                     // :282-283
  a->ch_each = no_trans ? output_dim : output_dim / a->num_classes;
  CNB_REQUIRE(a->ch_each > 0, CNB_EINVAL, "%s: output_dim %d < number of offset classes %d", who, output_dim,
              a->num_classes);
  return CNB_OK;
}

}  // namespace cnb

using namespace cnb;

extern "C" {

int cnb_psroi_pooling_forward(const float *data, const float *rois, const float *trans, float *out, float *top_count,
                              int b, int c, int h, int w, int num_rois, int channels_trans, int no_trans,
                              float spatial_scale, int output_dim, int group_size, int pooled_size, int part_size,
                              int sample_per_part, float trans_std, void *stream) {
  PsroiArgs a;
  int rc = fill_args("cnb_psroi_pooling_forward", &a, data, rois, trans, b, c, h, w, num_rois, channels_trans, no_trans,
                     spatial_scale, output_dim, group_size, pooled_size, part_size, sample_per_part, trans_std);
  if (rc != CNB_OK) return rc;
  CNB_REQUIRE(out && top_count, CNB_EINVAL, "cnb_psroi_pooling_forward: null output");
  const long long count = (long long)num_rois * output_dim * pooled_size * pooled_size;
  if (count == 0) return CNB_OK;
  k_psroi_forward<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, count, out, top_count);
  CNB_CHECK_LAUNCH("cnb_psroi_pooling_forward");
  count_launch();
  return CNB_OK;
}

int cnb_psroi_pooling_backward(const float *grad_out, const float *data, const float *rois, const float *trans,
                               const float *top_count, float *grad_data, float *grad_trans, int b, int c, int h, int w,
                               int num_rois, int channels_trans, int no_trans, float spatial_scale, int output_dim,
                               int group_size, int pooled_size, int part_size, int sample_per_part, float trans_std,
                               void *stream) {
  PsroiArgs a;
  int rc = fill_args("cnb_psroi_pooling_backward", &a, data, rois, trans, b, c, h, w, num_rois, channels_trans,
                     no_trans, spatial_scale, output_dim, group_size, pooled_size, part_size, sample_per_part, trans_std);
  if (rc != CNB_OK) return rc;
  CNB_REQUIRE(grad_out && top_count && grad_data && (no_trans || grad_trans), CNB_EINVAL,
              "cnb_psroi_pooling_backward: null pointer");
  const long long count = (long long)num_rois * output_dim * pooled_size * pooled_size;
  if (count == 0) return CNB_OK;
  k_psroi_backward<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, count, grad_out, top_count,
                                                                                      grad_data, grad_trans);
  CNB_CHECK_LAUNCH("cnb_psroi_pooling_backward");
  count_launch();
  return CNB_OK;
}

}  // extern "C"
