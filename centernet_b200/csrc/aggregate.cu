// ExtremeNet edge aggregation (what BASELINE.json calls "corner_pool t/l/b/r"):
// models/decode.py:17-77.  Along a row (left/right) or column (top/bottom)
//     ret[i] = heat[i] + ret[i-1] * [heat[i] >= heat[i-1]],   dir = ret - heat
//     _h_aggregate = w*left + w*right + heat,  _v_aggregate = w*top + w*bottom + heat.
// The fp32 additions are strictly sequential in the reference, so each row/column is
// scanned by ONE thread in that order (a tree/warp scan would re-associate the sums and
// break bit parity); parallelism comes from the B*C*H rows (or B*C*W columns).
//   rows   : one thread per row with 16-byte accesses (its 128-byte lines stay in L1 while it consumes
//            them); the reverse pass parks its result in `out`, the forward pass combines.
//   columns: thread-per-column is coalesced as is; `out` doubles as the scratch of the
//            bottom-up pass (it stays in L2 between the two passes).
#include "common.cuh"

namespace cnb {

constexpr int AG_R = 64;    // rows per CTA (one thread each)
constexpr int AG_CW = 128;  // columns per tile

// mode: 0 both directions weighted (+heat), 1 forward only (left/top), 2 reverse only (right/bottom)
__device__ __forceinline__ float combine(int mode, float w, float fwd, float rev, float h) {
  if (mode == 1) return fwd;
  if (mode == 2) return rev;
  return __fadd_rn(__fadd_rn(__fmul_rn(w, fwd), __fmul_rn(w, rev)), h);  // decode.py:72-73 order
}

// Rows: one thread walks one row with 16-byte accesses.  Neighbouring threads touch different 128-byte
// lines, but every thread consumes each of its lines completely over 8 consecutive steps, so the lines
// live in L1 for that long and DRAM still sees each byte once; 2048 threads per SM hide the latency.
template <bool VEC>
__global__ void __launch_bounds__(256) k_aggr_rows(const float *__restrict__ heat, float *__restrict__ out,
                                                   long long nrows, int W, float w, int mode) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  const float *hp = heat + row * W;
  float *op = out + row * W;
  float ret = 0.0f, hn = 0.0f;
  if (mode != 1) {  // right aggregate (decode.py:30-41): from the right edge inwards, parked in `out`
    if (VEC) {
      for (int c = W - 4; c >= 0; c -= 4) {
        const float4 h4 = __ldg(reinterpret_cast<const float4 *>(hp + c));
        float4 r4;
        float h;
        h = h4.w; ret = (c + 3 == W - 1) ? h : ((h >= hn) ? __fadd_rn(h, ret) : h); hn = h; r4.w = __fsub_rn(ret, h);
        h = h4.z; ret = (h >= hn) ? __fadd_rn(h, ret) : h; hn = h; r4.z = __fsub_rn(ret, h);
        h = h4.y; ret = (h >= hn) ? __fadd_rn(h, ret) : h; hn = h; r4.y = __fsub_rn(ret, h);
        h = h4.x; ret = (h >= hn) ? __fadd_rn(h, ret) : h; hn = h; r4.x = __fsub_rn(ret, h);
        *reinterpret_cast<float4 *>(op + c) = r4;
      }
    } else {
      for (int c = W - 1; c >= 0; --c) {
        const float h = __ldg(hp + c);
        ret = (c == W - 1) ? h : ((h >= hn) ? __fadd_rn(h, ret) : h);
        hn = h;
        op[c] = __fsub_rn(ret, h);
      }
    }
    if (mode == 2) return;
  }
  // left aggregate (decode.py:17-28) + combine
  if (VEC) {
    for (int c = 0; c < W; c += 4) {
      const float4 h4 = __ldg(reinterpret_cast<const float4 *>(hp + c));
      float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mode == 0) rv = *reinterpret_cast<const float4 *>(op + c);
      float4 o4;
      float h;
      h = h4.x; ret = (c == 0) ? h : ((h >= hn) ? __fadd_rn(h, ret) : h); hn = h; o4.x = combine(mode, w, __fsub_rn(ret, h), rv.x, h);
      h = h4.y; ret = (h >= hn) ? __fadd_rn(h, ret) : h; hn = h; o4.y = combine(mode, w, __fsub_rn(ret, h), rv.y, h);
      h = h4.z; ret = (h >= hn) ? __fadd_rn(h, ret) : h; hn = h; o4.z = combine(mode, w, __fsub_rn(ret, h), rv.z, h);
      h = h4.w; ret = (h >= hn) ? __fadd_rn(h, ret) : h; hn = h; o4.w = combine(mode, w, __fsub_rn(ret, h), rv.w, h);
      *reinterpret_cast<float4 *>(op + c) = o4;
    }
  } else {
    for (int c = 0; c < W; ++c) {
      const float h = __ldg(hp + c);
      ret = (c == 0) ? h : ((h >= hn) ? __fadd_rn(h, ret) : h);
      hn = h;
      const float rv = (mode == 0) ? op[c] : 0.0f;
      op[c] = combine(mode, w, __fsub_rn(ret, h), rv, h);
    }
  }
}

__global__ void __launch_bounds__(128) k_aggr_cols(const float *__restrict__ heat, float *__restrict__ out,
                                                   long long planes, int H, int W, float w, int mode) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int xchunks = (W + 127) / 128;
  const long long plane = blockIdx.x / xchunks;
  const int x = (int)(blockIdx.x - plane * xchunks) * 128 + threadIdx.x;
  (void)gid;
  if (plane >= planes || x >= W) return;
  const float *hp = heat + plane * (long long)H * W + x;
  float *op = out + plane * (long long)H * W + x;
  float ret = 0.0f, hn = 0.0f;
  if (mode != 1) {  // bottom aggregate (decode.py:57-69)
#pragma unroll 4
    for (int y = H - 1; y >= 0; --y) {
      const float h = __ldg(hp + (long long)y * W);
      ret = (y == H - 1) ? h : ((h >= hn) ? __fadd_rn(h, ret) : h);
      hn = h;
      op[(long long)y * W] = __fsub_rn(ret, h);
    }
    if (mode == 2) return;
  }
#pragma unroll 4
  for (int y = 0; y < H; ++y) {  // top aggregate (decode.py:43-55) + combine
    const float h = __ldg(hp + (long long)y * W);
    ret = (y == 0) ? h : ((h >= hn) ? __fadd_rn(h, ret) : h);
    hn = h;
    const float rev = (mode == 0) ? op[(long long)y * W] : 0.0f;
    op[(long long)y * W] = combine(mode, w, __fsub_rn(ret, h), rev, h);
  }
}

// horizontal: 1 = _h_aggregate, 0 = _v_aggregate, 2 = left, 3 = right, 4 = top, 5 = bottom (raw ret - heat)
int launch_edge_aggregate(const float *heat, float *out, int n, int c, int h, int w, float weight, int horizontal,
                          cudaStream_t stream) {
  const long long planes = (long long)n * c;
  const bool rows = (horizontal == 1 || horizontal == 2 || horizontal == 3);
  const int mode = (horizontal <= 1) ? 0 : ((horizontal == 2 || horizontal == 4) ? 1 : 2);
  if (rows) {
    const long long nrows = planes * h;
    const bool vec = (w % 4 == 0) && ((((uintptr_t)heat | (uintptr_t)out) & 15u) == 0);
    if (vec)
      k_aggr_rows<true><<<(unsigned)((nrows + 255) / 256), 256, 0, stream>>>(heat, out, nrows, w, weight, mode);
    else
      k_aggr_rows<false><<<(unsigned)((nrows + 255) / 256), 256, 0, stream>>>(heat, out, nrows, w, weight, mode);
  } else {
    const long long blocks = planes * ((w + 127) / 128);
    k_aggr_cols<<<(unsigned)blocks, 128, 0, stream>>>(heat, out, planes, h, w, weight, mode);
  }
  CNB_CHECK_LAUNCH("cnb_edge_aggregate");
  count_launch();
  return CNB_OK;
}

}  // namespace cnb

extern "C" int cnb_edge_aggregate(const float *heat, float *out, int n, int c, int h, int w, float aggr_weight,
                                  int horizontal, void *stream) {
  CNB_REQUIRE(heat && out, CNB_EINVAL, "cnb_edge_aggregate: null pointer");
  CNB_REQUIRE(heat != out, CNB_EINVAL, "cnb_edge_aggregate: in-place operation is not supported");
  CNB_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, CNB_EINVAL, "cnb_edge_aggregate: non-positive dimension");
  CNB_REQUIRE(horizontal >= 0 && horizontal <= 5, CNB_EINVAL, "cnb_edge_aggregate: bad direction code %d", horizontal);
  return cnb::launch_edge_aggregate(heat, out, n, c, h, w, aggr_weight, horizontal, (cudaStream_t)stream);
}
