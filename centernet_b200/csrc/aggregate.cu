// ExtremeNet edge aggregation (what BASELINE.json calls "corner_pool t/l/b/r"):
// models/decode.py:17-77.  Along a row (left/right) or column (top/bottom)
//     ret[i] = heat[i] + ret[i-1] * [heat[i] >= heat[i-1]],   dir = ret - heat
//     _h_aggregate = w*left + w*right + heat,  _v_aggregate = w*top + w*bottom + heat.
// The fp32 additions are strictly sequential in the reference, so each row/column is
// scanned by ONE thread in that order (a tree/warp scan would re-associate the sums and
// break bit parity); parallelism comes from the B*C*H rows (or B*C*W columns).
//   rows   : 64-row x 128-column tiles are staged through padded shared memory so global
//            accesses stay coalesced while each thread walks its row; the reverse pass
//            parks its result in `out`, the forward pass combines.
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

__global__ void __launch_bounds__(AG_R) k_aggr_rows(const float *__restrict__ heat, float *__restrict__ out,
                                                     long long nrows, int W, float w, int mode) {
  extern __shared__ float ag_sm[];
  float (*th)[AG_CW + 1] = reinterpret_cast<float (*)[AG_CW + 1]>(ag_sm);                         // heat tile
  float (*tr)[AG_CW + 1] = reinterpret_cast<float (*)[AG_CW + 1]>(ag_sm + AG_R * (AG_CW + 1));  // reverse / result
  const long long row0 = (long long)blockIdx.x * AG_R;
  const int nr = (int)min((long long)AG_R, nrows - row0);
  const int r = threadIdx.x;
  const int nchunk = (W + AG_CW - 1) / AG_CW;
  // ---- reverse pass (right aggregate, decode.py:30-41): chunks from the right edge inwards
  float ret = 0.0f, hnext = 0.0f;
  if (mode != 1) {
    for (int ch = nchunk - 1; ch >= 0; --ch) {
      const int c0 = ch * AG_CW, cw = min(AG_CW, W - c0);
      for (int i = threadIdx.x; i < nr * cw; i += AG_R) {
        const int rr = i / cw, cc = i - rr * cw;
        th[rr][cc] = heat[(row0 + rr) * W + c0 + cc];
      }
      __syncthreads();
      if (r < nr) {
        for (int cc = cw - 1; cc >= 0; --cc) {
          const float h = th[r][cc];
          if (c0 + cc == W - 1) {
            ret = h;
          } else {
            ret = (h >= hnext) ? __fadd_rn(h, ret) : h;
          }
          hnext = h;
          tr[r][cc] = __fsub_rn(ret, h);
        }
      }
      __syncthreads();
      for (int i = threadIdx.x; i < nr * cw; i += AG_R) {
        const int rr = i / cw, cc = i - rr * cw;
        out[(row0 + rr) * W + c0 + cc] = tr[rr][cc];
      }
      __syncthreads();
    }
    if (mode == 2) return;
  }
  // ---- forward pass (left aggregate, decode.py:17-28) + combine
  float hprev = 0.0f;
  for (int ch = 0; ch < nchunk; ++ch) {
    const int c0 = ch * AG_CW, cw = min(AG_CW, W - c0);
    for (int i = threadIdx.x; i < nr * cw; i += AG_R) {
      const int rr = i / cw, cc = i - rr * cw;
      th[rr][cc] = heat[(row0 + rr) * W + c0 + cc];
      tr[rr][cc] = (mode == 0) ? out[(row0 + rr) * W + c0 + cc] : 0.0f;
    }
    __syncthreads();
    if (r < nr) {
      for (int cc = 0; cc < cw; ++cc) {
        const float h = th[r][cc];
        if (c0 + cc == 0) {
          ret = h;
        } else {
          ret = (h >= hprev) ? __fadd_rn(h, ret) : h;
        }
        hprev = h;
        tr[r][cc] = combine(mode, w, __fsub_rn(ret, h), tr[r][cc], h);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nr * cw; i += AG_R) {
      const int rr = i / cw, cc = i - rr * cw;
      out[(row0 + rr) * W + c0 + cc] = tr[rr][cc];
    }
    __syncthreads();
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
    const size_t smem = (size_t)2 * AG_R * (AG_CW + 1) * sizeof(float);
    CNB_CUDA(cudaFuncSetAttribute(k_aggr_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_aggr_rows<<<(unsigned)((nrows + AG_R - 1) / AG_R), AG_R, smem, stream>>>(heat, out, nrows, w, weight, mode);
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
