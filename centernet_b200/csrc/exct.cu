// exct_decode / agnex_ct_decode (models/decode.py:273-424, :122-271) -- ExtremeNet grouping.
//
//   1. optional edge aggregation of the t,l,b,r maps (:287-291, aggregate.cu)
//   2. fused NMS + clamp-to-1 + top-K on each of the four maps (:294-307, select.cu)
//   3. k_exct_tuples: every (t_i, l_j, b_k, r_m) tuple is scored in registers
//        score = (ts+ls+bs+rs+2*ct)/6 - [score/centre thresholds] - [class mismatch]
//                - [top] - [left] - [bottom] - [right geometry]            (:333-360)
//      and streamed through an exact top-`num_dets` selection (64-bit keys: order-preserving
//      score bits | ~tuple index, candidate buffer + bitonic prune) -- the reference
//      materialises ~30 [B, K^4] tensors instead (10 MB each at K=40, 400 MB at K=100);
//   4. k_exct_merge: merges the per-group segments of an image, adds the regression offsets
//      (or 0.5) and writes the 14-column rows (:366-421).
// Tie order: score desc, tuple index ((i*K+j)*K+k)*K+m asc (torch.topk leaves it open).
#include "select.cuh"

namespace cnb {

typedef unsigned long long u64;
int launch_edge_aggregate(const float *heat, float *out, int n, int c, int h, int w, float weight, int horizontal,
                          cudaStream_t stream);

constexpr int EX_THREADS = 512;
constexpr int EX_CAP = 4096;       // candidate keys per CTA
constexpr int EX_STEP = 2048;      // tuples scored between two capacity checks
constexpr int EX_MAX_DETS = 2048;  // num_dets limit (CAP - STEP)
constexpr int EX_MAX_GROUPS = 16;   // segments per image (the merge sorts groups * num_dets keys in shared memory: 128 KB at 16 x 1000)

struct RawList {
  float *scores; int64_t *inds; int32_t *clses; float *ys; float *xs;
};

struct ExctArgs {
  RawList t, l, b, r;
  const float *ct_heat;     // [B, CC, H, W]   (non-agnostic: CC == C)
  const float *ct_agn;      // [B, H*W] max over classes (agnostic)
  const int32_t *ct_cls;    // [B, H*W] argmax over classes (agnostic)
  const float *t_regr, *l_regr, *b_regr, *r_regr;
  int B, CC, H, W, K, num_dets, groups, agnostic;
  float scores_thresh, center_thresh;
  const uint32_t *ct_max;   // [B] order-preserving bits of the largest centre-map value of the image (pruning bound)
  u64 *gthr;                // [B] best num_dets-th key any group of the image has reached (shared pruning threshold)
  u64 *seg;                 // [B, groups, num_dets]
  int *seg_cnt;             // [B, groups]
  float *dets;              // [B, num_dets, 14]
};

__device__ __forceinline__ uint32_t mono_bits(float v) {  // order-preserving map of any float
  const uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float mono_inv(uint32_t m) {
  return __uint_as_float((m & 0x80000000u) ? (m & 0x7fffffffu) : ~m);
}

__device__ __forceinline__ void sort_desc(u64 *buf, int n) {
  for (int k = 2; k <= n; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const u64 a = buf[i], b = buf[i + j];
        if ((a < b) == ((i & k) == 0)) { buf[i] = b; buf[i + j] = a; }
      }
      __syncthreads();
    }
}
__device__ __forceinline__ int np2(int v) { int n = 2; while (n < v) n <<= 1; return n; }

// max / argmax over classes of the centre map (decode.py:159)
__global__ void __launch_bounds__(256) k_ct_agnostic(const float *__restrict__ ct, int B, int CC, long long HW,
                                                     float *__restrict__ agn, int32_t *__restrict__ cls) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)B * HW) return;
  const long long b = i / HW, sp = i - b * HW;
  const float *p = ct + b * CC * HW + sp;
  float best = p[0];
  int arg = 0;
  for (int c = 1; c < CC; ++c) {
    const float v = p[(long long)c * HW];
    if (v > best) { best = v; arg = c; }
  }
  agn[i] = best;
  cls[i] = arg;
}

// out[b] = max over the image's centre map, as order-preserving bits (out zero-filled: below every float)
__global__ void __launch_bounds__(256) k_ct_max(const float *__restrict__ ct, long long per_image, uint32_t *__restrict__ out) {
  __shared__ uint32_t red[8];
  const float *p = ct + (long long)blockIdx.y * per_image;
  uint32_t best = 0u;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_image; i += (long long)gridDim.x * blockDim.x)
    best = max(best, mono_bits(__ldg(p + i)));
  for (int k = 16; k > 0; k >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, k));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) best = max(best, red[i]);
    atomicMax(out + blockIdx.y, best);
  }
}

__global__ void __launch_bounds__(EX_THREADS, 1) k_exct_tuples(const ExctArgs a) {
  extern __shared__ __align__(16) unsigned char ex_smem[];
  u64 *buf = reinterpret_cast<u64 *>(ex_smem);                     // [EX_CAP]
  float *ls = reinterpret_cast<float *>(buf + EX_CAP);             // 4 lists x (score, y, x, cls) x K
  __shared__ int s_cnt;
  __shared__ u64 s_g;
  const int K = a.K, b = blockIdx.x / a.groups, grp = blockIdx.x - b * a.groups;
  const int tid = threadIdx.x;
  float *S[4], *Y[4], *X[4];
  int *CL[4];
  for (int q = 0; q < 4; ++q) {
    S[q] = ls + (q * 4 + 0) * K; Y[q] = ls + (q * 4 + 1) * K; X[q] = ls + (q * 4 + 2) * K;
    CL[q] = reinterpret_cast<int *>(ls + (q * 4 + 3) * K);
  }
  const RawList *L[4] = {&a.t, &a.l, &a.b, &a.r};
  for (int idx = tid; idx < 4 * K; idx += blockDim.x) {
    const int q = idx / K, i = idx - q * K;
    const size_t o = (size_t)b * K + i;
    S[q][i] = L[q]->scores[o]; Y[q][i] = L[q]->ys[o]; X[q][i] = L[q]->xs[o]; CL[q][i] = L[q]->clses[o];
  }
  if (tid == 0) s_cnt = 0;
  __syncthreads();

  const long long HW = (long long)a.H * a.W;
  const int K2 = K * K;
  const long long K3 = (long long)K2 * K;
  const int nd = a.num_dets;
  // upper bounds for the pruning test: the best bottom / right scores and the largest centre value of the image
  float bmax = S[2][0], rmax = S[3][0];
  for (int q = 1; q < K; ++q) { bmax = fmaxf(bmax, S[2][q]); rmax = fmaxf(rmax, S[3][q]); }
  const float ctmax = mono_inv(a.ct_max[b]);
  u64 thr = 0ull;
  // the groups of an image take the top candidates round-robin (every group gets the same mix of strong and weak ones)
  // and share their thresholds through gthr, so they finish together; the merge orders the keys, not the groups
  for (int i = grp; i < K; i += a.groups) {
    const float ts = S[0][i], ty = Y[0][i], tx = X[0][i];
    const int ci = CL[0][i];
    for (int j = 0; j < K; ++j) {
      const float lsx = S[1][j], ly = Y[1][j], lx = X[1][j];
      // Exact pruning of the whole (i, j, *, *) block: every penalty already decided by (i, j) costs 1.0, the score
      // before penalties is at most the same expression over the best bottom / right / centre values, and every step
      // of the computation is monotone in fp32 -- so if even that bound cannot beat the current num_dets-th key the
      // K^2 tuples need not be scored (CTA-uniform: no barrier is skipped by only some threads).
      {
        int lb = 0;
        lb += ((ts < a.scores_thresh) || (lsx < a.scores_thresh)) ? 1 : 0;
        lb += (!a.agnostic && ci != CL[1][j]) ? 1 : 0;
        lb += (ty > ly) ? 1 : 0;
        lb += (lx > tx) ? 1 : 0;
        float ub = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(ts, lsx), bmax), rmax), __fmul_rn(2.0f, ctmax));
        ub = __fdiv_rn(ub, 6.0f);
        for (int q = 0; q < lb; ++q) ub = __fsub_rn(ub, 1.0f);
        ub = __fadd_rn(ub, 0.0f);      // -0 -> +0: the key order puts -0 below +0
        if (mono_bits(ub) < (uint32_t)(thr >> 32)) continue;
      }
      for (int base = 0; base < K2; base += EX_STEP) {
        // the num_dets-th key of ANY group of the image bounds the result from below: adopt the best one published
        if (tid == 0) s_g = *reinterpret_cast<volatile u64 *>(a.gthr + b);
        __syncthreads();
        const int cnt = s_cnt;           // every push of the previous step is in
        const u64 gth = s_g;
        __syncthreads();                 // nobody pushes before everyone has read: the branches are uniform
        if (gth > thr) thr = gth;
        if (cnt + EX_STEP > EX_CAP) {    // sort, keep the num_dets best, raise the threshold
          const int n = np2(cnt);
          for (int t = cnt + tid; t < n; t += blockDim.x) buf[t] = 0ull;
          __syncthreads();
          sort_desc(buf, n);
          if (cnt >= nd && buf[nd - 1] > thr) thr = buf[nd - 1];
          __syncthreads();
          if (tid == 0) {
            if (cnt > nd) s_cnt = nd;
            if (cnt >= nd) atomicMax(a.gthr + b, thr);
          }
          __syncthreads();
        }
        for (int e = base + tid; e < min(base + EX_STEP, K2); e += blockDim.x) {
          const int k = e / K, m = e - k * K;
          const float bs = S[2][k], rs = S[3][m];
          const float by = Y[2][k], bx = X[2][k], ry = Y[3][m], rx = X[3][m];
          const int cx = ((int)lx + (int)rx) >> 1;   // ((l_x + r_x + 0.5) / 2).long(), decode.py:322
          const int cy = ((int)ty + (int)by) >> 1;   // :323
          float ct;
          if (a.agnostic) ct = __ldg(a.ct_agn + (long long)b * HW + cy * a.W + cx);
          else ct = __ldg(a.ct_heat + ((long long)b * a.CC + ci) * HW + cy * a.W + cx);
          float sc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(ts, lsx), bs), rs), __fmul_rn(2.0f, ct));
          sc = __fdiv_rn(sc, 6.0f);                                                            // :333
          const bool sc_bad = (ts < a.scores_thresh) || (lsx < a.scores_thresh) || (bs < a.scores_thresh) ||
                              (rs < a.scores_thresh) || (ct < a.center_thresh);                // :350-355
          const bool top_bad = (ty > ly) || (ty > by) || (ty > ry);                            // :341-348
          const bool left_bad = (lx > tx) || (lx > bx) || (lx > rx);
          const bool bot_bad = (by < ty) || (by < ly) || (by < ry);
          const bool right_bad = (rx < tx) || (rx < lx) || (rx < bx);
          sc = __fsub_rn(sc, sc_bad ? 1.0f : 0.0f);                                            // :355-360 order
          if (!a.agnostic) {
            const bool cls_bad = (ci != CL[1][j]) || (ci != CL[2][k]) || (ci != CL[3][m]);
            sc = __fsub_rn(sc, cls_bad ? 1.0f : 0.0f);
          }
          sc = __fsub_rn(sc, top_bad ? 1.0f : 0.0f);
          sc = __fsub_rn(sc, left_bad ? 1.0f : 0.0f);
          sc = __fsub_rn(sc, bot_bad ? 1.0f : 0.0f);
          sc = __fsub_rn(sc, right_bad ? 1.0f : 0.0f);
          const uint32_t tuple = (uint32_t)((long long)i * K3 + (long long)j * K2 + e);
          const u64 key = ((u64)mono_bits(sc) << 32) | (u64)(0xffffffffu - tuple);
          if (key > thr) buf[atomicAdd(&s_cnt, 1)] = key;
        }
      }
    }
  }
  // final cut of this group's segment
  __syncthreads();
  const int cnt = s_cnt;
  const int n = np2(cnt);
  for (int t = cnt + tid; t < n; t += blockDim.x) buf[t] = 0ull;
  __syncthreads();
  sort_desc(buf, n);
  const int n_out = min(cnt, nd);
  u64 *dst = a.seg + ((size_t)b * a.groups + grp) * nd;
  for (int t = tid; t < n_out; t += blockDim.x) dst[t] = buf[t];
  if (tid == 0) a.seg_cnt[(size_t)b * a.groups + grp] = n_out;
}

__global__ void __launch_bounds__(1024, 1) k_exct_merge(const ExctArgs a) {
  extern __shared__ __align__(16) unsigned char ex_smem[];
  u64 *sbuf = reinterpret_cast<u64 *>(ex_smem);
  const int b = blockIdx.x, tid = threadIdx.x, K = a.K, nd = a.num_dets;
  int total = 0;
  for (int g = 0; g < a.groups; ++g) {
    const int n = a.seg_cnt[(size_t)b * a.groups + g];
    const u64 *seg = a.seg + ((size_t)b * a.groups + g) * nd;
    for (int t = tid; t < n; t += blockDim.x) sbuf[total + t] = seg[t];
    total += n;
  }
  const int n = np2(total);
  for (int t = total + tid; t < n; t += blockDim.x) sbuf[t] = 0ull;
  __syncthreads();
  if (a.groups > 1) sort_desc(sbuf, n);
  const long long HW = (long long)a.H * a.W;
  const bool has_reg = a.t_regr && a.l_regr && a.b_regr && a.r_regr;   // decode.py:366-367
  for (int r = tid; r < nd; r += blockDim.x) {
    const u64 key = sbuf[r];
    const float score = mono_inv((uint32_t)(key >> 32));
    const uint32_t tuple = 0xffffffffu - (uint32_t)key;
    const int m = tuple % K, k = (tuple / K) % K, j = (tuple / (K * K)) % K, i = tuple / (K * K * K);
    const size_t oi = (size_t)b * K + i, oj = (size_t)b * K + j, ok = (size_t)b * K + k, om = (size_t)b * K + m;
    float tx = a.t.xs[oi], ty = a.t.ys[oi], lx = a.l.xs[oj], ly = a.l.ys[oj];
    float bx = a.b.xs[ok], by = a.b.ys[ok], rx = a.r.xs[om], ry = a.r.ys[om];
    float cls;
    if (a.agnostic) {
      const int cx = ((int)lx + (int)rx) >> 1, cy = ((int)ty + (int)by) >> 1;
      cls = (float)a.ct_cls[(long long)b * HW + cy * a.W + cx];
    } else {
      cls = (float)a.t.clses[oi];
    }
    if (has_reg) {  // :366-384
      const long long st = a.t.inds[oi], sl = a.l.inds[oj], sb = a.b.inds[ok], sr = a.r.inds[om];
      tx += a.t_regr[((long long)b * 2) * HW + st]; ty += a.t_regr[((long long)b * 2 + 1) * HW + st];
      lx += a.l_regr[((long long)b * 2) * HW + sl]; ly += a.l_regr[((long long)b * 2 + 1) * HW + sl];
      bx += a.b_regr[((long long)b * 2) * HW + sb]; by += a.b_regr[((long long)b * 2 + 1) * HW + sb];
      rx += a.r_regr[((long long)b * 2) * HW + sr]; ry += a.r_regr[((long long)b * 2 + 1) * HW + sr];
    } else {        // :385-393
      tx += 0.5f; ty += 0.5f; lx += 0.5f; ly += 0.5f; bx += 0.5f; by += 0.5f; rx += 0.5f; ry += 0.5f;
    }
    float *d = a.dets + ((size_t)b * nd + r) * 14;   // :395-421
    d[0] = lx; d[1] = ty; d[2] = rx; d[3] = by; d[4] = score;
    d[5] = tx; d[6] = ty; d[7] = lx; d[8] = ly; d[9] = bx; d[10] = by; d[11] = rx; d[12] = ry; d[13] = cls;
  }
}

static size_t raw_sz(long long n) { return align_up((size_t)n * 8, 256) + 4 * align_up((size_t)n * 4, 256); }
static RawList carve(char *&p, long long n) {
  RawList r;
  r.inds = reinterpret_cast<int64_t *>(p); p += align_up((size_t)n * 8, 256);
  r.scores = reinterpret_cast<float *>(p); p += align_up((size_t)n * 4, 256);
  r.clses = reinterpret_cast<int32_t *>(p); p += align_up((size_t)n * 4, 256);
  r.ys = reinterpret_cast<float *>(p); p += align_up((size_t)n * 4, 256);
  r.xs = reinterpret_cast<float *>(p); p += align_up((size_t)n * 4, 256);
  return r;
}
static int exct_groups(int b, int k) {
  int g = (2 * num_sms() + b - 1) / b;
  if (g > EX_MAX_GROUPS) g = EX_MAX_GROUPS;
  if (g > k) g = k;
  return g < 1 ? 1 : g;
}

}  // namespace cnb

using namespace cnb;

extern "C" {

size_t cnb_exct_workspace_bytes(int b, int c, int cc, int h, int w, int k, int num_dets) {
  SelectPlan pl;
  if (make_select_plan(nullptr, b, c, h, w, k, 1, &pl) != CNB_OK) return 0;
  const int g = exct_groups(b, k);
  const size_t maps = 4 * align_up((size_t)b * c * h * w * 4, 256);       // aggregated maps (aggr_weight > 0)
  const size_t agn = 2 * align_up((size_t)b * h * w * 4, 256);            // class-agnostic centre map
  (void)cc;
  return select_workspace_bytes(pl) + 4 * raw_sz((long long)b * k) + maps + agn +
         align_up((size_t)b * g * num_dets * 8, 256) + align_up((size_t)b * g * 4, 256) + align_up((size_t)b * 4, 256) +
         align_up((size_t)b * 8, 256);
}

int cnb_exct_decode(const float *t_heat, const float *l_heat, const float *b_heat, const float *r_heat,
                    const float *ct_heat, const float *t_regr, const float *l_regr, const float *b_regr,
                    const float *r_regr, int b, int c, int cc, int h, int w, int k, float scores_thresh,
                    float center_thresh, float aggr_weight, int num_dets, int agnostic, float *dets, void *workspace,
                    size_t workspace_bytes, void *stream_) {
  CNB_REQUIRE(t_heat && l_heat && b_heat && r_heat && ct_heat && dets && workspace, CNB_EINVAL,
              "cnb_exct_decode: null pointer");
  CNB_REQUIRE(b > 0 && c > 0 && cc > 0 && h > 0 && w > 0 && k > 0 && num_dets > 0, CNB_EINVAL,
              "cnb_exct_decode: non-positive dimension");
  CNB_REQUIRE(agnostic || cc == c, CNB_EINVAL, "cnb_exct_decode: centre map has %d classes, extreme maps %d", cc, c);
  CNB_REQUIRE(k <= 255, CNB_EUNSUPPORTED, "cnb_exct_decode: k=%d (k^4 must fit 32 bits)", k);
  const long long k4 = (long long)k * k * k * k;
  CNB_REQUIRE(num_dets <= k4, CNB_EINVAL, "cnb_exct_decode: num_dets=%d exceeds k^4=%lld (torch.topk would raise)",
              num_dets, k4);
  CNB_REQUIRE(num_dets <= EX_MAX_DETS, CNB_EUNSUPPORTED, "cnb_exct_decode: num_dets=%d > %d", num_dets, EX_MAX_DETS);
  CNB_REQUIRE(workspace_bytes >= cnb_exct_workspace_bytes(b, c, cc, h, w, k, num_dets), CNB_EWORKSPACE,
              "cnb_exct_decode: workspace %zu < %zu", workspace_bytes,
              cnb_exct_workspace_bytes(b, c, cc, h, w, k, num_dets));
  cudaStream_t stream = (cudaStream_t)stream_;
  SelectPlan pl;
  int rc = make_select_plan(t_heat, b, c, h, w, k, 1, &pl);
  if (rc != CNB_OK) return rc;
  pl.clamp_one = 1;   // decode.py:299-302: values above 1 are clamped before _topk
  char *p = reinterpret_cast<char *>(workspace);
  void *sel_ws = p; p += select_workspace_bytes(pl);
  ExctArgs a;
  a.t = carve(p, (long long)b * k); a.l = carve(p, (long long)b * k);
  a.b = carve(p, (long long)b * k); a.r = carve(p, (long long)b * k);
  const size_t map_bytes = align_up((size_t)b * c * h * w * 4, 256);
  float *agg[4];
  for (int q = 0; q < 4; ++q) { agg[q] = reinterpret_cast<float *>(p); p += map_bytes; }
  float *agn = reinterpret_cast<float *>(p); p += align_up((size_t)b * h * w * 4, 256);
  int32_t *agn_cls = reinterpret_cast<int32_t *>(p); p += align_up((size_t)b * h * w * 4, 256);
  const int groups = exct_groups(b, k);
  a.seg = reinterpret_cast<u64 *>(p); p += align_up((size_t)b * groups * num_dets * 8, 256);
  a.seg_cnt = reinterpret_cast<int *>(p); p += align_up((size_t)b * groups * 4, 256);
  uint32_t *ct_max = reinterpret_cast<uint32_t *>(p); p += align_up((size_t)b * 4, 256);
  a.ct_max = ct_max;
  a.gthr = reinterpret_cast<u64 *>(p);
  CNB_CUDA(cudaMemsetAsync(a.gthr, 0, (size_t)b * 8, stream));

  const float *maps[4] = {t_heat, l_heat, b_heat, r_heat};
  if (aggr_weight > 0.0f) {   // :287-291: t,b horizontal; l,r vertical
    const int horiz[4] = {1, 0, 1, 0};
    for (int q = 0; q < 4; ++q) {
      rc = launch_edge_aggregate(maps[q], agg[q], b, c, h, w, aggr_weight, horiz[q], stream);
      if (rc != CNB_OK) return rc;
      maps[q] = agg[q];
    }
  }
  RawList *lists[4] = {&a.t, &a.l, &a.b, &a.r};
  for (int q = 0; q < 4; ++q) {
    SelectPlan pq = pl;
    pq.use_tma = pq.use_tma && ((reinterpret_cast<uintptr_t>(maps[q]) & 15u) == 0);
    if (!pq.use_tma) { pq.hot = 0; pq.seg_cap = pq.K; }
    FinalizeOut o = {lists[q]->scores, lists[q]->inds, lists[q]->clses, lists[q]->ys, lists[q]->xs,
                     nullptr, nullptr, 0, nullptr};
    rc = run_select(maps[q], pq, o, sel_ws, stream);
    if (rc != CNB_OK) return rc;
  }
  a.ct_heat = ct_heat; a.ct_agn = nullptr; a.ct_cls = nullptr;
  if (agnostic) {
    const long long n = (long long)b * h * w;
    k_ct_agnostic<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(ct_heat, b, cc, (long long)h * w, agn, agn_cls);
    CNB_CHECK_LAUNCH("cnb_exct_decode agnostic centre");
    count_launch();
    a.ct_agn = agn; a.ct_cls = agn_cls;
  }
  {   // largest centre value per image: the pruning bound of k_exct_tuples
    CNB_CUDA(cudaMemsetAsync(ct_max, 0, (size_t)b * 4, stream));
    const long long per = agnostic ? (long long)h * w : (long long)cc * h * w;
    int blocks = (int)((per + 256 * 8 - 1) / (256 * 8));
    if (blocks > 64) blocks = 64;
    if (blocks < 1) blocks = 1;
    k_ct_max<<<dim3((unsigned)blocks, (unsigned)b), 256, 0, stream>>>(agnostic ? agn : ct_heat, per, ct_max);
    CNB_CHECK_LAUNCH("cnb_exct_decode centre maximum");
    count_launch();
  }
  a.t_regr = t_regr; a.l_regr = l_regr; a.b_regr = b_regr; a.r_regr = r_regr;
  a.B = b; a.CC = cc; a.H = h; a.W = w; a.K = k; a.num_dets = num_dets; a.groups = groups; a.agnostic = agnostic;
  a.scores_thresh = scores_thresh; a.center_thresh = center_thresh; a.dets = dets;
  const size_t smem1 = (size_t)EX_CAP * 8 + (size_t)16 * k * 4;
  CNB_CUDA(cudaFuncSetAttribute(k_exct_tuples, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  k_exct_tuples<<<b * groups, EX_THREADS, smem1, stream>>>(a);
  CNB_CHECK_LAUNCH("cnb_exct_decode tuples");
  int n2 = 2;
  while (n2 < groups * num_dets) n2 <<= 1;
  const size_t smem2 = (size_t)n2 * 8;
  CNB_CUDA(cudaFuncSetAttribute(k_exct_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  k_exct_merge<<<b, 1024, smem2, stream>>>(a);
  CNB_CHECK_LAUNCH("cnb_exct_decode merge");
  count_launch(2);
  return CNB_OK;
}

}  // extern "C"
