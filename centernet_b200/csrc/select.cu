// Fused 3x3 peak-NMS + exact top-K (score desc, flat index asc) for sm_100a.
//
// Reference semantics (models/decode.py):
//   _nms   :9-15     keep = (maxpool3x3(heat) == heat); heat*keep
//   _topk  :103-119  per-class top-K, then top-K over the C*K survivors
// Under the total order (score desc, flat index asc) the two-stage top-K equals ONE
// global top-K over the NMS'd map, so the kernel streams every plane exactly once:
//
//   stage 1 (persistent, one CTA per SM, 1024 threads)
//     * the B*C planes are cut into n_cta contiguous, equal ranges (perfect balance);
//     * a 3-deep ring of 64 KiB shared-memory stages is filled by TMA bulk copies
//       (cp.async.bulk + mbarrier, UBLKCP in SASS) -- a 128x128 fp32 plane is one stage;
//     * each warp sweeps rows with a 3-row register window: vertical max3 on float4,
//       horizontal neighbours via warp shuffles, `==` peak test, threshold test, and a
//       ballot bitmask of the qualifying pixels (no per-pixel atomics);
//     * qualifying pixels are expanded into a 2048-entry key buffer; when it fills it
//       is bitonic-sorted and cut to K, which raises the running threshold so later
//       planes of the same image contribute only a handful of candidates;
//     * at an image boundary the CTA's best <=K keys go to that image's segment.
//   finalize (one CTA per image): merge the <=max_slots segments, sort, emit the K best
//     with the ctdet gathers / box assembly fused (decode.py:472-493).
//
// Non-positive scores (zero fillers, negative peaks) only matter when an image has < K
// positive peaks; finalize handles that exactly with an ordered rescan (rare slow path).
#include "select.cuh"

namespace cnb {

typedef unsigned long long u64;

__host__ __device__ __forceinline__ long long cta_first_plane(long long cta, long long P, int n_cta) {
  return cta * P / n_cta;
}
// CTA whose range contains plane p: largest i with floor(i*P/n) <= p.
__host__ __device__ __forceinline__ int cta_of_plane(long long p, long long P, int n_cta) {
  return (int)(((p + 1) * (long long)n_cta - 1) / P);
}

// ------------------------------------------------------------------ CTA-wide bitonic sort
// Sorts buf[0..n) descending; n is a power of two; all threads of the CTA must call.
__device__ __forceinline__ void cta_sort_desc(u64 *buf, int n) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (n >> 1); t += nt) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const u64 a = buf[i], b = buf[i + j];
        const bool desc = (i & k) == 0;
        if ((a < b) == desc) {
          buf[i] = b;
          buf[i + j] = a;
        }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ int next_pow2(int v) {
  int n = 2;
  while (n < v) n <<= 1;
  return n;
}

// Sort the candidate buffer and keep the K best.  Returns the new 64-bit threshold
// (key of the K-th best, or 0 when fewer than K are known).  CTA-uniform.
__device__ __forceinline__ u64 cta_prune(u64 *buf, int *s_cnt, int K) {
  __syncthreads();
  const int cnt = *s_cnt;
  const int n = next_pow2(cnt);
  for (int t = cnt + threadIdx.x; t < n; t += blockDim.x) buf[t] = 0ull;
  __syncthreads();
  cta_sort_desc(buf, n);
  u64 thr = 0ull;
  if (cnt >= K) thr = buf[K - 1];
  __syncthreads();
  if (threadIdx.x == 0 && cnt > K) *s_cnt = K;
  __syncthreads();
  return thr;
}

// ------------------------------------------------------------------ stage 1
struct UnitGeom {
  long long plane;
  int img, c;
  int r0, r1;  // rows owned by the unit
  int ra, rb;  // rows resident in the stage (halo included)
};

__device__ __forceinline__ UnitGeom unit_geom(const SelectPlan &pl, long long p_begin, int u) {
  UnitGeom g;
  const int pi = u / pl.upp, strip = u - pi * pl.upp;
  g.plane = p_begin + pi;
  g.img = (int)(g.plane / pl.C);
  g.c = (int)(g.plane - (long long)g.img * pl.C);
  g.r0 = strip * pl.rb;
  g.r1 = min(pl.H, g.r0 + pl.rb);
  g.ra = max(g.r0 - 1, 0);
  g.rb = min(g.r1 + 1, pl.H);
  return g;
}

template <bool NMS, bool TMA>
__global__ void __launch_bounds__(SEL_THREADS, 1)
k_select_stage1(const float *__restrict__ src, const SelectPlan pl, u64 *__restrict__ cand,
                int *__restrict__ cand_cnt) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *stages = reinterpret_cast<float *>(smem_raw);
  u64 *buf = reinterpret_cast<u64 *>(smem_raw + (size_t)SEL_STAGES * SEL_STAGE_BYTES);
  uint32_t *masks = reinterpret_cast<uint32_t *>(buf + SEL_CAP);
  uint64_t *full = reinterpret_cast<uint64_t *>(masks + SEL_MASK_WORDS);
  int *s_cnt = reinterpret_cast<int *>(full + SEL_STAGES);  // [0] buffer count, [1..2] unit counts

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = pl.H, W = pl.W, Wp = pl.Wp, ncb = pl.ncb, K = pl.K;
  const long long HW = (long long)H * W;
  const long long p_begin = cta_first_plane(blockIdx.x, pl.P, pl.n_cta);
  const long long p_end = cta_first_plane(blockIdx.x + 1, pl.P, pl.n_cta);
  const int total_units = (int)(p_end - p_begin) * pl.upp;

  auto issue = [&](int u) {  // thread 0 only (TMA path)
    const UnitGeom g = unit_geom(pl, p_begin, u);
    const int s = u % SEL_STAGES;
    const uint32_t bytes = (uint32_t)(g.rb - g.ra) * (uint32_t)W * 4u;
    const char *gsrc = reinterpret_cast<const char *>(src + (g.plane * H + g.ra) * (long long)W);
    char *dst = reinterpret_cast<char *>(stages) + (size_t)s * SEL_STAGE_BYTES;
    mbar_expect_tx(&full[s], bytes);
    for (uint32_t off = 0; off < bytes; off += 16384u)
      bulk_g2s(dst + off, gsrc + off, min(16384u, bytes - off), &full[s]);
  };

  if (tid == 0) {
    s_cnt[0] = 0;
    s_cnt[1] = 0;
    s_cnt[2] = 0;
    if (TMA) {
      for (int s = 0; s < SEL_STAGES; ++s) mbar_init(&full[s], 1);
      mbar_fence_init();
    }
  }
  __syncthreads();
  if (TMA && tid == 0) {
    for (int u = 0; u < SEL_STAGES && u < total_units; ++u) issue(u);
  }

  u64 thr64 = 0ull;
  int cur_img = -1;

  auto flush = [&](int img) {
    cta_prune(buf, s_cnt, K);
    const int n = min(s_cnt[0], K);
    const int slot = (int)blockIdx.x - cta_of_plane((long long)img * pl.C, pl.P, pl.n_cta);
    u64 *dst = cand + ((size_t)img * pl.max_slots + slot) * K;
    for (int t = tid; t < n; t += SEL_THREADS) dst[t] = buf[t];
    if (tid == 0) cand_cnt[(size_t)img * pl.max_slots + slot] = n;
    __syncthreads();
    if (tid == 0) s_cnt[0] = 0;
    __syncthreads();
  };

  for (int u = 0; u < total_units; ++u) {
    const UnitGeom g = unit_geom(pl, p_begin, u);
    const int s = TMA ? (u % SEL_STAGES) : 0;
    float *st = stages + (size_t)s * (SEL_STAGE_BYTES / 4);
    if (g.img != cur_img) {
      if (cur_img >= 0) flush(cur_img);
      cur_img = g.img;
      thr64 = 0ull;
    }
    if (TMA) {
      mbar_wait(&full[s], (uint32_t)((u / SEL_STAGES) & 1));
    } else {
      // generic loader (W % 4 != 0 or unaligned base): pad columns are -inf
      const int nrow = g.rb - g.ra;
      const float *gsrc = src + (g.plane * H + g.ra) * (long long)W;
      for (int i = tid; i < nrow * Wp; i += SEL_THREADS) {
        const int r = i / Wp, col = i - r * Wp;
        st[i] = (col < W) ? __ldg(gsrc + (long long)r * W + col) : CNB_NEG_INF;
      }
      __syncthreads();
    }

    // ---------------- phase A: peak + threshold test, bitmask of qualifying pixels
    const uint32_t thr_hi = key_bits(thr64);
    const int rows = g.r1 - g.r0;
    const int ngroups = (rows + SEL_RW - 1) / SEL_RW;
    int local = 0;
    const float NI = CNB_NEG_INF;
    for (int item = warp; item < ngroups * ncb; item += SEL_WARPS) {
      const int grp = item / ncb, cb = item - grp * ncb;
      const int y0 = g.r0 + grp * SEL_RW, yend = min(y0 + SEL_RW, g.r1);
      const int x0 = cb * 128 + lane * 4;
      const bool xin = x0 < Wp;
      auto ldrow = [&](int y) -> float4 {
        if (y < 0 || y >= H || !xin) return make_float4(NI, NI, NI, NI);
        return *reinterpret_cast<const float4 *>(st + (size_t)(y - g.ra) * Wp + x0);
      };
      auto vcol = [&](int y, int x) -> float {  // vertical max3 of one column (warp-edge lanes)
        float m = st[(size_t)(y - g.ra) * Wp + x];
        if (y - 1 >= 0) m = fmaxf(m, st[(size_t)(y - 1 - g.ra) * Wp + x]);
        if (y + 1 < H) m = fmaxf(m, st[(size_t)(y + 1 - g.ra) * Wp + x]);
        return m;
      };
      float4 a = NMS ? ldrow(y0 - 1) : make_float4(NI, NI, NI, NI);
      float4 b = ldrow(y0);
#pragma unroll 2
      for (int y = y0; y < yend; ++y) {
        bool q0, q1, q2, q3;
        const uint32_t b0 = __float_as_uint(b.x), b1 = __float_as_uint(b.y);
        const uint32_t b2 = __float_as_uint(b.z), b3 = __float_as_uint(b.w);
        // strictly positive, non-NaN, and not below the running threshold
        q0 = (b0 - 1u < 0x7f800000u) && (b0 >= thr_hi);
        q1 = (b1 - 1u < 0x7f800000u) && (b1 >= thr_hi);
        q2 = (b2 - 1u < 0x7f800000u) && (b2 >= thr_hi);
        q3 = (b3 - 1u < 0x7f800000u) && (b3 >= thr_hi);
        float4 c = make_float4(NI, NI, NI, NI);
        if (NMS) {
          c = ldrow(y + 1);
          const float v0 = fmax3(a.x, b.x, c.x), v1 = fmax3(a.y, b.y, c.y);
          const float v2 = fmax3(a.z, b.z, c.z), v3 = fmax3(a.w, b.w, c.w);
          float l = __shfl_up_sync(0xffffffffu, v3, 1);
          float r = __shfl_down_sync(0xffffffffu, v0, 1);
          if (lane == 0) l = (cb > 0) ? vcol(y, x0 - 1) : NI;
          if (lane == 31) r = (x0 + 4 < Wp) ? vcol(y, x0 + 4) : NI;
          q0 = q0 && (b.x == fmax3(l, v0, v1));
          q1 = q1 && (b.y == fmax3(v0, v1, v2));
          q2 = q2 && (b.z == fmax3(v1, v2, v3));
          q3 = q3 && (b.w == fmax3(v2, v3, r));
        }
        const uint32_t m0 = __ballot_sync(0xffffffffu, q0), m1 = __ballot_sync(0xffffffffu, q1);
        const uint32_t m2 = __ballot_sync(0xffffffffu, q2), m3 = __ballot_sync(0xffffffffu, q3);
        if (lane < 4) {
          const uint32_t mv = lane == 0 ? m0 : (lane == 1 ? m1 : (lane == 2 ? m2 : m3));
          masks[((y - g.r0) * ncb + cb) * 4 + lane] = mv;
        }
        local += __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
        a = b;
        b = NMS ? c : ldrow(y + 1);
      }
    }
    if (lane == 0 && local) atomicAdd(&s_cnt[1 + (u & 1)], local);
    __syncthreads();

    // ---------------- phase B: expand bitmasks into the key buffer
    const int uc = s_cnt[1 + (u & 1)];
    const int n0 = s_cnt[0];
    const int words = rows * ncb * 4;
    const uint32_t cbase = (uint32_t)((long long)g.c * HW);
    auto push_bit = [&](int word, int bit) {
      const int j = word & 3, rc = word >> 2;
      const int row = rc / ncb, cb = rc - row * ncb;
      const int y = g.r0 + row, x = cb * 128 + bit * 4 + j;
      const float v = st[(size_t)(y - g.ra) * Wp + x];
      const u64 key = make_key(__float_as_uint(v), cbase + (uint32_t)(y * W + x));
      if (key > thr64) {
        const int slot = atomicAdd(&s_cnt[0], 1);
        buf[slot] = key;
      }
    };
    if (n0 + uc <= SEL_CAP) {
      if (uc > 0) {
        for (int t = tid; t < words; t += SEL_THREADS) {
          uint32_t m = masks[t];
          while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            push_bit(t, bit);
          }
        }
      }
    } else {
      // overflow-safe path: <=1024 candidates per round, prune whenever the buffer
      // could not take a full round (K <= 1024 <= CAP - 1024).
      for (int base = 0; base < words; base += 32) {
        __syncthreads();
        if (s_cnt[0] + 1024 > SEL_CAP) thr64 = cta_prune(buf, s_cnt, K);
        const int word = base + (tid >> 5);
        if (word < words && ((masks[word] >> (tid & 31)) & 1u)) push_bit(word, tid & 31);
      }
    }
    __syncthreads();
    if (tid == 0) s_cnt[1 + ((u + 1) & 1)] = 0;
    if (s_cnt[0] > SEL_CAP / 2) thr64 = cta_prune(buf, s_cnt, K);
    // the stage is free again: prefetch the unit that will reuse it
    if (TMA && tid == 0 && u + SEL_STAGES < total_units) issue(u + SEL_STAGES);
    if (!TMA) __syncthreads();
  }
  if (cur_img >= 0) flush(cur_img);
}

// ------------------------------------------------------------------ finalize
// v'(i): the reference's heat*keep value of flat pixel i (keep only in NMS mode).
template <bool NMS>
__device__ __forceinline__ float nms_value(const float *__restrict__ img, int C, int H, int W, long long i) {
  const float v = img[i];
  if (!NMS) return v;
  const long long HW = (long long)H * W;
  const long long c = i / HW, sp = i - c * HW;
  const int y = (int)(sp / W), x = (int)(sp - (long long)y * W);
  const float *pl = img + c * HW;
  float m = v;
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= W) continue;
      m = fmaxf(m, pl[(long long)yy * W + xx]);
    }
  }
  return (m == v) ? v : 0.0f;
}

template <bool NMS>
__device__ void finalize_fill(const float *__restrict__ img, const SelectPlan &pl, u64 *sbuf, int have, int K,
                              int *s_tmp, u64 *s_red) {
  // Need K-have more entries ranked (value desc, index asc) among NON-positive v'.
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const long long N = (long long)pl.C * pl.H * pl.W;
  int found = 0;
  const int need = K - have;
  // step 1: zeros, in index order
  for (long long base = 0; base < N && found < need; base += blockDim.x) {
    const long long i = base + tid;
    bool pred = false;
    if (i < N) pred = (nms_value<NMS>(img, pl.C, pl.H, pl.W, i) == 0.0f);
    const uint32_t bal = __ballot_sync(0xffffffffu, pred);
    if (lane == 0) s_tmp[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w2 = 0; w2 < nw; ++w2) {
      const int cw = s_tmp[w2];
      if (w2 < warp) before += cw;
      total += cw;
    }
    const int pos = found + before + __popc(bal & ((1u << lane) - 1u));
    if (pred && pos < need) sbuf[have + pos] = make_key(0u, (uint32_t)i);
    found += total;
    __syncthreads();
  }
  found = min(found, need);
  // step 2: negative values, largest first (brute force; pathological inputs only)
  u64 prev = ~0ull;
  while (found < need) {
    u64 best = 0ull;
    for (long long i = tid; i < N; i += blockDim.x) {
      const float v = nms_value<NMS>(img, pl.C, pl.H, pl.W, i);
      if (v < 0.0f) {
        const u64 key = ((u64)(~__float_as_uint(v)) << 32) | (u64)(0xffffffffu - (uint32_t)i);
        if (key < prev && key > best) best = key;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const u64 other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0) s_red[warp] = best;
    __syncthreads();
    best = 0ull;
    for (int w2 = 0; w2 < nw; ++w2) best = s_red[w2] > best ? s_red[w2] : best;
    __syncthreads();
    if (best == 0ull) {  // nothing left (NaNs only): pad with pixel 0 / score 0
      if (tid == 0) sbuf[have + found] = make_key(0u, 0u);
    } else {
      if (tid == 0) sbuf[have + found] = make_key(~(uint32_t)(best >> 32), key_idx(best));
      prev = best;
    }
    ++found;
  }
  __syncthreads();
}

template <bool NMS>
__global__ void __launch_bounds__(1024, 1)
k_select_finalize(const float *__restrict__ src, const SelectPlan pl, const u64 *__restrict__ cand,
                  const int *__restrict__ cand_cnt, const FinalizeOut out) {
  extern __shared__ __align__(16) unsigned char fin_raw[];
  u64 *sbuf = reinterpret_cast<u64 *>(fin_raw);
  __shared__ int s_tmp[32];
  __shared__ u64 s_red[32];
  const int tid = threadIdx.x;
  const int b = blockIdx.x, K = pl.K;
  const int i0 = cta_of_plane((long long)b * pl.C, pl.P, pl.n_cta);
  const int i1 = cta_of_plane((long long)(b + 1) * pl.C - 1, pl.P, pl.n_cta);
  int total = 0;
  for (int slot = 0; slot <= i1 - i0; ++slot) {
    const int n = cand_cnt[(size_t)b * pl.max_slots + slot];
    const u64 *seg = cand + ((size_t)b * pl.max_slots + slot) * K;
    for (int t = tid; t < n; t += blockDim.x) sbuf[total + t] = seg[t];
    total += n;
  }
  const int n = next_pow2(max(total, K));
  for (int t = total + tid; t < n; t += blockDim.x) sbuf[t] = 0ull;
  __syncthreads();
  if (i1 > i0) cta_sort_desc(sbuf, n);  // a single segment arrives sorted already
  if (total < K) {
    const long long N = (long long)pl.C * pl.H * pl.W;
    finalize_fill<NMS>(src + (long long)b * N, pl, sbuf, total, K, s_tmp, s_red);
  }
  const long long HW = (long long)pl.H * pl.W;
  for (int k = tid; k < K; k += blockDim.x) {
    const u64 key = sbuf[k];
    const float score = __uint_as_float(key_bits(key));
    const uint32_t flat = key_idx(key);
    const int cls = (int)(flat / (uint32_t)HW);
    const int sp = (int)(flat - (uint32_t)cls * (uint32_t)HW);
    const int yi = sp / pl.W, xi = sp - yi * pl.W;
    const size_t o = (size_t)b * K + k;
    if (out.scores) out.scores[o] = score;
    if (out.inds) out.inds[o] = (int64_t)sp;
    if (out.clses) out.clses[o] = cls;
    if (out.ys) out.ys[o] = (float)yi;
    if (out.xs) out.xs[o] = (float)xi;
    if (out.dets) {
      float xs = (float)xi, ys = (float)yi;
      if (out.reg) {  // decode.py:472-476
        const float *r = out.reg + (size_t)b * 2 * HW;
        xs += r[sp];
        ys += r[HW + sp];
      } else {  // :477-479
        xs += 0.5f;
        ys += 0.5f;
      }
      const int whc = out.cat_spec_wh ? 2 * pl.C : 2;  // :480-486
      const float *wp = out.wh + ((size_t)b * whc + (out.cat_spec_wh ? 2 * cls : 0)) * HW;
      const float hw_ = wp[sp] * 0.5f, hh_ = wp[HW + sp] * 0.5f;
      float *d = out.dets + o * 6;  // :487-493
      d[0] = xs - hw_;
      d[1] = ys - hh_;
      d[2] = xs + hw_;
      d[3] = ys + hh_;
      d[4] = score;
      d[5] = (float)cls;
    }
  }
}

// ------------------------------------------------------------------ host
static size_t stage1_smem_bytes() {
  return (size_t)SEL_STAGES * SEL_STAGE_BYTES + (size_t)SEL_CAP * 8 + (size_t)SEL_MASK_WORDS * 4 +
         SEL_STAGES * 8 + 16;
}

static int host_next_pow2(int v) {
  int n = 2;
  while (n < v) n <<= 1;
  return n;
}

int make_select_plan(const float *src, int n_img, int C, int H, int W, int K, int nms, SelectPlan *pl) {
  CNB_REQUIRE(n_img > 0 && C > 0 && H > 0 && W > 0 && K > 0, CNB_EINVAL,
              "top-k: non-positive dimension (n_img=%d c=%d h=%d w=%d k=%d)", n_img, C, H, W, K);
  CNB_REQUIRE((long long)K <= (long long)H * W, CNB_EINVAL,
              "top-k: k=%d exceeds h*w=%lld (torch.topk would raise, decode.py:106)", K, (long long)H * W);
  CNB_REQUIRE(K <= SEL_MAX_K, CNB_EUNSUPPORTED, "top-k: k=%d > %d not implemented", K, SEL_MAX_K);
  CNB_REQUIRE((long long)C * H * W < (1ll << 32), CNB_EUNSUPPORTED, "top-k: c*h*w must be < 2^32");
  pl->n_img = n_img; pl->C = C; pl->H = H; pl->W = W; pl->K = K; pl->nms = nms;
  pl->Wp = (W + 3) / 4 * 4;
  pl->ncb = (W + 127) / 128;
  pl->P = (long long)n_img * C;
  pl->use_tma = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
  const int max_rows = SEL_STAGE_BYTES / (pl->Wp * 4);
  CNB_REQUIRE(max_rows >= 3 && pl->ncb <= 64, CNB_EUNSUPPORTED, "top-k: w=%d too wide for a 64 KiB stage", W);
  int rb_cap = SEL_MASK_WORDS / (4 * pl->ncb);
  if (H <= max_rows && H <= rb_cap) {
    pl->rb = H;
  } else {
    pl->rb = max_rows - 2 < rb_cap ? max_rows - 2 : rb_cap;
  }
  pl->upp = (H + pl->rb - 1) / pl->rb;
  int n_cta = num_sms();
  if ((long long)n_cta > pl->P) n_cta = (int)pl->P;
  for (;;) {
    int ms = 1;
    for (int b = 0; b < n_img; ++b) {
      const int i0 = cta_of_plane((long long)b * C, pl->P, n_cta);
      const int i1 = cta_of_plane((long long)(b + 1) * C - 1, pl->P, n_cta);
      if (i1 - i0 + 1 > ms) ms = i1 - i0 + 1;
    }
    if ((long long)ms * K <= SEL_FIN_MAX || n_cta == 1) {
      pl->max_slots = ms;
      break;
    }
    n_cta = (n_cta * 3) / 4 > 0 ? (n_cta * 3) / 4 : 1;
  }
  pl->n_cta = n_cta;
  return CNB_OK;
}

size_t select_workspace_bytes(const SelectPlan &pl) {
  const size_t keys = align_up((size_t)pl.n_img * pl.max_slots * pl.K * 8, 256);
  const size_t cnts = align_up((size_t)pl.n_img * pl.max_slots * 4, 256);
  return keys + cnts;
}

template <bool NMS>
static int launch_select(const float *src, const SelectPlan &pl, const FinalizeOut &out, u64 *cand, int *cnt,
                         cudaStream_t stream) {
  const size_t smem1 = stage1_smem_bytes();
  const size_t smem2 = (size_t)host_next_pow2(pl.max_slots * pl.K > pl.K ? pl.max_slots * pl.K : pl.K) * 8;
  if (pl.use_tma) {
    CNB_CUDA(cudaFuncSetAttribute(k_select_stage1<NMS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem1));
    k_select_stage1<NMS, true><<<pl.n_cta, SEL_THREADS, smem1, stream>>>(src, pl, cand, cnt);
  } else {
    CNB_CUDA(cudaFuncSetAttribute(k_select_stage1<NMS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem1));
    k_select_stage1<NMS, false><<<pl.n_cta, SEL_THREADS, smem1, stream>>>(src, pl, cand, cnt);
  }
  CNB_CHECK_LAUNCH("select stage 1");
  CNB_CUDA(cudaFuncSetAttribute(k_select_finalize<NMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  k_select_finalize<NMS><<<pl.n_img, 1024, smem2, stream>>>(src, pl, cand, cnt, out);
  CNB_CHECK_LAUNCH("select finalize");
  count_launch(2);
  return CNB_OK;
}

int run_select(const float *src, const SelectPlan &pl, const FinalizeOut &out, void *ws, cudaStream_t stream) {
  u64 *cand = reinterpret_cast<u64 *>(ws);
  int *cnt = reinterpret_cast<int *>(reinterpret_cast<char *>(ws) +
                                     align_up((size_t)pl.n_img * pl.max_slots * pl.K * 8, 256));
  return pl.nms ? launch_select<true>(src, pl, out, cand, cnt, stream)
                : launch_select<false>(src, pl, out, cand, cnt, stream);
}

}  // namespace cnb
