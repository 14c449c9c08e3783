// DCNv2 forward with the dense contraction on the 5th-generation tensor cores (tcgen05 / UMMA).
//
// Same decomposition as dcnv2.cu (a CTA owns 128 output pixels x 64|128 output channels, the
// bilinear geometry of every (tap, pixel) is computed once, the sampled column tile of 8 input
// channels is built in shared memory -- never in HBM), but the contraction
//     D[128 pixels, Cout_tile] += col[128, 72] * W[Cout_tile, 72]^T
// is issued by ONE thread as tcgen05.mma.kind::tf32 instructions (UTCHMMA in SASS), A and B both
// K-major in the canonical no-swizzle shared-memory layout, the accumulator in TMEM (128 lanes x
// Cout_tile fp32 columns), read back with tcgen05.ld for the bias + store epilogue.
//
// Precision: the reference contraction is an fp32 SGEMM (dcn_v2_cuda.c:93-96).  Plain TF32 would
// miss the 1e-4 bar for K = Cin*9 up to 4608, so every operand is split x = hi + lo with hi = x
// truncated to TF32 (lo is exact in fp32) and three MMAs are accumulated: hi*hi + hi*lo + lo*hi
// ("3xTF32"); the dropped lo*lo term and the TF32 rounding of lo are both ~2^-22 relative.
//
// Sampling: the input is first re-laid channels-last (k_dcn_nhwc: [B][H*W][chunk][8 channels], one
// coalesced pass), so one bilinear corner of a (pixel, tap) brings the 8 channels of the chunk with two
// 16-byte loads from one 32-byte sector -- the NCHW layout needs 8 scalar loads from 8 different planes.
// A work item is a (pixel, tap): 8 vector gathers, 32 FMAs, the TF32 split, four 16-byte stores into
// the UMMA tile (K is ordered tap-major, k = tap * 8 + channel, so the 8 channels are two k-chunks).
//
// The weight tiles are pre-split and pre-tiled once per call by k_dcn_prep_weights into the
// caller-provided workspace, so every chunk's B operand (hi + lo) arrives by two TMA bulk copies
// (cp.async.bulk, UBLKCP) that overlap with the sampling of the A tile.
#include "common.cuh"

namespace cnb {

constexpr int TC_TP = 128;        // pixels per CTA (UMMA M)
constexpr int TC_CB = 32;         // input channels per chunk = one 128-byte line of the channels-last copy
constexpr int TC_TG = 1;          // taps per chunk
constexpr int TC_NTG = 9;         // tap groups (3x3 kernel)
constexpr int TC_K = TC_CB * TC_TG;   // 32: K per chunk, k = channel_local
constexpr int TC_KC = TC_K / 4;   // 16-byte k-chunks
constexpr int TC_THREADS = 512;       // 16 warps, two CTAs per SM: one samples while the other's MMAs run
constexpr uint32_t TC_LBO = 128;            // bytes between consecutive k-chunks (one 8 x 16 B core matrix)
constexpr uint32_t TC_SBO = TC_KC * 128;    // bytes between 8-row groups
constexpr int TC_A_BYTES = TC_TP * TC_K * 4;  // 16384

struct DcnShapeTc {
  int B, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, dg, Ho, Wo;
  int co_t;       // output channels per CTA (64 or 128)
  int cbs_pg;     // 32-channel blocks per deformable group
  int n_chunks;   // chunks = dg * 9 taps * cbs_pg
};

struct __align__(16) TapMetaTc {
  int o[4];
  float w[4];
};

__device__ __forceinline__ uint32_t tc_tile_off(int row, int k) {  // byte offset of element (row, k) in a K-major tile
  return (uint32_t)(row >> 3) * TC_SBO + (uint32_t)(k >> 2) * TC_LBO + (uint32_t)(row & 7) * 16u + (uint32_t)(k & 3) * 4u;
}
__device__ __forceinline__ uint64_t tc_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((TC_LBO >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((TC_SBO >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell); layout_type 0 = no swizzle
  return d;
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// W[Cout][Cin][KT] -> per (cout tile, chunk): hi tile then lo tile, each [co_t][32] in the UMMA layout.
// chunk = (group g, tap t, channel block cbi) in that nesting order; k = channel within the block.
__global__ void __launch_bounds__(256) k_dcn_prep_weights(const float *__restrict__ w, const DcnShapeTc s,
                                                          float *__restrict__ ws) {
  const int KT = s.kh * s.kw;
  const int cpg = s.Cin / s.dg;
  const int tile = blockIdx.x;                 // (cout tile, chunk)
  const int cot = tile / s.n_chunks, ch = tile - cot * s.n_chunks;
  const int per_g = TC_NTG * s.cbs_pg;
  const int g = ch / per_g, r = ch - g * per_g;
  const int tg = r / s.cbs_pg, cbi = r - tg * s.cbs_pg;
  const size_t tile_floats = (size_t)s.co_t * TC_K;
  unsigned char *hi = reinterpret_cast<unsigned char *>(ws + (size_t)tile * 2 * tile_floats);
  unsigned char *lo = hi + tile_floats * 4;
  for (int i = threadIdx.x; i < s.co_t * TC_K; i += blockDim.x) {
    const int o = i / TC_K, k = i - o * TC_K;
    const int t = tg, cw = cbi * TC_CB + k;                    // tap, channel within the group
    float v = 0.f;
    const int oc = cot * s.co_t + o;
    if (oc < s.Cout && cw < cpg && t < KT) v = w[((size_t)oc * s.Cin + g * cpg + cw) * KT + t];
    const float h = tf32_hi(v);
    *reinterpret_cast<float *>(hi + tc_tile_off(o, k)) = h;
    *reinterpret_cast<float *>(lo + tc_tile_off(o, k)) = v - h;
  }
}

// x [B][Cin][HW] -> xt [B][HW][dg * cbs_pg][32]; channel ch of group g sits in block g*cbs_pg + (ch - g*cpg)/32,
// slot (ch - g*cpg) % 32 (identity when Cin/dg is a multiple of 32; pad slots are zero-filled by the caller).
// grid (pixel blocks of 32, channel blocks of 64, B), 256 threads.
__global__ void __launch_bounds__(256) k_dcn_nhwc(const float *__restrict__ x, float *__restrict__ xt,
                                                  const DcnShapeTc s) {
  __shared__ float tile[64][33];
  const long long HW = (long long)s.H * s.W;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 64, b = blockIdx.z;
  const int Cp = s.dg * s.cbs_pg * TC_CB;
  const int cpg = s.Cin / s.dg;
  {
    const int px = threadIdx.x & 31, cr = threadIdx.x >> 5;
    for (int c = cr; c < 64; c += 8) {
      float v = 0.f;
      if (c0 + c < s.Cin && p0 + px < HW) v = __ldg(x + ((long long)b * s.Cin + c0 + c) * HW + p0 + px);
      tile[c][px] = v;
    }
  }
  __syncthreads();
  {
    const int c = threadIdx.x & 63, pr = threadIdx.x >> 6;
    const int ch = c0 + c;
    if (ch < s.Cin) {
      const int g = ch / cpg, within = ch - g * cpg;
      const int slot = (g * s.cbs_pg + within / TC_CB) * TC_CB + within % TC_CB;
      for (int px = pr; px < 32; px += 4)
        if (p0 + px < HW) xt[((long long)b * HW + p0 + px) * Cp + slot] = tile[c][px];
    }
  }
}

// cycle breakdown of CTA (0,0), thread 0 (tools/dbg_dcn_stats.py): total, meta, wait-mma, sample+sync, wait-B, issue, epilogue
__device__ long long g_dcn_dbg[8];

template <int CO_T>
__global__ void __launch_bounds__(TC_THREADS, 2)
k_dcn_forward_tc(const float *__restrict__ xt, const float *__restrict__ offset, const float *__restrict__ mask,
                 const float *__restrict__ wtiles, const float *__restrict__ bias, float *__restrict__ y,
                 const DcnShapeTc s) {
  extern __shared__ __align__(128) unsigned char tc_smem[];
  constexpr int B_BYTES = CO_T * TC_K * 4;
  unsigned char *a_hi = tc_smem;
  unsigned char *a_lo = a_hi + TC_A_BYTES;
  unsigned char *b_hi = a_lo + TC_A_BYTES;
  unsigned char *b_lo = b_hi + B_BYTES;
  TapMetaTc *meta = reinterpret_cast<TapMetaTc *>(b_lo + B_BYTES);     // [9][128]
  __shared__ __align__(8) uint64_t bar_b, bar_mma;
  __shared__ uint32_t tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KT = s.kh * s.kw;
  const long long HWo = (long long)s.Ho * s.Wo, HW = (long long)s.H * s.W;
  const int tiles = (int)((HWo + TC_TP - 1) / TC_TP);
  const int b = blockIdx.x / tiles;
  const long long p_base = (long long)(blockIdx.x - b * tiles) * TC_TP;
  const int cot = blockIdx.y;
  const int Cp = s.dg * s.cbs_pg * TC_CB;       // channel pitch of the channels-last copy

  if (tid == 0) {
    mbar_init(&bar_b, 1);
    mbar_init(&bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)),
                 "n"(CO_T));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CO_T >> 3) << 17) | ((uint32_t)(TC_TP >> 4) << 24);

  const bool dbg = (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0);
  long long d_meta = 0, d_wmma = 0, d_samp = 0, d_wb = 0, d_issue = 0, d_epi = 0, tq;
  const long long d_start = clock64();
  // sampler geometry: a warp gather covers 4 pixels x 32 channels with lane = pixel * 8 + quad, so each
  // quarter-warp (the unit the L1 data pipe serves for 16-byte accesses) reads ONE fully used 128-byte
  // line: 4 wavefronts per gather instead of one per 32-byte sector.  This is what makes the 36x re-read
  // of the input (9 taps x 4 corners) affordable.  The UMMA tile wants the opposite lane order (a
  // quarter-warp = 8 rows of one k-chunk, 128 contiguous bytes), so the reduced values are transposed
  // across the warp with shuffles before the conflict-free 16-byte stores.
  const int sp = lane >> 3, sq = lane & 7;
  const float *xbat = xt + (long long)b * HW * Cp + sq * 4;
  int chunk = 0;
  for (int g = 0; g < s.dg; ++g) {
    tq = clock64();
    // ---- bilinear geometry of every (tap, pixel) of this deformable group (mask folded into the weights)
    if (chunk > 0) mbar_wait(&bar_mma, (uint32_t)((chunk - 1) & 1));   // (meta is not read by the MMAs, but keep the order simple)
    __syncthreads();
    for (int idx = tid; idx < 9 * TC_TP; idx += TC_THREADS) {
      const int t = idx / TC_TP, pp = idx - t * TC_TP;
      const long long p = p_base + pp;
      TapMetaTc mt;
#pragma unroll
      for (int q = 0; q < 4; ++q) { mt.o[q] = 0; mt.w[q] = 0.f; }
      if (p < HWo && t < KT) {
        const int ho = (int)(p / s.Wo), wo = (int)(p - (long long)ho * s.Wo);
        const int i = t / s.kw, j = t - i * s.kw;
        const float *op = offset + ((long long)b * s.dg + g) * 2 * KT * HWo;
        const float dy = __ldg(op + (2 * t) * HWo + p), dx = __ldg(op + (2 * t + 1) * HWo + p);
        const float m = __ldg(mask + (((long long)b * s.dg + g) * KT + t) * HWo + p);
        const float h_im = (float)(ho * s.sh - s.ph + i * s.dh) + dy;
        const float w_im = (float)(wo * s.sw - s.pw + j * s.dw) + dx;
        if (h_im > -1.f && w_im > -1.f && h_im < (float)s.H && w_im < (float)s.W) {   // dcn_v2_im2col_cuda.cu:165
          const float hf = floorf(h_im), wf = floorf(w_im);
          const int hl = (int)hf, wl = (int)wf;
          const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
          const int base = hl * s.W + wl;
          if (hl >= 0 && wl >= 0) { mt.o[0] = base; mt.w[0] = hh * hw * m; }                       // :31-41
          if (hl >= 0 && wl + 1 <= s.W - 1) { mt.o[1] = base + 1; mt.w[1] = hh * lw * m; }
          if (hl + 1 <= s.H - 1 && wl >= 0) { mt.o[2] = base + s.W; mt.w[2] = lh * hw * m; }
          if (hl + 1 <= s.H - 1 && wl + 1 <= s.W - 1) { mt.o[3] = base + s.W + 1; mt.w[3] = lh * lw * m; }
        }
      }
      meta[idx] = mt;
    }
    __syncthreads();
    d_meta += clock64() - tq;
    for (int tg = 0; tg < TC_NTG; ++tg) {     // one tap per chunk
      for (int cbi = 0; cbi < s.cbs_pg; ++cbi, ++chunk) {
        tq = clock64();
        // the previous chunk's MMAs must have consumed A and B before they are overwritten
        if (chunk > 0) mbar_wait(&bar_mma, (uint32_t)((chunk - 1) & 1));
        d_wmma += clock64() - tq;
        tq = clock64();
        if (tid == 0) {  // B operand (hi + lo tiles, contiguous in the workspace): TMA bulk copies
          const float *src = wtiles + ((size_t)cot * s.n_chunks + chunk) * 2 * (size_t)CO_T * TC_K;
          mbar_expect_tx(&bar_b, 2u * B_BYTES);
          for (uint32_t off = 0; off < 2u * B_BYTES; off += 8192u)
            bulk_g2s(b_hi + off, reinterpret_cast<const unsigned char *>(src) + off, 8192u, &bar_b);
        }
        // ---- A operand: sampled column tile of 32 channels of one tap, split into TF32 hi / lo, UMMA layout.
        //      warp item = 4 pixels: every lane gathers the 4 corners of its pixel as 16-byte pieces
        //      (4 channels) of the channels-last copy
        const float *xb = xbat + (long long)(g * s.cbs_pg + cbi) * TC_CB;
        //      both of a warp's items (pixels 8w..8w+3 and 8w+4..8w+7) are gathered before either is reduced
        static_assert(TC_TP / 4 == 2 * (TC_THREADS / 32), "two warp items per warp and chunk");
        const int pp0 = warp * 8 + sp, pp1 = pp0 + 4;
        const TapMetaTc m0 = meta[tg * TC_TP + pp0], m1 = meta[tg * TC_TP + pp1];
        float4 u0[4], u1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) u0[q] = __ldg(reinterpret_cast<const float4 *>(xb + (long long)m0.o[q] * Cp));
#pragma unroll
        for (int q = 0; q < 4; ++q) u1[q] = __ldg(reinterpret_cast<const float4 *>(xb + (long long)m1.o[q] * Cp));
        auto reduce = [&](const TapMetaTc &mt, const float4 (&u)[4]) {
          float4 v;
          v.x = fmaf(mt.w[3], u[3].x, fmaf(mt.w[2], u[2].x, fmaf(mt.w[1], u[1].x, mt.w[0] * u[0].x)));
          v.y = fmaf(mt.w[3], u[3].y, fmaf(mt.w[2], u[2].y, fmaf(mt.w[1], u[1].y, mt.w[0] * u[0].y)));
          v.z = fmaf(mt.w[3], u[3].z, fmaf(mt.w[2], u[2].z, fmaf(mt.w[1], u[1].z, mt.w[0] * u[0].z)));
          v.w = fmaf(mt.w[3], u[3].w, fmaf(mt.w[2], u[2].w, fmaf(mt.w[1], u[1].w, mt.w[0] * u[0].w)));
          return v;
        };
        const float4 v0 = reduce(m0, u0), v1 = reduce(m1, u1);   // (row sp, quad sq) and (row 4 + sp, quad sq)
        // transpose: store instruction j writes (row = lane & 7, quad = (lane >> 3) + 4 j); that value lives in
        // lane (row & 3) * 8 + quad, in v0 for rows 0-3 and v1 for rows 4-7
        const int row = lane & 7;
        const bool upper = row >= 4;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int quad = (lane >> 3) + 4 * j;
          const int src = (row & 3) * 8 + quad;
          float4 a, c;
          a.x = __shfl_sync(0xffffffffu, v0.x, src); c.x = __shfl_sync(0xffffffffu, v1.x, src);
          a.y = __shfl_sync(0xffffffffu, v0.y, src); c.y = __shfl_sync(0xffffffffu, v1.y, src);
          a.z = __shfl_sync(0xffffffffu, v0.z, src); c.z = __shfl_sync(0xffffffffu, v1.z, src);
          a.w = __shfl_sync(0xffffffffu, v0.w, src); c.w = __shfl_sync(0xffffffffu, v1.w, src);
          const float4 v = upper ? c : a;
          const float4 h4 = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
          const uint32_t off = tc_tile_off(warp * 8 + row, quad * 4);
          *reinterpret_cast<float4 *>(a_hi + off) = h4;
          *reinterpret_cast<float4 *>(a_lo + off) = make_float4(v.x - h4.x, v.y - h4.y, v.z - h4.z, v.w - h4.w);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy (UMMA reads)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        d_samp += clock64() - tq;
        if (tid == 0) {
          tq = clock64();
          mbar_wait(&bar_b, (uint32_t)(chunk & 1));
          d_wb += clock64() - tq;
          tq = clock64();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll 1
          for (int ks = 0; ks < TC_K / 8; ++ks) {   // one UMMA per 8 k (two 16-byte k-chunks), 3 per step (3xTF32)
            const uint32_t koff = (uint32_t)ks * 2u * TC_LBO;
            const uint64_t dah = tc_desc(ah + koff), dal = tc_desc(al + koff);
            const uint64_t dbh = tc_desc(bh + koff), dbl = tc_desc(bl + koff);
            const uint32_t acc0 = (chunk > 0 || ks > 0) ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(dal), "l"(dbh),
                         "r"(idesc), "r"(acc0) : "memory");      // lo * hi   (small terms first)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(dah), "l"(dbl),
                         "r"(idesc), "r"(1u) : "memory");        // hi * lo
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(dah), "l"(dbh),
                         "r"(idesc), "r"(1u) : "memory");        // hi * hi
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                           smem_u32(&bar_mma)) : "memory");
          d_issue += clock64() - tq;
        }
      }
    }
  }
  // ---- epilogue: TMEM -> registers -> bias -> global (coalesced along pixels)
  tq = clock64();
  mbar_wait(&bar_mma, (uint32_t)((chunk - 1) & 1));
  d_wmma += clock64() - tq;
  tq = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 8) {
    constexpr int HALF = CO_T / 2;                 // columns per warp group (warps 0-3: first half, 4-7: second)
    const int q = warp & 3, h = warp >> 2;
    const long long p = p_base + q * 32 + lane;
    const uint32_t taddr = tm + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * HALF);
#pragma unroll 1
    for (int c32 = 0; c32 < HALF; c32 += 32) {
      uint32_t v[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
          "%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr + (uint32_t)c32));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (p < HWo) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int o = cot * CO_T + h * HALF + c32 + j;
          if (o < s.Cout) {
            const float bv = bias ? __ldg(bias + o) : 0.f;
            y[((long long)b * s.Cout + o) * HWo + p] = __uint_as_float(v[j]) + bv;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  d_epi = clock64() - tq;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(CO_T));
  if (dbg) {
    g_dcn_dbg[0] = clock64() - d_start; g_dcn_dbg[1] = d_meta; g_dcn_dbg[2] = d_wmma; g_dcn_dbg[3] = d_samp;
    g_dcn_dbg[4] = d_wb; g_dcn_dbg[5] = d_issue; g_dcn_dbg[6] = d_epi; g_dcn_dbg[7] = chunk;
  }
}

static void fill_shape(DcnShapeTc *s, int b, int cin, int h, int w, int cout, int kh, int kw, int sh, int sw, int ph,
                       int pw, int dh, int dw, int dg) {
  s->B = b; s->Cin = cin; s->H = h; s->W = w; s->Cout = cout; s->kh = kh; s->kw = kw; s->sh = sh; s->sw = sw;
  s->ph = ph; s->pw = pw; s->dh = dh; s->dw = dw; s->dg = dg;
  s->Ho = (h + 2 * ph - (dh * (kh - 1) + 1)) / sh + 1;
  s->Wo = (w + 2 * pw - (dw * (kw - 1) + 1)) / sw + 1;
  s->co_t = cout > 64 ? 128 : 64;
  const int cpg = cin / dg;
  s->cbs_pg = (cpg + TC_CB - 1) / TC_CB;
  s->n_chunks = dg * TC_NTG * s->cbs_pg;
}

static size_t tc_weight_bytes(const DcnShapeTc &s) {
  const int cot = (s.Cout + s.co_t - 1) / s.co_t;
  return align_up((size_t)cot * s.n_chunks * 2 * s.co_t * TC_K * 4, 256);
}

// weights re-tiled (hi/lo) + the channels-last copy of the input
size_t dcn_tc_workspace_bytes(int b, int cin, int h, int w, int cout, int dg) {
  DcnShapeTc s;
  fill_shape(&s, b, cin, h, w, cout, 3, 3, 1, 1, 1, 1, 1, 1, dg);
  return tc_weight_bytes(s) + align_up((size_t)b * h * w * dg * s.cbs_pg * TC_CB * 4, 256);
}

// Tensor-core forward.  Requires kh*kw <= 9 (checked by the caller) and a workspace of
// dcn_tc_workspace_bytes(); enqueues the weight re-tiling, the channels-last copy and the fused forward.
int dcn_forward_tc(const float *input, const float *offset, const float *mask, const float *weight,
                   const float *bias, float *output, int b, int cin, int h, int w, int cout, int kh, int kw, int sh,
                   int sw, int ph, int pw, int dh, int dw, int dg, void *workspace, cudaStream_t stream) {
  DcnShapeTc s;
  fill_shape(&s, b, cin, h, w, cout, kh, kw, sh, sw, ph, pw, dh, dw, dg);
  const int cot = (cout + s.co_t - 1) / s.co_t;
  float *ws = reinterpret_cast<float *>(workspace);
  float *xt = reinterpret_cast<float *>(reinterpret_cast<char *>(workspace) + tc_weight_bytes(s));
  k_dcn_prep_weights<<<cot * s.n_chunks, 256, 0, stream>>>(weight, s, ws);
  CNB_CHECK_LAUNCH("cnb_dcnv2_forward weight tiles");
  const long long HW = (long long)h * w;
  if ((cin / dg) % TC_CB != 0)   // pad slots of the last channel block of every group must read as zero
    CNB_CUDA(cudaMemsetAsync(xt, 0, (size_t)b * HW * dg * s.cbs_pg * TC_CB * 4, stream));
  dim3 tgrid((unsigned)((HW + 31) / 32), (unsigned)((cin + 63) / 64), (unsigned)b);
  k_dcn_nhwc<<<tgrid, 256, 0, stream>>>(input, xt, s);
  CNB_CHECK_LAUNCH("cnb_dcnv2_forward channels-last copy");
  const long long HWo = (long long)s.Ho * s.Wo;
  const int tiles = (int)((HWo + TC_TP - 1) / TC_TP);
  dim3 grid((unsigned)(b * tiles), (unsigned)cot);
  const size_t smem = 2 * (size_t)TC_A_BYTES + 2 * (size_t)s.co_t * TC_K * 4 + sizeof(TapMetaTc) * 9 * TC_TP;
  if (s.co_t == 64) {
    CNB_CUDA(cudaFuncSetAttribute(k_dcn_forward_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_dcn_forward_tc<64><<<grid, TC_THREADS, smem, stream>>>(xt, offset, mask, ws, bias, output, s);
  } else {
    CNB_CUDA(cudaFuncSetAttribute(k_dcn_forward_tc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_dcn_forward_tc<128><<<grid, TC_THREADS, smem, stream>>>(xt, offset, mask, ws, bias, output, s);
  }
  CNB_CHECK_LAUNCH("cnb_dcnv2_forward (tcgen05)");
  count_launch(3);
  return CNB_OK;
}

}  // namespace cnb

// debug: cycle breakdown of the last tensor-core forward (CTA 0, thread 0); synchronises the device
extern "C" int cnb_debug_dcn_stats(long long *out8) {
  return cudaMemcpyFromSymbol(out8, cnb::g_dcn_dbg, sizeof(long long) * 8) == cudaSuccess ? 0 : 1;
}
