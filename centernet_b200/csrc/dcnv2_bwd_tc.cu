// DCNv2 backward with both dense contractions on the 5th-generation tensor cores (tcgen05 / UMMA, 3xTF32).
// Reference: dcn_v2_cuda.c:104-241 (per sample: SGEMM columns = W^T . dY, col2im_coord, col2im, im2col again,
// SGEMM dW += dY . columns^T, SGEMV dBias) and dcn_v2_im2col_cuda.cu:182-312.
//
// The column gradient
//     dcol[b][p][q = (channel block, tap)][32 ch] = sum_o dY[b][o][p] * W[o][ch][tap]
// is a plain GEMM (M = pixels, N = Cin*9, K = Cout).  k_dcn_bwd_dcol_tc runs it as a persistent TMA -> UMMA ->
// TMEM pipeline (both operands arrive pre-split into TF32 hi/lo tiles; no thread touches an operand) and writes
// dcol channels-last, so that everything after it moves whole 128-byte lines:
//   k_dcn_bwd_offmask   one (tap, 4 pixels) per warp, lane = (pixel, 4-channel quad): reads the dcol line and the
//                       four bilinear corner lines of the channels-last input, reduces dMask / dOffset over the
//                       channels in a fixed order (deterministic; dcn_v2_im2col_cuda.cu:241-312) and scatters
//                       dX = dcol * mask * corner weight with 16-byte vector reductions (red.global.add.v4.f32:
//                       a quarter of the reference's per-element atomicAdds, :182-239) into a channels-last dX
//   k_dcn_bwd_dx_nchw   adds the channels-last dX into the caller's NCHW gradient
// The weight gradient dW[o][ch][tap] = sum_{b,p} dY[b][o][p] * col[b][p][ch][tap] is a GEMM with K = pixels:
// k_dcn_bwd_weight_tc rebuilds the masked column tile in shared memory exactly as the forward sampler does
// (never in HBM), keeps a [128 rows = 4 chunks x 32 ch] x Cout accumulator in TMEM over its whole pixel range and
// writes one partial per CTA; k_dcn_bwd_wreduce adds the partials in a fixed order (deterministic).
#include <stdlib.h>
#include "dcnv2_tc.cuh"

namespace cnb {

constexpr int BW_STAGES = 3;
constexpr int BW_STAGE_BYTES = 4 * TC_A_BYTES;     // A hi, A lo, B hi, B lo: 4 x (128 rows x 32 k x 4 B) = 64 KB
constexpr int BW_TILE_FLOATS = 2 * TC_TP * TC_K;   // one operand tile, hi + lo
constexpr int BW_THREADS = 192;                    // warps 0-3 epilogue, 4 MMA, 5 TMA

struct BwdGeom {
  int KS;      // 32-wide slices of Cout (K of the dcol GEMM)
  int Q;       // chunks (channel block, tap) = nb * 9
  int RG;      // groups of 4 chunks = 128 UMMA rows
  int Qp;      // RG * 4: chunk pitch of dcol
  int tiles;   // pixel tiles per image
};
static BwdGeom bwd_geom(const DcnShapeTc &s) {
  BwdGeom g;
  g.KS = (s.Cout + TC_K - 1) / TC_K;
  g.Q = s.nb * TC_NT;
  g.RG = (g.Q + 3) / 4;
  g.Qp = g.RG * 4;
  g.tiles = s.tiles_x * s.tiles_y;
  return g;
}

// W[Cout][Cin][KT] -> per (row group rg, slice ks): [128 rows = (chunk ql, ch)][32 k = output channel], hi then lo
__global__ void __launch_bounds__(256) k_bwd_prep_w(const float *__restrict__ w, const DcnShapeTc s, const BwdGeom g,
                                                    float *__restrict__ wtt) {
  const int KT = s.kh * s.kw, cpg = s.Cin / s.dg;
  const int rg = blockIdx.x / g.KS, ks = blockIdx.x - rg * g.KS;
  unsigned char *hi = reinterpret_cast<unsigned char *>(wtt + (size_t)blockIdx.x * BW_TILE_FLOATS);
  unsigned char *lo = hi + TC_A_BYTES;
  for (int i = threadIdx.x; i < TC_TP * TC_K; i += blockDim.x) {
    const int k = i & (TC_K - 1), row = i >> 5;
    const int q = rg * 4 + (row >> 5), ch = row & 31;
    const int bi = q / TC_NT, tap = q - bi * TC_NT;
    const int gi = bi / s.cbs_pg, cw = (bi - gi * s.cbs_pg) * TC_CB + ch;
    const int o = ks * TC_K + k;
    float v = 0.f;
    if (q < g.Q && o < s.Cout && cw < cpg && tap < KT) v = __ldg(w + ((size_t)o * s.Cin + gi * cpg + cw) * KT + tap);
    const float h = tf32_hi(v);
    *reinterpret_cast<float *>(hi + tc_tile_off(row, k)) = h;
    *reinterpret_cast<float *>(lo + tc_tile_off(row, k)) = v - h;
  }
}

// dY [B][Cout][HWo] -> per (image, pixel tile, slice ks): [128 px][32 k = output channel], hi then lo.  grid (tiles, KS, B)
__global__ void __launch_bounds__(256) k_bwd_prep_gy(const float *__restrict__ gy, const DcnShapeTc s, const BwdGeom g,
                                                     float *__restrict__ gyt) {
  const int t = blockIdx.x, ks = blockIdx.y, b = blockIdx.z;
  const int pp = threadIdx.x & (TC_TP - 1), half = threadIdx.x >> 7;
  const int ho = (t / s.tiles_x) * s.th + pp / s.tw, wo = (t % s.tiles_x) * s.tw + pp % s.tw;
  const bool ok = ho < s.Ho && wo < s.Wo;
  const long long HWo = (long long)s.Ho * s.Wo, p = (long long)ho * s.Wo + wo;
  unsigned char *hi = reinterpret_cast<unsigned char *>(gyt + ((size_t)((size_t)b * g.tiles + t) * g.KS + ks) * BW_TILE_FLOATS);
  unsigned char *lo = hi + TC_A_BYTES;
#pragma unroll
  for (int kc = 0; kc < 4; ++kc) {
    const int kq = half * 4 + kc;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int o = ks * TC_K + kq * 4 + j;
      v[j] = (ok && o < s.Cout) ? __ldg(gy + ((long long)b * s.Cout + o) * HWo + p) : 0.f;
    }
    const float4 h4 = make_float4(tf32_hi(v[0]), tf32_hi(v[1]), tf32_hi(v[2]), tf32_hi(v[3]));
    const uint32_t off = tc_tile_off(pp, kq * 4);
    *reinterpret_cast<float4 *>(hi + off) = h4;
    *reinterpret_cast<float4 *>(lo + off) = make_float4(v[0] - h4.x, v[1] - h4.y, v[2] - h4.z, v[3] - h4.w);
  }
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc),
               "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
      "%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ dcol = dY^T . W   (item = image, pixel tile, row group)
__global__ void __launch_bounds__(BW_THREADS, 1)
k_dcn_bwd_dcol_tc(const float *__restrict__ gyt, const float *__restrict__ wtt, float *__restrict__ dcol,
                  const DcnShapeTc s, const BwdGeom g, const int n_items) {
  extern __shared__ __align__(128) unsigned char tc_smem[];
  __shared__ __align__(8) uint64_t full[BW_STAGES], empty[BW_STAGES], tm_full[2], tm_empty[2];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < BW_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tm_full[i], 1); mbar_init(&tm_empty[i], TC_EPI_WARPS); }
    mbar_fence_init();
  }
  if (warp == TC_WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;

  if (warp == TC_WARP_TMA) {
    if (lane == 0) {
      int stage = 0, ph = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int rg = item % g.RG, bt = item / g.RG;          // bt = image * tiles + tile
        for (int ks = 0; ks < g.KS; ++ks) {
          mbar_wait(&empty[stage], (uint32_t)(ph ^ 1));
          unsigned char *dst = tc_smem + (size_t)stage * BW_STAGE_BYTES;
          const unsigned char *a = reinterpret_cast<const unsigned char *>(gyt + ((size_t)bt * g.KS + ks) * BW_TILE_FLOATS);
          const unsigned char *bsrc = reinterpret_cast<const unsigned char *>(wtt + ((size_t)rg * g.KS + ks) * BW_TILE_FLOATS);
          mbar_expect_tx(&full[stage], (uint32_t)BW_STAGE_BYTES);
          for (uint32_t off = 0; off < 2u * TC_A_BYTES; off += 8192u) bulk_g2s(dst + off, a + off, 8192u, &full[stage]);
          for (uint32_t off = 0; off < 2u * TC_A_BYTES; off += 8192u)
            bulk_g2s(dst + 2 * TC_A_BYTES + off, bsrc + off, 8192u, &full[stage]);
          if (++stage == BW_STAGES) { stage = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == TC_WARP_MMA) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(TC_TP >> 4) << 24);
      int stage = 0, ph = 0, acc = 0, aph = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        mbar_wait(&tm_empty[acc], (uint32_t)(aph ^ 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tm + (uint32_t)(acc * 128);
        for (int ks = 0; ks < g.KS; ++ks) {
          mbar_wait(&full[stage], (uint32_t)ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t ah = smem_u32(tc_smem + (size_t)stage * BW_STAGE_BYTES), al = ah + TC_A_BYTES;
          const uint32_t bh = al + TC_A_BYTES, bl = bh + TC_A_BYTES;
#pragma unroll
          for (int k8 = 0; k8 < TC_K / 8; ++k8) {
            const uint32_t koff = (uint32_t)k8 * 2u * TC_LBO;
            const uint64_t dah = tc_desc(ah + koff), dal = tc_desc(al + koff);
            const uint64_t dbh = tc_desc(bh + koff), dbl = tc_desc(bl + koff);
            umma_tf32(d_tmem, dal, dbh, idesc, (ks > 0 || k8 > 0) ? 1u : 0u);   // lo * hi  (small terms first)
            umma_tf32(d_tmem, dah, dbl, idesc, 1u);                             // hi * lo
            umma_tf32(d_tmem, dah, dbh, idesc, 1u);                             // hi * hi
          }
          umma_commit(&empty[stage]);
          if (++stage == BW_STAGES) { stage = 0; ph ^= 1; }
        }
        umma_commit(&tm_full[acc]);
        if (++acc == 2) { acc = 0; aph ^= 1; }
      }
    }
  } else {
    // epilogue: TMEM lane = pixel, 32 columns = the 32 channels of one chunk = one 128-byte line of dcol.  A thread
    // holding a whole line would make every store instruction touch 32 lines (16 bytes each); the warp's 32 x 128 B
    // block goes through shared memory instead (16-byte chunks XOR-swizzled by the pixel: conflict-free both ways)
    // and leaves as 8 stores of 4 complete lines.
    const int qw = warp;
    const long long HWo = (long long)s.Ho * s.Wo;
    unsigned char *stg = tc_smem + (size_t)BW_STAGES * BW_STAGE_BYTES + (size_t)qw * 4096;
    const int ej = lane >> 3, ec = lane & 7;
    int acc = 0, aph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int rg = item % g.RG, bt = item / g.RG;
      const int t = bt % g.tiles, b = bt / g.tiles;
      const int ty0 = (t / s.tiles_x) * s.th, tx0 = (t % s.tiles_x) * s.tw;
      float *dimg = dcol + (long long)b * HWo * g.Qp * TC_CB + (long long)rg * 4 * TC_CB + ec * 4;
      mbar_wait(&tm_full[acc], (uint32_t)aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tm + ((uint32_t)(qw * 32) << 16) + (uint32_t)(acc * 128);
#pragma unroll 1
      for (int ql = 0; ql < 4; ++ql) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)(ql * 32), v);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<float4 *>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) =
              make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]),
                          __uint_as_float(v[4 * c + 3]));
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int pl = it * 4 + ej;                       // pixel within the warp's 32
          const float4 val = *reinterpret_cast<const float4 *>(stg + pl * 128 + ((ec ^ (pl & 7)) << 4));
          const int pp = qw * 32 + pl;
          const int ho = ty0 + pp / s.tw, wo = tx0 + pp % s.tw;
          if (ho < s.Ho && wo < s.Wo)
            *reinterpret_cast<float4 *>(dimg + ((long long)ho * s.Wo + wo) * g.Qp * TC_CB + ql * TC_CB) = val;
        }
        __syncwarp();
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_tc(&tm_empty[acc]);
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == TC_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(256));
}

// ------------------------------------------------------------------ dMask / dOffset (+ the dX scatter)
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d));
}
__device__ __forceinline__ float dot4(const float4 a, const float4 b) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}

// Geometry of one (tap, pixel): fractional parts, mask and (hl*W + wl + W + 1) << 7 | flags
// (bit 0-3: corner in range, 4: sampling point inside (-1,H)x(-1,W), 5: pixel inside the output map,
//  6: "near" -- the top-left corner lies within [-GX_RD, GX_RD + 1] of the pixel's own input position, so the
//  sample's dX contribution is collected by the deterministic gather (k_dcn_bwd_dx_gather) instead of scattered).
constexpr int GX_RD = 2;                       // offsets of magnitude < 2 pixels are always near
constexpr int GX_WIN = 2 * GX_RD + 3;          // candidate output rows / columns per input position
constexpr int GX_TH = 8, GX_TW = 16;           // input positions per CTA
constexpr int GX_WH = GX_TH + GX_WIN - 1, GX_WW = GX_TW + GX_WIN - 1;
constexpr int GX_NC = GX_WIN * GX_WIN * TC_NT; // candidates per position (441)
constexpr int GX_ROUNDS = (GX_NC + 31) / 32;
constexpr int GX_LIST = GX_ROUNDS * 32;        // hit list capacity per warp (every candidate a hit)
struct __align__(16) BwdMetaTc {
  float lh, lw, m;
  int packed;
};
__device__ __forceinline__ BwdMetaTc bwd_meta(const DcnShapeTc &s, const float *__restrict__ off_img,
                                              const float *__restrict__ mask_img, int gi, int tap, int ho, int wo,
                                              bool gather = false) {
  const int KT = s.kh * s.kw;
  const long long HWo = (long long)s.Ho * s.Wo;
  BwdMetaTc mt;
  mt.lh = mt.lw = mt.m = 0.f;
  mt.packed = 0;
  if (ho < s.Ho && wo < s.Wo && tap < KT) {
    const long long p = (long long)ho * s.Wo + wo;
    const int ki = tap / s.kw, kj = tap - ki * s.kw;
    const float *op = off_img + ((long long)gi * 2 * KT + 2 * tap) * HWo + p;
    const float dy = __ldg(op), dx = __ldg(op + HWo);
    mt.m = __ldg(mask_img + ((long long)gi * KT + tap) * HWo + p);
    const float h_im = (float)(ho * s.sh - s.ph + ki * s.dh) + dy;
    const float w_im = (float)(wo * s.sw - s.pw + kj * s.dw) + dx;
    int flags = 32;
    int base = 0;
    if (h_im > -1.f && w_im > -1.f && h_im < (float)s.H && w_im < (float)s.W) {   // dcn_v2_im2col_cuda.cu:286-289
      const float hf = floorf(h_im), wf = floorf(w_im);
      const int hl = (int)hf, wl = (int)wf;
      mt.lh = h_im - hf;
      mt.lw = w_im - wf;
      flags |= 16;
      if (hl >= 0 && wl >= 0) flags |= 1;
      if (hl >= 0 && wl + 1 <= s.W - 1) flags |= 2;
      if (hl + 1 <= s.H - 1 && wl >= 0) flags |= 4;
      if (hl + 1 <= s.H - 1 && wl + 1 <= s.W - 1) flags |= 8;
      base = hl * s.W + wl + s.W + 1;
      const int dh_ = hl - (ho * s.sh - s.ph), dw_ = wl - (wo * s.sw - s.pw);
      if (gather && dh_ >= -GX_RD && dh_ <= GX_RD + 1 && dw_ >= -GX_RD && dw_ <= GX_RD + 1) flags |= 64;
    }
    mt.packed = (base << 7) | flags;
  }
  return mt;
}

// grid = images x pixel tiles, 8 warps.  Per deformable group the CTA first lays the geometry of its 9 x 128 sampling
// points into shared memory (every point computed once), then each warp walks (tap, 4 pixels) items with
// lane = (pixel j, 4-channel quad c4): 16-byte loads of the dcol line and of the four corner lines, next channel
// block in flight while the current one is reduced.
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
k_dcn_bwd_offmask(const float *__restrict__ xt, const float *__restrict__ offset, const float *__restrict__ mask,
                  const float *__restrict__ dcol, float *__restrict__ dxt, float *__restrict__ goff,
                  float *__restrict__ gmask, float *__restrict__ gx_far, const DcnShapeTc s, const BwdGeom g,
                  const int b_first) {
  // dxt != NULL: every sample is scattered into the channels-last dX (16-byte reductions);
  // gx_far != NULL: near samples are left to k_dcn_bwd_dx_gather, far ones go to the NCHW gradient with scalar atomics
  __shared__ BwdMetaTc meta[TC_NT * TC_TP];
  const int KT = s.kh * s.kw;
  const long long HWo = (long long)s.Ho * s.Wo, HW = (long long)s.H * s.W;
  const int Cp = s.nb * TC_CB;
  const int t = blockIdx.x % g.tiles, bl = blockIdx.x / g.tiles;   // bl: image within this batch chunk (dcol index)
  const int b = b_first + bl;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = lane >> 3, c4 = lane & 7;
  const int ty0 = (t / s.tiles_x) * s.th, tx0 = (t % s.tiles_x) * s.tw;
  const int tws = s.tw == 16 ? 4 : 3;                               // tile width is 16 or 8 (fill_shape)
  const float *off_img = offset + (long long)b * s.off_bs, *mask_img = mask + (long long)b * s.mask_bs;
  const float *x_img = xt + (long long)b * HW * Cp + c4 * 4;
  float *dx_img = dxt ? dxt + (long long)b * HW * Cp + c4 * 4 : nullptr;
  const float *dc_img = dcol + (long long)bl * HWo * g.Qp * TC_CB + c4 * 4;
  float *gmask_img = gmask ? gmask + (long long)b * s.mask_bs : nullptr;
  float *goff_img = goff ? goff + (long long)b * s.off_bs : nullptr;
  const int xs1 = Cp, xs2 = s.W * Cp, xs3 = (s.W + 1) * Cp;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int gi = 0; gi < s.dg; ++gi) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < KT * TC_TP; idx += 256) {
      const int tap = idx >> 7, pp = idx & (TC_TP - 1);
      meta[idx] = bwd_meta(s, off_img, mask_img, gi, tap, ty0 + (pp >> tws), tx0 + (pp & (s.tw - 1)), gx_far != nullptr);
    }
    __syncthreads();
    const int cb0 = gi * s.cbs_pg;
    // This warp's items k = 0 .. 4 KT - 1: tap k / 4, pixels (warp + 8 (k % 4)) * 4 + j.  The five 16-byte loads of a
    // (item, channel block) step are issued one step ahead -- across the item boundary too -- so the only exposed
    // latency is the very first step's.
    const int n_items = KT * 4;
    struct Item { BwdMetaTc mt; int p, xo, dco; };
    auto item = [&](int k) {
      Item it;
      const int tap = k >> 2, pp = (warp + 8 * (k & 3)) * 4 + j;
      it.mt = meta[tap * TC_TP + pp];
      it.p = (ty0 + (pp >> tws)) * s.Wo + tx0 + (pp & (s.tw - 1));            // 32-bit offsets within one image (host-checked)
      it.xo = ((it.mt.packed >> 7) - (s.W + 1)) * Cp + cb0 * TC_CB;
      it.dco = (it.p * g.Qp + cb0 * TC_NT + tap) * TC_CB;
      return it;
    };
    float4 d, x1, x2, x3, x4;
    auto issue = [&](int flags, int xo, int dco) {      // a corner out of range (or a sample outside: all flags clear) is not loaded
      d = (flags & 16) ? __ldg(reinterpret_cast<const float4 *>(dc_img + dco)) : z4;
      x1 = (flags & 1) ? __ldg(reinterpret_cast<const float4 *>(x_img + xo)) : z4;
      x2 = (flags & 2) ? __ldg(reinterpret_cast<const float4 *>(x_img + (xo + xs1))) : z4;
      x3 = (flags & 4) ? __ldg(reinterpret_cast<const float4 *>(x_img + (xo + xs2))) : z4;
      x4 = (flags & 8) ? __ldg(reinterpret_cast<const float4 *>(x_img + (xo + xs3))) : z4;
    };
    Item cur = item(0);
    issue(cur.mt.packed, cur.xo, cur.dco);
#pragma unroll 1
    for (int k = 0; k < n_items; ++k) {
      const int tap = k >> 2;
      Item nxt = cur;
      if (k + 1 < n_items) nxt = item(k + 1);
      const int flags = cur.mt.packed & 127;
      const bool f1 = flags & 1, f2 = flags & 2, f3 = flags & 4, f4 = flags & 8;
      const float lh = cur.mt.lh, lw = cur.mt.lw, hh = 1.f - lh, hw = 1.f - lw, m = cur.mt.m;
      const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
      const int base = (cur.mt.packed >> 7) - (s.W + 1);
      int xo = cur.xo, dco = cur.dco;
      float am = 0.f, ah = 0.f, aw = 0.f;
#pragma unroll 1
      for (int cbi = 0; cbi < s.cbs_pg; ++cbi) {
        const float4 cd = d, c1 = x1, c2 = x2, c3 = x3, c4v = x4;
        if (cbi + 1 < s.cbs_pg) issue(flags, xo + TC_CB, dco + TC_NT * TC_CB);
        else if (k + 1 < n_items) issue(nxt.mt.packed, nxt.xo, nxt.dco);
        if (flags & 16) {
          // dMask (:297): dcol * unmasked bilinear sample
          float4 v;
          v.x = w1 * c1.x + w2 * c2.x + w3 * c3.x + w4 * c4v.x;
          v.y = w1 * c1.y + w2 * c2.y + w3 * c3.y + w4 * c4v.y;
          v.z = w1 * c1.z + w2 * c2.z + w3 * c3.z + w4 * c4v.z;
          v.w = w1 * c1.w + w2 * c2.w + w3 * c3.w + w4 * c4v.w;
          am += dot4(cd, v);
          // dOffset (:75-116, :299-302): dcol * mask * d(sample)/d{h,w}
          const float4 dm = make_float4(cd.x * m, cd.y * m, cd.z * m, cd.w * m);
          float4 gh, gw;
          gh.x = -hw * c1.x - lw * c2.x + hw * c3.x + lw * c4v.x;
          gh.y = -hw * c1.y - lw * c2.y + hw * c3.y + lw * c4v.y;
          gh.z = -hw * c1.z - lw * c2.z + hw * c3.z + lw * c4v.z;
          gh.w = -hw * c1.w - lw * c2.w + hw * c3.w + lw * c4v.w;
          gw.x = -hh * c1.x + hh * c2.x - lh * c3.x + lh * c4v.x;
          gw.y = -hh * c1.y + hh * c2.y - lh * c3.y + lh * c4v.y;
          gw.z = -hh * c1.z + hh * c2.z - lh * c3.z + lh * c4v.z;
          gw.w = -hh * c1.w + hh * c2.w - lh * c3.w + lh * c4v.w;
          ah += dot4(dm, gh);
          aw += dot4(dm, gw);
          // dX (:182-239): the four corners receive dcol * mask * corner weight
          if (dx_img) {
            if (f1 && w1 != 0.f) red_add_v4(dx_img + xo, dm.x * w1, dm.y * w1, dm.z * w1, dm.w * w1);
            if (f2 && w2 != 0.f) red_add_v4(dx_img + (xo + xs1), dm.x * w2, dm.y * w2, dm.z * w2, dm.w * w2);
            if (f3 && w3 != 0.f) red_add_v4(dx_img + (xo + xs2), dm.x * w3, dm.y * w3, dm.z * w3, dm.w * w3);
            if (f4 && w4 != 0.f) red_add_v4(dx_img + (xo + xs3), dm.x * w4, dm.y * w4, dm.z * w4, dm.w * w4);
          } else if (gx_far && !(flags & 64)) {
            const int cw = cbi * TC_CB + c4 * 4, cpg = s.Cin / s.dg;     // channel within the group (pad slots skipped)
            float *gp = gx_far + ((long long)b * s.Cin + gi * cpg + cw) * HW + base;
            const float dv[4] = {dm.x, dm.y, dm.z, dm.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (cw + i >= cpg) break;
              float *q = gp + (long long)i * HW;
              if (f1 && w1 != 0.f) atomicAdd(q, dv[i] * w1);
              if (f2 && w2 != 0.f) atomicAdd(q + 1, dv[i] * w2);
              if (f3 && w3 != 0.f) atomicAdd(q + s.W, dv[i] * w3);
              if (f4 && w4 != 0.f) atomicAdd(q + s.W + 1, dv[i] * w4);
            }
          }
        }
        xo += TC_CB;
        dco += TC_NT * TC_CB;
      }
      // fixed-order reduction over the 8 channel quads of a pixel
#pragma unroll
      for (int sh = 4; sh > 0; sh >>= 1) {
        am += __shfl_xor_sync(0xffffffffu, am, sh);
        ah += __shfl_xor_sync(0xffffffffu, ah, sh);
        aw += __shfl_xor_sync(0xffffffffu, aw, sh);
      }
      // the gradients arrive zero-filled (dcn_v2_func.py:44-48) and every (tap, pixel) has exactly one writer:
      // a plain store is the accumulation
      if ((flags & 32) && c4 == 0) {
        const int go = (gi * KT + tap) * (int)HWo + cur.p;
        if (gmask_img) gmask_img[go] = am;
        if (goff_img) {
          goff_img[2 * go - cur.p] = ah;                 // channel 2 * (gi*KT + tap)
          goff_img[2 * go - cur.p + (int)HWo] = aw;
        }
      }
      cur = nxt;
    }
  }
}

// Deterministic dX for the near samples.  CTA = 8 x 16 input positions of one image; it lays the geometry of every
// output pixel whose near samples can land in the tile (14 x 22 pixels x 9 taps) into shared memory, then one warp per
// position walks the position's 441 candidates (tap, row, column) 32 at a time in a FIXED order: a lane tests whether
// one corner of its candidate is this position; the hits are taken in lane order and every lane (= channel) adds
// weight * dcol of the hit.  No atomics: two runs give the same bits.  The tile is transposed through shared memory
// and added to the NCHW gradient (which already holds the far samples' scatter).
template <int NCB>
__global__ void __launch_bounds__(256)
k_dcn_bwd_dx_gather(const float *__restrict__ offset, const float *__restrict__ mask, const float *__restrict__ dcol,
                    float *__restrict__ gx, const DcnShapeTc s, const BwdGeom g, const int b_first, const int ptx,
                    const int pty) {
  extern __shared__ __align__(16) unsigned char gx_smem[];
  BwdMetaTc *meta = reinterpret_cast<BwdMetaTc *>(gx_smem);                                  // [9][14][22]
  float(*tile)[GX_TH * GX_TW + 1] =
      reinterpret_cast<float(*)[GX_TH * GX_TW + 1]>(gx_smem + sizeof(BwdMetaTc) * TC_NT * GX_WH * GX_WW);   // [NCB*32][129]
  float2 *mylist = reinterpret_cast<float2 *>(gx_smem + sizeof(BwdMetaTc) * TC_NT * GX_WH * GX_WW +
                                              (size_t)NCB * TC_CB * (GX_TH * GX_TW + 1) * 4) + (threadIdx.x >> 5) * GX_LIST;   // per warp
  const int KT = s.kh * s.kw, cpg = s.Cin / s.dg;
  const long long HWo = (long long)s.Ho * s.Wo, HW = (long long)s.H * s.W;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tix = blockIdx.x % ptx, tiy = (blockIdx.x / ptx) % pty, bl = blockIdx.x / (ptx * pty);
  const int b = b_first + bl;
  const int y0 = tiy * GX_TH, x0 = tix * GX_TW;
  const int ho0 = y0 + s.ph - GX_RD - 2, wo0 = x0 + s.pw - GX_RD - 2;     // first output row / column of the window (stride 1)
  const float *off_img = offset + (long long)b * s.off_bs, *mask_img = mask + (long long)b * s.mask_bs;
  const float *dc_img = dcol + (long long)bl * HWo * g.Qp * TC_CB + lane;

  // candidate (tap, r, c) of this lane in every round: its index into the geometry window and into dcol
  int moff[GX_ROUNDS], doff[GX_ROUNDS];
#pragma unroll
  for (int rnd = 0; rnd < GX_ROUNDS; ++rnd) {
    const int cand = rnd * 32 + lane;
    const int tap = cand / (GX_WIN * GX_WIN), rc = cand - tap * (GX_WIN * GX_WIN);
    const int r = rc / GX_WIN, c = rc - r * GX_WIN;
    moff[rnd] = cand < GX_NC ? (tap * GX_WH + r) * GX_WW + c : -1;
    doff[rnd] = ((r * s.Wo + c) * g.Qp + tap) * TC_CB;
  }

  for (int gi = 0; gi < s.dg; ++gi) {
    __syncthreads();
    for (int idx = tid; idx < TC_NT * GX_WH * GX_WW; idx += 256) {
      const int tap = idx / (GX_WH * GX_WW), rem = idx - tap * (GX_WH * GX_WW);
      const int ho = ho0 + rem / GX_WW, wo = wo0 + rem % GX_WW;
      BwdMetaTc mt;
      mt.lh = mt.lw = mt.m = 0.f;
      mt.packed = 0;
      if (ho >= 0 && wo >= 0 && tap < KT) mt = bwd_meta(s, off_img, mask_img, gi, tap, ho, wo, true);
      meta[idx] = mt;
    }
    __syncthreads();
    for (int cbg = 0; cbg < s.cbs_pg; cbg += NCB) {
      const int ncb = min(NCB, s.cbs_pg - cbg);
      // warp = one row of the position tile
      for (int px = 0; px < GX_TW; ++px) {
        const int y = y0 + warp, x = x0 + px;
        float acc[NCB];
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb) acc[cb] = 0.f;
        if (y < s.H && x < s.W) {
          const int mbase = warp * GX_WW + px;
          const int posk = y * s.W + x + s.W + 1;
          const int dbase = (((ho0 + warp) * s.Wo + wo0 + px) * g.Qp + (gi * s.cbs_pg + cbg) * TC_NT) * TC_CB;
          // pass 1: the hits of this position, in candidate order, as (weight, dcol offset) pairs
          int cnt = 0;
#pragma unroll
          for (int rnd = 0; rnd < GX_ROUNDS; ++rnd) {
            float wgt = 0.f;
            bool hit = false;
            if (moff[rnd] >= 0) {
              const BwdMetaTc e = meta[moff[rnd] + mbase];
              if (e.packed & 64) {
                const int diff = posk - (e.packed >> 7);
                const float hh = 1.f - e.lh, hw = 1.f - e.lw;
                if (diff == 0) { hit = e.packed & 1; wgt = hh * hw; }
                else if (diff == 1) { hit = e.packed & 2; wgt = hh * e.lw; }
                else if (diff == s.W) { hit = e.packed & 4; wgt = e.lh * hw; }
                else if (diff == s.W + 1) { hit = e.packed & 8; wgt = e.lh * e.lw; }
                wgt *= e.m;
              }
            }
            const unsigned hits = __ballot_sync(0xffffffffu, hit);
            if (hit) mylist[cnt + __popc(hits & ((1u << lane) - 1u))] = make_float2(wgt, __int_as_float(doff[rnd] + dbase));
            cnt += __popc(hits);
          }
          __syncwarp();
          // pass 2: eight hits' lines in flight at a time, added in list order (every lane = one channel)
          for (int i0 = 0; i0 < cnt; i0 += 8) {
            const int n = min(8, cnt - i0);
            float2 e[8];
            float v[8][NCB];
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (k < n) {
                e[k] = mylist[i0 + k];
#pragma unroll
                for (int cb = 0; cb < NCB; ++cb)
                  if (cb < ncb) v[k][cb] = __ldg(dc_img + (__float_as_int(e[k].y) + cb * TC_NT * TC_CB));
              }
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (k < n) {
#pragma unroll
                for (int cb = 0; cb < NCB; ++cb)
                  if (cb < ncb) acc[cb] = fmaf(e[k].x, v[k][cb], acc[cb]);
              }
          }
          __syncwarp();
        }
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb) tile[cb * TC_CB + lane][warp * GX_TW + px] = acc[cb];
      }
      __syncthreads();
      // tile[channel][position] -> gx[b][channel][y][x], 16 consecutive x per row
      for (int i = tid; i < ncb * TC_CB * GX_TH * GX_TW; i += 256) {
        const int pp = i & (GX_TH * GX_TW - 1), cl = i >> 7;
        const int cw = cbg * TC_CB + cl;                   // channel within the group
        const int y = y0 + (pp >> 4), x = x0 + (pp & 15);
        if (cw < cpg && y < s.H && x < s.W) gx[((long long)b * s.Cin + gi * cpg + cw) * HW + (long long)y * s.W + x] += tile[cl][pp];
      }
      __syncthreads();
    }
  }
}

// gx[b][ch][pos] += dxt[b][pos][slot(ch)]   (inverse of k_dcn_nhwc).  grid (pixel blocks of 32, channel blocks of 64, B)
__global__ void __launch_bounds__(256) k_dcn_bwd_dx_nchw(const float *__restrict__ dxt, float *__restrict__ gx,
                                                         const DcnShapeTc s) {
  __shared__ float tile[64][33];
  const long long HW = (long long)s.H * s.W;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 64, b = blockIdx.z;
  const int Cp = s.nb * TC_CB, cpg = s.Cin / s.dg;
  {
    const int c = threadIdx.x & 63, pr = threadIdx.x >> 6;
    const int ch = c0 + c;
    if (ch < s.Cin) {
      const int gi = ch / cpg, within = ch - gi * cpg;
      const int slot = (gi * s.cbs_pg + within / TC_CB) * TC_CB + within % TC_CB;
      for (int px = pr; px < 32; px += 4)
        tile[c][px] = (p0 + px < HW) ? __ldg(dxt + ((long long)b * HW + p0 + px) * Cp + slot) : 0.f;
    }
  }
  __syncthreads();
  {
    const int px = threadIdx.x & 31, cr = threadIdx.x >> 5;
    for (int c = cr; c < 64; c += 8)
      if (c0 + c < s.Cin && p0 + px < HW) gx[((long long)b * s.Cin + c0 + c) * HW + p0 + px] += tile[c][px];
  }
}

// ------------------------------------------------------------------ dW = dY . col^T   (K = pixels)
constexpr int WG_SAMPLERS = 16;
constexpr int WG_GEOM = 4;                                    // warps 0-3: geometry producers (4 taps x 32 pixels per slice), then epilogue
constexpr int WG_MR = 4;                                      // slices of geometry in flight
constexpr int WG_THREADS = (TC_WARP_S0 + WG_SAMPLERS) * 32;   // warps 0-3 geometry + epilogue, 4 MMA, 5 TMA, 6-21 samplers
constexpr uint32_t WG_SBO = 1040;                             // pitch of the A tile's 8-row groups (bank-conflict-free scalar stores)
constexpr int WG_A_BYTES = (TC_TP / 8) * WG_SBO;              // 16640 per hi / lo tile
__device__ __forceinline__ uint64_t wg_desc_a(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((TC_LBO >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((WG_SBO >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

struct WgGeom {
  int co_r;          // Cout rounded up to 16 (UMMA N)
  int spi;           // 32-pixel slices per image
  int n_slices;      // B * spi
  int splits;        // pixel-range splits (CTAs per row group)
  int stages;
  int stage_bytes;   // 2 * WG_A_BYTES + 2 * co_r * 128
  unsigned wo_magic; // floor(2^32 / Wo) + 1: p / Wo == umulhi(p, magic) for p < 2^32 / Wo; 0 = divide
};
static WgGeom wg_geom(const DcnShapeTc &s, const BwdGeom &g) {
  WgGeom w;
  w.co_r = (s.Cout + 15) / 16 * 16;
  w.spi = (int)(((long long)s.Ho * s.Wo + TC_K - 1) / TC_K);
  w.n_slices = s.B * w.spi;
  int splits = num_sms() / g.RG;          // one CTA per SM at most (a CTA owns ~200 KB of shared memory): no second wave
  if (splits > w.n_slices) splits = w.n_slices;
  if (splits < 1) splits = 1;
  w.splits = splits;
  w.stage_bytes = 2 * WG_A_BYTES + 2 * w.co_r * TC_K * 4;
  w.stages = w.stage_bytes <= 66048 ? 3 : 2;
  w.wo_magic = ((unsigned long long)s.Ho * s.Wo * s.Wo < (1ull << 32)) ? (unsigned)((1ull << 32) / (unsigned)s.Wo + 1) : 0u;
  return w;
}

// dY [B][Cout][HWo] -> per (image, 32-pixel slice): [co_r rows = output channel][32 k = pixel], hi then lo.  grid (spi, B)
// A warp owns one output channel of the slice at a time, so the bias gradient's partial sums (dcn_v2_cuda.c:225-230)
// fall out of the same pass: bpart[slice][o] = sum of the 32 pixels, fixed shuffle order.
__global__ void __launch_bounds__(256) k_bwd_prep_gyk(const float *__restrict__ gy, const DcnShapeTc s, const WgGeom wg,
                                                      float *__restrict__ gyk, float *__restrict__ bpart) {
  const int sl = blockIdx.x, b = blockIdx.y;
  const long long HWo = (long long)s.Ho * s.Wo, p0 = (long long)sl * TC_K;
  const size_t tile = (size_t)wg.co_r * TC_K;
  unsigned char *hi = reinterpret_cast<unsigned char *>(gyk + ((size_t)b * wg.spi + sl) * 2 * tile);
  unsigned char *lo = hi + tile * 4;
  for (int i = threadIdx.x; i < wg.co_r * TC_K; i += blockDim.x) {
    const int k = i & (TC_K - 1), o = i >> 5;
    const float v = (o < s.Cout && p0 + k < HWo) ? __ldg(gy + ((long long)b * s.Cout + o) * HWo + p0 + k) : 0.f;
    const float h = tf32_hi(v);
    *reinterpret_cast<float *>(hi + tc_tile_off(o, k)) = h;
    *reinterpret_cast<float *>(lo + tc_tile_off(o, k)) = v - h;
    if (bpart) {
      float sum = v;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
      if (k == 0) bpart[((size_t)b * wg.spi + sl) * wg.co_r + o] = sum;
    }
  }
}

// gb[o] += sum over slices of bpart[slice][o]: strided per-thread sums, then a fixed tree.  grid = Cout
__global__ void __launch_bounds__(256) k_dcn_bwd_bias_reduce(const float *__restrict__ bpart, float *__restrict__ gb,
                                                             const WgGeom wg) {
  __shared__ float red[256];
  const int o = blockIdx.x;
  float sum = 0.f;
  for (int i = threadIdx.x; i < wg.n_slices; i += 256) sum += __ldg(bpart + (size_t)i * wg.co_r + o);
  red[threadIdx.x] = sum;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if (threadIdx.x < d) red[threadIdx.x] += red[threadIdx.x + d];
    __syncthreads();
  }
  if (threadIdx.x == 0) gb[o] += red[0];
}

// CTA = (row group rg = 4 chunks x 32 channels, split): D[128 rows][Cout] accumulates over the CTA's slices in TMEM.
__global__ void __launch_bounds__(WG_THREADS, 1)
k_dcn_bwd_weight_tc(const float *__restrict__ xt, const float *__restrict__ offset, const float *__restrict__ mask,
                    const float *__restrict__ gyk, float *__restrict__ wpart, const DcnShapeTc s, const BwdGeom g,
                    const WgGeom wg) {
  extern __shared__ __align__(128) unsigned char tc_smem[];
  __shared__ __align__(8) uint64_t a_full[3], b_full[3], empty[3], tm_full, m_full[WG_MR], m_empty[WG_MR];
  __shared__ uint32_t tmem_base;
  __shared__ BwdMetaTc wmeta[WG_MR][128];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rg = blockIdx.x % g.RG, split = blockIdx.x / g.RG;
  const int sl0 = (int)((long long)split * wg.n_slices / wg.splits), sl1 = (int)((long long)(split + 1) * wg.n_slices / wg.splits);
  const int KT = s.kh * s.kw;
  const long long HWo = (long long)s.Ho * s.Wo, HW = (long long)s.H * s.W;
  const int Cp = s.nb * TC_CB;
  const int b_bytes = wg.co_r * TC_K * 4;

  if (tid == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(&a_full[i], WG_SAMPLERS); mbar_init(&b_full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&tm_full, 1);
    for (int i = 0; i < WG_MR; ++i) { mbar_init(&m_full[i], WG_GEOM); mbar_init(&m_empty[i], WG_SAMPLERS); }
    mbar_fence_init();
  }
  if (warp == TC_WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;

  if (warp >= TC_WARP_S0) {
    // =========================================================== samplers: the masked column tile, rows = channels
    // task = (chunk ql, 4 pixels): lane = (pixel j, channel quad c4) gathers 4 corners x 16 bytes and stores its
    // 4 channels x 1 pixel as scalars -- the 8-row groups of the A tile are pitched WG_SBO = 1040 bytes so that the 32
    // lanes of a store hit 32 different banks.  Warps run free of each other: the geometry comes through a ring, the
    // tile goes out through the 3-stage A ring.
    const int sw = warp - TC_WARP_S0;
    const int j = lane >> 3, c4 = lane & 7;
    int t_meta[2], t_xc[2];
    uint32_t t_dst[2];
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int task = sw + it * WG_SAMPLERS;      // chunk ql = task / 8, pixel group pg = task % 8
      const int ql = task >> 3, pg = task & 7;
      const int q = rg * 4 + ql;
      const int bi = q < g.Q ? q / TC_NT : 0;
      t_meta[it] = ql * 32 + pg * 4 + j;
      t_xc[it] = bi * TC_CB + c4 * 4;
      const int row = ql * 32 + c4 * 4;            // + i (i < 4 stays inside one 8-row group)
      t_dst[it] = (uint32_t)(row >> 3) * WG_SBO + (uint32_t)pg * TC_LBO + (uint32_t)(row & 7) * 16u + (uint32_t)j * 4u;
    }
    const int xs1 = Cp, xs2 = s.W * Cp, xs3 = (s.W + 1) * Cp;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    int stage = 0, ph = 0, ms = 0, mph = 0;
    int b = sl0 / wg.spi, sll = sl0 - b * wg.spi;      // image and slice within the image (of the slice being LOADED)
    BwdMetaTc mt[2];
    float4 xc[2][4];
    // take the geometry of the next slice from the ring and issue its eight corner loads (a corner that is out of range,
    // or a sample that is outside altogether -- all four flags clear, weights zero -- is simply not loaded)
    auto load_slice = [&]() {
      mbar_wait(&m_full[ms], (uint32_t)mph);
#pragma unroll
      for (int it = 0; it < 2; ++it) mt[it] = wmeta[ms][t_meta[it]];
      __syncwarp();
      if (lane == 0) mbar_arrive_tc(&m_empty[ms]);
      if (++ms == WG_MR) { ms = 0; mph ^= 1; }
      const float *x_img = xt + (long long)b * HW * Cp;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int flags = mt[it].packed;
        const int xo = ((flags >> 7) - (s.W + 1)) * Cp + t_xc[it];
        xc[it][0] = (flags & 1) ? __ldg(reinterpret_cast<const float4 *>(x_img + xo)) : z4;
        xc[it][1] = (flags & 2) ? __ldg(reinterpret_cast<const float4 *>(x_img + (xo + xs1))) : z4;
        xc[it][2] = (flags & 4) ? __ldg(reinterpret_cast<const float4 *>(x_img + (xo + xs2))) : z4;
        xc[it][3] = (flags & 8) ? __ldg(reinterpret_cast<const float4 *>(x_img + (xo + xs3))) : z4;
      }
      if (++sll == wg.spi) { sll = 0; ++b; }
    };
    for (int sl = sl0; sl < sl1; ++sl) {
      load_slice();
      float4 val[2];
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const float lh = mt[it].lh, lw = mt[it].lw, mm = mt[it].m, hh = 1.f - lh, hw = 1.f - lw;
        const float w1 = hh * hw * mm, w2 = hh * lw * mm, w3 = lh * hw * mm, w4 = lh * lw * mm;   // as the forward sampler
        const float4 x1 = xc[it][0], x2 = xc[it][1], x3 = xc[it][2], x4 = xc[it][3];
        float4 v;
        v.x = fmaf(w4, x4.x, fmaf(w3, x3.x, fmaf(w2, x2.x, w1 * x1.x)));
        v.y = fmaf(w4, x4.y, fmaf(w3, x3.y, fmaf(w2, x2.y, w1 * x1.y)));
        v.z = fmaf(w4, x4.z, fmaf(w3, x3.z, fmaf(w2, x2.z, w1 * x1.z)));
        v.w = fmaf(w4, x4.w, fmaf(w3, x3.w, fmaf(w2, x2.w, w1 * x1.w)));
        val[it] = v;
      }
      mbar_wait(&empty[stage], (uint32_t)(ph ^ 1));
      unsigned char *a_hi = tc_smem + (size_t)stage * wg.stage_bytes;
      unsigned char *a_lo = a_hi + WG_A_BYTES;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const float4 v = val[it];
        const float4 h4 = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
        float *dh = reinterpret_cast<float *>(a_hi + t_dst[it]), *dl = reinterpret_cast<float *>(a_lo + t_dst[it]);
        dh[0] = h4.x; dh[4] = h4.y; dh[8] = h4.z; dh[12] = h4.w;                          // rows c4*4 + 0..3: 16 bytes apart
        dl[0] = v.x - h4.x; dl[4] = v.y - h4.y; dl[8] = v.z - h4.z; dl[12] = v.w - h4.w;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_tc(&a_full[stage]);
      if (++stage == wg.stages) { stage = 0; ph ^= 1; }
    }
  } else if (warp == TC_WARP_TMA) {
    if (lane == 0) {
      int stage = 0, ph = 0;
      for (int sl = sl0; sl < sl1; ++sl) {
        mbar_wait(&empty[stage], (uint32_t)(ph ^ 1));
        unsigned char *dst = tc_smem + (size_t)stage * wg.stage_bytes + 2 * WG_A_BYTES;
        const unsigned char *src = reinterpret_cast<const unsigned char *>(gyk) + (size_t)sl * 2 * b_bytes;
        mbar_expect_tx(&b_full[stage], 2u * (uint32_t)b_bytes);
        for (uint32_t off = 0; off < 2u * (uint32_t)b_bytes; off += 2048u) bulk_g2s(dst + off, src + off, 2048u, &b_full[stage]);
        if (++stage == wg.stages) { stage = 0; ph ^= 1; }
      }
    }
  } else if (warp == TC_WARP_MMA) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(wg.co_r >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int stage = 0, ph = 0;
      for (int sl = sl0; sl < sl1; ++sl) {
        mbar_wait(&a_full[stage], (uint32_t)ph);
        mbar_wait(&b_full[stage], (uint32_t)ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ah = smem_u32(tc_smem + (size_t)stage * wg.stage_bytes), al = ah + WG_A_BYTES;
        const uint32_t bh = al + WG_A_BYTES, bl = bh + (uint32_t)b_bytes;
#pragma unroll
        for (int k8 = 0; k8 < TC_K / 8; ++k8) {
          const uint32_t koff = (uint32_t)k8 * 2u * TC_LBO;
          const uint64_t dah = wg_desc_a(ah + koff), dal = wg_desc_a(al + koff);
          const uint64_t dbh = tc_desc(bh + koff), dbl = tc_desc(bl + koff);
          umma_tf32(tm, dal, dbh, idesc, (sl > sl0 || k8 > 0) ? 1u : 0u);
          umma_tf32(tm, dah, dbl, idesc, 1u);
          umma_tf32(tm, dah, dbh, idesc, 1u);
        }
        umma_commit(&empty[stage]);
        if (++stage == wg.stages) { stage = 0; ph ^= 1; }
      }
      umma_commit(&tm_full);
    }
  } else {
    // =========================================================== warps 0-3 first produce the geometry: 4 taps x 32 pixels of a slice, one
    // entry per thread, WG_MR slices ahead of the samplers (ring in shared memory, mbarrier hand-over); the offsets and
    // mask of the next slice are fetched before the current entry is finished
    const int st = tid;
    const int m_q = rg * 4 + (st >> 5);
    const int m_bi = m_q / TC_NT, m_tap = m_q < g.Q ? m_q - m_bi * TC_NT : TC_NT;     // TC_NT = "no such tap"
    const int m_gi = m_bi / s.cbs_pg;
    const bool m_ok = m_tap < KT;
    const int ki = m_ok ? m_tap / s.kw : 0, kj = m_tap - ki * s.kw;
    const int hb = ki * s.dh - s.ph, wb = kj * s.dw - s.pw;
    const long long o_dy = ((long long)m_gi * 2 * KT + 2 * (m_ok ? m_tap : 0)) * HWo, o_m = ((long long)m_gi * KT + (m_ok ? m_tap : 0)) * HWo;
    const int HWo_i = (int)HWo;
    int b = sl0 / wg.spi, sll = sl0 - b * wg.spi;
    float dy = 0.f, dx = 0.f, mk = 0.f;
    auto fetch = [&](int bb, int sl_local) {
      const int p = sl_local * TC_K + (st & 31);
      dy = dx = mk = 0.f;
      if (m_ok && p < HWo_i) {
        const float *op = offset + (long long)bb * s.off_bs + o_dy + p;
        dy = __ldg(op);
        dx = __ldg(op + HWo);
        mk = __ldg(mask + (long long)bb * s.mask_bs + o_m + p);
      }
    };
    if (sl0 < sl1) fetch(b, sll);
    int ms = 0, mph = 0;
    for (int sl = sl0; sl < sl1; ++sl) {
      const float cdy = dy, cdx = dx, cm = mk;
      const int p = sll * TC_K + (st & 31);
      int nb_ = b, nsl = sll + 1;
      if (nsl == wg.spi) { nsl = 0; ++nb_; }
      if (sl + 1 < sl1) fetch(nb_, nsl);
      BwdMetaTc mt;
      mt.lh = mt.lw = mt.m = 0.f;
      mt.packed = 0;
      if (m_ok && p < HWo_i) {
        const int ho = wg.wo_magic ? (int)__umulhi((unsigned)p, wg.wo_magic) : p / s.Wo;
        const int wo = p - ho * s.Wo;
        const float h_im = (float)(ho * s.sh + hb) + cdy, w_im = (float)(wo * s.sw + wb) + cdx;
        if (h_im > -1.f && w_im > -1.f && h_im < (float)s.H && w_im < (float)s.W) {     // dcn_v2_im2col_cuda.cu:165
          const float hf = floorf(h_im), wf = floorf(w_im);
          const int hl = (int)hf, wl = (int)wf;
          int flags = 16;
          if (hl >= 0 && wl >= 0) flags |= 1;
          if (hl >= 0 && wl + 1 <= s.W - 1) flags |= 2;
          if (hl + 1 <= s.H - 1 && wl >= 0) flags |= 4;
          if (hl + 1 <= s.H - 1 && wl + 1 <= s.W - 1) flags |= 8;
          mt.lh = h_im - hf;
          mt.lw = w_im - wf;
          mt.m = cm;
          mt.packed = ((hl * s.W + wl + s.W + 1) << 7) | flags;
        }
      }
      mbar_wait(&m_empty[ms], (uint32_t)(mph ^ 1));
      wmeta[ms][st] = mt;
      __syncwarp();
      if (lane == 0) mbar_arrive_tc(&m_full[ms]);
      if (++ms == WG_MR) { ms = 0; mph ^= 1; }
      b = nb_;
      sll = nsl;
    }
    // epilogue: TMEM lane = row (chunk, channel), columns = output channels -> wpart[split][rg][row][co_r]
    const int qw = warp;
    float *dst = wpart + (((size_t)split * g.RG + rg) * 128 + qw * 32 + lane) * wg.co_r;
    if (sl1 > sl0) {
      mbar_wait(&tm_full, 0u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const uint32_t taddr = tm + ((uint32_t)(qw * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < wg.co_r; c0 += 32) {
      uint32_t v[32];
      if (sl1 > sl0) {
        tmem_ld32(taddr + (uint32_t)c0, v);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        if (c0 + i < wg.co_r)
          *reinterpret_cast<float4 *>(dst + c0 + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                                                   __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == TC_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(256));
}

// gw[o][ch][tap] += sum over splits (ascending) of wpart[split][rg][row][o]
__global__ void __launch_bounds__(256) k_dcn_bwd_wreduce(const float *__restrict__ wpart, float *__restrict__ gw,
                                                         const DcnShapeTc s, const BwdGeom g, const WgGeom wg) {
  const int KT = s.kh * s.kw, cpg = s.Cin / s.dg;
  const long long total = (long long)g.RG * 128 * wg.co_r;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i % wg.co_r);
    const int row = (int)((i / wg.co_r) % 128), rg = (int)(i / ((long long)wg.co_r * 128));
    const int q = rg * 4 + (row >> 5), ch = row & 31;
    const int bi = q / TC_NT, tap = q - bi * TC_NT;
    const int gi = bi / s.cbs_pg, cw = (bi - gi * s.cbs_pg) * TC_CB + ch;
    if (q >= g.Q || o >= s.Cout || cw >= cpg || tap >= KT) continue;
    float acc = 0.f;
    for (int sp = 0; sp < wg.splits; ++sp) acc += __ldg(wpart + (size_t)sp * total + i);
    gw[((size_t)o * s.Cin + gi * cpg + cw) * KT + tap] += acc;
  }
}

// ------------------------------------------------------------------ host
std::atomic<int> g_dcn_deterministic{0};
static size_t bw_xt_bytes(const DcnShapeTc &s) { return align_up((size_t)s.B * s.H * s.W * s.nb * TC_CB * 4, 256); }
static size_t bw_wtt_bytes(const BwdGeom &g) { return align_up((size_t)g.RG * g.KS * BW_TILE_FLOATS * 4, 256); }
static size_t bw_gyt_bytes_per_image(const BwdGeom &g) { return (size_t)g.tiles * g.KS * BW_TILE_FLOATS * 4; }
static size_t bw_dcol_bytes_per_image(const DcnShapeTc &s, const BwdGeom &g) {
  return align_up((size_t)s.Ho * s.Wo * g.Qp * TC_CB * 4, 256);
}
// images per pass: the column gradient of a pass is held to ~1.5 GiB (B=16 of the largest dla_34 layer is 0.67 GB)
static int bw_batch_chunk(const DcnShapeTc &s, const BwdGeom &g) {
  const size_t per = bw_dcol_bytes_per_image(s, g) + bw_gyt_bytes_per_image(g);
  long long bc = (long long)((1536ull << 20) / per);
  if (bc < 1) bc = 1;
  if (bc > s.B) bc = s.B;
  return (int)bc;
}

static size_t bw_gyk_bytes(const WgGeom &wg) { return align_up((size_t)wg.n_slices * 2 * wg.co_r * TC_K * 4, 256); }
static size_t bw_wpart_bytes(const BwdGeom &g, const WgGeom &wg) {
  return align_up((size_t)wg.splits * g.RG * 128 * wg.co_r * 4, 256);
}
static size_t bw_bpart_bytes(const WgGeom &wg) { return align_up((size_t)wg.n_slices * wg.co_r * 4, 256); }

// [xt][dxt][wtt][ max( gyt + dcol of one batch chunk , gyk + wpart ) ]: the data and the weight passes run one after
// the other on the stream and share the tail
size_t dcn_tc_bwd_workspace_bytes(int b, int cin, int h, int w, int cout, int kh, int kw, int sh, int ph, int dh, int dg) {
  DcnShapeTc s;
  fill_shape(&s, b, cin, h, w, cout, kh, kw, sh, sh, ph, ph, dh, dh, dg);
  const BwdGeom g = bwd_geom(s);
  // the kernels index within one image with 32 bits (and pack hl*W+wl into 24): larger images use the fp32 path
  const unsigned long long lim = 1ull << 31;
  if ((unsigned long long)(h + 2) * (w + 2) >= (1ull << 24) || (unsigned long long)(h + 2) * (w + 2) * s.nb * TC_CB >= lim ||
      (unsigned long long)s.Ho * s.Wo * g.Qp * TC_CB >= lim || (unsigned long long)s.Ho * s.Wo * dg * 2 * kh * kw >= lim)
    return 0;
  const WgGeom wg = wg_geom(s, g);
  const int bc = bw_batch_chunk(s, g);
  const size_t data = align_up((size_t)bc * bw_gyt_bytes_per_image(g), 256) + (size_t)bc * bw_dcol_bytes_per_image(s, g);
  const size_t wgt = bw_gyk_bytes(wg) + bw_wpart_bytes(g, wg) + bw_bpart_bytes(wg);
  return 2 * bw_xt_bytes(s) + bw_wtt_bytes(g) + (data > wgt ? data : wgt);
}

// dX, dOffset, dMask (each nullable; accumulated into, as cnb_dcnv2_backward documents).
int dcn_backward_data_tc(const float *input, const float *offset, const float *mask, const float *weight,
                         const float *grad_output, float *grad_input, float *grad_offset, float *grad_mask, int b,
                         int cin, int h, int w, int cout, int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw,
                         int dg, void *workspace, cudaStream_t stream, int xt_ready, int layout) {
  // layout bit 0: `input` already is channels-last [b][h*w][cin]; bit 1: so is grad_input (both need cin/dg % 32 == 0,
  // checked by the caller): the layout passes and the channels-last scratch copies are skipped
  DcnShapeTc s;
  fill_shape(&s, b, cin, h, w, cout, kh, kw, sh, sw, ph, pw, dh, dw, dg);
  const BwdGeom g = bwd_geom(s);
  const int bc = bw_batch_chunk(s, g);
  char *wp = reinterpret_cast<char *>(workspace);
  float *xt = reinterpret_cast<float *>(wp);  wp += bw_xt_bytes(s);
  float *dxt = reinterpret_cast<float *>(wp); wp += bw_xt_bytes(s);
  const float *xsrc = (layout & 1) ? input : xt;
  if (layout & 2) dxt = grad_input;
  float *wtt = reinterpret_cast<float *>(wp); wp += bw_wtt_bytes(g);
  float *gyt = reinterpret_cast<float *>(wp); wp += align_up((size_t)bc * bw_gyt_bytes_per_image(g), 256);
  float *dcol = reinterpret_cast<float *>(wp);
  const long long HW = (long long)h * w, HWo = (long long)s.Ho * s.Wo;

  int rc = CNB_OK;
  if (!xt_ready && !(layout & 1)) rc = dcn_to_channels_last(input, xt, s, stream);
  if (rc != CNB_OK) return rc;
  // default: every sample is scattered into a channels-last dX (16-byte vector reductions) and added to grad_input at
  // the end; deterministic mode (cnb_dcnv2_set_deterministic, stride 1): near samples by the gather, the rest by atomics
  // straight into grad_input
  const bool gather = grad_input && sh == 1 && sw == 1 && w >= 2 && g_dcn_deterministic.load(std::memory_order_relaxed) != 0;
  if (grad_input && !gather && !(layout & 2)) CNB_CUDA(cudaMemsetAsync(dxt, 0, (size_t)b * HW * s.nb * TC_CB * 4, stream));
  k_bwd_prep_w<<<g.RG * g.KS, 256, 0, stream>>>(weight, s, g, wtt);
  CNB_CHECK_LAUNCH("cnb_dcnv2_backward weight tiles");
  count_launch();
  const size_t smem = (size_t)BW_STAGES * BW_STAGE_BYTES + 4 * 4096;     // operand ring + the epilogue warps' staging
  static thread_local int dev_done = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev_done != dev) {
    CNB_CUDA(cudaFuncSetAttribute(k_dcn_bwd_dcol_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dev_done = dev;
  }
  for (int b0 = 0; b0 < b; b0 += bc) {
    const int nbimg = (b - b0 < bc) ? b - b0 : bc;
    dim3 pgrid((unsigned)g.tiles, (unsigned)g.KS, (unsigned)nbimg);
    k_bwd_prep_gy<<<pgrid, 256, 0, stream>>>(grad_output + (long long)b0 * cout * HWo, s, g, gyt);
    CNB_CHECK_LAUNCH("cnb_dcnv2_backward grad_output tiles");
    const int n_items = nbimg * g.tiles * g.RG;
    const int grid = n_items < num_sms() ? n_items : num_sms();
    k_dcn_bwd_dcol_tc<<<grid, BW_THREADS, smem, stream>>>(gyt, wtt, dcol, s, g, n_items);
    CNB_CHECK_LAUNCH("cnb_dcnv2_backward column gradient (tcgen05)");
    k_dcn_bwd_offmask<3><<<nbimg * g.tiles, 256, 0, stream>>>(xsrc, offset, mask, dcol, (grad_input && !gather) ? dxt : nullptr,
                                                              grad_offset, grad_mask, gather ? grad_input : nullptr, s, g, b0);
    CNB_CHECK_LAUNCH("cnb_dcnv2_backward offset/mask gradient");
    count_launch(3);
    if (gather) {
      const int ptx = (w + GX_TW - 1) / GX_TW, pty = (h + GX_TH - 1) / GX_TH;
      const int ncb = s.cbs_pg >= 2 ? 2 : 1;
      const size_t gsm = sizeof(BwdMetaTc) * TC_NT * GX_WH * GX_WW + (size_t)ncb * TC_CB * (GX_TH * GX_TW + 1) * 4 +
                         (size_t)8 * GX_LIST * sizeof(float2);
      static thread_local int gdev_done = -1;
      if (gdev_done != dev) {
        CNB_CUDA(cudaFuncSetAttribute(k_dcn_bwd_dx_gather<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        CNB_CUDA(cudaFuncSetAttribute(k_dcn_bwd_dx_gather<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        gdev_done = dev;
      }
      if (ncb == 2)
        k_dcn_bwd_dx_gather<2><<<nbimg * ptx * pty, 256, gsm, stream>>>(offset, mask, dcol, grad_input, s, g, b0, ptx, pty);
      else
        k_dcn_bwd_dx_gather<1><<<nbimg * ptx * pty, 256, gsm, stream>>>(offset, mask, dcol, grad_input, s, g, b0, ptx, pty);
      CNB_CHECK_LAUNCH("cnb_dcnv2_backward input gradient (gather)");
      count_launch();
    }
  }
  if (grad_input && !gather && !(layout & 2)) {
    dim3 tgrid((unsigned)((HW + 31) / 32), (unsigned)((cin + 63) / 64), (unsigned)b);
    k_dcn_bwd_dx_nchw<<<tgrid, 256, 0, stream>>>(dxt, grad_input, s);
    CNB_CHECK_LAUNCH("cnb_dcnv2_backward input gradient");
    count_launch();
  }
  return CNB_OK;
}

// dW and dBias (accumulated into grad_weight / grad_bias; grad_bias nullable).  xt_ready != 0: the channels-last input copy at the head of the workspace was
// already written by dcn_backward_data_tc on this stream.
int dcn_backward_weight_tc(const float *input, const float *offset, const float *mask, const float *grad_output,
                           float *grad_weight, float *grad_bias, int b, int cin, int h, int w, int cout, int kh, int kw, int sh, int sw,
                           int ph, int pw, int dh, int dw, int dg, void *workspace, cudaStream_t stream, int xt_ready,
                           int layout) {
  DcnShapeTc s;
  fill_shape(&s, b, cin, h, w, cout, kh, kw, sh, sw, ph, pw, dh, dw, dg);
  const BwdGeom g = bwd_geom(s);
  const WgGeom wg = wg_geom(s, g);
  char *wp = reinterpret_cast<char *>(workspace);
  float *xt = reinterpret_cast<float *>(wp);  wp += 2 * bw_xt_bytes(s) + bw_wtt_bytes(g);
  float *gyk = reinterpret_cast<float *>(wp); wp += bw_gyk_bytes(wg);
  float *wpart = reinterpret_cast<float *>(wp); wp += bw_wpart_bytes(g, wg);
  float *bpart = reinterpret_cast<float *>(wp);
  const float *xsrc = (layout & 1) ? input : xt;
  if (!xt_ready && !(layout & 1)) {
    const int rc = dcn_to_channels_last(input, xt, s, stream);
    if (rc != CNB_OK) return rc;
  }
  dim3 pgrid((unsigned)wg.spi, (unsigned)b);
  k_bwd_prep_gyk<<<pgrid, 256, 0, stream>>>(grad_output, s, wg, gyk, grad_bias ? bpart : nullptr);
  CNB_CHECK_LAUNCH("cnb_dcnv2_backward grad_output tiles (weight pass)");
  const size_t smem = (size_t)wg.stages * wg.stage_bytes;
  static thread_local int dev_done = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev_done != dev) {
    CNB_CUDA(cudaFuncSetAttribute(k_dcn_bwd_weight_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 66048));
    dev_done = dev;
  }
  k_dcn_bwd_weight_tc<<<g.RG * wg.splits, WG_THREADS, smem, stream>>>(xsrc, offset, mask, gyk, wpart, s, g, wg);
  CNB_CHECK_LAUNCH("cnb_dcnv2_backward weight gradient (tcgen05)");
  const long long total = (long long)g.RG * 128 * wg.co_r;
  k_dcn_bwd_wreduce<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(wpart, grad_weight, s, g, wg);
  CNB_CHECK_LAUNCH("cnb_dcnv2_backward weight gradient reduce");
  count_launch(3);
  if (grad_bias) {
    k_dcn_bwd_bias_reduce<<<cout, 256, 0, stream>>>(bpart, grad_bias, wg);
    CNB_CHECK_LAUNCH("cnb_dcnv2_backward bias gradient reduce");
    count_launch();
  }
  return CNB_OK;
}

}  // namespace cnb
