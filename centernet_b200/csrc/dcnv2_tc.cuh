// Shared by the tcgen05 DCNv2 kernels (dcnv2_tc.cu: forward, dcnv2_bwd_tc.cu: backward): the UMMA tile
// layout (K-major, no swizzle), the instruction descriptor pieces and the shape / work decomposition.
#pragma once
#include "common.cuh"

namespace cnb {

constexpr int TC_TP = 128;        // pixels per item (UMMA M)
constexpr int TC_CB = 32;         // input channels per chunk = one 128-byte line of the channels-last copy
constexpr int TC_NT = 9;          // taps (3x3 kernel; fewer taps leave zero weight tiles)
constexpr int TC_K = TC_CB;       // K per chunk
constexpr int TC_KC = TC_K / 4;   // 16-byte k-chunks
constexpr int TC_EPI_WARPS = 4;   // warps 0-3: TMEM lane quarter = warp index
constexpr int TC_WARP_MMA = 4;
constexpr int TC_WARP_TMA = 5;
constexpr int TC_WARP_S0 = 6;     // first sampler warp
constexpr int TC_SAMPLERS = 16;   // sampler warps: 2 warp items (4 pixels each) per warp and chunk
constexpr int TC_THREADS = (TC_WARP_S0 + TC_SAMPLERS) * 32;   // 704
constexpr uint32_t TC_LBO = 128;            // bytes between consecutive k-chunks (one 8 x 16 B core matrix)
constexpr uint32_t TC_SBO = TC_KC * 128;    // bytes between 8-row groups
constexpr int TC_A_BYTES = TC_TP * TC_K * 4;  // 16384 per hi / lo tile

// A tile of the forward sampler: the same K-major core-matrix layout with the k-chunks pitched 144 bytes apart
constexpr uint32_t FW_LBO = 144;                    // bytes between consecutive k-chunks (8 rows x 16 B + 16 B of padding)
constexpr uint32_t FW_SBO = TC_KC * FW_LBO;         // bytes between 8-row groups (1152)
constexpr int FW_A_BYTES = (TC_TP / 8) * FW_SBO;    // 18432 per hi / lo tile

struct DcnShapeTc {
  int B, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, dg, Ho, Wo;
  int co_t;       // output channels per item (64 or 128)
  int n_cot;      // Cout tiles
  int cbs_pg;     // 32-channel blocks per deformable group
  int nb;         // channel blocks in total = dg * cbs_pg
  int tw, th;     // pixel tile (tw * th == 128)
  int tiles_x, tiles_y;
  int splits;     // K splits (channel-block ranges) per item; > 1 -> partial sums + k_dcn_reduce
  int n_items;    // B * tiles * n_cot * splits
  // fused prologue / epilogue (N3, dcn_v2.py:64-70 + pose_dla_dcn.py:345-357)
  long long off_bs, mask_bs;   // batch strides (floats) of the offset / mask tensors
  int mask_logit;              // mask holds the raw conv_offset_mask output: sigmoid applied on the fly
  int relu;                    // epilogue: y = relu(scale[o] * (acc + bias[o]) + shift[o])
  const float *epi_scale, *epi_shift;   // folded inference BatchNorm (nullable)
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
__device__ __forceinline__ uint64_t fw_desc_a(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((FW_LBO >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((FW_SBO >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ void mbar_arrive_tc(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

inline void fill_shape(DcnShapeTc *s, int b, int cin, int h, int w, int cout, int kh, int kw, int sh, int sw, int ph,
                       int pw, int dh, int dw, int dg) {
  s->B = b; s->Cin = cin; s->H = h; s->W = w; s->Cout = cout; s->kh = kh; s->kw = kw; s->sh = sh; s->sw = sw;
  s->ph = ph; s->pw = pw; s->dh = dh; s->dw = dw; s->dg = dg;
  s->Ho = (h + 2 * ph - (dh * (kh - 1) + 1)) / sh + 1;
  s->Wo = (w + 2 * pw - (dw * (kw - 1) + 1)) / sw + 1;
  s->co_t = cout > 64 ? 128 : 64;
  s->n_cot = (cout + s->co_t - 1) / s->co_t;
  const int cpg = cin / dg;
  s->cbs_pg = (cpg + TC_CB - 1) / TC_CB;
  s->nb = dg * s->cbs_pg;
  s->tw = s->Wo >= 12 ? 16 : 8;          // 8 x 16 pixel tiles; 16 x 8 on narrow maps
  s->th = TC_TP / s->tw;
  s->tiles_x = (s->Wo + s->tw - 1) / s->tw;
  s->tiles_y = (s->Ho + s->th - 1) / s->th;
  const long long base_items = (long long)b * s->tiles_x * s->tiles_y * s->n_cot;
  int splits = 1;
  const int sms = num_sms();
  if (base_items < sms) {                 // small maps: split the channel blocks so every SM gets an item
    splits = (int)((sms + base_items - 1) / base_items);
    if (splits > s->nb) splits = s->nb;
    if (splits > 8) splits = 8;
    if (splits < 1) splits = 1;
  }
  s->splits = splits;
  s->n_items = (int)(base_items * splits);
  s->off_bs = (long long)dg * 2 * kh * kw * s->Ho * s->Wo;
  s->mask_bs = (long long)dg * kh * kw * s->Ho * s->Wo;
  s->mask_logit = 0;
  s->relu = 0;
  s->epi_scale = s->epi_shift = nullptr;
}

// x [B][Cin][HW] -> channels-last copy xt [B][HW][nb][32] (dcnv2_tc.cu)
int dcn_to_channels_last(const float *x, float *xt, const DcnShapeTc &s, cudaStream_t stream);

}  // namespace cnb
