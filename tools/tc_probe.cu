// Probe for the tcgen05 (UMMA) path used by the DCNv2 contraction: D[128 x N] = A[128 x K] * B[N x K]^T
// with kind::tf32, both operands K-major in shared memory in the canonical no-swizzle ("interleave")
// layout, accumulator in TMEM.  Small integers make the result exact.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tc_probe tools/tc_probe.cu && /tmp/tc_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 72;           // first chunk shape of the DCN kernel (8 channels x 9 taps); production now uses
                                                 // K = 32 per chunk with the same descriptor fields (LBO 128 B, SBO = K/4 * 128 B)
constexpr int KC = K / 4;                        // 16-byte k-chunks (4 tf32 each)
constexpr uint32_t LBO = 128;                    // bytes between consecutive k-chunks (one 8x16B core matrix)
constexpr uint32_t SBO_A = KC * 128;             // bytes between 8-row groups
constexpr uint32_t SBO_B = KC * 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // layout_type = 0 (no swizzle), base_offset = 0
}
__device__ __forceinline__ uint32_t tile_off(int row, int k, uint32_t sbo) {  // byte offset of element (row, k)
  return (uint32_t)(row >> 3) * sbo + (uint32_t)(k >> 2) * LBO + (uint32_t)(row & 7) * 16u + (uint32_t)(k & 3) * 4u;
}

__global__ void __launch_bounds__(128) probe(const float *A, const float *B, float *D) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *sa = smem;                       // 128 x 72 x 4 = 36864
  unsigned char *sb = smem + 36864;               // 64 x 72 x 4 = 18432
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < M * K; i += 128) { int r = i / K, k = i % K; *(float *)(sa + tile_off(r, k, SBO_A)) = A[i]; }
  for (int i = tid; i < N * K; i += 128) { int r = i / K, k = i % K; *(float *)(sb + tile_off(r, k, SBO_B)) = B[i]; }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic smem writes -> async proxy (UMMA reads)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int ks = 0; ks < K / 8; ++ks) {          // one MMA per 8 k (two 16-byte k-chunks)
      const uint64_t da = make_desc(smem_u32(sa) + ks * 2 * LBO, LBO, SBO_A);
      const uint64_t db = make_desc(smem_u32(sb) + ks * 2 * LBO, LBO, SBO_B);
      const uint32_t acc = ks > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  // wait for the MMAs
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: warp w reads TMEM lanes [32w, 32w+32), 64 columns (two x32 loads)
  uint32_t v[64];
  const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
#define LD32(off, base)                                                                                      \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                     \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24," \
               "%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                        \
               : "=r"(v[base + 0]), "=r"(v[base + 1]), "=r"(v[base + 2]), "=r"(v[base + 3]), "=r"(v[base + 4]),  \
                 "=r"(v[base + 5]), "=r"(v[base + 6]), "=r"(v[base + 7]), "=r"(v[base + 8]), "=r"(v[base + 9]),  \
                 "=r"(v[base + 10]), "=r"(v[base + 11]), "=r"(v[base + 12]), "=r"(v[base + 13]),                  \
                 "=r"(v[base + 14]), "=r"(v[base + 15]), "=r"(v[base + 16]), "=r"(v[base + 17]),                  \
                 "=r"(v[base + 18]), "=r"(v[base + 19]), "=r"(v[base + 20]), "=r"(v[base + 21]),                  \
                 "=r"(v[base + 22]), "=r"(v[base + 23]), "=r"(v[base + 24]), "=r"(v[base + 25]),                  \
                 "=r"(v[base + 26]), "=r"(v[base + 27]), "=r"(v[base + 28]), "=r"(v[base + 29]),                  \
                 "=r"(v[base + 30]), "=r"(v[base + 31])                                                           \
               : "r"(taddr + off))
  LD32(0, 0);
  LD32(32, 32);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int n = 0; n < N; ++n) D[(size_t)tid * N + n] = __uint_as_float(v[n]);   // thread = TMEM lane = row m
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tm));
}

int main() {
  std::vector<float> A(M * K), B(N * K), D(M * N), R(M * N, 0.f);
  srand(1);
  for (auto &x : A) x = (float)(rand() % 15 - 7);
  for (auto &x : B) x = (float)(rand() % 9 - 4) * 0.5f;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
      R[m * N + n] = s;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 36864 + 18432);
  probe<<<1, 128, 36864 + 18432>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  int bad = 0;
  for (int i = 0; i < M * N; ++i) {
    double d = fabs((double)D[i] - R[i]);
    if (d > maxerr) maxerr = d;
    if (d > 1e-3 && bad < 5) { printf("mismatch m=%d n=%d got %f want %f\n", i / N, i % N, D[i], R[i]); ++bad; }
  }
  printf("max abs err %g  (D[0]=%f R[0]=%f, D[last]=%f R[last]=%f)\n", maxerr, D[0], R[0], D[M * N - 1], R[M * N - 1]);
  return maxerr < 1e-3 ? 0 : 1;
}
