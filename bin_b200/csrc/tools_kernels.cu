// bin_b200 -- measurement tooling: tcgen05 issue-rate microbenchmarks and the role-timeline reader.
// NOT part of the product: compiled only into libbin_b200_tools.so (python -m bin_b200.build --tools), which is the
// product sources built with -DBIN_B200_TOOLS (timeline hooks in the kernels, per-call option re-reads) plus this
// file.  Its entry points are declared in tools_abi.h, not in include/bin_b200.h.
#include "common.cuh"
#include "internal.h"
#include "tools_abi.h"

namespace binb {

// ------------------------------------------------------------------ tcgen05 issue-rate microbenchmark
// mode bits: [0,4) independent accumulators cycled round-robin; [4,6) A layout; [6,8) B layout
// (0 = no-swizzle K-major, 1 = SWIZZLE_128B, 2 = SWIZZLE_64B, 3 = SWIZZLE_32B); bit 8: shift the A
// start by one row per MMA; bit 9: M=64 instead of 128.  Operand contents are zeros (timing only).
__device__ __forceinline__ uint64_t bench_desc(uint32_t addr, int layout, uint32_t noswz_lbo) {
  uint64_t d = (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 46);
  if (layout == 0) {
    d |= (uint64_t)((noswz_lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(128 >> 4) << 32;
  } else {
    const uint32_t sbo = layout == 1 ? 1024 : layout == 2 ? 512 : 256;
    const uint64_t type = layout == 1 ? 2 : layout == 2 ? 4 : 6;
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= type << 61;
  }
  return d;
}
__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int n, int iters, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  if (threadIdx.x < 32) {                     // whole warp converged; one elected lane issues
    const int M = (mode & 0x200) ? 64 : 128;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 48 * 1024);
    int nacc = mode & 0xf;
    if (nacc < 1) nacc = 1;
    if (nacc * n > 512) nacc = 512 / n;
    const int la = (mode >> 4) & 3, lb = (mode >> 6) & 3;
    const bool shift = (mode & 0x100) != 0;
    const uint32_t row16 = (la == 0 ? 16 : la == 1 ? 128 : la == 2 ? 64 : 32) >> 4;
    const uint64_t ad0 = bench_desc(a0, la, 320 * 16);
    const uint64_t bd0 = bench_desc(b0, lb, (uint32_t)n * 16);
    const uint32_t dstep = (nacc > 1) ? (uint32_t)n : 0u;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
      if (elect_one()) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint64_t ad = ad0 + (shift ? (uint64_t)(u * row16) : 0ull);
          umma_f16_ss(tb + (uint32_t)(u % 2) * dstep, ad, bd0, idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tb, 512);
  }
}

// Replays the exact descriptor sequence of the x-stacked RDB conv0 MMA loop (3 resident weight
// chunks, 3 activation stages, 2 accumulators x 3 ky taps x 2 k-steps, N=96) with no TMA and no
// epilogue.  vary bit0: A addresses as in the kernel (else one fixed tile); bit1: B addresses as in the
// kernel (else one fixed slab); bit2: fill smem with non-zero data.
__global__ void __launch_bounds__(128, 1) mma_pattern_kernel(int tiles, int vary, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = (vary & 4) ? 0x3c003800u + (i * 2654435761u & 0x03ff03ffu) : 0u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  if (threadIdx.x < 32) {
    constexpr int NM = 96, A_PLANE = 10 * 32 * 16, W_TAP = 4 * NM * 16, W_CHUNK = 3 * W_TAP, STAGE = 4 * A_PLANE;
    constexpr uint32_t idesc = umma_idesc_f16(128, NM);
    constexpr uint32_t HI = (128u >> 4) | (1u << 14);
    const uint32_t wbase = smem_u32(smem), sbase = smem_u32(smem + 3 * W_CHUNK);
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      for (int c = 0; c < 3; ++c) {
        const uint32_t a_lo = (((sbase + ((vary & 1) ? c * STAGE : 0)) >> 4) & 0x3FFFu) | ((uint32_t)(A_PLANE >> 4) << 16);
        const uint32_t b_lo = (((wbase + ((vary & 2) ? c * W_CHUNK : 0)) >> 4) & 0x3FFFu) | ((uint32_t)(NM * 16 >> 4) << 16);
        if (elect_one()) {
#pragma unroll
          for (int m = 0; m < 2; ++m) {
#pragma unroll
            for (int tp = 0; tp < 3; ++tp) {
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const uint32_t ao = (vary & 1) ? (uint32_t)(m * 128 + tp * 32) + j * 2 * (A_PLANE >> 4) : 0u;
                const uint32_t bo = (vary & 2) ? (uint32_t)(tp * W_TAP + j * 2 * NM * 16) / 16 : 0u;
                umma_f16_ss(tb + (t & 1) * 192 + m * NM, ((uint64_t)HI << 32) | (a_lo + ao),
                            ((uint64_t)HI << 32) | (b_lo + bo), idesc, 1u);
              }
            }
          }
        }
        __syncwarp();
      }
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tb, 512);
  }
}

// 2-CTA (cta_group::2) issue-rate probe: a CTA pair shares one 256 x N x 16 MMA (A: 128 rows from each CTA's smem,
// B: N/2 rows from each), issued by the leader CTA.  Measures cycles per MMA to size the benefit for round 2.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma2_bench_kernel(int n, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  if (rank == 0 && threadIdx.x < 32) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 48 * 1024);
    const uint64_t ad0 = umma_desc_kmajor_noswz(a0, 320 * 16, 128);
    const uint64_t bd0 = umma_desc_kmajor_noswz(b0, (uint32_t)(n / 2) * 16, 128);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
      if (elect_one()) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb + (uint32_t)(u % 2) * (uint32_t)n),
              "l"(ad0 + (uint64_t)u), "l"(bd0), "r"(idesc), "r"(1u)
              : "memory");
        }
      }
      __syncwarp();
    }
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x / 2] = t1 - t0;
  }
  tc_fence_before();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
  }
}

int run_mma_bench(int n, int iters, int mode, float* cycles_host) {
  if (n < 16 || n > 256 || n % 16) return fail(BIN_ERR_ARG, "microbench: N must be a multiple of 16 in [16,256]");
  long long* d = nullptr;
  const int grid = 148;
  BIN_CUDA_OK(cudaMalloc(&d, grid * sizeof(long long)));
  BIN_CUDA_OK(cudaMemset(d, 0, grid * sizeof(long long)));
  if (mode & 0x2000) {                       // 2-CTA pair probe
    BIN_CUDA_OK(cudaFuncSetAttribute(mma2_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    mma2_bench_kernel<<<grid, 128, 96 * 1024>>>(n, iters, d);
  } else if (mode & 0x1000) {                // conv-pattern replay: n = ignored, iters = tiles
    BIN_CUDA_OK(cudaFuncSetAttribute(mma_pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    mma_pattern_kernel<<<grid, 128, 160 * 1024>>>(iters, mode & 7, d);
    iters *= 36;
  } else {
  BIN_CUDA_OK(cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  mma_bench_kernel<<<grid, 128, 96 * 1024>>>(n, iters, mode, d);
  }
  BIN_CUDA_OK(cudaGetLastError());
  BIN_CUDA_OK(cudaDeviceSynchronize());
  long long h[148];
  BIN_CUDA_OK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  cudaFree(d);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  *cycles_host = (float)mx / (float)iters;
  return BIN_OK;
}


}  // namespace binb

extern "C" {
int bin_tools_debug_timeline(long long* host, int n) {
  using namespace binb;
  if (!g_dbg) return fail(BIN_ERR_ARG, "no timeline recorded (set BIN_B200_DEBUG=8)");
  if (n > 3 * 4096) n = 3 * 4096;
  BIN_CUDA_OK(cudaDeviceSynchronize());
  BIN_CUDA_OK(cudaMemcpy(host, g_dbg, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost));
  return BIN_OK;
}
int bin_tools_microbench_mma(int n, int iters, int mode, float* cycles_host) {
  return binb::run_mma_bench(n, iters, mode, cycles_host);
}
}
