#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// probe: fill TMEM [128 lanes][64 cols] with lane*1000+col, issue `nshift` tcgen05.shift at (lane_off, col_off), read back
__global__ void __launch_bounds__(128, 1) shift_probe(int lane_off, int col_off, int nshift, float* out) {
  __shared__ uint32_t tbase_s;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase_s)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tbase_s;
  const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < 64; c0 += 8) {
    uint32_t v[8];
    for (int i = 0; i < 8; ++i) v[i] = __float_as_uint((float)(threadIdx.x * 1000 + c0 + i));
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr + c0), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int i = 0; i < nshift; ++i)
      asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(tbase + ((uint32_t)lane_off << 16) + col_off) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait
  {
    uint32_t ok = 0; int spins = 0;
    while (!ok && ++spins < 1000000) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
    if (!ok && lane == 0) printf("timeout warp %d\n", warp);
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < 64; c0 += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]),
                 "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[threadIdx.x * 64 + c0 + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(64) : "memory");
}
int main(int argc, char** argv) {
  float* d; cudaMalloc(&d, 128 * 64 * 4);
  static float h[128 * 64];
  int cfgs[][3] = {{0, 0, 1}, {0, 8, 1}, {32, 16, 1}, {0, 4, 1}, {0, 0, 2}, {64, 32, 3}};
  for (auto& c : cfgs) {
    shift_probe<<<1, 128>>>(c[0], c[1], c[2], d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cfg lane %d col %d n %d: %s\n", c[0], c[1], c[2], cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("== shift at lane %d col %d x%d: changed cells (lane,col: got = src lane/col)\n", c[0], c[1], c[2]);
    int nch = 0, cmin = 999, cmax = -1, lmin = 999, lmax = -1;
    for (int l = 0; l < 128; ++l) for (int cc = 0; cc < 64; ++cc) {
      float exp = l * 1000 + cc, got = h[l * 64 + cc];
      if (got != exp) { ++nch; if (cc < cmin) cmin = cc; if (cc > cmax) cmax = cc; if (l < lmin) lmin = l; if (l > lmax) lmax = l;
        if (nch <= 6 || (l % 32 == 0 && cc == cmin) || (l % 32 == 31 && cc == cmin)) printf("  (%d,%d) got %.0f -> src lane %d col %d\n", l, cc, got, (int)got / 1000, (int)got % 1000); }
    }
    printf("  total changed %d, lanes [%d,%d], cols [%d,%d]\n", nch, lmin, lmax, cmin, cmax);
  }
  return 0;
}
