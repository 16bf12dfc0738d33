/* bin_b200 measurement tooling ABI (libbin_b200_tools.so only; see tools_kernels.cu).  Not a product interface. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
/* Issue `iters` back-to-back tcgen05.mma (M=128, N=n, K=16, fp16) from one CTA per SM and return cycles per MMA in
 * *cycles_host (host pointer; synchronises).  mode: see tools_kernels.cu. */
int bin_tools_microbench_mma(int n, int iters, int mode, float* cycles_host);
/* With env BIN_B200_DEBUG=8 block 0 of the conv / rdb_tail kernels records clock64 at role milestones
 * ([role][iter][k] as 3x1024x4 int64); copies the last launch's timeline to the host. */
int bin_tools_debug_timeline(long long* host, int n);
#ifdef __cplusplus
}
#endif
