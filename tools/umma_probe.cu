// Micro-probe: sustained tcgen05.mma throughput of one CTA per SM with fixed shared-memory operands, no TMA, no
// epilogue.  Answers "what can kind::f16 cta_group::1 M=128 reach per clock for N = 64 / 128 / 256, and does the
// number of accumulators in rotation or the operand start alignment matter?"  Prints cycles per MMA and the fraction
// of the nominal 4096 MAC/clk/SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);

__global__ void __launch_bounds__(128, 1) probe(int N, int iters, int n_acc, int a_shift_rows, int n_issuers, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_tile = smem;                 // 144 rows x 128 B (room for shifted starts)
  uint8_t* b_tile = smem + 20 * 1024;     // 256 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 20 * 1024 + 32 * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (20 + 32) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // half 1.0
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(n_issuers) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  long long t0 = 0, t1 = 0;
  const int warp = threadIdx.x >> 5;
  if (warp < n_issuers) {
    // warp-uniform loop, one elected lane issues: addresses stay in uniform registers (1-2 SASS instructions per MMA)
    const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (8u << 24);
    const uint32_t a_lo = (((smem_u32(a_tile) + a_shift_rows * 128) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b_lo = ((smem_u32(b_tile) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint64_t hi = static_cast<uint64_t>(kDescHi) << 32;
    const uint32_t d_base = tmem + static_cast<uint32_t>(warp * n_acc * N);
    t0 = clock64();
    int acc = 0;
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = d_base + static_cast<uint32_t>(acc * N);
      uint32_t pred;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
      if (pred) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = hi | (a_lo + 2 * k), db = hi | (b_lo + 2 * k);
          asm volatile(
              "{\n\t.reg .pred p;\n\t"
              "setp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u)
              : "memory");
        }
      }
      __syncwarp();
      if (++acc == n_acc) acc = 0;
    }
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    if (pred) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    __syncwarp();
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(smem_u32(bar))
          : "memory");
    }
    t1 = clock64();
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

int main() {
  const int smem = 1024 + (20 + 32) * 1024 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  const int iters = 20000;
  printf("| grid | issuer warps | N | accumulators / issuer | A start row | cycles / MMA (K=16, per SM) | MAC/clk/SM | of 4096 |\n|---|---|---|---|---|---:|---:|---:|\n");
  const int cfgs[][5] = {{148, 1, 64, 1, 0}, {148, 1, 64, 4, 0}, {148, 1, 64, 4, 3}, {148, 2, 64, 2, 0}, {148, 2, 64, 4, 0}, {148, 1, 128, 2, 0},
                         {148, 2, 128, 2, 0}, {148, 1, 256, 2, 0}, {148, 2, 256, 1, 0}, {1, 2, 64, 2, 0},
                         // round 2: the widths the stacked-tap kernels issue (N = 16 x rows fed, up to 112; 192 for the triple group)
                         {148, 1, 16, 1, 0}, {148, 1, 32, 1, 0}, {148, 1, 48, 1, 0}, {148, 1, 80, 1, 0}, {148, 1, 96, 1, 0}, {148, 1, 112, 1, 0},
                         {148, 1, 112, 1, 3}, {148, 1, 112, 4, 0}, {148, 1, 192, 1, 0}, {148, 1, 192, 2, 0}};
  for (auto& c : cfgs) {
    if (c[1] * c[2] * c[3] > 512) continue;
    probe<<<c[0], 128, smem>>>(c[2], iters, c[3], c[4], c[1], d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148];
    cudaMemcpy(h, d, c[0] * sizeof(long long), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < c[0]; ++i) mx = h[i] > mx ? h[i] : mx;
    const double per = mx / (iters * 4.0 * c[1]);
    const double macs = 128.0 * c[2] * 16 / per;
    printf("| %d | %d | %d | %d | %d | %.1f | %.0f | %.1f %% |\n", c[0], c[1], c[2], c[3], c[4], per, macs, 100 * macs / 4096);
  }
  return 0;
}
