// tcgen05 implicit-GEMM convolution for sm_100a -- replaces cuDNN's fp32 `cudnnConvolutionForward`
// (caffe/src/caffe/layers/cudnn_conv_layer.cu:21-37) for the SegNet convolutions with 64 input channels
// (all of Basic except conv1 / the classifier; the 64-channel layers of Standard).
//
// GEMM view per CTA:  D[128 pixels x N couts] += A[128 pixels x 64 cin] * B[64 cin x N couts]  per filter tap,
// M = 128 consecutive pixels of one output row, two output rows (R = 2) per accumulator stage.
//
//  * Activations are NHWC half, so one pixel's 64 input channels are one 128-byte row: exactly a
//    SWIZZLE_128B K-major UMMA operand row.  A TMA box {64 ch, 128+K-1 px, 1 row} of the input lands one
//    *halo row* in shared memory; out-of-image coordinates are zero-filled by TMA = the conv's zero padding.
//  * The A operand of tap (kh, kw) for output row r is the same halo row shifted by kw pixels: the UMMA
//    shared-memory descriptor simply starts kw*128 bytes later (the swizzle phase follows the absolute address), so
//    each input row is fetched from L2 once per K rows of output instead of K*K times.
//  * A CTA walks down a 128-px-wide column strip two output rows at a time with a ring of halo rows:
//    every input row is loaded once per CTA (plus the K-1 overlap between vertically adjacent CTAs).
//  * Weights [tap][cout][cin] stream through a 4-stage TMA ring, one {64 cin x N cout} tile per tap.
//  * Accumulators live in TMEM (2 stages x R x N fp32 columns) so the epilogue of row pair j overlaps the
//    MMAs of pair j+1.  Warp roles: 0 = halo-row TMA producer, 1 = weight TMA producer, 2 = MMA issuer
//    (+ TMEM alloc), 3..6 = epilogue (tcgen05.ld -> bias / BN affine / ReLU / dropout -> half -> global).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <mutex>

#include "philox.cuh"
#include "segnet.h"

namespace sivo {

namespace {

constexpr int kTcThreads = 7 * 32;
constexpr int kRows = 2;               // output rows per accumulator stage
constexpr int kSlotBytes = 17 * 1024;  // one halo row: (128 + K - 1) px * 128 B, padded to a 1024-B multiple
constexpr int kBStages = 4;

struct TcParams {
  int H, W, N_batch;       // spatial size and batch of input == output
  int cout_total;          // channel stride of the output tensor
  int n_tile;              // UMMA N (64 or 128)
  int pairs_per_cta;       // row pairs one CTA walks
  int strips;              // ceil(W / 128)
  int relu, has_bn, has_drop;
  float slope;
  const float* bias;
  const float* bn_scale;
  const float* bn_shift;
  __half* out;
  // dropout fused into the epilogue (decoder convs of Basic: decdrop4 / decdrop3)
  uint64_t seed;
  const uint64_t* frame;
  int drop_layer;
  float drop_scale;
  int bo_mode;             // 0 (default): descriptor base_offset 0; 1: (addr >> 7) & 7 (experiment switch, wrong on B200)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug must surface as a CUDA error, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spin > (1u << 26)) __trap();
  }
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// K-major SWIZZLE_128B operand descriptor: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, int bo_mode) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);        // start address
  d |= static_cast<uint64_t>(1) << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                // stride byte offset
  d |= static_cast<uint64_t>(1) << 46;                        // descriptor version (sm_100)
  // base_offset (bits 49-51) stays 0: measured on B200, the swizzle phase comes from the absolute shared-memory
  // address bits [7:9], so a start address kw*128 B into a 1024-B-aligned row needs no correction
  // (bo_mode 1 = experiment switch that sets (addr >> 7) & 7; it produces wrong results -- profiles/r1_notes.md).
  if (bo_mode == 1) d |= static_cast<uint64_t>((saddr >> 7) & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;                        // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
        "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
        "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int K>
__global__ void __launch_bounds__(kTcThreads, 1)
k_conv_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams p) {
  constexpr int kSlots = K + kRows - 1 + kRows;  // rows live for one pair + the next pair's new rows
  constexpr int kPad = (K - 1) / 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_slots = smem;
  uint8_t* b_stages = smem + kSlots * kSlotBytes;
  const int b_bytes = p.n_tile * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_stages + kBStages * b_bytes);
  uint64_t* a_full = bars;                       // [kSlots]
  uint64_t* a_empty = a_full + kSlots;           // [kSlots]
  uint64_t* b_full = a_empty + kSlots;           // [kBStages]
  uint64_t* b_empty = b_full + kBStages;         // [kBStages]
  uint64_t* t_full = b_empty + kBStages;         // [2]
  uint64_t* t_empty = t_full + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int strip = blockIdx.x % p.strips, rowblk = blockIdx.x / p.strips;
  const int n0 = blockIdx.y * p.n_tile;
  const int img = blockIdx.z;
  const int x0 = strip * 128;
  const int total_pairs = (p.H + kRows - 1) / kRows;
  const int pair0 = rowblk * p.pairs_per_cta;
  const int npairs = min(p.pairs_per_cta, total_pairs - pair0);
  const int y_base = pair0 * kRows;              // first output row of this CTA
  const int n_units = npairs * kRows + K - 1;    // halo rows this CTA touches: y_base - pad ... (+ n_units - 1)
  const uint32_t tmem_cols = static_cast<uint32_t>(2 * kRows * p.n_tile);  // 256 or 512: a power of two >= 32

  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) { mbar_init(a_full + i, 1); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < kBStages; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== halo-row producer =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
      for (int u = 0; u < n_units; ++u) {
        const int slot = u % kSlots;
        const uint32_t round = static_cast<uint32_t>(u / kSlots);
        mbar_wait(a_empty + slot, (round & 1) ^ 1);
        mbar_expect_tx(a_full + slot, static_cast<uint32_t>((128 + K - 1) * 128));
        tma_load_4d(a_slots + slot * kSlotBytes, &map_a, a_full + slot, 0, x0 - kPad, y_base - kPad + u, img);
      }
    }
  } else if (warp == 1) {
    // ===== weight producer =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
      uint32_t it = 0;
      for (int j = 0; j < npairs; ++j)
        for (int tap = 0; tap < K * K; ++tap, ++it) {
          const int st = it % kBStages;
          mbar_wait(b_empty + st, ((it / kBStages) & 1) ^ 1);
          mbar_expect_tx(b_full + st, static_cast<uint32_t>(b_bytes));
          tma_load_3d(b_stages + st * b_bytes, &map_b, b_full + st, 0, n0, tap);
        }
    }
  } else if (warp == 2) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(p.n_tile >> 3) << 17) | (8u << 24);  // f16 x f16 -> f32, M = 128
      uint32_t it = 0;
      int waited = 0;  // halo units whose TMA has been observed
      for (int j = 0; j < npairs; ++j) {
        const int acc = j & 1;
        mbar_wait(t_empty + acc, ((j >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kh = 0; kh < K; ++kh) {
          while (waited <= j * kRows + kh + kRows - 1 && waited < n_units) {
            mbar_wait(a_full + waited % kSlots, (waited / kSlots) & 1);
            ++waited;
          }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int kw = 0; kw < K; ++kw, ++it) {
            const int st = it % kBStages;
            mbar_wait(b_full + st, (it / kBStages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t b_addr = smem_u32(b_stages + st * b_bytes);
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
              const int unit = j * kRows + r + kh;
              const uint32_t a_addr = smem_u32(a_slots + (unit % kSlots) * kSlotBytes) + kw * 128;
              const uint32_t d = tmem_base + static_cast<uint32_t>((acc * kRows + r) * p.n_tile);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(d, umma_desc(a_addr + k * 32, p.bo_mode), umma_desc(b_addr + k * 32, p.bo_mode), idesc,
                         (kh | kw | k) != 0);
            }
            umma_commit(b_empty + st);  // weight stage is free once these MMAs retire
          }
        }
        umma_commit(t_full + acc);
        // the first kRows halo rows of this pair are dead now
        for (int r = 0; r < kRows; ++r) umma_commit(a_empty + (j * kRows + r) % kSlots);
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> bias / BN / ReLU / dropout -> half -> global =====
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int x = x0 + q * 32 + lane;
    for (int j = 0; j < npairs; ++j) {
      const int acc = j & 1;
      mbar_wait(t_full + acc, (j >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const int y = y_base + j * kRows + r;
        uint32_t bits[4] = {0, 0, 0, 0};
        for (int cc = 0; cc < p.n_tile; cc += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>((acc * kRows + r) * p.n_tile + cc), v);
          if (y < p.H && x < p.W) {
            const int c0 = n0 + cc;
            if (p.has_drop && ((c0 & 127) == 0 || cc == 0))
              dropout_bits128(p.seed, *p.frame, p.drop_layer, img, static_cast<uint32_t>(y * p.W + x), c0 >> 7, bits);
            const uint32_t keep = p.has_drop ? bits[(c0 >> 5) & 3] : 0xFFFFFFFFu;
            uint32_t packed[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float f[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int c = c0 + i + e;
                float t = __fadd_rn(__uint_as_float(v[i + e]), __ldg(p.bias + c));
                if (p.has_bn) t = __fadd_rn(__fmul_rn(t, __ldg(p.bn_scale + c)), __ldg(p.bn_shift + c));
                if (p.relu) t = t > 0.f ? t : __fmul_rn(p.slope, t);
                f[e] = t;
              }
              __half2 h = __floats2half2_rn(f[0], f[1]);
              if (p.has_drop) {  // y = x * keep * 2 on the stored half value (exact)
                __half2 s = __floats2half2_rn((keep >> i) & 1u ? p.drop_scale : 0.f, (keep >> (i + 1)) & 1u ? p.drop_scale : 0.f);
                h = __hmul2(h, s);
              }
              packed[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
            }
            uint4* dst = reinterpret_cast<uint4*>(p.out + ((static_cast<size_t>(img) * p.H + y) * p.W + x) * p.cout_total + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty + acc);
    }
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---- host: tensor maps through the driver entry point (no -lcuda at link time)
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  });
  if (!fn) fail(SIVO_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  return fn;
}

void encode(CUtensorMap* m, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  cuuint32_t elem[5] = {1, 1, 1, 1, 1};
  CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, base, dims, strides_bytes, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(SIVO_ECUDA, "cuTensorMapEncodeTiled failed with %d", static_cast<int>(r));
}

}  // namespace

struct ConvTcPlan {
  CUtensorMap map_a, map_b;
  TcParams p;
  dim3 grid;
  size_t smem;
  int k;
};

bool conv_tc_supported(const Op& op, const TensorView& in, const TensorView& out) {
  if (in.dt != DType::F16 || out.dt != DType::F16) return false;
  if (op.k != 3 && op.k != 7) return false;
  if (in.cs != 64 || op.cin != 64) return false;        // one 64-channel K chunk per tap (first landing of the kernel)
  if (op.cout % 64 || out.cs != op.cout) return false;
  if (const char* e = std::getenv("SIVO_B200_NO_TC")) if (e[0] == '1') return false;
  return true;
}

std::shared_ptr<ConvTcPlan> conv_tc_plan(const Op& op, const TensorView& in, const TensorView& out, const void* w_tc) {
  auto plan = std::make_shared<ConvTcPlan>();
  const int K = op.k;
  {  // input: NHWC half, dims (C, W, H, N)
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(in.cs), static_cast<cuuint64_t>(in.w), static_cast<cuuint64_t>(in.h),
                          static_cast<cuuint64_t>(in.n)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(in.cs) * 2, static_cast<cuuint64_t>(in.w) * in.cs * 2,
                             static_cast<cuuint64_t>(in.h) * in.w * in.cs * 2};
    cuuint32_t box[4] = {64, static_cast<cuuint32_t>(128 + K - 1), 1, 1};
    encode(&plan->map_a, in.p, 4, dims, strides, box);
  }
  const int n_tile = (op.cout_p % 128 == 0) ? 128 : 64;
  {  // weights: [tap][cout_p][cin_p] half, dims (cin, cout, tap)
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(op.cin_p), static_cast<cuuint64_t>(op.cout_p), static_cast<cuuint64_t>(K * K)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(op.cin_p) * 2, static_cast<cuuint64_t>(op.cout_p) * op.cin_p * 2};
    cuuint32_t box[3] = {64, static_cast<cuuint32_t>(n_tile), 1};
    encode(&plan->map_b, const_cast<void*>(w_tc), 3, dims, strides, box);
  }
  TcParams& p = plan->p;
  p.H = in.h; p.W = in.w; p.N_batch = in.n;
  p.cout_total = out.cs;
  p.n_tile = n_tile;
  p.strips = ceil_div(in.w, 128);
  const int total_pairs = ceil_div(in.h, kRows);
  // enough CTAs for >= 2 waves of 148 SMs when the layer allows it, otherwise as many as there are
  const int columns = p.strips * in.n * (op.cout_p / n_tile);
  int want_blocks = ceil_div(2 * 148, columns);
  int ppc = std::max(1, total_pairs / std::max(1, want_blocks));
  ppc = std::min(ppc, 16);
  p.pairs_per_cta = ppc;
  p.relu = op.relu; p.has_bn = op.has_bn; p.slope = op.slope;
  p.bias = op.bias.as<float>();
  p.bn_scale = op.has_bn ? op.bn_scale.as<float>() : nullptr;
  p.bn_shift = op.has_bn ? op.bn_shift.as<float>() : nullptr;
  p.out = static_cast<__half*>(out.p);
  p.has_drop = 0; p.seed = 0; p.frame = nullptr; p.drop_layer = 0; p.drop_scale = 2.f;
  p.bo_mode = 0;
  if (const char* e = std::getenv("SIVO_B200_TC_BO")) p.bo_mode = atoi(e);
  plan->grid = dim3(p.strips * ceil_div(total_pairs, ppc), op.cout_p / n_tile, in.n);
  const int slots = K + kRows - 1 + kRows;
  plan->smem = 1024 + static_cast<size_t>(slots) * kSlotBytes + static_cast<size_t>(kBStages) * n_tile * 128 + (2 * slots + 2 * kBStages + 4) * 8 + 16;
  plan->k = K;
  auto set = [&](auto kern) {
    SIVO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(plan->smem)));
  };
  if (K == 7) set(k_conv_tc<7>); else set(k_conv_tc<3>);
  return plan;
}

void conv_tc_launch(const ConvTcPlan& plan, const Op& op, cudaStream_t s) {
  (void)op;
  if (plan.k == 7) k_conv_tc<7><<<plan.grid, kTcThreads, plan.smem, s>>>(plan.map_a, plan.map_b, plan.p);
  else k_conv_tc<3><<<plan.grid, kTcThreads, plan.smem, s>>>(plan.map_a, plan.map_b, plan.p);
  SIVO_CUDA(cudaGetLastError());
}

}  // namespace sivo
