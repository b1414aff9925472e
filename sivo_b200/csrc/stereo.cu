// Hamming stage of Frame::ComputeStereoMatches (src/orbslam/Frame.cc:452-533) -- SURVEY 8f row 1.
// One warp per left keypoint scans every right keypoint (<= ~2000, descriptors stay in L2): row band
// [floor(vR - 2 s), ceil(vR + 2 s)] must contain int(vL), octave within +-1, uR in [uL - maxD, uL - minD];
// distance = ORBmatcher::DescriptorDistance (ORBmatcher.cc:1582-1596) as 4 x __popcll.  The reference keeps
// the first strictly smaller distance walking candidates in increasing iR, i.e. (min dist, min iR).
#include "common.h"

namespace sivo {
namespace {
constexpr int kThHigh = 100;  // ORBmatcher::TH_HIGH (ORBmatcher.cc:37)

__global__ void k_stereo_hamming(const sivo_keypoint* __restrict__ kl, const uint8_t* __restrict__ dl, int nl,
                                 const sivo_keypoint* __restrict__ kr, const uint8_t* __restrict__ dr, int nr,
                                 const float* __restrict__ scale, int rows, float min_d, float max_d,
                                 int* __restrict__ best_idx, int* __restrict__ best_dist) {
  int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= nl) return;
  const sivo_keypoint L = kl[wid];
  const int row = static_cast<int>(L.y);
  const float min_u = __fsub_rn(L.x, max_d), max_u = __fsub_rn(L.x, min_d);
  const ulonglong4 a = *reinterpret_cast<const ulonglong4*>(dl + static_cast<size_t>(wid) * 32);
  unsigned best = (static_cast<unsigned>(kThHigh) << 20) | 0xFFFFFu;  // dist:12 | idx:20
  if (max_u >= 0.f && row >= 0 && row < rows) {
    for (int i = lane; i < nr; i += 32) {
      const sivo_keypoint R = kr[i];
      const float band = __fmul_rn(2.0f, scale[R.octave]);
      const int maxr = static_cast<int>(ceilf(__fadd_rn(R.y, band)));
      const int minr = static_cast<int>(floorf(__fsub_rn(R.y, band)));
      if (row < minr || row > maxr) continue;
      if (R.octave < L.octave - 1 || R.octave > L.octave + 1) continue;
      if (!(R.x >= min_u && R.x <= max_u)) continue;
      const ulonglong4 b = *reinterpret_cast<const ulonglong4*>(dr + static_cast<size_t>(i) * 32);
      unsigned d = __popcll(a.x ^ b.x) + __popcll(a.y ^ b.y) + __popcll(a.z ^ b.z) + __popcll(a.w ^ b.w);
      if (d < static_cast<unsigned>(kThHigh)) best = min(best, (d << 20) | static_cast<unsigned>(i));
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
  if (lane == 0) {
    int d = best >> 20;
    best_dist[wid] = d;
    best_idx[wid] = d < kThHigh ? static_cast<int>(best & 0xFFFFFu) : -1;
  }
}
}  // namespace

void stereo_hamming(int device, const sivo_keypoint* left, const uint8_t* dl, int nl, const sivo_keypoint* right,
                    const uint8_t* dr, int nr, const float* scale, int nlevels, int rows, float min_d, float max_d,
                    int* best_idx, int* best_dist) {
  if (nl < 0 || nr < 0 || nr >= (1 << 20)) fail(SIVO_EINVAL, "stereo: bad keypoint counts %d / %d", nl, nr);
  if (nl == 0) return;
  for (int i = 0; i < nl; ++i) if (left[i].octave < 0 || left[i].octave >= nlevels) fail(SIVO_EINVAL, "stereo: left octave out of range");
  for (int i = 0; i < nr; ++i) if (right[i].octave < 0 || right[i].octave >= nlevels) fail(SIVO_EINVAL, "stereo: right octave out of range");
  SIVO_CUDA(cudaSetDevice(device));
  DevBuf d_kl(nl * sizeof(sivo_keypoint)), d_dl(static_cast<size_t>(nl) * 32), d_kr(std::max(nr, 1) * sizeof(sivo_keypoint)),
      d_dr(static_cast<size_t>(std::max(nr, 1)) * 32), d_sc(nlevels * sizeof(float)), d_bi(nl * sizeof(int)), d_bd(nl * sizeof(int));
  SIVO_CUDA(cudaMemcpy(d_kl.p, left, nl * sizeof(sivo_keypoint), cudaMemcpyHostToDevice));
  SIVO_CUDA(cudaMemcpy(d_dl.p, dl, static_cast<size_t>(nl) * 32, cudaMemcpyHostToDevice));
  if (nr) {
    SIVO_CUDA(cudaMemcpy(d_kr.p, right, nr * sizeof(sivo_keypoint), cudaMemcpyHostToDevice));
    SIVO_CUDA(cudaMemcpy(d_dr.p, dr, static_cast<size_t>(nr) * 32, cudaMemcpyHostToDevice));
  }
  SIVO_CUDA(cudaMemcpy(d_sc.p, scale, nlevels * sizeof(float), cudaMemcpyHostToDevice));
  k_stereo_hamming<<<ceil_div(nl * 32, 128), 128>>>(d_kl.as<sivo_keypoint>(), d_dl.as<uint8_t>(), nl, d_kr.as<sivo_keypoint>(),
                                                    d_dr.as<uint8_t>(), nr, d_sc.as<float>(), rows, min_d, max_d,
                                                    d_bi.as<int>(), d_bd.as<int>());
  SIVO_CUDA(cudaGetLastError());
  SIVO_CUDA(cudaMemcpy(best_idx, d_bi.p, nl * sizeof(int), cudaMemcpyDeviceToHost));
  SIVO_CUDA(cudaMemcpy(best_dist, d_bd.p, nl * sizeof(int), cudaMemcpyDeviceToHost));
}

}  // namespace sivo
