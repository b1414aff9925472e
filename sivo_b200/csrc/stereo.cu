// Hamming stage of Frame::ComputeStereoMatches (src/orbslam/Frame.cc:452-533) -- SURVEY 8f row 1.
// One warp per left keypoint scans every right keypoint (<= ~2000, descriptors stay in L2): row band
// [floor(vR - 2 s), ceil(vR + 2 s)] must contain int(vL), octave within +-1, uR in [uL - maxD, uL - minD];
// distance = ORBmatcher::DescriptorDistance (ORBmatcher.cc:1582-1596) as 4 x __popcll.  The reference keeps
// the first strictly smaller distance walking candidates in increasing iR, i.e. (min dist, min iR).
#include <algorithm>
#include <vector>

#include "orb.h"

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

// ---- the whole of Frame::ComputeStereoMatches (Frame.cc:444-629) for one left keypoint per warp: Hamming search as
// above, then the 11x11 SAD slide over +-5 px on the keypoint's pyramid level (windows are centred-intensity
// differences of u8 pixels, so every L1 distance is an exact integer), parabola sub-pixel fit, disparity gate.
// The median-based outlier cut (:617-628) needs all matches and runs on the host over the returned SAD distances.
__device__ __forceinline__ float c_roundf(float v) { return v >= 0.f ? floorf(__fadd_rn(v, 0.5f)) : -floorf(__fadd_rn(-v, 0.5f)); }

__global__ void k_stereo_match(const sivo_keypoint* __restrict__ kl, const uint8_t* __restrict__ dl, int nl,
                               const sivo_keypoint* __restrict__ kr, const uint8_t* __restrict__ dr, int nr,
                               const float* __restrict__ scale, const float* __restrict__ inv_scale, const uint8_t* __restrict__ pyr_l,
                               OrbLevelTable lt_l, const uint8_t* __restrict__ pyr_r, OrbLevelTable lt_r, float mb, float mbf,
                               float* __restrict__ u_right, float* __restrict__ depth, int* __restrict__ sad_dist) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= nl) return;
  const sivo_keypoint L = kl[wid];
  const int rows = lt_l.lv[0].h;
  const int row = static_cast<int>(L.y);
  const float min_d = 0.f, max_d = __fdiv_rn(mbf, mb);
  const float min_u = __fsub_rn(L.x, max_d), max_u = __fsub_rn(L.x, min_d);
  const ulonglong4 a = *reinterpret_cast<const ulonglong4*>(dl + static_cast<size_t>(wid) * 32);
  unsigned best = (static_cast<unsigned>(kThHigh) << 20) | 0xFFFFFu;
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
  float out_u = -1.f, out_z = -1.f;
  int out_sad = -1;
  const int hd = best >> 20;
  if (hd < (kThHigh + 50) / 2) {  // thOrbDist = (TH_HIGH + TH_LOW) / 2 (:448)
    const int ir = best & 0xFFFFFu;
    const int oct = L.octave;
    const float sf = inv_scale[oct];
    const float su_l = c_roundf(__fmul_rn(L.x, sf)), sv_l = c_roundf(__fmul_rn(L.y, sf)), su_r0 = c_roundf(__fmul_rn(kr[ir].x, sf));
    const OrbLevel ll = lt_l.lv[oct], lr = lt_r.lv[oct];
    const float iniu = su_r0 + 5.f - 5.f, endu = su_r0 + 5.f + 5.f + 1.f;
    const int cx_l = static_cast<int>(su_l), cy = static_cast<int>(sv_l), cx_r = static_cast<int>(su_r0);
    if (!(iniu < 0.f || endu >= static_cast<float>(lr.w)) && cx_l - 5 >= -kEdge && cx_l + 5 < ll.w + kEdge && cy - 5 >= -kEdge &&
        cy + 5 < ll.h + kEdge) {
      const uint8_t* pl = pyr_l + ll.img_off + static_cast<size_t>(cy + kEdge) * ll.pitch + cx_l + kEdge;
      const uint8_t* pr = pyr_r + lr.img_off + static_cast<size_t>(cy + kEdge) * lr.pitch + cx_r + kEdge;
      const int lc = pl[0];
      int dists[11];
#pragma unroll
      for (int inc = -5; inc <= 5; ++inc) {
        const int rc = pr[inc];
        int s = 0;
        for (int e = lane; e < 121; e += 32) {
          const int dy = e / 11 - 5, dx = e % 11 - 5;
          const int vl = static_cast<int>(pl[dy * ll.pitch + dx]) - lc;
          const int vr = static_cast<int>(pr[dy * lr.pitch + dx + inc]) - rc;
          s += abs(vl - vr);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        dists[inc + 5] = s;
      }
      int bd = 0x7fffffff, binc = 0;
#pragma unroll
      for (int i = 0; i < 11; ++i)
        if (dists[i] < bd) { bd = dists[i]; binc = i - 5; }
      if (binc != -5 && binc != 5) {
        float d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
        for (int i = 1; i < 10; ++i)
          if (i - 5 == binc) { d1 = static_cast<float>(dists[i - 1]); d2 = static_cast<float>(dists[i]); d3 = static_cast<float>(dists[i + 1]); }
        const float delta = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
        if (!(delta < -1.f || delta > 1.f)) {
          float best_u = __fmul_rn(scale[oct], __fadd_rn(__fadd_rn(su_r0, static_cast<float>(binc)), delta));
          float disp = __fsub_rn(L.x, best_u);
          if (disp >= min_d && disp < max_d) {
            if (disp <= 0.f) {
              disp = 0.01f;
              best_u = static_cast<float>(static_cast<double>(L.x) - 0.01);
            }
            out_z = __fdiv_rn(mbf, disp);
            out_u = best_u;
            out_sad = bd;
          }
        }
      }
    }
  }
  if (lane == 0) {
    u_right[wid] = out_u;
    depth[wid] = out_z;
    sad_dist[wid] = out_sad;
  }
}

// Per-thread, per-device workspace of the stereo / Hamming entry points: one arena that only grows and one non-blocking stream.
// (Allocating and freeing half a dozen buffers per call and running on the legacy default stream synchronises the whole device,
// i.e. serialises every per-frame call against the SegNet graph and the extractors' streams.)
struct StereoWs {
  int device = -1;
  cudaStream_t stream = nullptr;
  DevBuf arena;
  size_t used = 0;
  void begin(int dev, size_t bytes_needed) {
    SIVO_CUDA(cudaSetDevice(dev));
    if (device != dev) {
      if (stream) { cudaSetDevice(device); cudaStreamDestroy(stream); cudaSetDevice(dev); }
      arena.release();
      SIVO_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
      device = dev;
    }
    if (arena.bytes < bytes_needed) {
      SIVO_CUDA(cudaStreamSynchronize(stream));
      arena.alloc(bytes_needed + bytes_needed / 2);
    }
    used = 0;
  }
  template <class T> T* take(size_t count) {
    used = (used + 255) & ~static_cast<size_t>(255);
    T* p = reinterpret_cast<T*>(arena.as<uint8_t>() + used);
    used += std::max<size_t>(count, 1) * sizeof(T);
    if (used > arena.bytes) fail(SIVO_ENOMEM, "stereo workspace accounting");
    return p;
  }
  ~StereoWs() { if (stream) { cudaSetDevice(device); cudaStreamDestroy(stream); } }
};
StereoWs& stereo_ws() {
  static thread_local StereoWs ws;
  return ws;
}
inline size_t pad256(size_t b) { return (std::max<size_t>(b, 1) + 255) & ~static_cast<size_t>(255); }
}  // namespace

void stereo_hamming(int device, const sivo_keypoint* left, const uint8_t* dl, int nl, const sivo_keypoint* right,
                    const uint8_t* dr, int nr, const float* scale, int nlevels, int rows, float min_d, float max_d,
                    int* best_idx, int* best_dist) {
  if (nl < 0 || nr < 0 || nr >= (1 << 20)) fail(SIVO_EINVAL, "stereo: bad keypoint counts %d / %d", nl, nr);
  if (nl == 0) return;
  for (int i = 0; i < nl; ++i) if (left[i].octave < 0 || left[i].octave >= nlevels) fail(SIVO_EINVAL, "stereo: left octave out of range");
  for (int i = 0; i < nr; ++i) if (right[i].octave < 0 || right[i].octave >= nlevels) fail(SIVO_EINVAL, "stereo: right octave out of range");
  StereoWs& ws = stereo_ws();
  ws.begin(device, pad256(nl * sizeof(sivo_keypoint)) + pad256(static_cast<size_t>(nl) * 32) + pad256(nr * sizeof(sivo_keypoint)) +
                       pad256(static_cast<size_t>(nr) * 32) + pad256(nlevels * sizeof(float)) + 2 * pad256(nl * sizeof(int)) + 4096);
  cudaStream_t st = ws.stream;
  sivo_keypoint* d_kl = ws.take<sivo_keypoint>(nl);
  uint8_t* d_dl = ws.take<uint8_t>(static_cast<size_t>(nl) * 32);
  sivo_keypoint* d_kr = ws.take<sivo_keypoint>(nr);
  uint8_t* d_dr = ws.take<uint8_t>(static_cast<size_t>(nr) * 32);
  float* d_sc = ws.take<float>(nlevels);
  int* d_bi = ws.take<int>(nl);
  int* d_bd = ws.take<int>(nl);
  SIVO_CUDA(cudaMemcpyAsync(d_kl, left, nl * sizeof(sivo_keypoint), cudaMemcpyHostToDevice, st));
  SIVO_CUDA(cudaMemcpyAsync(d_dl, dl, static_cast<size_t>(nl) * 32, cudaMemcpyHostToDevice, st));
  if (nr) {
    SIVO_CUDA(cudaMemcpyAsync(d_kr, right, nr * sizeof(sivo_keypoint), cudaMemcpyHostToDevice, st));
    SIVO_CUDA(cudaMemcpyAsync(d_dr, dr, static_cast<size_t>(nr) * 32, cudaMemcpyHostToDevice, st));
  }
  SIVO_CUDA(cudaMemcpyAsync(d_sc, scale, nlevels * sizeof(float), cudaMemcpyHostToDevice, st));
  k_stereo_hamming<<<ceil_div(nl * 32, 128), 128, 0, st>>>(d_kl, d_dl, nl, d_kr, d_dr, nr, d_sc, rows, min_d, max_d, d_bi, d_bd);
  SIVO_CUDA(cudaGetLastError());
  SIVO_CUDA(cudaMemcpyAsync(best_idx, d_bi, nl * sizeof(int), cudaMemcpyDeviceToHost, st));
  SIVO_CUDA(cudaMemcpyAsync(best_dist, d_bd, nl * sizeof(int), cudaMemcpyDeviceToHost, st));
  SIVO_CUDA(cudaStreamSynchronize(st));
}

// ---- best / second-best descriptor match over per-query candidate lists: the inner loop shared by
// ORBmatcher::SearchByProjection (ORBmatcher.cc:79-113 and :1278-), SearchForTriangulation (:631-) and SearchBySim3 --
//   bestDist = bestDist2 = 256, bestLevel = bestLevel2 = -1, bestIdx = -1;
//   for idx in candidates (in order): dist = DescriptorDistance(q, train[idx]);
//     if (dist < bestDist) { bestDist2 = bestDist; bestLevel2 = bestLevel; bestDist = dist; bestLevel = level[idx]; bestIdx = idx; }
//     else if (dist < bestDist2) { bestLevel2 = level[idx]; bestDist2 = dist; }
// One warp per query: the lanes compute 32 distances at a time, lane 0 folds them in candidate order so that ties fall
// exactly as in the sequential loop.
namespace {
__global__ void k_hamming_best2(const uint8_t* __restrict__ query, int nq, const uint8_t* __restrict__ train,
                                const int* __restrict__ cand_off, const int* __restrict__ cand_idx, const int* __restrict__ level,
                                int* __restrict__ out) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= nq) return;
  const ulonglong4 a = *reinterpret_cast<const ulonglong4*>(query + static_cast<size_t>(wid) * 32);
  const int b = cand_off[wid], e = cand_off[wid + 1];
  int best_dist = 256, best_level = -1, best_dist2 = 256, best_level2 = -1, best_idx = -1;
  for (int base = b; base < e; base += 32) {
    const int i = base + lane;
    int d = 0x7FFFFFFF, idx = -1, lvl = -1;
    if (i < e) {
      idx = cand_idx[i];
      const ulonglong4 t = *reinterpret_cast<const ulonglong4*>(train + static_cast<size_t>(idx) * 32);
      d = __popcll(a.x ^ t.x) + __popcll(a.y ^ t.y) + __popcll(a.z ^ t.z) + __popcll(a.w ^ t.w);
      lvl = level ? level[idx] : 0;
    }
    const int n = min(32, e - base);
    for (int k = 0; k < n; ++k) {  // warp-uniform fold in candidate order
      const int dk = __shfl_sync(0xffffffffu, d, k), ik = __shfl_sync(0xffffffffu, idx, k), lk = __shfl_sync(0xffffffffu, lvl, k);
      if (dk < best_dist) { best_dist2 = best_dist; best_level2 = best_level; best_dist = dk; best_level = lk; best_idx = ik; }
      else if (dk < best_dist2) { best_level2 = lk; best_dist2 = dk; }
    }
  }
  if (lane == 0) {
    int* o = out + static_cast<size_t>(wid) * 5;
    o[0] = best_idx; o[1] = best_dist; o[2] = best_level; o[3] = best_dist2; o[4] = best_level2;
  }
}
}  // namespace

void hamming_best2(int device, const uint8_t* query, int nq, const uint8_t* train, int nt, const int* cand_off, const int* cand_idx,
                   const int* train_level, int* out5) {
  if (nq < 0 || nt < 0) fail(SIVO_EINVAL, "hamming_best2: negative counts");
  if (nq == 0) return;
  if (cand_off[0] != 0) fail(SIVO_EINVAL, "hamming_best2: candidate offsets must start at 0");
  for (int i = 0; i < nq; ++i) if (cand_off[i + 1] < cand_off[i]) fail(SIVO_EINVAL, "hamming_best2: candidate offsets must not decrease");
  const int nc = cand_off[nq];
  for (int i = 0; i < nc; ++i) if (cand_idx[i] < 0 || cand_idx[i] >= nt) fail(SIVO_ERANGE, "hamming_best2: candidate %d names train descriptor %d of %d", i, cand_idx[i], nt);
  StereoWs& ws = stereo_ws();
  ws.begin(device, pad256(static_cast<size_t>(nq) * 32) + pad256(static_cast<size_t>(nt) * 32) + pad256((nq + 1) * sizeof(int)) +
                       pad256(nc * sizeof(int)) + pad256(nt * sizeof(int)) + pad256(static_cast<size_t>(nq) * 5 * sizeof(int)) + 4096);
  cudaStream_t st = ws.stream;
  uint8_t* d_q = ws.take<uint8_t>(static_cast<size_t>(nq) * 32);
  uint8_t* d_t = ws.take<uint8_t>(static_cast<size_t>(nt) * 32);
  int* d_off = ws.take<int>(nq + 1);
  int* d_idx = ws.take<int>(nc);
  int* d_lvl = ws.take<int>(nt);
  int* d_out = ws.take<int>(static_cast<size_t>(nq) * 5);
  SIVO_CUDA(cudaMemcpyAsync(d_q, query, static_cast<size_t>(nq) * 32, cudaMemcpyHostToDevice, st));
  if (nt) SIVO_CUDA(cudaMemcpyAsync(d_t, train, static_cast<size_t>(nt) * 32, cudaMemcpyHostToDevice, st));
  SIVO_CUDA(cudaMemcpyAsync(d_off, cand_off, (nq + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  if (nc) SIVO_CUDA(cudaMemcpyAsync(d_idx, cand_idx, nc * sizeof(int), cudaMemcpyHostToDevice, st));
  if (train_level && nt) SIVO_CUDA(cudaMemcpyAsync(d_lvl, train_level, nt * sizeof(int), cudaMemcpyHostToDevice, st));
  k_hamming_best2<<<ceil_div(nq * 32, 128), 128, 0, st>>>(d_q, nq, d_t, d_off, d_idx, train_level ? d_lvl : nullptr, d_out);
  SIVO_CUDA(cudaGetLastError());
  SIVO_CUDA(cudaMemcpyAsync(out5, d_out, static_cast<size_t>(nq) * 5 * sizeof(int), cudaMemcpyDeviceToHost, st));
  SIVO_CUDA(cudaStreamSynchronize(st));
}

}  // namespace sivo

namespace sivo {
// ComputeStereoMatches end to end; `left` / `right` are the extractors whose last runs produced the keypoints.
void stereo_match(const Orb& left, const Orb& right, const sivo_keypoint* kl, const uint8_t* dl, int nl, const sivo_keypoint* kr,
                  const uint8_t* dr, int nr, float mb, float mbf, float* u_right, float* depth) {
  if (!left.has_run() || !right.has_run()) fail(SIVO_EINVAL, "stereo: run both extractors first");
  if (left.device() != right.device() || left.nlevels() != right.nlevels()) fail(SIVO_EINVAL, "stereo: extractors differ");
  if (nl < 0 || nr < 0 || nr >= (1 << 20) || !(mb > 0.f) || !(mbf > 0.f)) fail(SIVO_EINVAL, "stereo: bad arguments");
  for (int i = 0; i < nl; ++i) { u_right[i] = -1.f; depth[i] = -1.f; }
  if (nl == 0 || nr == 0) return;
  const int nlev = left.nlevels();
  for (int i = 0; i < nl; ++i) if (kl[i].octave < 0 || kl[i].octave >= nlev) fail(SIVO_EINVAL, "stereo: left octave out of range");
  for (int i = 0; i < nr; ++i) if (kr[i].octave < 0 || kr[i].octave >= nlev) fail(SIVO_EINVAL, "stereo: right octave out of range");
  StereoWs& ws = stereo_ws();
  ws.begin(left.device(), pad256(nl * sizeof(sivo_keypoint)) + pad256(static_cast<size_t>(nl) * 32) + pad256(nr * sizeof(sivo_keypoint)) +
                              pad256(static_cast<size_t>(nr) * 32) + 2 * pad256(nlev * sizeof(float)) + 3 * pad256(nl * sizeof(float)) + 4096);
  cudaStream_t st = ws.stream;
  sivo_keypoint* d_kl = ws.take<sivo_keypoint>(nl);
  uint8_t* d_dl = ws.take<uint8_t>(static_cast<size_t>(nl) * 32);
  sivo_keypoint* d_kr = ws.take<sivo_keypoint>(nr);
  uint8_t* d_dr = ws.take<uint8_t>(static_cast<size_t>(nr) * 32);
  float* d_sc = ws.take<float>(nlev);
  float* d_isc = ws.take<float>(nlev);
  float* d_u = ws.take<float>(nl);
  float* d_z = ws.take<float>(nl);
  int* d_s = ws.take<int>(nl);
  SIVO_CUDA(cudaMemcpyAsync(d_kl, kl, nl * sizeof(sivo_keypoint), cudaMemcpyHostToDevice, st));
  SIVO_CUDA(cudaMemcpyAsync(d_dl, dl, static_cast<size_t>(nl) * 32, cudaMemcpyHostToDevice, st));
  SIVO_CUDA(cudaMemcpyAsync(d_kr, kr, nr * sizeof(sivo_keypoint), cudaMemcpyHostToDevice, st));
  SIVO_CUDA(cudaMemcpyAsync(d_dr, dr, static_cast<size_t>(nr) * 32, cudaMemcpyHostToDevice, st));
  SIVO_CUDA(cudaMemcpyAsync(d_sc, left.tables().scale.data(), nlev * sizeof(float), cudaMemcpyHostToDevice, st));
  SIVO_CUDA(cudaMemcpyAsync(d_isc, left.tables().inv_scale.data(), nlev * sizeof(float), cudaMemcpyHostToDevice, st));
  // the pyramids were written on the two extractors' own streams; their synchronous run() has returned, so they are complete
  k_stereo_match<<<ceil_div(nl * 32, 128), 128, 0, st>>>(d_kl, d_dl, nl, d_kr, d_dr, nr, d_sc, d_isc, left.dev_pyramid(), left.levels(),
                                                         right.dev_pyramid(), right.levels(), mb, mbf, d_u, d_z, d_s);
  SIVO_CUDA(cudaGetLastError());
  std::vector<int> sad(nl);
  SIVO_CUDA(cudaMemcpyAsync(u_right, d_u, nl * sizeof(float), cudaMemcpyDeviceToHost, st));
  SIVO_CUDA(cudaMemcpyAsync(depth, d_z, nl * sizeof(float), cudaMemcpyDeviceToHost, st));
  SIVO_CUDA(cudaMemcpyAsync(sad.data(), d_s, nl * sizeof(int), cudaMemcpyDeviceToHost, st));
  SIVO_CUDA(cudaStreamSynchronize(st));
  // median-based outlier cut (Frame.cc:617-628)
  std::vector<std::pair<int, int>> v;
  for (int i = 0; i < nl; ++i) if (sad[i] >= 0) v.emplace_back(sad[i], i);
  if (v.empty()) return;
  std::sort(v.begin(), v.end());
  const float median = static_cast<float>(v[v.size() / 2].first);
  const float th = 1.5f * 1.4f * median;
  for (int i = static_cast<int>(v.size()) - 1; i >= 0; --i) {
    if (static_cast<float>(v[i].first) < th) break;
    u_right[v[i].second] = -1.f;
    depth[v[i].second] = -1.f;
  }
}
}  // namespace sivo
