// DistributeOctTree (src/orbslam/ORBextractor.cc:544-750, ExtractorNode::DivideNode :488-542) on the device: one thread
// block per pyramid level, the whole tree in shared memory.  Same function as the host restatement `orb_distribute`
// (orb_host.cu) -- same node order, same documented tie rule (creation order instead of the reference's heap-pointer order in
// the last round) -- so the selected keypoints and their ORDER are identical; tests compare the two on random inputs and on
// whole images.  With the tree on the device the extractor has no host round trip: pyramid -> FAST -> cells -> tree ->
// describe is one asynchronous chain of launches (the host path remains for inputs that exceed the shared-memory caps).
//
// Representation.  Keys stay in place; `perm` holds key indices and every node owns a contiguous range of it (a split is a
// stable 4-way partition of that range, as in the host version).  The reference's std::list is an ARRAY in list order:
// a regular round visits the list front to back, replaces every splittable node by its non-empty children (each
// `push_front`), and leaves the others in place, so
//     new list = reverse(children in creation order) ++ (unsplit nodes in their old order),
// which a prefix sum over the old list computes in parallel (one warp per node does the partitions).  The last phase (largest
// nodes first, one at a time, stop as soon as the target is reached) is sequential by nature: one warp runs it, pushing
// children on a `front` stack (list = reverse(front) ++ base) and clearing the parents' alive flags.
#include <cuda_runtime.h>

#include <cstdint>

#include "orb.h"

namespace sivo {

namespace {

constexpr int kTreeThreads = 512, kTreeWarps = kTreeThreads / 32;

struct TreeNode {
  short ulx, uly, urx, bry;
  unsigned short begin, count, seq;
  unsigned char no_more, alive;
};

struct TreeSmem {
  unsigned short kx[kTreeKeyCap], ky[kTreeKeyCap];
  unsigned char kr[kTreeKeyCap];
  unsigned short perm[kTreeKeyCap], perm2[kTreeKeyCap];
  TreeNode node[kTreeNodeCap];
  unsigned short base[kTreeListCap], base2[kTreeListCap], front[kTreeListCap];
  unsigned short pend[kTreeListCap], pend2[kTreeListCap];   // splittable children in creation order (node ids)
  unsigned short seq2node[kTreeNodeCap];
  unsigned short cnt[kTreeListCap][4];                       // per list position: child sizes of this round's split
  unsigned int sortkey[kTreeListCap];
  int scan[kTreeWarps];
  int warp_cnt[kTreeWarps][16];
  unsigned short c4[4];
  int n_nodes, size, base_len, n_front, n_pend, n_pend2, seq, error, finish;
};

__device__ __forceinline__ int ceil_half(int d) { return (d + 1) >> 1; }  // ceil(d / 2.f) for d >= 0

// exclusive block scan of one int per thread (returns the prefix; *total = sum); all threads must call
__device__ int block_exscan(int v, int* scratch, int* total) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // scratch may still be read from a previous call
  if (lane == 31) scratch[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < kTreeWarps ? scratch[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    if (lane < kTreeWarps) scratch[lane] = w;
  }
  __syncthreads();
  const int warp_off = warp ? scratch[warp - 1] : 0;
  *total = scratch[kTreeWarps - 1];
  return warp_off + inc - v;
}

// One warp splits node `id` (ExtractorNode::DivideNode): stable 4-way partition of its key range by quadrant
// (UL, UR, BL, BR = !(x < mx) + 2 * !(y < my)), child sizes to out_cnt[4].  perm2 is scratch for the range.
__device__ void warp_split(TreeSmem& s, int id, unsigned short* out_cnt) {
  const int lane = threadIdx.x & 31;
  const TreeNode nd = s.node[id];
  const int mx = nd.ulx + ceil_half(nd.urx - nd.ulx), my = nd.uly + ceil_half(nd.bry - nd.uly);
  int c[4] = {0, 0, 0, 0};
  for (int i0 = 0; i0 < nd.count; i0 += 32) {
    const int i = i0 + lane;
    int q = -1;
    if (i < nd.count) {
      const int k = s.perm[nd.begin + i];
      q = static_cast<int>(!(s.kx[k] < mx)) + 2 * static_cast<int>(!(s.ky[k] < my));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] += __popc(__ballot_sync(0xffffffffu, q == j));
  }
  int off[4] = {0, c[0], c[0] + c[1], c[0] + c[1] + c[2]};
  for (int i0 = 0; i0 < nd.count; i0 += 32) {
    const int i = i0 + lane;
    int q = -1, k = 0;
    if (i < nd.count) {
      k = s.perm[nd.begin + i];
      q = static_cast<int>(!(s.kx[k] < mx)) + 2 * static_cast<int>(!(s.ky[k] < my));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned b = __ballot_sync(0xffffffffu, q == j);
      if (q == j) s.perm2[nd.begin + off[j] + __popc(b & ((1u << lane) - 1u))] = static_cast<unsigned short>(k);
      off[j] += __popc(b);
    }
  }
  __syncwarp();
  for (int i = lane; i < nd.count; i += 32) s.perm[nd.begin + i] = s.perm2[nd.begin + i];
  __syncwarp();
  if (lane < 4) out_cnt[lane] = static_cast<unsigned short>(c[lane]);
  __syncwarp();
}

__device__ __forceinline__ TreeNode child_of(const TreeNode& p, int q, int start, int c) {
  const int mx = p.ulx + ceil_half(p.urx - p.ulx), my = p.uly + ceil_half(p.bry - p.uly);
  TreeNode ch;
  ch.ulx = static_cast<short>((q & 1) ? mx : p.ulx);
  ch.urx = static_cast<short>((q & 1) ? p.urx : mx);
  ch.uly = static_cast<short>((q & 2) ? my : p.uly);
  ch.bry = static_cast<short>((q & 2) ? p.bry : my);
  ch.begin = static_cast<unsigned short>(p.begin + start);
  ch.count = static_cast<unsigned short>(c);
  ch.seq = 0;
  ch.no_more = c == 1;
  ch.alive = 1;
  return ch;
}

// cand: packed x | y << 12 | response << 24 (k_cells), all levels, level l at [level_off[l], level_off[l + 1]).
// out_sel[l * kTreeSelCap + i]: packed candidate of the i-th retained keypoint of level l (list order); out_count[l].
// A level that does not fit the shared-memory caps sets bit l of *out_error (the host then takes its own path).
__global__ void __launch_bounds__(kTreeThreads, 1)
k_distribute(const uint32_t* __restrict__ cand, const int* __restrict__ level_off, OrbTreeParams prm, uint32_t* __restrict__ out_sel,
             int* __restrict__ out_count, int* __restrict__ out_error) {
  extern __shared__ __align__(16) uint8_t tree_raw[];
  TreeSmem& s = *reinterpret_cast<TreeSmem*>(tree_raw);
  const int l = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b0 = level_off[l], m = level_off[l + 1] - b0;
  const int n_target = prm.n_target[l], n_ini = prm.n_ini[l];
  const float hx = prm.hx[l];
  if (m <= 0) { if (tid == 0) out_count[l] = 0; return; }
  if (m > kTreeKeyCap || n_ini > 16 || n_target + 8 > kTreeSelCap) {
    if (tid == 0) { out_count[l] = 0; atomicOr(out_error, 1 << l); }
    return;
  }
  // ---- keys; initial cells: stable counting sort by int(x / hx), clamped (ORBextractor.cc:551-575)
  for (int k = tid; k < m; k += kTreeThreads) {
    const uint32_t c = cand[b0 + k];
    s.kx[k] = static_cast<unsigned short>(c & 0xFFF);
    s.ky[k] = static_cast<unsigned short>((c >> 12) & 0xFFF);
    s.kr[k] = static_cast<unsigned char>(c >> 24);
  }
  if (tid < kTreeWarps * 16) (&s.warp_cnt[0][0])[tid] = 0;
  if (tid == 0) { s.error = 0; s.finish = 0; }
  __syncthreads();
  const int per_warp = (m + kTreeWarps - 1) / kTreeWarps;
  const int k_lo = min(m, warp * per_warp), k_hi = min(m, k_lo + per_warp);
  auto cell_of = [&](int k) {
    const int c = static_cast<int>(__fdiv_rn(static_cast<float>(s.kx[k]), hx));
    return c >= n_ini ? n_ini - 1 : c;
  };
  for (int k0 = k_lo; k0 < k_hi; k0 += 32) {
    const int k = k0 + lane;
    const int c = k < k_hi ? cell_of(k) : -1;
    for (int j = 0; j < n_ini; ++j) {
      const int n = __popc(__ballot_sync(0xffffffffu, c == j));
      if (lane == 0) s.warp_cnt[warp][j] += n;
    }
  }
  __syncthreads();
  if (tid == 0) {
    // nodes 0 .. n_ini - 1 in creation order (push_back); warp_cnt becomes the write cursor of each (warp, cell)
    int off = 0;
    for (int j = 0; j < n_ini; ++j) {
      TreeNode nd;
      nd.ulx = static_cast<short>(static_cast<int>(__fmul_rn(hx, static_cast<float>(j))));
      nd.urx = static_cast<short>(static_cast<int>(__fmul_rn(hx, static_cast<float>(j + 1))));
      nd.uly = 0;
      nd.bry = static_cast<short>(prm.height[l]);
      nd.begin = static_cast<unsigned short>(off);
      int c = 0;
      for (int w = 0; w < kTreeWarps; ++w) { const int t = s.warp_cnt[w][j]; s.warp_cnt[w][j] = off + c; c += t; }
      nd.count = static_cast<unsigned short>(c);
      nd.seq = 0;
      nd.no_more = c == 1;
      nd.alive = 1;
      s.node[j] = nd;
      off += c;
    }
    int sz = 0;
    for (int j = 0; j < n_ini; ++j) if (s.node[j].count > 0) s.base[sz++] = static_cast<unsigned short>(j);  // empty cells are erased
    s.size = s.base_len = sz;
    s.n_nodes = n_ini;
    s.seq = 0;
    s.n_front = 0;
    s.n_pend = 0;
  }
  __syncthreads();
  for (int k0 = k_lo; k0 < k_hi; k0 += 32) {
    const int k = k0 + lane;
    const int c = k < k_hi ? cell_of(k) : -1;
    for (int j = 0; j < n_ini; ++j) {
      const unsigned b = __ballot_sync(0xffffffffu, c == j);
      if (c == j) s.perm[s.warp_cnt[warp][j] + __popc(b & ((1u << lane) - 1u))] = static_cast<unsigned short>(k);
      __syncwarp();
      if (lane == 0) s.warp_cnt[warp][j] += __popc(b);
      __syncwarp();
    }
  }
  __syncthreads();

  // ---- regular rounds (:585-665)
  bool last_phase = false;
  for (int guard = 0; guard < 64; ++guard) {
    const int S = s.size;  // == base_len: the front stack is empty in this phase
    // (1) every splittable node of the list is partitioned, one warp per node
    for (int i = warp; i < S; i += kTreeWarps) {
      const int id = s.base[i];
      if (!s.node[id].no_more) warp_split(s, id, s.cnt[i]);
    }
    __syncthreads();
    const int n_nodes0 = s.n_nodes, seq0 = s.seq;
    // (2) pass A: number of children this round creates (they go in front of the nodes that stay)
    int G = 0;
    for (int base_i = 0; base_i < S; base_i += kTreeThreads) {
      const int i = base_i + tid;
      int c = 0;
      if (i < S && !s.node[s.base[i]].no_more)
        for (int q = 0; q < 4; ++q) c += s.cnt[i][q] > 0;
      int t;
      block_exscan(c, s.scan, &t);
      G += t;
    }
    int n_keep_total = 0;
    if (n_nodes0 + G > kTreeNodeCap || G + S > kTreeListCap) {
      if (tid == 0) { atomicOr(out_error, 1 << l); out_count[l] = 0; }
      return;
    }
    // (3) pass B: children in creation order = list order of the parents, quadrant order within a parent
    int tot_ch = 0, tot_sp = 0;
    for (int base_i = 0; base_i < S; base_i += kTreeThreads) {
      const int i = base_i + tid;
      int nch = 0, nsp = 0, nkeep = 0, pid = 0;
      TreeNode par;
      bool split = false;
      if (i < S) {
        pid = s.base[i];
        par = s.node[pid];
        split = !par.no_more;
        if (split) for (int q = 0; q < 4; ++q) { nch += s.cnt[i][q] > 0; nsp += s.cnt[i][q] > 1; }
        else nkeep = 1;
      }
      int t_ch, t_sp, t_keep;
      const int e_ch = block_exscan(nch, s.scan, &t_ch);
      const int e_sp = block_exscan(nsp, s.scan, &t_sp);
      const int e_keep = block_exscan(nkeep, s.scan, &t_keep);
      if (i < S) {
        if (split) {
          int g = tot_ch + e_ch, sp = tot_sp + e_sp, start = 0;
          for (int q = 0; q < 4; ++q) {
            const int c = s.cnt[i][q];
            if (c) {
              TreeNode ch = child_of(par, q, start, c);
              const int cid = n_nodes0 + g;
              if (c > 1) {
                ch.seq = static_cast<unsigned short>(seq0 + sp + 1);
                s.seq2node[ch.seq] = static_cast<unsigned short>(cid);
                s.pend[sp] = static_cast<unsigned short>(cid);
                ++sp;
              }
              s.node[cid] = ch;
              s.base2[G - 1 - g] = static_cast<unsigned short>(cid);
              ++g;
            }
            start += c;
          }
          s.node[pid].alive = 0;
        } else {
          s.base2[G + n_keep_total + e_keep] = static_cast<unsigned short>(pid);
        }
      }
      tot_ch += t_ch; tot_sp += t_sp; n_keep_total += t_keep;
    }
    __syncthreads();
    const int S_new = G + n_keep_total;
    for (int i = tid; i < S_new; i += kTreeThreads) s.base[i] = s.base2[i];
    __syncthreads();
    if (tid == 0) {
      s.n_nodes = n_nodes0 + G;
      s.seq = seq0 + tot_sp;
      s.size = s.base_len = S_new;
      s.n_pend = tot_sp;
    }
    __syncthreads();
    if (S_new >= n_target || S_new == S) break;                           // (:659-662)
    if (S_new + 3 * tot_sp > n_target) { last_phase = true; break; }      // (:663)
  }

  // ---- last phase (:665-732): largest nodes first, one split at a time, until the target is reached
  while (last_phase) {
    const int R = s.n_pend, prev = s.size;
    // sort the round by (count, seq) ascending (std::sort on (size, pointer) pairs in the reference; ties by creation order here)
    int P = 1;
    while (P < R) P <<= 1;
    for (int i = tid; i < P; i += kTreeThreads)
      s.sortkey[i] = i < R ? (static_cast<unsigned>(s.node[s.pend[i]].count) << 16) | s.node[s.pend[i]].seq : 0xFFFFFFFFu;
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < P; i += kTreeThreads) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const unsigned a = s.sortkey[i], b = s.sortkey[ixj];
            const bool up = (i & k) == 0;
            if ((a > b) == up) { s.sortkey[i] = b; s.sortkey[ixj] = a; }
          }
        }
        __syncthreads();
      }
    if (warp == 0) {
      int size = s.size, n_nodes = s.n_nodes, seq = s.seq, nf = s.n_front, np2 = 0;
      bool overflow = false;
      for (int j = R - 1; j >= 0; --j) {
        const int id = s.seq2node[s.sortkey[j] & 0xFFFFu];
        warp_split(s, id, s.c4);
        const TreeNode par = s.node[id];
        int start = 0, made = 0;
        for (int q = 0; q < 4; ++q) {
          const int c = s.c4[q];
          if (c) {
            if (n_nodes >= kTreeNodeCap || nf >= kTreeListCap || np2 >= kTreeListCap) { overflow = true; break; }
            if (lane == 0) {
              TreeNode ch = child_of(par, q, start, c);
              if (c > 1) {
                ch.seq = static_cast<unsigned short>(seq + 1);
                s.seq2node[seq + 1] = static_cast<unsigned short>(n_nodes);
                s.pend2[np2] = static_cast<unsigned short>(n_nodes);
              }
              s.node[n_nodes] = ch;
              s.front[nf] = static_cast<unsigned short>(n_nodes);  // push_front
            }
            if (c > 1) { ++seq; ++np2; }
            ++n_nodes; ++nf; ++made;
          }
          start += c;
        }
        if (overflow) break;
        if (lane == 0) s.node[id].alive = 0;  // erase
        size += made - 1;
        __syncwarp();
        if (size >= n_target) break;
      }
      if (lane == 0) {
        s.size = size; s.n_nodes = n_nodes; s.seq = seq; s.n_front = nf; s.n_pend2 = np2;
        if (overflow) s.error = 1;
        s.finish = (size >= n_target || size == prev || overflow) ? 1 : 0;
      }
    }
    __syncthreads();
    if (s.error) {
      if (tid == 0) { atomicOr(out_error, 1 << l); out_count[l] = 0; }
      return;
    }
    if (s.finish) break;
    const int np2 = s.n_pend2;
    __syncthreads();
    for (int i = tid; i < np2; i += kTreeThreads) s.pend[i] = s.pend2[i];
    if (tid == 0) s.n_pend = np2;
    __syncthreads();
  }

  // ---- output (:735-748): list = reverse(front) ++ base, alive nodes only; per node the first key with the largest response
  const int nf = s.n_front, E = nf + s.base_len;
  int written = 0;
  for (int base_e = 0; base_e < E; base_e += kTreeThreads) {
    const int e = base_e + tid;
    int id = -1;
    if (e < E) {
      id = e < nf ? s.front[nf - 1 - e] : s.base[e - nf];
      if (!s.node[id].alive) id = -1;
    }
    int t;
    const int pos = written + block_exscan(id >= 0 ? 1 : 0, s.scan, &t);
    if (id >= 0 && pos < kTreeSelCap) {
      const TreeNode nd = s.node[id];
      int best = s.perm[nd.begin];
      for (int k = 1; k < nd.count; ++k) {
        const int kk = s.perm[nd.begin + k];
        if (s.kr[kk] > s.kr[best]) best = kk;
      }
      out_sel[l * kTreeSelCap + pos] = static_cast<uint32_t>(s.kx[best]) | (static_cast<uint32_t>(s.ky[best]) << 12) |
                                       (static_cast<uint32_t>(s.kr[best]) << 24);
    }
    written += t;
  }
  if (tid == 0) {
    if (written > kTreeSelCap) { atomicOr(out_error, 1 << l); written = 0; }
    out_count[l] = written;
  }
}

// Level-major concatenation of the per-level selections: OrbSelected (level coordinates, for k_describe) and the finished
// keypoint records (ComputeKeyPointsOctTree :824-835 + the final scaling of operator() :1071-1078); angle is filled by k_describe.
__global__ void k_finalize_keypoints(const uint32_t* __restrict__ sel_packed, const int* __restrict__ level_count, OrbTreeParams prm,
                                     int nlevels, int cap, OrbSelected* __restrict__ sel, sivo_keypoint* __restrict__ kps,
                                     int* __restrict__ n_out, long long* __restrict__ n_out_i64, int* __restrict__ out_error) {
  __shared__ int off[kOrbMaxLevels + 1];
  if (threadIdx.x == 0) {
    int run = 0;
    for (int l = 0; l < nlevels; ++l) { off[l] = run; run += level_count[l]; }
    off[nlevels] = run;
    if (run > cap) { atomicOr(out_error, 1 << 30); run = 0; off[nlevels] = 0; }
    *n_out = run;
    if (n_out_i64) *n_out_i64 = run;
  }
  __syncthreads();
  if (off[nlevels] == 0) return;
  for (int l = 0; l < nlevels; ++l) {
    const int n = off[l + 1] - off[l];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t c = sel_packed[l * kTreeSelCap + i];
      const int x = static_cast<int>(c & 0xFFF) + prm.min_b, y = static_cast<int>((c >> 12) & 0xFFF) + prm.min_b;
      OrbSelected sk;
      sk.x = static_cast<short>(x); sk.y = static_cast<short>(y); sk.level = static_cast<short>(l); sk.pad = 0;
      sel[off[l] + i] = sk;
      sivo_keypoint kp;
      kp.x = static_cast<float>(x);
      kp.y = static_cast<float>(y);
      if (l != 0) { kp.x = __fmul_rn(kp.x, prm.scale[l]); kp.y = __fmul_rn(kp.y, prm.scale[l]); }
      kp.size = prm.size[l];
      kp.angle = -1.f;
      kp.response = static_cast<float>(c >> 24);
      kp.octave = l;
      kp.class_id = -1;
      kps[off[l] + i] = kp;
    }
  }
}

}  // namespace

void orb_tree_configure() {
  static_assert(sizeof(TreeSmem) <= 227 * 1024, "the tree must fit one CTA's shared memory");
  SIVO_CUDA(cudaFuncSetAttribute(k_distribute, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(TreeSmem))));
}

void orb_launch_distribute(const uint32_t* cand, const int* level_off, const OrbTreeParams& prm, int nlevels, uint32_t* sel_packed,
                           int* level_count, int* error, cudaStream_t s) {
  k_distribute<<<nlevels, kTreeThreads, sizeof(TreeSmem), s>>>(cand, level_off, prm, sel_packed, level_count, error);
  SIVO_CUDA(cudaGetLastError());
}

void orb_launch_finalize(const uint32_t* sel_packed, const int* level_count, const OrbTreeParams& prm, int nlevels, int cap,
                         OrbSelected* sel, sivo_keypoint* kps, int* n_out, long long* n_out_i64, int* error, cudaStream_t s) {
  k_finalize_keypoints<<<1, 256, 0, s>>>(sel_packed, level_count, prm, nlevels, cap, sel, kps, n_out, n_out_i64, error);
  SIVO_CUDA(cudaGetLastError());
}

}  // namespace sivo
