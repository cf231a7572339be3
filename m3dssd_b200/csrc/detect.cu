// Detection tail on the device: score top-K selection + sort, anchor decode,
// bitmask NMS (warp-ballot) with an on-device greedy sweep.
//
// Reference: lib/rpn_util.py:1444-1555 (im_detect_3d: de-normalise, decode,
// argsort(-score), top nms_topN_pre, gpu_nms, gather), lib/nms/nms_kernel.cu
// (devIoU :24-32, bitmask kernel :34-78, host sweep :124-139) and
// lib/nms/gpu_nms.pyx:16-31.  The reference handles one image per call with a
// host round trip (D2H of 3000 boxes, cudaMalloc/cudaFree, 1.1 MB mask D2H, CPU
// sweep); here everything stays on the device and is batched over images.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace m3d {

constexpr int kSelThreads = 1024;
constexpr int kMaxTopK = 4096;

// ------------------------------------------------------------------ top-K
// Exact top-K of `score` (descending; ties -> lower index first) for one image
// per block: 3-pass radix select on the fp32 bit pattern (scores are >= 0, so
// the unsigned pattern is order-preserving), ordered compaction, bitonic sort.
__device__ __forceinline__ uint32_t score_key(float s) {
  uint32_t u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // total order for any sign
}

// Descending scan of a (1 << BITS)-bin histogram in shared memory (all threads of a 1024-thread block): the digit d
// at which the running count from the top reaches `need`, how many are still needed inside d, and hist[d].
// Two levels: 32 warp partial sums, then inside the crossing group.
template <int BITS>
__device__ void pick_digit(const uint32_t* hist, int need, uint32_t* out_digit, int* out_need, int* out_count) {
  constexpr int NB = 1 << BITS;
  constexpr int PER = NB / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ uint32_t s_part[32];
  uint32_t v = 0;
  if (warp < 32)
    for (int i = lane; i < PER; i += 32) v += hist[warp * PER + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0 && warp < 32) s_part[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    int g = 31;
    for (; g > 0; --g) {
      if (acc + static_cast<int>(s_part[g]) >= need) break;
      acc += s_part[g];
    }
    int d = g * PER + PER - 1;
    for (; d > g * PER; --d) {
      if (acc + static_cast<int>(hist[d]) >= need) break;
      acc += hist[d];
    }
    *out_digit = static_cast<uint32_t>(d);
    *out_need = need - acc;  // how many still to take from inside digit d
    *out_count = static_cast<int>(hist[d]);
  }
  __syncthreads();
}

// One histogram pass over the image's scores.  8 elements per thread per trip (two 16-byte loads in
// flight), same-bin lanes of a warp combined with match_any before the shared-memory atomic.
template <int BITS>
__device__ void radix_pass(const float* __restrict__ score, int M, uint32_t prefix, uint32_t prefix_mask, int shift,
                           uint32_t* hist /*smem [1<<BITS]*/, int need, uint32_t* out_digit, int* out_need,
                           int* out_count) {
  constexpr int NB = 1 << BITS;
  for (int i = threadIdx.x; i < NB; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  auto add = [&](float v, bool live) {
    const uint32_t k = score_key(v);
    const bool in = live && (k & prefix_mask) == prefix;
    const uint32_t bin = (k >> shift) & (NB - 1);
    const uint32_t act = __ballot_sync(0xffffffffu, in);
    if (in) {
      const uint32_t peers = __match_any_sync(act, bin);
      if (lane == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
    }
  };
  const int M8 = M & ~7;
  for (int i0 = threadIdx.x * 8; i0 < ((M8 + blockDim.x * 8 - 1) / (blockDim.x * 8)) * (blockDim.x * 8);
       i0 += blockDim.x * 8) {
    const bool live = i0 < M8;  // uniform trip count: the warp-collective ops above need every lane
    float4 a = make_float4(0, 0, 0, 0), b = a;
    if (live) {
      a = __ldg(reinterpret_cast<const float4*>(score + i0));
      b = __ldg(reinterpret_cast<const float4*>(score + i0) + 1);
    }
    add(a.x, live), add(a.y, live), add(a.z, live), add(a.w, live);
    add(b.x, live), add(b.y, live), add(b.z, live), add(b.w, live);
  }
  for (int i0 = M8; i0 < M8 + 32; i0 += 32) {  // tail (< 8 elements)
    const int i = i0 + lane;
    const bool live = threadIdx.x < 32 && i < M;
    if (threadIdx.x < 32) add(live ? score[i] : 0.f, live);
  }
  __syncthreads();
  pick_digit<BITS>(hist, need, out_digit, out_need, out_count);
}

struct DecodeParams {
  const float* score;         // [N, M]
  const unsigned char* cls;   // [N, M]
  const float* bbox_2d;       // [N, M, 4]
  const float* bbox_3d;       // [N, M, 7]
  const float* heads;         // or (bbox_2d == nullptr): the NHWC head buffer [N, H, W, heads_cstride], column of output
  int heads_cstride;          //   j of anchor a = slot[j] * A + a (j: x,y,w,h,x3d,y3d,z3d,w3d,h3d,l3d,rY3d)
  int slot[11];
  const float* anchors;       // [A, 9]
  float means[11], stds[11];
  int M, A, H, W;
  float feat_stride, scale_factor;
  int topk;
  float* dets;   // [N, topk, 14]
  int* det_idx;  // [N, topk] flattened anchor index of each row (or -1)
  int* det_num;  // [N]
};

// Exact selection by three radix passes + one compaction pass over all M scores (any input).
__device__ void topk_select_full(const DecodeParams& p, int n, uint32_t* hist, unsigned long long* keys) {
  __shared__ uint32_t s_digit;
  __shared__ int s_need, s_count, s_cnt_gt, s_cnt_eq;
  __shared__ int s_warp_gt[32], s_warp_eq[32];
  const float* score = p.score + static_cast<long>(n) * p.M;
  const int K = min(p.topk, p.M);

  // ---- threshold key T: the K-th largest
  uint32_t prefix = 0, mask = 0;
  int need = K;
  radix_pass<11>(score, p.M, prefix, mask, 21, hist, need, &s_digit, &s_need, &s_count);
  prefix |= s_digit << 21, mask |= 0x7FFu << 21, need = s_need;
  radix_pass<11>(score, p.M, prefix, mask, 10, hist, need, &s_digit, &s_need, &s_count);
  prefix |= s_digit << 10, mask |= 0x7FFu << 10, need = s_need;
  radix_pass<10>(score, p.M, prefix, mask, 0, hist, need, &s_digit, &s_need, &s_count);
  const uint32_t T = prefix | s_digit;
  const int need_eq = s_need;        // number of elements equal to T to take (lowest indices)
  const int count_eq = s_count;      // number of elements equal to T in the image

  if (threadIdx.x == 0) s_cnt_gt = 0, s_cnt_eq = 0;
  for (int i = threadIdx.x; i < kMaxTopK; i += blockDim.x) keys[i] = 0ull;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (count_eq == need_eq) {
    // ---- fast path (no tie straddles the cut): every key >= T is selected, order is irrelevant here
    // because the bitonic sort below orders by (key, index).  Warp-aggregated slot allocation.
    const int M4 = p.M & ~3;
    for (int i0 = threadIdx.x * 4; i0 < ((p.M + blockDim.x * 4 - 1) / (blockDim.x * 4)) * (blockDim.x * 4);
         i0 += blockDim.x * 4) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (i0 < M4) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(score + i0));
        v[0] = f.x, v[1] = f.y, v[2] = f.z, v[3] = f.w;
      } else {
        for (int e = 0; e < 4; ++e)
          if (i0 + e < p.M) v[e] = score[i0 + e];
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = i0 + e;
        const uint32_t k = score_key(v[e]);
        const bool take = i < p.M && k >= T;
        const uint32_t b = __ballot_sync(0xffffffffu, take);
        if (b) {
          int base = 0;
          if (lane == 0) base = atomicAdd(&s_cnt_gt, __popc(b));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (take)
            keys[base + __popc(b & ((1u << lane) - 1))] =
                (static_cast<unsigned long long>(k) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(i));
        }
      }
    }
    __syncthreads();
  } else {
    // ---- ties straddle the cut: ordered compaction, keys > T (all) and == T (first need_eq by index)
    for (int base = 0; base < p.M; base += kSelThreads) {
      const int i = base + threadIdx.x;
      uint32_t k = 0;
      bool gt = false, eq = false;
      if (i < p.M) {
        k = score_key(score[i]);
        gt = k > T;
        eq = k == T;
      }
      const uint32_t bgt = __ballot_sync(0xffffffffu, gt), beq = __ballot_sync(0xffffffffu, eq);
      if (lane == 0) s_warp_gt[warp] = __popc(bgt), s_warp_eq[warp] = __popc(beq);
      __syncthreads();
      int off_gt = s_cnt_gt, off_eq = s_cnt_eq;
      for (int w = 0; w < warp; ++w) off_gt += s_warp_gt[w], off_eq += s_warp_eq[w];
      const uint32_t lt = (1u << lane) - 1;
      const unsigned long long packed =
          (static_cast<unsigned long long>(k) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(i));
      if (gt) keys[off_gt + __popc(bgt & lt)] = packed;  // < K - need_eq by construction
      if (eq) {
        const int r = off_eq + __popc(beq & lt);
        if (r < need_eq) keys[(K - need_eq) + r] = packed;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int tg = 0, te = 0;
        for (int w = 0; w < kSelThreads / 32; ++w) tg += s_warp_gt[w], te += s_warp_eq[w];
        s_cnt_gt += tg, s_cnt_eq += te;
      }
      __syncthreads();
    }
  }

}

__device__ void topk_sort_decode(const DecodeParams& p, int n, unsigned long long* keys) {
  const float* score = p.score + static_cast<long>(n) * p.M;
  const int K = min(p.topk, p.M);
  // ---- bitonic sort, descending on (key, ~index): higher score first, lower index first on ties
  for (int size = 2; size <= kMaxTopK; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < kMaxTopK / 2; t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a < b) == desc) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
      __syncthreads();
    }
  }

  // ---- decode the K selected boxes (lib/rpn_util.py:1462-1521, 1137-1186)
  if (threadIdx.x == 0) p.det_num[n] = K;
  const int HW = p.H * p.W;
  for (int r = threadIdx.x; r < p.topk; r += blockDim.x) {
    float* d = p.dets + (static_cast<long>(n) * p.topk + r) * 14;
    if (r >= K) {
      for (int j = 0; j < 14; ++j) d[j] = 0.f;
      p.det_idx[static_cast<long>(n) * p.topk + r] = -1;
      continue;
    }
    const int idx = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(keys[r] & 0xFFFFFFFFull));
    p.det_idx[static_cast<long>(n) * p.topk + r] = idx;
    const int a = idx / HW, h = (idx / p.W) % p.H, w = idx % p.W;
    const float* an = p.anchors + a * 9;
    // rois (locate_anchors, lib/rpn_util.py:1345-1386): float64 shift + float32 anchor, rounded to float32
    const float rx1 = static_cast<float>(static_cast<double>(w) * p.feat_stride + static_cast<double>(an[0]));
    const float ry1 = static_cast<float>(static_cast<double>(h) * p.feat_stride + static_cast<double>(an[1]));
    const float rx2 = static_cast<float>(static_cast<double>(w) * p.feat_stride + static_cast<double>(an[2]));
    const float ry2 = static_cast<float>(static_cast<double>(h) * p.feat_stride + static_cast<double>(an[3]));
    const float widths = rx2 - rx1 + 1.0f, heights = ry2 - ry1 + 1.0f;
    const float ctr_x = rx1 + 0.5f * widths, ctr_y = ry1 + 0.5f * heights;
    float b2[4], b3[7];
    if (p.bbox_2d != nullptr) {
      const float* s2 = p.bbox_2d + (static_cast<long>(n) * p.M + idx) * 4;
      const float* s3 = p.bbox_3d + (static_cast<long>(n) * p.M + idx) * 7;
      for (int j = 0; j < 4; ++j) b2[j] = s2[j];
      for (int j = 0; j < 7; ++j) b3[j] = s3[j];
    } else {  // the same values, before flatten_tensor (lib/rpn_util.py:892-901) moved them
      const float* hp = p.heads + ((static_cast<long>(n) * p.H + h) * p.W + w) * p.heads_cstride + a;
      for (int j = 0; j < 4; ++j) b2[j] = hp[p.slot[j] * p.A];
      for (int j = 0; j < 7; ++j) b3[j] = hp[p.slot[4 + j] * p.A];
    }
    float t3[7];
    for (int j = 0; j < 7; ++j) t3[j] = b3[j] * p.stds[4 + j] + p.means[4 + j];
    const float x3d = t3[0] * widths + ctr_x;
    const float y3d = t3[1] * heights + ctr_y;
    const float z3d = an[4] + t3[2];
    const float w3d = expf(t3[3]) * an[5];
    const float h3d = expf(t3[4]) * an[6];
    const float l3d = expf(t3[5]) * an[7];
    const float ry3d = an[8] + t3[6];
    const float dx = b2[0] * p.stds[0] + p.means[0];
    const float dy = b2[1] * p.stds[1] + p.means[1];
    const float dw = b2[2] * p.stds[2] + p.means[2];
    const float dh = b2[3] * p.stds[3] + p.means[3];
    const float pcx = dx * widths + ctr_x, pcy = dy * heights + ctr_y;
    const float pw = expf(dw) * widths, ph = expf(dh) * heights;
    d[0] = (pcx - 0.5f * pw) / p.scale_factor;
    d[1] = (pcy - 0.5f * ph) / p.scale_factor;
    d[2] = (pcx + 0.5f * pw) / p.scale_factor;
    d[3] = (pcy + 0.5f * ph) / p.scale_factor;
    d[4] = score[idx];
    d[5] = static_cast<float>(p.cls[static_cast<long>(n) * p.M + idx]);
    d[6] = x3d / p.scale_factor;
    d[7] = y3d / p.scale_factor;
    d[8] = z3d, d[9] = w3d, d[10] = h3d, d[11] = l3d, d[12] = ry3d;
    d[13] = static_cast<float>(a);
  }
}

__global__ void __launch_bounds__(kSelThreads) topk_decode_kernel(const DecodeParams p) {
  __shared__ uint32_t hist[2048];
  __shared__ unsigned long long keys[kMaxTopK];
  topk_select_full(p, blockIdx.x, hist, keys);
  topk_sort_decode(p, blockIdx.x, keys);
}

// ---------------------------------------------------------------- multi-CTA top-K
// The single-CTA kernel above reads the image's 276 480 scores four times from one SM (0.68 ms for a batch of 8,
// 8 SMs busy).  Here the two full passes are spread over the device and the rest works on a short candidate list:
//   1. topk_hist_kernel<0>  (S slices x N images): histogram of the keys' top 11 bits -> global
//      topk_hist_kernel<1>  (S x N): histogram of bits [20:10] of the keys inside the threshold digit d1
//   2. topk_compact_kernel  (S x N): every CTA re-derives the 22-bit threshold prefix from the global histograms and
//                           appends its keys >= that prefix to the image's candidate list (packed key|~index): K plus
//                           at most one 2^-14-wide bin of scores
//   3. topk_finish_kernel   (N): radix pass 3 over the candidates of that bin, selection of every key >= T,
//                           bitonic sort, decode.  Falls back to the full single-CTA selection when the candidate
//                           list or the tie group at T overflows (pathological score distributions): always exact.
constexpr int kTopkSlices = 16;
constexpr int kCandCap = 32768;

template <int LEVEL>
__global__ void __launch_bounds__(kSelThreads) topk_hist_kernel(const float* __restrict__ score_all, int M, int topk,
                                                                uint32_t* __restrict__ hist_g) {
  __shared__ uint32_t hist[2048];
  __shared__ uint32_t s_digit;
  __shared__ int s_need, s_count;
  const int n = blockIdx.y;
  const float* score = score_all + static_cast<long>(n) * M;
  uint32_t d1 = 0;
  if (LEVEL == 1) {  // re-derive the first digit from the finished level-0 histogram
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) hist[i] = hist_g[(n * 2) * 2048 + i];
    __syncthreads();
    pick_digit<11>(hist, min(topk, M), &s_digit, &s_need, &s_count);
    d1 = s_digit;
  }
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const int chunk = (((M + kTopkSlices - 1) / kTopkSlices) + 3) & ~3;
  const int lo = blockIdx.x * chunk, hi = min(M, lo + chunk);
  for (int i0 = lo + threadIdx.x * 4; i0 < hi; i0 += blockDim.x * 4) {
    float v[4];
    int cnt = min(4, hi - i0);
    if (cnt == 4 && ((reinterpret_cast<uintptr_t>(score + i0) & 15) == 0)) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(score + i0));
      v[0] = f.x, v[1] = f.y, v[2] = f.z, v[3] = f.w;
    } else {
      for (int e = 0; e < 4; ++e) v[e] = e < cnt ? score[i0 + e] : 0.f;
    }
    for (int e = 0; e < cnt; ++e) {
      const uint32_t k = score_key(v[e]);
      if (LEVEL == 0)
        atomicAdd(&hist[k >> 21], 1u);
      else if ((k >> 21) == d1)
        atomicAdd(&hist[(k >> 10) & 0x7FFu], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += blockDim.x)
    if (hist[i]) atomicAdd(&hist_g[(n * 2 + LEVEL) * 2048 + i], hist[i]);
}

__global__ void __launch_bounds__(kSelThreads) topk_compact_kernel(const float* __restrict__ score_all, int M, int topk,
                                                                   const uint32_t* __restrict__ hist_g,
                                                                   uint32_t* __restrict__ count_g,
                                                                   unsigned long long* __restrict__ cand_g) {
  __shared__ uint32_t hist[2048];
  __shared__ uint32_t s_digit;
  __shared__ int s_need, s_count;
  const int n = blockIdx.y;
  const float* score = score_all + static_cast<long>(n) * M;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) hist[i] = hist_g[(n * 2) * 2048 + i];
  __syncthreads();
  pick_digit<11>(hist, min(topk, M), &s_digit, &s_need, &s_count);
  const uint32_t d1 = s_digit;
  const int need1 = s_need;
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) hist[i] = hist_g[(n * 2 + 1) * 2048 + i];
  __syncthreads();
  pick_digit<11>(hist, need1, &s_digit, &s_need, &s_count);
  const uint32_t p22 = (d1 << 11) | s_digit;  // threshold on the keys' top 22 bits
  const int lane = threadIdx.x & 31;
  unsigned long long* cand = cand_g + static_cast<long>(n) * kCandCap;
  const int chunk = (((M + kTopkSlices - 1) / kTopkSlices) + 3) & ~3;
  const int lo = blockIdx.x * chunk, hi = min(M, lo + chunk);
  const int span = blockDim.x * 4;
  for (int base = lo; base < hi; base += span) {  // uniform trip count per block: the ballots need every lane
    const int i0 = base + threadIdx.x * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const int cnt = max(0, min(4, hi - i0));
    if (cnt == 4 && ((reinterpret_cast<uintptr_t>(score + i0) & 15) == 0)) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(score + i0));
      v[0] = f.x, v[1] = f.y, v[2] = f.z, v[3] = f.w;
    } else {
      for (int e = 0; e < cnt; ++e) v[e] = score[i0 + e];
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint32_t k = score_key(v[e]);
      const bool take = e < cnt && (k >> 10) >= p22;
      const uint32_t b = __ballot_sync(0xffffffffu, take);
      if (b) {
        uint32_t slot = 0;
        if (lane == 0) slot = atomicAdd(&count_g[n], static_cast<uint32_t>(__popc(b)));
        slot = __shfl_sync(0xffffffffu, slot, 0) + __popc(b & ((1u << lane) - 1));
        if (take && slot < kCandCap)
          cand[slot] = (static_cast<unsigned long long>(k) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(i0 + e));
      }
    }
  }
}

__global__ void __launch_bounds__(kSelThreads) topk_finish_kernel(const DecodeParams p, const uint32_t* __restrict__ hist_g,
                                                                  const uint32_t* __restrict__ count_g,
                                                                  const unsigned long long* __restrict__ cand_g) {
  __shared__ uint32_t hist[2048];
  __shared__ unsigned long long keys[kMaxTopK];
  __shared__ uint32_t s_digit;
  __shared__ int s_need, s_count, s_slots;
  const int n = blockIdx.x;
  const int K = min(p.topk, p.M);
  const uint32_t nc = count_g[n];
  bool fallback = nc > static_cast<uint32_t>(kCandCap);  // block-uniform
  if (!fallback) {
    const unsigned long long* cand = cand_g + static_cast<long>(n) * kCandCap;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) hist[i] = hist_g[(n * 2) * 2048 + i];
    __syncthreads();
    pick_digit<11>(hist, K, &s_digit, &s_need, &s_count);
    uint32_t prefix = s_digit << 21, mask = 0x7FFu << 21;
    int need = s_need;
    __syncthreads();
    // pass 2 was taken device-wide (level-1 histogram)
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) hist[i] = hist_g[(n * 2 + 1) * 2048 + i];
    __syncthreads();
    pick_digit<11>(hist, need, &s_digit, &s_need, &s_count);
    prefix |= s_digit << 10, mask |= 0x7FFu << 10, need = s_need;
    __syncthreads();
    // pass 3: bits [9:0]
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nc; i += blockDim.x) {
      const uint32_t k = static_cast<uint32_t>(cand[i] >> 32);
      if ((k & mask) == prefix) atomicAdd(&hist[k & 0x3FFu], 1u);
    }
    __syncthreads();
    pick_digit<10>(hist, need, &s_digit, &s_need, &s_count);
    const uint32_t T = prefix | s_digit;
    const int need_eq = s_need, count_eq = s_count;
    // every key >= T goes to the sort buffer (ties at T included: the sort orders them by index and the first K
    // rows are the answer); a tie group that does not fit falls back to the ordered full selection
    fallback = (K - need_eq) + count_eq > kMaxTopK;
    if (!fallback) {
      if (threadIdx.x == 0) s_slots = 0;
      for (int i = threadIdx.x; i < kMaxTopK; i += blockDim.x) keys[i] = 0ull;
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < nc; i += blockDim.x) {
        const unsigned long long c = cand[i];
        if (static_cast<uint32_t>(c >> 32) >= T) keys[atomicAdd(&s_slots, 1)] = c;
      }
      __syncthreads();
    }
  }
  if (fallback) topk_select_full(p, n, hist, keys);
  topk_sort_decode(p, n, keys);
}

// scratch of the multi-CTA top-K: hist [N][2][2048] + [N] counts, then cand [N][kCandCap].  CALLER-owned (the engine
// allocates it next to its activations): a process-global buffer would be baked into captured CUDA graphs and freed
// under them as soon as a second engine with a larger batch resized it.
static size_t topk_hist_bytes(int batch) { return (static_cast<size_t>(batch) * 4097 * sizeof(uint32_t) + 255) & ~size_t(255); }

// -------------------------------------------------------------------- NMS
// IoU with the +1 pixel convention, in the exact operation order the
// reference's kernel has when built with nvcc's default -fmad=true
// (FMUL, FFMA(wb, hb, Sa), FMUL, FADD, IEEE divide) so `> thresh` never flips.
__device__ __forceinline__ float dev_iou(const float4 a, const float4 b) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float width = fmaxf(__fadd_rn(__fsub_rn(right, left), 1.f), 0.f);
  const float height = fmaxf(__fadd_rn(__fsub_rn(bottom, top), 1.f), 0.f);
  const float interS = __fmul_rn(width, height);
  const float Sa = __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), 1.f), __fadd_rn(__fsub_rn(a.w, a.y), 1.f));
  const float SaSb = __fmaf_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f), Sa);
  return __fdiv_rn(interS, __fsub_rn(SaSb, interS));
}

// IoU(a, b) > thresh, bit for bit like dev_iou(a, b) > thresh: boxes that do not intersect have interS == 0, hence
// IoU == +0 (the union is >= 1 pixel) and the comparison is false for any thresh >= 0 -- which skips the IEEE
// division for the vast majority of pairs (the mask kernel is issue bound on it).
__device__ __forceinline__ bool iou_above(const float4 a, const float4 b, float thresh) {
  const float width = fmaxf(__fadd_rn(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 1.f), 0.f);
  const float height = fmaxf(__fadd_rn(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 1.f), 0.f);
  if (thresh >= 0.f && (width == 0.f || height == 0.f)) return false;
  return dev_iou(a, b) > thresh;
}

// mask[n][i][cb] bit k  <=>  j = cb*64 + k > i  and  IoU(box_i, box_j) > thresh.
// Block = 4 warps; block (cb, rb, n) covers rows rb*64..+63 against columns
// cb*64..+63 (upper triangle only).  Lanes own columns (2 per lane), rows are
// broadcast from shared memory, one ballot yields 32 mask bits.
__global__ void __launch_bounds__(128) nms_mask_kernel(const float* __restrict__ boxes, int box_stride,
                                                       const int* __restrict__ num, int max_n, float thresh,
                                                       unsigned long long* __restrict__ mask, int col_blocks) {
  const int cb = blockIdx.x, rb = blockIdx.y, n = blockIdx.z;
  if (cb < rb) return;
  const int nb = num ? num[n] : max_n;
  if (rb * 64 >= nb || cb * 64 >= nb) return;
  __shared__ float4 rows[64];
  const float* b = boxes + static_cast<long>(n) * max_n * box_stride;
  if (threadIdx.x < 64) {
    const int i = rb * 64 + threadIdx.x;
    rows[threadIdx.x] = i < nb ? make_float4(b[i * box_stride], b[i * box_stride + 1], b[i * box_stride + 2], b[i * box_stride + 3])
                               : make_float4(0, 0, 0, 0);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j0 = cb * 64 + lane, j1 = j0 + 32;
  const float4 c0 = j0 < nb ? make_float4(b[j0 * box_stride], b[j0 * box_stride + 1], b[j0 * box_stride + 2], b[j0 * box_stride + 3])
                            : make_float4(0, 0, 0, 0);
  const float4 c1 = j1 < nb ? make_float4(b[j1 * box_stride], b[j1 * box_stride + 1], b[j1 * box_stride + 2], b[j1 * box_stride + 3])
                            : make_float4(0, 0, 0, 0);
  __syncthreads();
  for (int r = warp; r < 64; r += 4) {
    const int i = rb * 64 + r;
    if (i >= nb) break;
    const float4 a = rows[r];
    const bool s0 = (j0 < nb) && (j0 > i) && iou_above(a, c0, thresh);
    const bool s1 = (j1 < nb) && (j1 > i) && iou_above(a, c1, thresh);
    const uint32_t lo = __ballot_sync(0xffffffffu, s0), hi = __ballot_sync(0xffffffffu, s1);
    if (lane == 0)
      mask[(static_cast<long>(n) * max_n + i) * col_blocks + cb] = (static_cast<unsigned long long>(hi) << 32) | lo;
  }
}

// Greedy sweep (nms_kernel.cu:124-139), one 1024-thread block per image, 256 boxes (4 mask words) per round:
//   1. the round's 256 x 256 diagonal sub-matrix (256 rows x 4 words) is staged in shared memory (fetched under the
//      previous round's OR phase);
//   2. thread 0 resolves the 256 boxes in order from shared memory and registers only -- the dependent chain has
//      no global memory in it (the first version paid an L2 round trip per kept box: 0.5 ms per image);
//   3. all threads OR the kept rows' words of the LATER rounds into the shared "removed" words: thread ->
//      (word, row group), the kept rows come from a shared list, so a thread's loads are independent.
constexpr int kSweepThreads = 1024;
constexpr int kSweepRound = 4;  // 64-box blocks per round

__global__ void __launch_bounds__(kSweepThreads) nms_sweep_kernel(const unsigned long long* __restrict__ mask,
                                                                   const int* __restrict__ num, int max_n,
                                                                   int col_blocks, int* __restrict__ keep,
                                                                   int* __restrict__ num_keep) {
  __shared__ unsigned long long remv[64];
  __shared__ unsigned long long diag[64 * kSweepRound][kSweepRound];
  __shared__ unsigned long long s_keptbits[kSweepRound];
  __shared__ int s_kept;
  __shared__ short s_list[64 * kSweepRound];
  const int n = blockIdx.x, tid = threadIdx.x;
  const int nb = num ? num[n] : max_n;
  const unsigned long long* m = mask + static_cast<long>(n) * max_n * col_blocks;
  int* kp = keep + static_cast<long>(n) * max_n;
  const int nblocks = (nb + 63) / 64;
  const int nrounds = (nblocks + kSweepRound - 1) / kSweepRound;
  if (tid < 64) remv[tid] = 0ull;
  if (tid == 0) s_kept = 0;
  // diagonal words of round 0: thread -> (row, word)
  const int drow = tid >> 2, dw = tid & 3;  // 256 rows x 4 words = 1024 threads
  auto load_diag = [&](int round) -> unsigned long long {
    const int r = round * 64 * kSweepRound + drow, w = round * kSweepRound + dw;
    return (round < nrounds && r < nb && w < col_blocks) ? m[static_cast<long>(r) * col_blocks + w] : 0ull;
  };
  unsigned long long next_diag = load_diag(0);
  __syncthreads();
  for (int round = 0; round < nrounds; ++round) {
    diag[drow][dw] = next_diag;
    __syncthreads();
    next_diag = load_diag(round + 1);  // independent of this round's outcome
    if (tid == 0) {
      unsigned long long cur[kSweepRound];
#pragma unroll
      for (int q = 0; q < kSweepRound; ++q) cur[q] = round * kSweepRound + q < 64 ? remv[(round * kSweepRound + q) & 63] : ~0ull;
#pragma unroll
      for (int q = 0; q < kSweepRound; ++q) {
        const int base = (round * kSweepRound + q) * 64;
        const int lim = min(64, nb - base);
        unsigned long long kept = 0ull;
        unsigned long long c = cur[q];
#pragma unroll 8
        for (int b = 0; b < 64; ++b) {
          if (b < lim && !((c >> b) & 1ull)) {
            kept |= 1ull << b;
            c |= diag[q * 64 + b][q];
#pragma unroll
            for (int q2 = q + 1; q2 < kSweepRound; ++q2) cur[q2] |= diag[q * 64 + b][q2];
          }
        }
        s_keptbits[q] = kept;
      }
    }
    __syncthreads();
    // kept indices in order: thread (q, b) writes its own slot in the global list and the round-local row list
    const int base0 = s_kept;
    int nk = 0;
    {
      int before = 0;
#pragma unroll
      for (int q = 0; q < kSweepRound; ++q) {
        const unsigned long long kq = s_keptbits[q];
        if (tid >= q * 64 && tid < q * 64 + 64 && ((kq >> (tid & 63)) & 1ull)) {
          const int pos = before + __popcll(kq & ((1ull << (tid & 63)) - 1));
          kp[base0 + pos] = (round * kSweepRound + q) * 64 + (tid & 63);
          s_list[pos] = static_cast<short>(q * 64 + (tid & 63));
        }
        before += __popcll(kq);
      }
      nk = before;
    }
    __syncthreads();
    // OR the kept rows into the words of the later rounds
    const int w = tid & 63, g = tid >> 6;
    if (w >= (round + 1) * kSweepRound && w < col_blocks) {
      unsigned long long acc = 0ull;
      const unsigned long long* mrow = m + static_cast<long>(round) * 64 * kSweepRound * col_blocks + w;
#pragma unroll 4
      for (int i = g; i < nk; i += 16) acc |= mrow[static_cast<long>(s_list[i]) * col_blocks];
      if (acc) atomicOr(&remv[w], acc);
    }
    __syncthreads();
    if (tid == 0) s_kept = base0 + nk;
  }
  __syncthreads();
  if (tid == 0) num_keep[n] = s_kept;
}

// Greedy sweep for MORE than 4096 boxes per image (the fast kernel above keeps the removed-set in 64 shared words):
// the reference's host loop (nms_kernel.cu:124-139) on the device, one block per image, the removed-set in global
// memory.  One block barrier per kept box: slow, exact, and only reached through m3d_nms / m3d_nms_batched with
// max_n > 4096 (the reference accepts any box count; the engine's top-K is 3000).
__global__ void __launch_bounds__(256) nms_sweep_generic_kernel(const unsigned long long* __restrict__ mask,
                                                                const int* __restrict__ num, int max_n, int col_blocks,
                                                                unsigned long long* __restrict__ remv_all,
                                                                int* __restrict__ keep, int* __restrict__ num_keep) {
  const int n = blockIdx.x, tid = threadIdx.x;
  const int nb = num ? num[n] : max_n;
  const unsigned long long* m = mask + static_cast<long>(n) * max_n * col_blocks;
  unsigned long long* remv = remv_all + static_cast<long>(n) * col_blocks;
  int* kp = keep + static_cast<long>(n) * max_n;
  for (int w = tid; w < col_blocks; w += blockDim.x) remv[w] = 0ull;
  __syncthreads();
  int kept = 0;
  for (int i = 0; i < nb; ++i) {
    const bool removed = (remv[i >> 6] >> (i & 63)) & 1ull;  // same value in every thread (written before the barrier)
    if (!removed) {
      if (tid == 0) kp[kept] = i;
      ++kept;
      const unsigned long long* row = m + static_cast<long>(i) * col_blocks;
      for (int w = (i >> 6) + tid; w < col_blocks; w += blockDim.x) remv[w] |= row[w];
      __syncthreads();
    }
  }
  if (tid == 0) num_keep[n] = kept;
}

__global__ void gather_kept_kernel(const float* __restrict__ dets, int row_len, int max_n, const int* __restrict__ keep,
                                   const int* __restrict__ num_keep, int max_out, float* __restrict__ out) {
  const int n = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= max_out * row_len) return;
  const int r = i / row_len, c = i % row_len;
  float v = 0.f;
  if (r < num_keep[n]) v = dets[(static_cast<long>(n) * max_n + keep[static_cast<long>(n) * max_n + r]) * row_len + c];
  out[(static_cast<long>(n) * max_out + r) * row_len + c] = v;
}

// persistent device workspace for the host-pointer gpu_nms entry
struct NmsWorkspace {
  std::mutex mu;
  float* boxes = nullptr;
  unsigned long long* mask = nullptr;
  int* keep = nullptr;
  int* num_keep = nullptr;
  unsigned long long* remv = nullptr;  // removed-set of the generic sweep (> 4096 boxes)
  int cap = 0, device = -1;
};
static NmsWorkspace g_nms_ws;

}  // namespace m3d

using namespace m3d;

static inline cudaStream_t S(m3d_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" size_t m3d_nms_workspace_bytes(int batch, int max_n) {
  const size_t cb = (max_n + 63) / 64;
  // suppression bitmask [batch][max_n][cb] + (more than 4096 boxes per image: generic sweep) removed-set [batch][cb]
  return static_cast<size_t>(batch) * max_n * cb * sizeof(unsigned long long) +
         (cb > 64 ? static_cast<size_t>(batch) * cb * sizeof(unsigned long long) : 0);
}

extern "C" int m3d_nms_batched(const float* boxes, int box_stride, const int* num, int batch, int max_n, float thresh,
                               void* workspace, size_t workspace_bytes, int* keep, int* num_keep, m3d_stream_t stream) {
  M3D_REQUIRE(boxes && keep && num_keep && workspace, "NULL pointer");
  M3D_REQUIRE(batch >= 1 && max_n >= 1 && box_stride >= 4, "bad geometry");
  const int cb = (max_n + 63) / 64;
  if (workspace_bytes < m3d_nms_workspace_bytes(batch, max_n)) {
    set_last_error("NMS workspace too small: %zu < %zu", workspace_bytes, m3d_nms_workspace_bytes(batch, max_n));
    return M3D_ERR_WORKSPACE;
  }
  unsigned long long* mask = static_cast<unsigned long long*>(workspace);
  dim3 grid(cb, cb, batch);
  nms_mask_kernel<<<grid, 128, 0, S(stream)>>>(boxes, box_stride, num, max_n, thresh, mask, cb);
  M3D_CUDA_OK(cudaGetLastError());
  if (cb <= 64) {
    nms_sweep_kernel<<<batch, kSweepThreads, 0, S(stream)>>>(mask, num, max_n, cb, keep, num_keep);
  } else {
    unsigned long long* remv = mask + static_cast<size_t>(batch) * max_n * cb;
    nms_sweep_generic_kernel<<<batch, 256, 0, S(stream)>>>(mask, num, max_n, cb, remv, keep, num_keep);
  }
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}

// Drop-in for `_nms` (lib/nms/gpu_nms.hpp:1-2): HOST pointers in and out, boxes
// already sorted by score, synchronous.  Device buffers persist between calls.
extern "C" int m3d_nms(int* keep_out, int* num_out, const float* boxes_host, int boxes_num, int boxes_dim,
                       float nms_overlap_thresh, int device_id) {
  M3D_REQUIRE(keep_out && num_out, "NULL pointer");
  if (boxes_num <= 0) {
    *num_out = 0;
    return M3D_OK;
  }
  M3D_REQUIRE(boxes_host != nullptr && boxes_dim >= 4, "bad boxes");
  int cur = 0;
  M3D_CUDA_OK(cudaGetDevice(&cur));
  if (cur != device_id) M3D_CUDA_OK(cudaSetDevice(device_id));
  std::lock_guard<std::mutex> lock(g_nms_ws.mu);
  NmsWorkspace& ws = g_nms_ws;
  if (ws.device != device_id || ws.cap < boxes_num) {
    if (ws.boxes) cudaFree(ws.boxes), cudaFree(ws.mask), cudaFree(ws.keep), cudaFree(ws.num_keep), cudaFree(ws.remv);
    ws.cap = std::max(boxes_num, 3072);
    ws.device = device_id;
    const size_t cb = (ws.cap + 63) / 64;
    M3D_CUDA_OK(cudaMalloc(&ws.boxes, sizeof(float) * ws.cap * 16));
    M3D_CUDA_OK(cudaMalloc(&ws.mask, sizeof(unsigned long long) * ws.cap * cb));
    M3D_CUDA_OK(cudaMalloc(&ws.keep, sizeof(int) * ws.cap));
    M3D_CUDA_OK(cudaMalloc(&ws.num_keep, sizeof(int)));
    M3D_CUDA_OK(cudaMalloc(&ws.remv, sizeof(unsigned long long) * cb));
  }
  M3D_REQUIRE(boxes_dim <= 16, "boxes_dim %d > 16", boxes_dim);
  cudaStream_t st = cudaStreamPerThread;
  M3D_CUDA_OK(cudaMemcpyAsync(ws.boxes, boxes_host, sizeof(float) * boxes_num * boxes_dim, cudaMemcpyHostToDevice, st));
  const int cb = (boxes_num + 63) / 64;
  dim3 grid(cb, cb, 1);
  nms_mask_kernel<<<grid, 128, 0, st>>>(ws.boxes, boxes_dim, nullptr, boxes_num, nms_overlap_thresh, ws.mask, cb);
  M3D_CUDA_OK(cudaGetLastError());
  if (cb <= 64)
    nms_sweep_kernel<<<1, kSweepThreads, 0, st>>>(ws.mask, nullptr, boxes_num, cb, ws.keep, ws.num_keep);
  else  // more than 4096 boxes: exact generic sweep (the reference accepts any box count)
    nms_sweep_generic_kernel<<<1, 256, 0, st>>>(ws.mask, nullptr, boxes_num, cb, ws.remv, ws.keep, ws.num_keep);
  M3D_CUDA_OK(cudaGetLastError());
  M3D_CUDA_OK(cudaMemcpyAsync(num_out, ws.num_keep, sizeof(int), cudaMemcpyDeviceToHost, st));
  M3D_CUDA_OK(cudaStreamSynchronize(st));
  M3D_CUDA_OK(cudaMemcpyAsync(keep_out, ws.keep, sizeof(int) * (*num_out), cudaMemcpyDeviceToHost, st));
  M3D_CUDA_OK(cudaStreamSynchronize(st));
  if (cur != device_id) M3D_CUDA_OK(cudaSetDevice(cur));
  return M3D_OK;
}

extern "C" size_t m3d_decode_topk_workspace(int batch) {
  return topk_hist_bytes(batch) + static_cast<size_t>(batch) * kCandCap * sizeof(unsigned long long);
}

static int decode_topk_impl(const float* score, const unsigned char* cls_pred, const float* bbox_2d,
                            const float* bbox_3d, const float* heads, int heads_cstride, const int* slot_of_output,
                            const float* anchors, const float* means11, const float* stds11, int batch, int A, int H,
                            int W, float feat_stride, float scale_factor, int topk, float* dets, int* det_idx,
                            int* det_num, void* workspace, size_t workspace_bytes, m3d_stream_t stream) {
  M3D_REQUIRE(score && cls_pred && anchors && means11 && stds11 && dets && det_idx && det_num, "NULL pointer");
  M3D_REQUIRE(topk >= 1 && topk <= kMaxTopK, "topk=%d out of range (1..%d)", topk, kMaxTopK);
  DecodeParams p;
  p.score = score, p.cls = cls_pred, p.bbox_2d = bbox_2d, p.bbox_3d = bbox_3d, p.anchors = anchors;
  p.heads = heads, p.heads_cstride = heads_cstride;
  for (int i = 0; i < 11; ++i) p.slot[i] = slot_of_output ? slot_of_output[i] : 0;
  for (int i = 0; i < 11; ++i) p.means[i] = means11[i], p.stds[i] = stds11[i];
  p.M = A * H * W, p.A = A, p.H = H, p.W = W;
  p.feat_stride = feat_stride, p.scale_factor = scale_factor, p.topk = topk;
  p.dets = dets, p.det_idx = det_idx, p.det_num = det_num;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (getenv("M3D_TOPK_SINGLE") == nullptr) {
    if (workspace == nullptr || workspace_bytes < m3d_decode_topk_workspace(batch) ||
        (reinterpret_cast<uintptr_t>(workspace) & 15) != 0) {
      set_last_error("m3d_decode_topk: 16-byte aligned workspace of %zu bytes needed (m3d_decode_topk_workspace), got %zu",
                     m3d_decode_topk_workspace(batch), workspace_bytes);
      return M3D_ERR_WORKSPACE;
    }
    struct { uint32_t* hist; unsigned long long* cand; } g_topk;
    g_topk.hist = static_cast<uint32_t*>(workspace);
    g_topk.cand = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) + topk_hist_bytes(batch));
    uint32_t* count = g_topk.hist + static_cast<size_t>(batch) * 4096;
    M3D_CUDA_OK(cudaMemsetAsync(g_topk.hist, 0, static_cast<size_t>(batch) * 4097 * sizeof(uint32_t), st));
    topk_hist_kernel<0><<<dim3(kTopkSlices, batch), kSelThreads, 0, st>>>(score, p.M, topk, g_topk.hist);
    M3D_CUDA_OK(cudaGetLastError());
    topk_hist_kernel<1><<<dim3(kTopkSlices, batch), kSelThreads, 0, st>>>(score, p.M, topk, g_topk.hist);
    M3D_CUDA_OK(cudaGetLastError());
    topk_compact_kernel<<<dim3(kTopkSlices, batch), kSelThreads, 0, st>>>(score, p.M, topk, g_topk.hist, count, g_topk.cand);
    M3D_CUDA_OK(cudaGetLastError());
    topk_finish_kernel<<<batch, kSelThreads, 0, st>>>(p, g_topk.hist, count, g_topk.cand);
    M3D_CUDA_OK(cudaGetLastError());
    return M3D_OK;
  }
  topk_decode_kernel<<<batch, kSelThreads, 0, st>>>(p);
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}

extern "C" int m3d_decode_topk(const float* score, const unsigned char* cls_pred, const float* bbox_2d,
                               const float* bbox_3d, const float* anchors, const float* means11, const float* stds11,
                               int batch, int A, int H, int W, float feat_stride, float scale_factor, int topk,
                               float* dets, int* det_idx, int* det_num, void* workspace, size_t workspace_bytes,
                               m3d_stream_t stream) {
  M3D_REQUIRE(bbox_2d && bbox_3d, "NULL pointer");
  return decode_topk_impl(score, cls_pred, bbox_2d, bbox_3d, nullptr, 0, nullptr, anchors, means11, stds11, batch, A, H, W,
                          feat_stride, scale_factor, topk, dets, det_idx, det_num, workspace, workspace_bytes, stream);
}

extern "C" int m3d_decode_topk_heads(const float* score, const unsigned char* cls_pred, const float* heads,
                                     int heads_cstride, const int* slot_of_output, const float* anchors,
                                     const float* means11, const float* stds11, int batch, int A, int H, int W,
                                     float feat_stride, float scale_factor, int topk, float* dets, int* det_idx,
                                     int* det_num, void* workspace, size_t workspace_bytes, m3d_stream_t stream) {
  M3D_REQUIRE(heads && slot_of_output, "NULL pointer");
  for (int i = 0; i < 11; ++i)
    M3D_REQUIRE(slot_of_output[i] >= 0 && (slot_of_output[i] + 1) * A <= heads_cstride, "head slot %d outside the buffer",
                slot_of_output[i]);
  return decode_topk_impl(score, cls_pred, nullptr, nullptr, heads, heads_cstride, slot_of_output, anchors, means11, stds11,
                          batch, A, H, W, feat_stride, scale_factor, topk, dets, det_idx, det_num, workspace,
                          workspace_bytes, stream);
}

extern "C" int m3d_gather_kept(const float* dets, int row_len, int batch, int max_n, const int* keep, const int* num_keep,
                               int max_out, float* out, m3d_stream_t stream) {
  M3D_REQUIRE(dets && keep && num_keep && out, "NULL pointer");
  dim3 grid((max_out * row_len + 255) / 256, batch);
  gather_kept_kernel<<<grid, 256, 0, S(stream)>>>(dets, row_len, max_n, keep, num_keep, max_out, out);
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}
