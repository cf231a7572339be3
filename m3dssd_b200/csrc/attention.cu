// ANAB asymmetric non-local attention (model/module/attention.py:120-216).
//
//   Q      = conv1x1(x)                         [HW, ck]        (ck = 168)
//   att    = sigmoid(conv1x1(x))                [HW, L]         (L = 4 pyramid levels)
//   K_tok  = cat_l adaptive_avg_pool_{s_l}(conv1x1_k(x) * att_l)   [T, ck]   (T = 1+16+64+256 = 337)
//   V_tok  = cat_l adaptive_avg_pool_{s_l}(conv1x1_v(x) * att_l)   [T, cv]   (cv = 128)
//   out    = LeakyReLU(BN(softmax(Q K_tok^T) V_tok + x))           (no 1/sqrt(d) scaling)
//
// The three 1x1 convolutions run on the tcgen05 conv kernel; this file holds
// the token pooling (two deterministic passes) and the attention itself.
#include <cuda_bf16.h>

#include <cstdint>
#include <cstdlib>

#include "common.cuh"

namespace m3d {

constexpr int kMaxLevels = 4;
constexpr int kRowsPerChunk = 4;

struct PoolGeom {
  int nlev;
  int size[kMaxLevels];      // pyramid bin counts per side
  int tok0[kMaxLevels + 1];  // first token of each level
  int max_chunks;            // scratch slots per token
};

// adaptive_avg_pool2d bin [start, end) along a dimension of length n with s bins
__device__ __forceinline__ void bin_range(int i, int s, int n, int* a, int* b) {
  *a = (i * n) / s;
  *b = ((i + 1) * n + s - 1) / s;
}

// Pass 1: block (token, chunk, image) sums kvs[pix][c] * sigmoid(kvs[pix][ck+cv+level]) over
// up to kRowsPerChunk rows of the token's bin.  thread = channel.
__global__ void __launch_bounds__(320) anab_pool_partial_kernel(const float* __restrict__ kvs, int cs, int H, int W,
                                                                int ck, int cv, const PoolGeom g,
                                                                float* __restrict__ scratch) {
  const int tok = blockIdx.x, chunk = blockIdx.y, n = blockIdx.z;
  const int C = ck + cv;
  int lev = 0;
  while (lev + 1 < g.nlev && tok >= g.tok0[lev + 1]) ++lev;
  const int s = g.size[lev];
  const int b = tok - g.tok0[lev];
  int r0, r1, c0, c1;
  bin_range(b / s, s, H, &r0, &r1);
  bin_range(b % s, s, W, &c0, &c1);
  const int ra = r0 + chunk * kRowsPerChunk;
  const int rb = min(r1, ra + kRowsPerChunk);
  float* dst = scratch + ((static_cast<long>(n) * gridDim.x + tok) * g.max_chunks + chunk) * C;
  const int c = threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (int r = ra; r < rb; ++r) {
    const float* row = kvs + ((static_cast<long>(n) * H + r) * W) * cs;
    for (int x = c0; x < c1; ++x) {
      const float a = 1.f / (1.f + expf(-__ldg(row + static_cast<long>(x) * cs + C + lev)));
      acc = fmaf(__ldg(row + static_cast<long>(x) * cs + c), a, acc);
    }
  }
  dst[c] = acc;  // zero when the chunk lies beyond the bin
}

// Pass 2: sum the chunks in order, divide by the bin area, write fp32 tokens.
__global__ void anab_pool_finish_kernel(const float* __restrict__ scratch, int H, int W, int ck, int cv, const PoolGeom g,
                                        int T, float* __restrict__ ktok, float* __restrict__ vtok) {
  const int tok = blockIdx.x, n = blockIdx.y;
  const int C = ck + cv;
  int lev = 0;
  while (lev + 1 < g.nlev && tok >= g.tok0[lev + 1]) ++lev;
  const int s = g.size[lev];
  const int b = tok - g.tok0[lev];
  int r0, r1, c0, c1;
  bin_range(b / s, s, H, &r0, &r1);
  bin_range(b % s, s, W, &c0, &c1);
  const float inv = 1.f / static_cast<float>((r1 - r0) * (c1 - c0));
  const int nchunks = (r1 - r0 + kRowsPerChunk - 1) / kRowsPerChunk;
  const float* src = scratch + (static_cast<long>(n) * T + tok) * g.max_chunks * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < nchunks; ++k) acc += src[k * C + c];
    acc *= inv;
    if (c < ck)
      ktok[(static_cast<long>(n) * T + tok) * ck + c] = acc;
    else
      vtok[(static_cast<long>(n) * T + tok) * cv + (c - ck)] = acc;
  }
}

// ---------------------------------------------------------------------------
// Fast pooling when every pyramid size divides the largest one (S) and S divides H and W (the benchmark shape:
// 48 x 160, sizes 1/4/8/16): the bins are exact, so every level's bins are unions of the S x S finest regions.
// Pass 1: block (region, image) reads its pixels ONCE (the general kernels read the map once per level and take
// a sigmoid per (pixel, channel)), computes the nlev sigmoids per pixel into shared memory and writes the
// per-level weighted channel sums of the region.  Pass 2: token = fixed-order sum of its regions / bin area.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(320) anab_pool_region_kernel(const float* __restrict__ kvs, int cs, int H, int W,
                                                               int C, int nlev, int S, float* __restrict__ part) {
  __shared__ float s_sig[64 * kMaxLevels];
  const int reg = blockIdx.x, n = blockIdx.y;
  const int rh = H / S, rw = W / S, npix = rh * rw;  // host: npix <= 64
  const int y0 = (reg / S) * rh, x0 = (reg % S) * rw;
  for (int i = threadIdx.x; i < npix * nlev; i += blockDim.x) {
    const int px = i / nlev, l = i - px * nlev;
    const int y = y0 + px / rw, x = x0 + px % rw;
    s_sig[i] = 1.f / (1.f + expf(-__ldg(kvs + ((static_cast<long>(n) * H + y) * W + x) * cs + C + l)));
  }
  __syncthreads();
  const int c = threadIdx.x;
  if (c >= C) return;
  float acc[kMaxLevels] = {0.f, 0.f, 0.f, 0.f};
  // rows of rw pixels; within a row the pixel stride is constant, so five loads are issued before their first use
  for (int ry = 0; ry < rh; ++ry) {
    const float* row = kvs + ((static_cast<long>(n) * H + y0 + ry) * W + x0) * cs + c;
    int rx = 0;
    for (; rx + 5 <= rw; rx += 5) {
      float v[5];
#pragma unroll
      for (int u = 0; u < 5; ++u) v[u] = __ldg(row + static_cast<long>(rx + u) * cs);
#pragma unroll
      for (int u = 0; u < 5; ++u)
#pragma unroll
        for (int l = 0; l < kMaxLevels; ++l)
          if (l < nlev) acc[l] = fmaf(v[u], s_sig[(ry * rw + rx + u) * nlev + l], acc[l]);
    }
    for (; rx < rw; ++rx) {
      const float v = __ldg(row + static_cast<long>(rx) * cs);
#pragma unroll
      for (int l = 0; l < kMaxLevels; ++l)
        if (l < nlev) acc[l] = fmaf(v, s_sig[(ry * rw + rx) * nlev + l], acc[l]);
    }
  }
  float* dst = part + ((static_cast<long>(n) * S * S + reg) * nlev) * C;
  for (int l = 0; l < nlev; ++l) dst[l * C + c] = acc[l];
}

// Block (token, image), thread = (channel, slice): the f x f finest regions of the token's bin are dealt round-robin to
// kFinSlices slices (the single token of a 1 x 1 level sums S * S = 256 regions: walked by one thread per channel, one
// L2 round trip after the other, it was the straggler of the whole pooling, ~100 us), each slice keeps four loads in
// flight, and the slices are combined through shared memory in a fixed order (deterministic).
constexpr int kFinSlices = 3;
constexpr int kFinThreads = 320;  // >= ck + cv (host check)

__global__ void __launch_bounds__(kFinThreads* kFinSlices) anab_pool_region_finish_kernel(
    const float* __restrict__ part, int H, int W, int ck, int cv, const PoolGeom g, int S, int T,
    float* __restrict__ ktok, float* __restrict__ vtok) {
  __shared__ float s_part[kFinSlices - 1][kFinThreads];
  const int tok = blockIdx.x, n = blockIdx.y;
  const int C = ck + cv;
  int lev = 0;
  while (lev + 1 < g.nlev && tok >= g.tok0[lev + 1]) ++lev;
  const int s = g.size[lev];
  const int b = tok - g.tok0[lev];
  const int f = S / s;  // finest regions per bin side
  const int ry0 = (b / s) * f, rx0 = (b % s) * f;
  const float inv = 1.f / static_cast<float>((H / s) * (W / s));
  const float* src = part + (static_cast<long>(n) * S * S * g.nlev + lev) * C;
  const long rstride = static_cast<long>(g.nlev) * C;
  const int c = threadIdx.x % kFinThreads, slice = threadIdx.x / kFinThreads;
  const int nreg = f * f;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (c < C) {
    auto at = [&](int i) { return __ldg(src + static_cast<long>((ry0 + i / f) * S + rx0 + i % f) * rstride + c); };
    int i = slice;
    for (; i + 3 * kFinSlices < nreg; i += 4 * kFinSlices) {
      const float v0 = at(i), v1 = at(i + kFinSlices), v2 = at(i + 2 * kFinSlices), v3 = at(i + 3 * kFinSlices);
      a0 += v0, a1 += v1, a2 += v2, a3 += v3;
    }
    for (; i < nreg; i += kFinSlices) a0 += at(i);
  }
  float acc = (a0 + a1) + (a2 + a3);
  if (slice > 0) s_part[slice - 1][c] = acc;
  __syncthreads();
  if (slice == 0 && c < C) {
#pragma unroll
    for (int k = 0; k < kFinSlices - 1; ++k) acc += s_part[k][c];
    acc *= inv;
    if (c < ck)
      ktok[(static_cast<long>(n) * T + tok) * ck + c] = acc;
    else
      vtok[(static_cast<long>(n) * T + tok) * cv + (c - ck)] = acc;
  }
}

// ---------------------------------------------------------------------------
// Attention, SIMT fp32: block = QB queries.  K tokens are streamed through
// shared memory in chunks for the logits; softmax per query by one warp;
// P V accumulated with V read straight from L2 (coalesced over channels).
// ---------------------------------------------------------------------------
constexpr int QB = 16;
constexpr int KCHUNK = 32;

template <typename T>
__device__ __forceinline__ float ldf(const T* p) {
  return static_cast<float>(*p);
}

template <typename TQ, typename TX>
__global__ void __launch_bounds__(256) anab_attention_kernel(const TQ* __restrict__ q, int q_cs,
                                                             const float* __restrict__ ktok,
                                                             const float* __restrict__ vtok, const TX* __restrict__ x,
                                                             int x_cs, const float* __restrict__ scale,
                                                             const float* __restrict__ shift, float slope,
                                                             TX* __restrict__ out, int out_cs, int HW, int ck, int cv,
                                                             int T) {
  extern __shared__ float sm[];
  float* s_q = sm;                      // [QB][ck]
  float* s_k = s_q + QB * ck;           // [KCHUNK][ck + 1]
  float* s_p = s_k + KCHUNK * (ck + 1); // [QB][T]
  const int n = blockIdx.y;
  const int q0 = blockIdx.x * QB;
  const int nq = min(QB, HW - q0);
  const int tid = threadIdx.x;
  for (int i = tid; i < QB * ck; i += blockDim.x) {
    const int qi = i / ck, c = i - qi * ck;
    s_q[i] = qi < nq ? ldf(q + (static_cast<long>(n) * HW + q0 + qi) * q_cs + c) : 0.f;
  }
  const float* kt = ktok + static_cast<long>(n) * T * ck;
  for (int j0 = 0; j0 < T; j0 += KCHUNK) {
    const int nj = min(KCHUNK, T - j0);
    __syncthreads();
    for (int i = tid; i < nj * ck; i += blockDim.x) {
      const int j = i / ck, c = i - j * ck;
      s_k[j * (ck + 1) + c] = __ldg(kt + static_cast<long>(j0 + j) * ck + c);
    }
    __syncthreads();
    for (int i = tid; i < QB * KCHUNK; i += blockDim.x) {
      const int j = i % KCHUNK, qi = i / KCHUNK;
      if (j < nj) {
        float acc = 0.f;
        const float* qa = s_q + qi * ck;
        const float* kb = s_k + j * (ck + 1);
        for (int c = 0; c < ck; ++c) acc = fmaf(qa[c], kb[c], acc);
        s_p[qi * T + j0 + j] = acc;
      }
    }
  }
  __syncthreads();
  // softmax over the T tokens of each query (one warp per query, queries strided over warps)
  const int lane = tid & 31, warp = tid >> 5;
  for (int qi = warp; qi < QB; qi += 8) {
    float* p = s_p + qi * T;
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) mx = fmaxf(mx, p[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float e = expf(p[j] - mx);
      p[j] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    for (int j = lane; j < T; j += 32) p[j] *= inv;
  }
  __syncthreads();
  // out[qi][c] = sum_j p[qi][j] * V[j][c] + x ; BN ; LeakyReLU
  const float* vt = vtok + static_cast<long>(n) * T * cv;
  for (int i = tid; i < QB * cv; i += blockDim.x) {
    const int c = i % cv, qi = i / cv;
    if (qi >= nq) continue;
    const float* p = s_p + qi * T;
    float acc = 0.f;
    for (int j = 0; j < T; ++j) acc = fmaf(p[j], __ldg(vt + static_cast<long>(j) * cv + c), acc);
    const long pix = static_cast<long>(n) * HW + q0 + qi;
    float v = acc + ldf(x + pix * x_cs + c);
    v = v * scale[c] + shift[c];
    v = v > 0.f ? v : v * slope;
    out[pix * out_cs + c] = static_cast<TX>(v);
  }
}

}  // namespace m3d

namespace m3d {
bool anab_tc_supported(int ck, int cv, int T, int q_cs, int x_cs, int out_cs);
size_t anab_tc_workspace_bytes(int N);
int launch_anab_attention_tc(const void* q, int q_cs, const float* ktok, const float* vtok, const void* x, int x_cs,
                             const float* scale, const float* shift, float slope, void* out, int out_cs, int N, int HW,
                             int ck, int cv, int T, void* workspace, size_t workspace_bytes, cudaStream_t stream);
}  // namespace m3d

using namespace m3d;

static inline cudaStream_t S(m3d_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static int make_geom(int nlev, const int* sizes, int H, PoolGeom* g) {
  if (nlev < 1 || nlev > kMaxLevels) return -1;
  g->nlev = nlev;
  int t = 0, maxc = 1;
  for (int l = 0; l < nlev; ++l) {
    if (sizes[l] < 1) return -1;
    g->size[l] = sizes[l];
    g->tok0[l] = t;
    t += sizes[l] * sizes[l];
    // tallest bin of this level: ceil(H/s) + 1 rows at most
    const int rows = (H + sizes[l] - 1) / sizes[l] + 1;
    const int ch = (rows + kRowsPerChunk - 1) / kRowsPerChunk;
    if (ch > maxc) maxc = ch;
  }
  g->tok0[nlev] = t;
  for (int l = nlev + 1; l <= kMaxLevels; ++l) g->tok0[l] = t;
  g->max_chunks = maxc;
  return t;
}

extern "C" size_t m3d_anab_pool_workspace(int N, int H, int nlev, const int* sizes, int ck, int cv) {
  PoolGeom g;
  const int T = make_geom(nlev, sizes, H, &g);
  if (T < 0) return 0;
  return static_cast<size_t>(N) * T * g.max_chunks * (ck + cv) * sizeof(float);
}

extern "C" int m3d_anab_pool(const float* kvs, int kvs_cstride, int N, int H, int W, int ck, int cv, int nlev,
                             const int* sizes, void* workspace, size_t workspace_bytes, float* ktok, float* vtok,
                             m3d_stream_t stream) {
  M3D_REQUIRE(kvs && sizes && workspace && ktok && vtok, "NULL pointer");
  M3D_REQUIRE(ck + cv <= 320, "ck + cv = %d exceeds 320", ck + cv);
  M3D_REQUIRE(kvs_cstride >= ck + cv + nlev, "kvs channel stride too small");
  PoolGeom g;
  const int T = make_geom(nlev, sizes, H, &g);
  M3D_REQUIRE(T > 0, "bad pyramid sizes");
  if (workspace_bytes < m3d_anab_pool_workspace(N, H, nlev, sizes, ck, cv)) {
    set_last_error("ANAB pooling workspace too small");
    return M3D_ERR_WORKSPACE;
  }
  float* scratch = static_cast<float*>(workspace);
  {
    // exact-bin fast path
    int SS = 0;
    for (int l = 0; l < nlev; ++l) SS = sizes[l] > SS ? sizes[l] : SS;
    bool exact = H % SS == 0 && W % SS == 0 && (H / SS) * (W / SS) <= 64;
    for (int l = 0; l < nlev; ++l) exact = exact && SS % sizes[l] == 0;
    const size_t need = static_cast<size_t>(N) * SS * SS * nlev * (ck + cv) * sizeof(float);
    if (exact && need <= workspace_bytes && getenv("M3D_ANAB_SIMT") == nullptr) {
      anab_pool_region_kernel<<<dim3(SS * SS, N), 320, 0, S(stream)>>>(kvs, kvs_cstride, H, W, ck + cv, nlev, SS, scratch);
      M3D_CUDA_OK(cudaGetLastError());
      anab_pool_region_finish_kernel<<<dim3(T, N), kFinThreads * kFinSlices, 0, S(stream)>>>(scratch, H, W, ck, cv, g, SS, T, ktok, vtok);
      M3D_CUDA_OK(cudaGetLastError());
      return M3D_OK;
    }
  }
  dim3 grid(T, g.max_chunks, N);
  anab_pool_partial_kernel<<<grid, 320, 0, S(stream)>>>(kvs, kvs_cstride, H, W, ck, cv, g, scratch);
  M3D_CUDA_OK(cudaGetLastError());
  anab_pool_finish_kernel<<<dim3(T, N), 128, 0, S(stream)>>>(scratch, H, W, ck, cv, g, T, ktok, vtok);
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}

template <typename TQ, typename TX>
static int launch_attn(const void* q, int q_cs, const float* ktok, const float* vtok, const void* x, int x_cs,
                       const float* scale, const float* shift, float slope, void* out, int out_cs, int N, int HW, int ck,
                       int cv, int T, cudaStream_t st) {
  const size_t smem = sizeof(float) * (static_cast<size_t>(QB) * ck + KCHUNK * (ck + 1) + static_cast<size_t>(QB) * T);
  auto kern = anab_attention_kernel<TQ, TX>;
  M3D_ONCE_PER_DEVICE_BEGIN
    M3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  M3D_ONCE_PER_DEVICE_END
  M3D_REQUIRE(smem <= 160 * 1024, "attention tile does not fit shared memory");
  dim3 grid((HW + QB - 1) / QB, N);
  kern<<<grid, 256, smem, st>>>(static_cast<const TQ*>(q), q_cs, ktok, vtok, static_cast<const TX*>(x), x_cs, scale,
                                shift, slope, static_cast<TX*>(out), out_cs, HW, ck, cv, T);
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}

extern "C" size_t m3d_anab_attention_workspace(int N, int act_dtype) {
  return act_dtype == M3D_BF16 ? anab_tc_workspace_bytes(N) : 0;
}

extern "C" int m3d_anab_attention(const void* q, int q_cstride, const float* ktok, const float* vtok, const void* x,
                                  int x_cstride, int act_dtype, const float* scale, const float* shift, float slope,
                                  void* out, int out_cstride, int N, int HW, int ck, int cv, int T, void* workspace,
                                  size_t workspace_bytes, m3d_stream_t stream) {
  M3D_REQUIRE(q && ktok && vtok && x && scale && shift && out, "NULL pointer");
  if (act_dtype == M3D_BF16 && anab_tc_supported(ck, cv, T, q_cstride, x_cstride, out_cstride) &&
      getenv("M3D_ANAB_SIMT") == nullptr)
    return launch_anab_attention_tc(q, q_cstride, ktok, vtok, x, x_cstride, scale, shift, slope, out, out_cstride, N, HW,
                                    ck, cv, T, workspace, workspace_bytes, S(stream));
  if (act_dtype == M3D_BF16)
    return launch_attn<__nv_bfloat16, __nv_bfloat16>(q, q_cstride, ktok, vtok, x, x_cstride, scale, shift, slope, out,
                                                     out_cstride, N, HW, ck, cv, T, S(stream));
  return launch_attn<float, float>(q, q_cstride, ktok, vtok, x, x_cstride, scale, shift, slope, out, out_cstride, N, HW,
                                   ck, cv, T, S(stream));
}
