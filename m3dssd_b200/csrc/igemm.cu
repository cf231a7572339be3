// tcgen05 implicit-GEMM convolution kernels; see igemm.cuh for the design.
#include "igemm.cuh"

#include <cstdio>

#include "common.cuh"
#include "ptx.cuh"

namespace m3d {

// ------------------------------------------------------------------ helpers
struct TileCoord {
  int g, nt, n, p0, q0;
};

__device__ __forceinline__ TileCoord decode_tile(int tile, int n_tiles, int tiles_w, int tiles_h, int N, int TW,
                                                 int TH) {
  TileCoord t;
  t.nt = tile % n_tiles;
  int r = tile / n_tiles;
  int tw = r % tiles_w;
  r /= tiles_w;
  int th = r % tiles_h;
  r /= tiles_h;
  t.n = r % N;
  t.g = r / N;
  t.p0 = th * TH;
  t.q0 = tw * TW;
  return t;
}

template <typename T>
struct Out16;

// Epilogue for 16 accumulator columns of one output pixel: + bias, + residual,
// LeakyReLU, convert, store.  `nvalid` = number of real channels in the chunk.
template <typename OutT, typename ResT>
__device__ __forceinline__ void epilogue_chunk16(const uint32_t (&acc)[16], const float* __restrict__ bias,
                                                 const ResT* __restrict__ res, OutT* __restrict__ out, int nvalid,
                                                 float slope) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(acc[i]);
  if (bias != nullptr) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nvalid) v[i] += __ldg(bias + i);
  }
  if (res != nullptr) {
    if (nvalid == 16 && (reinterpret_cast<uintptr_t>(res) & 15) == 0) {
      if constexpr (sizeof(ResT) == 2) {
        const uint4* r4 = reinterpret_cast<const uint4*>(res);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint4 u = __ldg(r4 + h);
          uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[h * 8 + 2 * i] += __uint_as_float(w[i] << 16);
            v[h * 8 + 2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
          }
        }
      } else {
        const float4* r4 = reinterpret_cast<const float4*>(res);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float4 u = __ldg(r4 + h);
          v[4 * h] += u.x;
          v[4 * h + 1] += u.y;
          v[4 * h + 2] += u.z;
          v[4 * h + 3] += u.w;
        }
      }
    } else {
      for (int i = 0; i < nvalid; ++i) v[i] += static_cast<float>(res[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * slope;

  if (nvalid == 16 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    if constexpr (sizeof(OutT) == 2) {
      uint32_t w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&b);
      }
      uint4* o4 = reinterpret_cast<uint4*>(out);
      o4[0] = make_uint4(w[0], w[1], w[2], w[3]);
      o4[1] = make_uint4(w[4], w[5], w[6], w[7]);
    } else {
      float4* o4 = reinterpret_cast<float4*>(out);
#pragma unroll
      for (int h = 0; h < 4; ++h) o4[h] = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
    }
  } else {
    for (int i = 0; i < nvalid; ++i) out[i] = static_cast<OutT>(v[i]);
  }
}

// TMEM columns per accumulator stage (power of two >= 32).
__host__ __device__ constexpr int acc_cols(int bn) { return bn <= 32 ? 32 : (bn <= 64 ? 64 : (bn <= 128 ? 128 : 256)); }

// Drain one 128 x BN accumulator tile: TMEM -> registers -> global (NHWC).
template <int BN, typename OutT, typename ResT>
__device__ __forceinline__ void epilogue_tile(uint32_t tmem_acc, int quarter, int lane, int n, int p0, int q0, int TW,
                                              int P, int Q, int col_base, int cout, const float* bias, const ResT* res,
                                              int res_cstride, OutT* out, int out_cstride, float slope) {
  const int row = quarter * 32 + lane;
  const int p = p0 + row / TW;
  const int q = q0 + row % TW;
  const bool pix_ok = (p < P) && (q < Q);
  const long pix = (static_cast<long>(n) * P + p) * Q + q;
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 16) {
    uint32_t acc[16];
    tmem_ld16(tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + c0, acc);  // warp-collective
    tmem_ld_wait();
    const int col = col_base + c0;
    int nvalid = cout - col;
    nvalid = nvalid > 16 ? 16 : nvalid;
    if (pix_ok && nvalid > 0) {
      epilogue_chunk16<OutT, ResT>(acc, bias ? bias + col : nullptr, res ? res + pix * res_cstride + col : nullptr,
                                   out + pix * out_cstride + col, nvalid, slope);
    }
  }
}

// =========================================================================
// Plain convolution: TMA-fed A operand.
//   warp 0: TMA producer (A window + weight tile per k-block)
//   warp 1: TMEM allocator + MMA issuer
//   warps 2-5: epilogue (TMEM lane quarter = warp % 4)
// =========================================================================
template <int BN, int BK>
struct TmaCfg {
  static constexpr int ROW_BYTES = BK * 2;
  static constexpr int A_BYTES = kTileM * ROW_BYTES;
  static constexpr int B_BYTES = BN * ROW_BYTES;
  static constexpr int B_STRIDE = (B_BYTES + 1023) & ~1023;
  static constexpr int STAGE = A_BYTES + B_STRIDE;
  static constexpr int STAGES = (STAGE * 6 <= 96 * 1024) ? 6 : ((STAGE * 4 <= 200 * 1024) ? 4 : 3);
  static constexpr int SMEM = STAGES * STAGE + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int ACC = acc_cols(BN);
};

template <int BN, int BK, typename OutT>
__global__ void __launch_bounds__(192, 1) conv_tma_kernel(const __grid_constant__ ConvTmaParams p) {
  using Cfg = TmaCfg<BN, BK>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 128);
    }
    fence_barrier_init();
    for (int i = 0; i < p.num_inputs; ++i) prefetch_tmap(&p.tmap_a[i]);
    prefetch_tmap(&p.tmap_b);
  }
  if (warp == 1) tmem_alloc<2 * Cfg::ACC>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int taps = p.R * p.S;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(tile, p.n_tiles, p.tiles_w, p.tiles_h, p.N, p.TW, p.TH);
        const int brow = t.g * p.b_goff + t.nt * BN;
        int kcol = 0;
        for (int i = 0; i < p.num_inputs; ++i) {
          const int c_base = p.a_coff[i] + t.g * p.a_goff[i];
          for (int tap = 0; tap < taps; ++tap) {
            const int r = tap / p.S, s = tap % p.S;
            const int h0 = t.p0 * p.stride - p.pad + r * p.dil;
            const int w0 = t.q0 * p.stride - p.pad + s * p.dil;
            for (int c = 0; c < p.chunks[i]; ++c) {
              mbar_wait(&empty[stage], phase ^ 1);
              mbar_arrive_expect_tx(&full[stage], Cfg::A_BYTES + Cfg::B_BYTES);
              tma_load_4d(smem_a + stage * Cfg::A_BYTES, &p.tmap_a[i], &full[stage], c_base + c * BK, w0, h0, t.n);
              tma_load_2d(smem_b + stage * Cfg::B_STRIDE, &p.tmap_b, &full[stage], kcol, brow);
              kcol += BK;
              if (++stage == STAGES) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BN);
      int total_kb = 0;
      for (int i = 0; i < p.num_inputs; ++i) total_kb += taps * p.chunks[i];
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + as * Cfg::ACC;
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_smem_desc<Cfg::ROW_BYTES>(smem_u32(smem_a + stage * Cfg::A_BYTES));
          const uint64_t db = umma_smem_desc<Cfg::ROW_BYTES>(smem_u32(smem_b + stage * Cfg::B_STRIDE));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_acc, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[as]);
      }
    }
  } else {
    const int quarter = warp & 3;
    int local = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
      const TileCoord t = decode_tile(tile, p.n_tiles, p.tiles_w, p.tiles_h, p.N, p.TW, p.TH);
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const float* bias = p.bias ? p.bias + t.g * p.bias_goff : nullptr;
      const __nv_bfloat16* res =
          p.res ? static_cast<const __nv_bfloat16*>(p.res) + p.res_coff + t.g * p.res_goff : nullptr;
      OutT* out = static_cast<OutT*>(p.out) + p.out_coff + t.g * p.out_goff;
      epilogue_tile<BN, OutT, __nv_bfloat16>(tmem_base + as * Cfg::ACC, quarter, lane, t.n, t.p0, t.q0, p.TW, p.P, p.Q,
                                             t.nt * BN, p.Cout, bias, res, p.res_cstride, out, p.out_cstride, p.slope);
      tc_fence_before();
      mbar_arrive(&tempty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2 * Cfg::ACC>(tmem_base);
  }
}

// =========================================================================
// Deformable (DCNv2) / software-gather convolution.
//   warps 0-7 : A producers (bilinear gather -> swizzled bf16 tile)
//   warp 8    : weight TMA producer
//   warp 9    : TMEM allocator + MMA issuer
//   warps 10-13: epilogue
// InT = bf16: one MMA per k16.  InT = float ("split"): activations and weights
// are carried as three bf16 parts (hi, mid, lo: 3 x 8 = 24 mantissa bits) and
// each k16 issues six MMAs -- hi*hi, hi*mid, mid*hi, mid*mid, hi*lo, lo*hi --
// i.e. every term down to 2^-16 of the product; the dropped ones are <= 2^-24.
// That reproduces fp32 products to ~2^-23 with fp32 accumulation in TMEM.
// =========================================================================
constexpr int kGatherBK = 64;
constexpr int kProducerThreads = 256;

template <int BN, bool SPLIT>
struct GatherCfg {
  static constexpr int ROW_BYTES = 128;
  static constexpr int A_BYTES = kTileM * ROW_BYTES;  // per part
  static constexpr int B_BYTES = BN * ROW_BYTES;      // per part
  static constexpr int PARTS = SPLIT ? 3 : 1;
  static constexpr int STAGE = PARTS * (A_BYTES + B_BYTES);
  static constexpr int OM_BYTES = kTileM * 28 * 4;
  static constexpr int STAGES = (STAGE * 4 + OM_BYTES <= 200 * 1024) ? 4 : ((STAGE * 3 + OM_BYTES <= 210 * 1024) ? 3 : 2);
  static_assert(STAGES * STAGE + OM_BYTES + 1280 <= 227 * 1024, "gather tile does not fit shared memory");
  static constexpr int SMEM = STAGES * STAGE + OM_BYTES + 1024 + 256;
  static constexpr int ACC = acc_cols(BN);
};

struct SampleInfo {
  int o00, o01, o10, o11;  // element offsets of the four corner pixels (channel 0 of this input)
  float w00, w01, w10, w11;
  float mask;
};

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}

__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&b);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int BN, typename InT, typename OutT>
__global__ void __launch_bounds__(448, 1) conv_gather_kernel(const __grid_constant__ ConvGatherParams p) {
  constexpr bool SPLIT = sizeof(InT) == 4;
  using Cfg = GatherCfg<BN, SPLIT>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BK = kGatherBK;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // stage layout: [A_hi][A_mid][A_lo][B_hi][B_mid][B_lo]  (one part each when !SPLIT)
  float* om_s = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE + Cfg::OM_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], kProducerThreads + 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 128);
    }
    fence_barrier_init();
    prefetch_tmap(&p.tmap_b);
    if (SPLIT) {
      prefetch_tmap(&p.tmap_b_mid);
      prefetch_tmap(&p.tmap_b_lo);
    }
  }
  if (warp == 9) tmem_alloc<2 * Cfg::ACC>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int taps = p.R * p.S;
  int total_kb = 0;
  for (int i = 0; i < p.num_inputs; ++i) total_kb += taps * p.chunks[i];

  if (warp < 8) {
    // ------------------------------------------------------------ A producers
    const int pt = threadIdx.x;
    const int j = pt & 7;        // 16-byte chunk (8 channels) inside the 64-channel k-block
    const int rbase = pt >> 3;   // rows rbase + 32*i
    const int omc = 3 * taps;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(tile, p.n_tiles, p.tiles_w, p.tiles_h, p.N, p.TW, p.TH);
      if (p.om != nullptr) {
        // stage this tile's offsets / masks: om_s[row][omc]
        asm volatile("bar.sync 1, 256;" ::: "memory");  // previous tile's readers are done
        for (int idx = pt; idx < kTileM * omc; idx += kProducerThreads) {
          const int row = idx / omc, k = idx - row * omc;
          const int pp = t.p0 + row / p.TW, qq = t.q0 + row % p.TW;
          float v = 0.f;
          if (pp < p.P && qq < p.Q) v = __ldg(p.om + ((static_cast<long>(t.n) * p.P + pp) * p.Q + qq) * p.om_cstride + k);
          om_s[row * omc + k] = v;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      for (int i = 0; i < p.num_inputs; ++i) {
        const InT* in = static_cast<const InT*>(p.in[i]) + p.in_coff[i];
        const int cs = p.in_cstride[i];
        for (int tap = 0; tap < taps; ++tap) {
          const int r = tap / p.S, s = tap % p.S;
          SampleInfo si[4];
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) {
            const int row = rbase + 32 * ii;
            const int pp = t.p0 + row / p.TW, qq = t.q0 + row % p.TW;
            SampleInfo& x = si[ii];
            x.o00 = x.o01 = x.o10 = x.o11 = 0;
            x.w00 = x.w01 = x.w10 = x.w11 = 0.f;
            x.mask = 0.f;
            if (pp < p.P && qq < p.Q) {
              float hf = static_cast<float>(pp * p.stride - p.pad + r * p.dil);
              float wf = static_cast<float>(qq * p.stride - p.pad + s * p.dil);
              float m = 1.f;
              if (p.om != nullptr) {
                hf += om_s[row * omc + 2 * tap];
                wf += om_s[row * omc + 2 * tap + 1];
                m = om_s[row * omc + 2 * taps + tap];
                if (p.sigmoid_mask) m = 1.f / (1.f + expf(-m));
              }
              if (hf > -1.f && wf > -1.f && hf < static_cast<float>(p.H) && wf < static_cast<float>(p.W)) {
                const float hl = floorf(hf), wl = floorf(wf);
                const int h_low = static_cast<int>(hl), w_low = static_cast<int>(wl);
                const int h_high = h_low + 1, w_high = w_low + 1;
                const float lh = hf - hl, lw = wf - wl;
                const float hh = 1.f - lh, hw = 1.f - lw;
                const bool hl_ok = h_low >= 0, wl_ok = w_low >= 0;
                const bool hh_ok = h_high <= p.H - 1, wh_ok = w_high <= p.W - 1;
                const int rl = (t.n * p.H + (hl_ok ? h_low : 0)) * p.W;
                const int rh = (t.n * p.H + (hh_ok ? h_high : 0)) * p.W;
                const int cl = wl_ok ? w_low : 0, ch = wh_ok ? w_high : 0;
                x.o00 = (rl + cl) * cs;
                x.o01 = (rl + ch) * cs;
                x.o10 = (rh + cl) * cs;
                x.o11 = (rh + ch) * cs;
                x.w00 = (hl_ok && wl_ok) ? hh * hw : 0.f;
                x.w01 = (hl_ok && wh_ok) ? hh * lw : 0.f;
                x.w10 = (hh_ok && wl_ok) ? lh * hw : 0.f;
                x.w11 = (hh_ok && wh_ok) ? lh * lw : 0.f;
                x.mask = m;
              }
            }
          }
          for (int c = 0; c < p.chunks[i]; ++c) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* a_hi = smem + stage * Cfg::STAGE;
            const int coff = c * BK + j * 8;
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
              const int row = rbase + 32 * ii;
              const SampleInfo& x = si[ii];
              float acc[8];
              if (x.mask == 0.f && x.w00 == 0.f && x.w01 == 0.f && x.w10 == 0.f && x.w11 == 0.f) {
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = 0.f;
              } else if (x.w00 == 1.f) {
                // integer sample position (plain convolution): one corner
                load8(in + x.o00 + coff, acc);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] *= x.mask;
              } else {
                float v1[8], v2[8], v3[8], v4[8];
                load8(in + x.o00 + coff, v1);
                load8(in + x.o01 + coff, v2);
                load8(in + x.o10 + coff, v3);
                load8(in + x.o11 + coff, v4);
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  acc[e] = (x.w00 * v1[e] + x.w01 * v2[e] + x.w10 * v3[e] + x.w11 * v4[e]) * x.mask;
              }
              const uint32_t off = swizzled_offset<128>(row, j);
              if constexpr (SPLIT) {
                float hi[8], mid[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  hi[e] = __bfloat162float(__float2bfloat16_rn(acc[e]));
                  const float r1 = acc[e] - hi[e];  // exact
                  mid[e] = __bfloat162float(__float2bfloat16_rn(r1));
                  lo[e] = r1 - mid[e];              // exact; rounded to bf16 by pack8
                }
                *reinterpret_cast<uint4*>(a_hi + off) = pack8(hi);
                *reinterpret_cast<uint4*>(a_hi + Cfg::A_BYTES + off) = pack8(mid);
                *reinterpret_cast<uint4*>(a_hi + 2 * Cfg::A_BYTES + off) = pack8(lo);
              } else {
                *reinterpret_cast<uint4*>(a_hi + off) = pack8(acc);
              }
            }
            fence_proxy_async_smem();
            mbar_arrive(&full[stage]);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 8) {
    // ---------------------------------------------------- weight TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(tile, p.n_tiles, p.tiles_w, p.tiles_h, p.N, p.TW, p.TH);
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* b_hi = smem + stage * Cfg::STAGE + Cfg::PARTS * Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full[stage], Cfg::PARTS * Cfg::B_BYTES);
          tma_load_2d(b_hi, &p.tmap_b, &full[stage], kb * BK, t.nt * BN);
          if constexpr (SPLIT) {
            tma_load_2d(b_hi + Cfg::B_BYTES, &p.tmap_b_mid, &full[stage], kb * BK, t.nt * BN);
            tma_load_2d(b_hi + 2 * Cfg::B_BYTES, &p.tmap_b_lo, &full[stage], kb * BK, t.nt * BN);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BN);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + as * Cfg::ACC;
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + stage * Cfg::STAGE);
          const uint32_t b_hi = a_hi + Cfg::PARTS * Cfg::A_BYTES;
          const uint64_t da = umma_smem_desc<128>(a_hi);
          const uint64_t db = umma_smem_desc<128>(b_hi);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if constexpr (SPLIT) {
              const uint64_t dam = umma_smem_desc<128>(a_hi + Cfg::A_BYTES) + 2 * k;
              const uint64_t dal = umma_smem_desc<128>(a_hi + 2 * Cfg::A_BYTES) + 2 * k;
              const uint64_t dbm = umma_smem_desc<128>(b_hi + Cfg::B_BYTES) + 2 * k;
              const uint64_t dbl = umma_smem_desc<128>(b_hi + 2 * Cfg::B_BYTES) + 2 * k;
              // smallest terms first, the dominant hi*hi last
              umma_f16(tmem_acc, da + 2 * k, dbl, idesc, (kb | k) != 0);
              umma_f16(tmem_acc, dal, db + 2 * k, idesc, 1);
              umma_f16(tmem_acc, dam, dbm, idesc, 1);
              umma_f16(tmem_acc, da + 2 * k, dbm, idesc, 1);
              umma_f16(tmem_acc, dam, db + 2 * k, idesc, 1);
              umma_f16(tmem_acc, da + 2 * k, db + 2 * k, idesc, 1);
            } else {
              umma_f16(tmem_acc, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            }
          }
          umma_commit(&empty[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[as]);
      }
    }
  } else {
    // --------------------------------------------------------------- epilogue
    const int quarter = warp & 3;
    int local = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
      const TileCoord t = decode_tile(tile, p.n_tiles, p.tiles_w, p.tiles_h, p.N, p.TW, p.TH);
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const InT* res = p.res ? static_cast<const InT*>(p.res) + p.res_coff : nullptr;
      OutT* out = static_cast<OutT*>(p.out) + p.out_coff;
      epilogue_tile<BN, OutT, InT>(tmem_base + as * Cfg::ACC, quarter, lane, t.n, t.p0, t.q0, p.TW, p.P, p.Q, t.nt * BN,
                                   p.Cout, p.bias, res, p.res_cstride, out, p.out_cstride, p.slope);
      tc_fence_before();
      mbar_arrive(&tempty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<2 * Cfg::ACC>(tmem_base);
  }
}

// ------------------------------------------------------------ host launchers
static int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}

template <int BN, int BK, typename OutT>
static int launch_tma_t(const ConvTmaParams& p, cudaStream_t stream) {
  using Cfg = TmaCfg<BN, BK>;
  auto kern = conv_tma_kernel<BN, BK, OutT>;
  static bool configured = false;
  if (!configured) {
    M3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    configured = true;
  }
  // Two co-resident CTAs per SM when shared memory allows (small tiles).
  const int per_sm = (2 * (Cfg::SMEM + 1024) <= 227 * 1024 && 4 * Cfg::ACC <= 512) ? 2 : 1;
  int grid = num_sms() * per_sm;
  if (grid > p.total_tiles) grid = p.total_tiles;
  kern<<<grid, 192, Cfg::SMEM, stream>>>(p);
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}

int launch_conv_tma(const ConvTmaParams& p, int BN, int BK, int out_dtype, cudaStream_t stream) {
#define M3D_TMA_CASE(bn, bk)                                                        \
  if (BN == bn && BK == bk) {                                                       \
    return out_dtype == DT_BF16 ? launch_tma_t<bn, bk, __nv_bfloat16>(p, stream)    \
                                : launch_tma_t<bn, bk, float>(p, stream);           \
  }
  M3D_TMA_CASE(16, 16)
  M3D_TMA_CASE(32, 16)
  M3D_TMA_CASE(32, 32)
  M3D_TMA_CASE(64, 32)
  M3D_TMA_CASE(32, 64)
  M3D_TMA_CASE(48, 64)
  M3D_TMA_CASE(64, 64)
  M3D_TMA_CASE(128, 64)
  M3D_TMA_CASE(256, 64)
#undef M3D_TMA_CASE
  return M3D_ERR_UNSUPPORTED;
}

template <int BN, typename InT, typename OutT>
static int launch_gather_t(const ConvGatherParams& p, cudaStream_t stream) {
  using Cfg = GatherCfg<BN, sizeof(InT) == 4>;
  auto kern = conv_gather_kernel<BN, InT, OutT>;
  static bool configured = false;
  if (!configured) {
    M3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    configured = true;
  }
  int grid = num_sms();
  if (grid > p.total_tiles) grid = p.total_tiles;
  kern<<<grid, 448, Cfg::SMEM, stream>>>(p);
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}

int launch_conv_gather(const ConvGatherParams& p, int BN, int in_dtype, int out_dtype, cudaStream_t stream) {
#define M3D_G_CASE(bn)                                                                           \
  if (BN == bn) {                                                                                \
    if (in_dtype == DT_BF16)                                                                     \
      return out_dtype == DT_BF16 ? launch_gather_t<bn, __nv_bfloat16, __nv_bfloat16>(p, stream) \
                                  : launch_gather_t<bn, __nv_bfloat16, float>(p, stream);        \
    if constexpr (bn <= 128) return launch_gather_t<bn, float, float>(p, stream);                \
    return M3D_ERR_UNSUPPORTED;                                                                  \
  }
  M3D_G_CASE(16)
  M3D_G_CASE(32)
  M3D_G_CASE(48)
  M3D_G_CASE(64)
  M3D_G_CASE(128)
  M3D_G_CASE(256)
#undef M3D_G_CASE
  return M3D_ERR_UNSUPPORTED;
}

}  // namespace m3d
