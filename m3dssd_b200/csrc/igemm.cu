// tcgen05 implicit-GEMM convolution kernels; see igemm.cuh for the design.
#include "igemm.cuh"

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "epilogue.cuh"
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

// TMEM columns per accumulator stage (power of two >= 32).
__host__ __device__ constexpr int acc_cols(int bn) { return bn <= 32 ? 32 : (bn <= 64 ? 64 : (bn <= 128 ? 128 : 256)); }

// =========================================================================
// Plain convolution: TMA-fed A operand.
//   warp 0: TMA producer (A window + weight tile per k-block)
//   warp 1: TMEM allocator + MMA issuer
//   warps 2-5: epilogue (TMEM lane quarter = warp % 4)
// =========================================================================
// A pipeline stage holds KSUB k-blocks (KSUB x BK channels of K): the producer / MMA hand-shake
// (mbarrier wait + arrive, ~100 clocks each on a lone warp, plus ~80 clocks per TMA issue) is paid once
// per stage, so a stage must carry several hundred tensor-pipe clocks (M128 x N x K16 = N/2 clocks) for
// the issuing warps to stay ahead of the MMAs.
template <int BN, int BK, int KSUB, bool STAGED>
struct TmaCfg {
  static constexpr int ROW_BYTES = BK * 2;
  static constexpr int A_BYTES = kTileM * ROW_BYTES;  // one k-block of A
  static constexpr int B_BYTES = BN * ROW_BYTES;
  static constexpr int B_STRIDE = (B_BYTES + 1023) & ~1023;
  static constexpr int STAGE = KSUB * (A_BYTES + B_STRIDE);
  static constexpr int EXTRA = STAGED ? (2 * kSlabBytes + 1024) : 0;  // staging slabs + bias
  static constexpr int BUDGET = 225 * 1024 + 512 - EXTRA;
  static constexpr int FIT = BUDGET / STAGE;
  static constexpr int STAGES = (STAGE * 6 <= 96 * 1024) ? 6 : (FIT >= 5 ? 5 : FIT);
  static constexpr int SMEM = STAGES * STAGE + EXTRA + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int ACC = acc_cols(BN);
  static_assert(STAGES >= 2, "conv tile does not fit shared memory");
  static_assert(SMEM <= 227 * 1024, "conv tile does not fit shared memory");
  static_assert(KSUB == 1 || B_STRIDE == B_BYTES, "multi-k-block stages need 1024-byte weight sub-tiles");
};

// Position of a k-block in the (input, tap, channel chunk) walk of K.
struct KWalk {
  int i, r, sx, c;
  __device__ __forceinline__ void reset() { i = 0, r = 0, sx = 0, c = 0; }
  __device__ __forceinline__ void advance(int n, const ConvTmaParams& p) {
    c += n;
    if (c >= p.chunks[i]) {
      c = 0;
      if (++sx == p.S) {
        sx = 0;
        if (++r == p.R) r = 0, ++i;
      }
    }
  }
};

// Epilogue warps: 8 for the staged (bf16, TMA-store) epilogue -- small-N full-resolution layers are bounded by
// the accumulator drain, and two warps per scheduler hide its TMEM-load / shared-memory latencies -- else 4.
template <bool STAGED>
constexpr int conv_epi_warps() { return STAGED ? 8 : 4; }

template <int BN, int BK, int KSUB, typename OutT, bool STAGED>
__global__ void __launch_bounds__(64 + 32 * conv_epi_warps<STAGED>(), 1)
    conv_tma_kernel(const __grid_constant__ ConvTmaParams p) {
  using Cfg = TmaCfg<BN, BK, KSUB, STAGED>;
  constexpr int NW = conv_epi_warps<STAGED>();
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // stage layout: [A k-block 0..KSUB-1][B k-block 0..KSUB-1]
  uint8_t* stage_out = smem + STAGES * Cfg::STAGE;  // STAGED: 2 slabs + bias (1 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE + Cfg::EXTRA);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* res_bar = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 32 * NW);
      mbar_init(&res_bar[s], 1);
    }
    fence_barrier_init();
    for (int i = 0; i < p.num_inputs; ++i) prefetch_tmap(&p.tmap_a[i]);
    prefetch_tmap(&p.tmap_b);
    if (STAGED) prefetch_tmap(&p.tmap_out);
  }
  if (warp == 1) tmem_alloc<2 * Cfg::ACC>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_sync();  // everything above overlapped the previous kernel's tail

  int total_kb = 0;
  for (int i = 0; i < p.num_inputs; ++i) total_kb += p.R * p.S * p.chunks[i];
  const int n_stages = total_kb / KSUB;  // host guarantees divisibility

  if (warp == 0) {
    // The whole warp walks the loop (convergent control flow keeps addresses and descriptors in
    // uniform registers); one elected lane issues the TMA operations.
    int stage = 0;
    uint32_t phase = 0;
    const bool wide_a = KSUB > 1 && p.a_wide;  // one 5-D box brings the stage's KSUB channel chunks of one tap
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(tile, p.n_tiles, p.tiles_w, p.tiles_h, p.N, p.TW, p.TH);
      const int brow = t.g * p.b_goff + t.nt * BN;
      const int hbase = t.p0 * p.stride - p.pad, wbase = t.q0 * p.stride - p.pad;
      KWalk kw;
      kw.reset();
      for (int st = 0; st < n_stages; ++st) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE;
        uint8_t* sb = sa + KSUB * Cfg::A_BYTES;
        if (wide_a) {
          const int cb = p.a_coff[kw.i] + t.g * p.a_goff[kw.i];
          if (elect_one()) {
            mbar_arrive_expect_tx(&full[stage], KSUB * (Cfg::A_BYTES + Cfg::B_BYTES));
            tma_load_5d(sa, &p.tmap_a[kw.i], &full[stage], cb & (BK - 1), wbase + kw.sx * p.dil, hbase + kw.r * p.dil,
                        t.n, cb / BK + kw.c);
            tma_load_3d(sb, &p.tmap_b, &full[stage], 0, brow, st * KSUB);
          }
          kw.advance(KSUB, p);
        } else {
          const bool leader = elect_one();
          if (leader) {
            mbar_arrive_expect_tx(&full[stage], KSUB * (Cfg::A_BYTES + Cfg::B_BYTES));
            tma_load_3d(sb, &p.tmap_b, &full[stage], 0, brow, st * KSUB);
          }
#pragma unroll
          for (int u = 0; u < KSUB; ++u) {
            const int cb = p.a_coff[kw.i] + t.g * p.a_goff[kw.i] + kw.c * BK;
            if (leader)
              tma_load_4d(sa + u * Cfg::A_BYTES, &p.tmap_a[kw.i], &full[stage], cb, wbase + kw.sx * p.dil,
                          hbase + kw.r * p.dil, t.n);
            kw.advance(1, p);
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
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
      for (int st = 0; st < n_stages; ++st) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE);
          const uint64_t da = umma_smem_desc<Cfg::ROW_BYTES>(sa);
          const uint64_t db = umma_smem_desc<Cfg::ROW_BYTES>(sa + KSUB * Cfg::A_BYTES);
#pragma unroll
          for (int u = 0; u < KSUB; ++u) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_f16(tmem_acc, da + (u * Cfg::A_BYTES >> 4) + 2 * k, db + (u * Cfg::B_BYTES >> 4) + 2 * k, idesc,
                       (st | u | k) != 0);
          }
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(&tfull[as]);
      __syncwarp();
    }
  } else {
    const int quarter = warp & 3;
    const int ep_tid = threadIdx.x - 64;
    StagedEpilogue st;
    if constexpr (STAGED) st.init(stage_out, reinterpret_cast<float*>(stage_out + 2 * kSlabBytes), res_bar);
    int local = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
      const TileCoord t = decode_tile(tile, p.n_tiles, p.tiles_w, p.tiles_h, p.N, p.TW, p.TH);
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const float* bias = p.bias ? p.bias + t.g * p.bias_goff : nullptr;
      if constexpr (STAGED && sizeof(OutT) == 4) {
        const int col0 = t.nt * BN;
        epilogue_tile_staged_f32<BN, NW>(st, tmem_base + as * Cfg::ACC, quarter, lane, ep_tid, t.n, t.p0, t.q0,
                                         &p.tmap_out, p.out_coff + t.g * p.out_goff + col0, bias ? bias + col0 : nullptr,
                                         p.Cout - col0, p.slope, [&]() {
                                           tc_fence_before();
                                           mbar_arrive(&tempty[as]);
                                         });
      } else if constexpr (STAGED) {
        const int col0 = t.nt * BN;
        epilogue_tile_staged<BN, NW>(st, tmem_base + as * Cfg::ACC, quarter, lane, ep_tid, t.n, t.p0, t.q0, &p.tmap_out,
                                 p.out_coff + t.g * p.out_goff + col0, p.res ? &p.tmap_res : nullptr,
                                 p.res_coff + t.g * p.res_goff + col0, bias ? bias + col0 : nullptr, p.Cout - col0,
                                 p.slope, [&]() {
                                   tc_fence_before();
                                   mbar_arrive(&tempty[as]);
                                 });
      } else {
        const __nv_bfloat16* res =
            p.res ? static_cast<const __nv_bfloat16*>(p.res) + p.res_coff + t.g * p.res_goff : nullptr;
        OutT* out = static_cast<OutT*>(p.out) + p.out_coff + t.g * p.out_goff;
        epilogue_tile_direct<BN, OutT, __nv_bfloat16>(tmem_base + as * Cfg::ACC, quarter, lane, t.n, t.p0, t.q0, p.TW,
                                                      p.P, p.Q, t.nt * BN, p.Cout, bias, res, p.res_cstride, out,
                                                      p.out_cstride, p.slope);
        tc_fence_before();
        mbar_arrive(&tempty[as]);
      }
    }
    if (STAGED && ep_tid == 0) tma_store_wait_all();
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

// Phase timeline probe of the gather kernel (tools/probe_stem.py): compile with -DM3D_PROBE.  Stamps go to shared
// memory (a global store per stamp would stall the stamping warp's next MEMBAR) and are copied out at the end.
#ifdef M3D_PROBE
__device__ long long g_gather_dbg[6 * 32];
#define GDBG(slot) do { if (blockIdx.x == 0 && lane == 0 && local < 6) s_gdbg[local * 32 + (slot)] = clock64(); } while (0)
#else
#define GDBG(slot) do { } while (0)
#endif

template <int BN, bool SPLIT, bool STAGED>
struct GatherCfg {
  static constexpr int ROW_BYTES = 128;
  static constexpr int A_BYTES = kTileM * ROW_BYTES;  // per part
  static constexpr int B_BYTES = BN * ROW_BYTES;      // per part
  static constexpr int PARTS = SPLIT ? 3 : 1;
  static constexpr int STAGE = PARTS * (A_BYTES + B_BYTES);
  // producer scratch: SPLIT: the tile's offsets/masks [128][27+1] fp32; bf16: the per-tile sample table
  // [128 rows][9 taps] x {int4 corner offsets, float4 corner weights * mask} (also the stem's image tile)
  static constexpr int OM_BYTES = SPLIT ? kTileM * 28 * 4 : kTileM * 9 * 32;
  static constexpr int EXTRA = STAGED ? (2 * kSlabBytes + 1024) : 0;
  static constexpr int BUDGET = 224 * 1024 - OM_BYTES - EXTRA;
  static constexpr int STAGES = BUDGET / STAGE >= 4 ? 4 : (BUDGET / STAGE >= 3 ? 3 : 2);
  static constexpr int SMEM = STAGES * STAGE + EXTRA + OM_BYTES + 1024 + 256;
  static_assert(SMEM <= 227 * 1024, "gather tile does not fit shared memory");
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

template <int BN, typename InT, typename OutT, bool STAGED>
__global__ void __launch_bounds__(448, 1) conv_gather_kernel(const __grid_constant__ ConvGatherParams p) {
#ifdef M3D_PROBE
  __shared__ long long s_gdbg[6 * 32];
  if (threadIdx.x < 6 * 32) s_gdbg[threadIdx.x] = 0;
#endif
  constexpr bool SPLIT = sizeof(InT) == 4;
  using Cfg = GatherCfg<BN, SPLIT, STAGED>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BK = kGatherBK;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // stage layout: [A_hi][A_mid][A_lo][B_hi][B_mid][B_lo]  (one part each when !SPLIT)
  uint8_t* stage_out = smem + STAGES * Cfg::STAGE;  // STAGED: 2 slabs + bias (1 KB)
  float* om_s = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE + Cfg::EXTRA);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE + Cfg::EXTRA + Cfg::OM_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* res_bar = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], kProducerThreads / 32 + 1);  // one arrival per producer warp + the weight TMA
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 128);
      mbar_init(&res_bar[s], 1);
    }
    fence_barrier_init();
    prefetch_tmap(&p.tmap_b);
    if (STAGED) prefetch_tmap(&p.tmap_out);
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
  grid_dep_sync();

  const int taps = p.R * p.S;
  int total_kb = 0;
  for (int i = 0; i < p.num_inputs; ++i) total_kb += taps * p.chunks[i];

  if (warp < 8) {
    // ------------------------------------------------------------ A producers
    const int pt = threadIdx.x;
    const int j = pt & 7;        // 16-byte chunk (8 channels) inside the 64-channel k-block
    const int rbase = pt >> 3;   // rows rbase + 32*i
    const int omc = 3 * taps;
    const int tw_shift = 31 - __clz(p.TW);  // TW is a power of two
    int stage = 0;
    uint32_t phase = 0;
    if (!SPLIT && p.stem_img != nullptr) {
      // ---- stem: 7x7 conv of the fp32 NCHW image in 2x2 space-to-depth form.  Row (Y, X) of the tile
      // is the 8x8 window at (2Y-3, 2X-3); k-block kb = image channel kb, 16-byte chunk j = window row
      // j, its 8 elements = 8 consecutive image columns.  The image tile (3 x (2TH+6) x (2TW+6)) is
      // staged once per tile in shared memory with coalesced loads.
      const int th2 = 2 * p.TH + 6, ld = 2 * p.TW + 8;  // TMA box: ld x th2 x 3 fp32, zero outside the image
      const uint32_t img_bytes = static_cast<uint32_t>(3 * th2 * ld * 4);
      // double-buffered (host checks the fit).  The buffers are addressed as offsets from the shared base (an array
      // of pointers made the loads generic LD.E: long-scoreboard stalls on every window read).
      const int img_stride = static_cast<int>((img_bytes + 127) / 128) * 32;
      auto load_img = [&](int tl, int buf) {
        const TileCoord tn = decode_tile(tl, p.n_tiles, p.tiles_w, p.tiles_h, p.N, p.TW, p.TH);
        mbar_arrive_expect_tx(&res_bar[buf], img_bytes);
        // x origin 2*q0 - 4 keeps the innermost coordinate 16-byte aligned (window columns start at +1)
        tma_load_4d(om_s + buf * img_stride, &p.tmap_img, &res_bar[buf], 2 * tn.q0 - 4, 2 * tn.p0 - 3, 0, tn.n);
      };
      if (pt == 0 && blockIdx.x < p.total_tiles) load_img(blockIdx.x, 0);
      int local = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
        const int buf = local & 1;
        named_bar_sync(1, kProducerThreads);  // the previous tile's readers are done with the other buffer
        if (warp == 0) GDBG(0);
        if (pt == 0 && tile + gridDim.x < p.total_tiles) load_img(tile + gridDim.x, buf ^ 1);  // next tile's image
        mbar_wait(&res_bar[buf], (local >> 1) & 1);
        if (warp == 0) GDBG(1);
        const uint32_t s_img = smem_u32(om_s) + static_cast<uint32_t>(buf * img_stride) * 4u;
        for (int kb = 0; kb < 3; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* a_hi = smem + stage * Cfg::STAGE;
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) {
            const int row = rbase + 32 * ii;
            // window = floats 1..8 of an 8-byte aligned span: ld.shared 32 + 3 x 64 + 32
            const uint32_t src =
                s_img + static_cast<uint32_t>((kb * th2 + 2 * (row >> tw_shift) + j) * ld + 2 * (row & (p.TW - 1))) * 4u;
            float v[8];
            asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(v[0]) : "r"(src));
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+8];" : "=f"(v[1]), "=f"(v[2]) : "r"(src));
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+16];" : "=f"(v[3]), "=f"(v[4]) : "r"(src));
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+24];" : "=f"(v[5]), "=f"(v[6]) : "r"(src));
            asm volatile("ld.shared.f32 %0, [%1+32];" : "=f"(v[7]) : "r"(src));
            *reinterpret_cast<uint4*>(a_hi + swizzled_offset<128>(row, j)) = pack8(v);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[stage]);
          if (warp == 0) GDBG(2 + kb);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if constexpr (!SPLIT) {
      // ---- bf16 deformable gather.  Phase 0 per tile: the 256 producers fill a shared-memory table with, for
      // every (row, tap), the four clamped corner offsets and the four bilinear weights already multiplied by
      // validity and by the modulation mask (dcn_v2_im2col_cuda.cu:18-47,151-175).  Main loop: software
      // pipeline over half k-blocks (2 of the thread's 4 rows): the 8 corner loads of unit u+1 are issued
      // before unit u is blended, so the L1/L2 latency overlaps the fp32 blend instead of serialising.
      struct Entry {
        int4 off;
        float4 w;
      };
      Entry* table = reinterpret_cast<Entry*>(om_s);
      const InT* in0 = static_cast<const InT*>(p.in[0]) + p.in_coff[0];
      const int cs = p.in_cstride[0];
      const int nchunk = p.chunks[0];
      const int units = taps * nchunk * 2;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(tile, p.n_tiles, p.tiles_w, p.tiles_h, p.N, p.TW, p.TH);
        named_bar_sync(1, kProducerThreads);  // previous tile's table readers are done
        // thread pt owns row pt/2 and the taps of parity pt%2: at most 5 entries, their 15 offset / mask
        // loads issued together
        const int trow = pt >> 1;
        const int tpp = t.p0 + (trow >> tw_shift), tqq = t.q0 + (trow & (p.TW - 1));
        const bool tok = tpp < p.P && tqq < p.Q;
        const float* om_px = p.om + ((static_cast<long>(t.n) * p.P + tpp) * p.Q + tqq) * p.om_cstride;
        float o_h[5], o_w[5], o_m[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const int tap = (pt & 1) + 2 * k;
          o_h[k] = o_w[k] = 0.f, o_m[k] = 1.f;
          if (tok && tap < taps && p.om != nullptr) {
            o_h[k] = __ldg(om_px + 2 * tap);
            o_w[k] = __ldg(om_px + 2 * tap + 1);
            o_m[k] = __ldg(om_px + 2 * taps + tap);
          }
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const int tap = (pt & 1) + 2 * k;
          if (tap >= taps) break;
          const int row = trow, pp = tpp, qq = tqq;
          Entry e;
          e.off = make_int4(0, 0, 0, 0);
          e.w = make_float4(0.f, 0.f, 0.f, 0.f);
          if (tok) {
            const int r = taps == 9 ? tap / 3 : tap / p.S, sx = tap - r * p.S;
            float hf = static_cast<float>(pp * p.stride - p.pad + r * p.dil) + o_h[k];
            float wf = static_cast<float>(qq * p.stride - p.pad + sx * p.dil) + o_w[k];
            float m = o_m[k];
            if (p.om != nullptr && p.sigmoid_mask) m = 1.f / (1.f + __expf(-m));
            {
            }
            if (hf > -1.f && wf > -1.f && hf < static_cast<float>(p.H) && wf < static_cast<float>(p.W)) {
              const float hl = floorf(hf), wl = floorf(wf);
              const int h_low = static_cast<int>(hl), w_low = static_cast<int>(wl);
              const int h_high = h_low + 1, w_high = w_low + 1;
              const float lh = hf - hl, lw = wf - wl, hh = 1.f - lh, hw = 1.f - lw;
              const bool hl_ok = h_low >= 0, wl_ok = w_low >= 0, hh_ok = h_high <= p.H - 1, wh_ok = w_high <= p.W - 1;
              const int rl = (t.n * p.H + (hl_ok ? h_low : 0)) * p.W, rh = (t.n * p.H + (hh_ok ? h_high : 0)) * p.W;
              const int cl = wl_ok ? w_low : 0, ch = wh_ok ? w_high : 0;
              e.off = make_int4((rl + cl) * cs, (rl + ch) * cs, (rh + cl) * cs, (rh + ch) * cs);
              e.w = make_float4((hl_ok && wl_ok) ? hh * hw * m : 0.f, (hl_ok && wh_ok) ? hh * lw * m : 0.f,
                                (hh_ok && wl_ok) ? lh * hw * m : 0.f, (hh_ok && wh_ok) ? lh * lw * m : 0.f);
            }
          }
          table[row * taps + tap] = e;
        }
        named_bar_sync(1, kProducerThreads);

        // unit u -> (tap, chunk, half); rows rbase + 32 * (2*half + {0,1})
        uint4 cv[2][2][4];
        float4 cw[2][2];
        int nx_tap = 0, nx_c = 0;  // (tap, chunk) of the next unit to issue
        auto issue = [&](int u, int buf) {
          const int half = u & 1;
          const int tap = nx_tap, c = nx_c;
          if (half) {  // chunk-major walk of K: the 9 taps of one 64-channel chunk share their input window
            if (++nx_tap == taps) nx_tap = 0, ++nx_c;
          }
          const int coff = c * BK + j * 8;
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const Entry& e = table[(rbase + 32 * (2 * half + rr)) * taps + tap];  // taps is 1 or 9: cheap multiply
            const int4 o = e.off;
            cw[buf][rr] = e.w;
            cv[buf][rr][0] = __ldg(reinterpret_cast<const uint4*>(in0 + o.x + coff));
            cv[buf][rr][1] = __ldg(reinterpret_cast<const uint4*>(in0 + o.y + coff));
            cv[buf][rr][2] = __ldg(reinterpret_cast<const uint4*>(in0 + o.z + coff));
            cv[buf][rr][3] = __ldg(reinterpret_cast<const uint4*>(in0 + o.w + coff));
          }
        };
        auto blend = [&](int u, int buf, uint8_t* a_hi) {
          const int half = u & 1;
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const float4 w4 = cw[buf][rr];
            const uint32_t* q0 = reinterpret_cast<const uint32_t*>(&cv[buf][rr][0]);
            const uint32_t* q1 = reinterpret_cast<const uint32_t*>(&cv[buf][rr][1]);
            const uint32_t* q2 = reinterpret_cast<const uint32_t*>(&cv[buf][rr][2]);
            const uint32_t* q3 = reinterpret_cast<const uint32_t*>(&cv[buf][rr][3]);
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float lo = w4.x * __uint_as_float(q0[e] << 16) + w4.y * __uint_as_float(q1[e] << 16) +
                               w4.z * __uint_as_float(q2[e] << 16) + w4.w * __uint_as_float(q3[e] << 16);
              const float hi = w4.x * __uint_as_float(q0[e] & 0xffff0000u) + w4.y * __uint_as_float(q1[e] & 0xffff0000u) +
                               w4.z * __uint_as_float(q2[e] & 0xffff0000u) + w4.w * __uint_as_float(q3[e] & 0xffff0000u);
              __nv_bfloat162 tt = __floats2bfloat162_rn(lo, hi);
              w[e] = *reinterpret_cast<uint32_t*>(&tt);
            }
            *reinterpret_cast<uint4*>(a_hi + swizzled_offset<128>(rbase + 32 * (2 * half + rr), j)) =
                make_uint4(w[0], w[1], w[2], w[3]);
          }
        };
        issue(0, 0);
#pragma unroll 1
        for (int u = 0; u < units; u += 2) {
          // half 0 of k-block u/2 (buffer 0), then half 1 (buffer 1)
          issue(u + 1, 1);
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* a_hi = smem + stage * Cfg::STAGE;
          blend(u, 0, a_hi);
          if (u + 2 < units) issue(u + 2, 0);
          blend(u + 1, 1, a_hi);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else
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
            if constexpr (!SPLIT) {
              // all 16 corner loads of this thread's 4 rows go out before the first blend: invalid corners
              // carry weight 0 and a clamped (in-range) address, so there is nothing to branch on
              uint4 cv[4][4];
#pragma unroll
              for (int ii = 0; ii < 4; ++ii) {
                cv[ii][0] = __ldg(reinterpret_cast<const uint4*>(in + si[ii].o00 + coff));
                cv[ii][1] = __ldg(reinterpret_cast<const uint4*>(in + si[ii].o01 + coff));
                cv[ii][2] = __ldg(reinterpret_cast<const uint4*>(in + si[ii].o10 + coff));
                cv[ii][3] = __ldg(reinterpret_cast<const uint4*>(in + si[ii].o11 + coff));
              }
#pragma unroll
              for (int ii = 0; ii < 4; ++ii) {
                const SampleInfo& x = si[ii];
                const uint32_t* q0 = reinterpret_cast<const uint32_t*>(&cv[ii][0]);
                const uint32_t* q1 = reinterpret_cast<const uint32_t*>(&cv[ii][1]);
                const uint32_t* q2 = reinterpret_cast<const uint32_t*>(&cv[ii][2]);
                const uint32_t* q3 = reinterpret_cast<const uint32_t*>(&cv[ii][3]);
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float lo = (x.w00 * __uint_as_float(q0[e] << 16) + x.w01 * __uint_as_float(q1[e] << 16) +
                                    x.w10 * __uint_as_float(q2[e] << 16) + x.w11 * __uint_as_float(q3[e] << 16)) * x.mask;
                  const float hi = (x.w00 * __uint_as_float(q0[e] & 0xffff0000u) + x.w01 * __uint_as_float(q1[e] & 0xffff0000u) +
                                    x.w10 * __uint_as_float(q2[e] & 0xffff0000u) + x.w11 * __uint_as_float(q3[e] & 0xffff0000u)) * x.mask;
                  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
                  w[e] = *reinterpret_cast<uint32_t*>(&t);
                }
                *reinterpret_cast<uint4*>(a_hi + swizzled_offset<128>(rbase + 32 * ii, j)) = make_uint4(w[0], w[1], w[2], w[3]);
              }
            } else {
#pragma unroll 1
              for (int ii = 0; ii < 4; ++ii) {
                const int row = rbase + 32 * ii;
                const SampleInfo& x = si[ii];
                float v1[8], v2[8], v3[8], v4[8], acc[8];
                load8(in + x.o00 + coff, v1);
                load8(in + x.o01 + coff, v2);
                load8(in + x.o10 + coff, v3);
                load8(in + x.o11 + coff, v4);
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  acc[e] = (x.w00 * v1[e] + x.w01 * v2[e] + x.w10 * v3[e] + x.w11 * v4[e]) * x.mask;
                const uint32_t off = swizzled_offset<128>(row, j);
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
              }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[stage]);
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
    const bool chunk_major = !SPLIT && p.stem_img == nullptr;  // must match the A producers' walk of K
    const int nchunk0 = p.chunks[0];
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(tile, p.n_tiles, p.tiles_w, p.tiles_h, p.N, p.TW, p.TH);
      int tap = 0, c = 0;
      for (int kb = 0; kb < total_kb; ++kb) {
        const int kblk = chunk_major ? tap * nchunk0 + c : kb;
        if (++tap == taps) tap = 0, ++c;
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* b_hi = smem + stage * Cfg::STAGE + Cfg::PARTS * Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full[stage], Cfg::PARTS * Cfg::B_BYTES);
          tma_load_2d(b_hi, &p.tmap_b, &full[stage], kblk * BK, t.nt * BN);
          if constexpr (SPLIT) {
            tma_load_2d(b_hi + Cfg::B_BYTES, &p.tmap_b_mid, &full[stage], kblk * BK, t.nt * BN);
            tma_load_2d(b_hi + 2 * Cfg::B_BYTES, &p.tmap_b_lo, &full[stage], kblk * BK, t.nt * BN);
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(BN);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      mbar_wait(&tempty[as], aphase ^ 1);
      GDBG(8);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + as * Cfg::ACC;
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        if (kb < 3) GDBG(9 + kb);
        tc_fence_after();
        if (elect_one()) {
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
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(&tfull[as]);
      __syncwarp();
      GDBG(12);
    }
  } else {
    // --------------------------------------------------------------- epilogue
    const int quarter = warp & 3;
    const int ep_tid = threadIdx.x - 320;
    StagedEpilogue st;
    if constexpr (STAGED) st.init(stage_out, reinterpret_cast<float*>(stage_out + 2 * kSlabBytes), res_bar);
    int local = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
      const TileCoord t = decode_tile(tile, p.n_tiles, p.tiles_w, p.tiles_h, p.N, p.TW, p.TH);
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      mbar_wait(&tfull[as], aphase);
      if (warp == 10) GDBG(16);
      tc_fence_after();
      if constexpr (STAGED) {
        const int col0 = t.nt * BN;
        epilogue_tile_staged<BN, 4>(st, tmem_base + as * Cfg::ACC, quarter, lane, ep_tid, t.n, t.p0, t.q0, &p.tmap_out,
                                 p.out_coff + col0, p.res ? &p.tmap_res : nullptr, p.res_coff + col0,
                                 p.bias ? p.bias + col0 : nullptr, p.Cout - col0, p.slope, [&]() {
                                   tc_fence_before();
                                   mbar_arrive(&tempty[as]);
                                 });
      } else {
        const InT* res = p.res ? static_cast<const InT*>(p.res) + p.res_coff : nullptr;
        OutT* out = static_cast<OutT*>(p.out) + p.out_coff;
        epilogue_tile_direct<BN, OutT, InT>(tmem_base + as * Cfg::ACC, quarter, lane, t.n, t.p0, t.q0, p.TW, p.P, p.Q,
                                            t.nt * BN, p.Cout, p.bias, res, p.res_cstride, out, p.out_cstride, p.slope);
        tc_fence_before();
        mbar_arrive(&tempty[as]);
      }
      if (warp == 10) GDBG(17);
    }
    if (STAGED && ep_tid == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
#ifdef M3D_PROBE
  if (blockIdx.x == 0 && threadIdx.x < 6 * 32) g_gather_dbg[threadIdx.x] = s_gdbg[threadIdx.x];
#endif
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<2 * Cfg::ACC>(tmem_base);
  }
}

// ------------------------------------------------------------ host launchers

template <int BN, int BK, int KSUB, typename OutT, bool STAGED>
static int launch_tma_t(const ConvTmaParams& p, cudaStream_t stream) {
  using Cfg = TmaCfg<BN, BK, KSUB, STAGED>;
  auto kern = conv_tma_kernel<BN, BK, KSUB, OutT, STAGED>;
  set_last_kernel("conv_tma_kernel<%d,%d,%d,%s,%d>", BN, BK, KSUB, sizeof(OutT) == 4 ? "f32" : "bf16", int(STAGED));
  M3D_ONCE_PER_DEVICE_BEGIN
    M3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  M3D_ONCE_PER_DEVICE_END
  // Two co-resident CTAs per SM when shared memory allows (small tiles).
  const int per_sm = (2 * (Cfg::SMEM + 1024) <= 227 * 1024 && 4 * Cfg::ACC <= 512) ? 2 : 1;
  int grid = persistent_sms() * per_sm;
  if (grid > p.total_tiles) grid = p.total_tiles;
  M3D_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(64 + 32 * conv_epi_warps<STAGED>()), Cfg::SMEM, stream, p));
  return M3D_OK;
}

int launch_conv_tma(const ConvTmaParams& p, int BN, int BK, int ksub, int out_dtype, bool staged,
                    cudaStream_t stream) {
#define M3D_TMA_CASE(bn, bk)                                                                  \
  if (BN == bn && BK == bk && ksub == 1) {                                                    \
    return out_dtype == DT_BF16 ? launch_tma_t<bn, bk, 1, __nv_bfloat16, false>(p, stream)    \
                                : launch_tma_t<bn, bk, 1, float, false>(p, stream);           \
  }
#define M3D_TMA_STAGED(bn, ks)                                                             \
  if (staged && BN == bn && BK == 64 && ksub == ks && out_dtype == DT_BF16)                \
    return launch_tma_t<bn, 64, ks, __nv_bfloat16, true>(p, stream);
  if (staged && BN == 256 && BK == 64 && ksub == 1 && out_dtype == DT_F32)  // fp32 logits through staging + TMA stores
    return launch_tma_t<256, 64, 1, float, true>(p, stream);
  if (staged && BN == 64 && BK == 32 && ksub == 1 && out_dtype == DT_BF16)
    return launch_tma_t<64, 32, 1, __nv_bfloat16, true>(p, stream);
  M3D_TMA_STAGED(64, 1)
  M3D_TMA_STAGED(64, 2)
  M3D_TMA_STAGED(64, 3)
  M3D_TMA_STAGED(64, 4)
  M3D_TMA_STAGED(128, 1)
  M3D_TMA_STAGED(128, 2)
  M3D_TMA_STAGED(256, 1)
  M3D_TMA_STAGED(256, 2)
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
#undef M3D_TMA_STAGED
  return M3D_ERR_UNSUPPORTED;
}

template <int BN, typename InT, typename OutT, bool STAGED>
static int launch_gather_t(const ConvGatherParams& p, cudaStream_t stream) {
  using Cfg = GatherCfg<BN, sizeof(InT) == 4, STAGED>;
  auto kern = conv_gather_kernel<BN, InT, OutT, STAGED>;
  set_last_kernel("conv_gather_kernel<%d,%s,%s,%d>", BN, sizeof(InT) == 4 ? "f32" : "bf16", sizeof(OutT) == 4 ? "f32" : "bf16", int(STAGED));
  M3D_ONCE_PER_DEVICE_BEGIN
    M3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  M3D_ONCE_PER_DEVICE_END
  int grid = persistent_sms();
  if (grid > p.total_tiles) grid = p.total_tiles;
  M3D_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(448), Cfg::SMEM, stream, p));
  return M3D_OK;
}

int launch_conv_gather(const ConvGatherParams& p, int BN, int in_dtype, int out_dtype, bool staged,
                       cudaStream_t stream) {
#define M3D_G_STAGED(bn)                                                    \
  if (staged && BN == bn && in_dtype == DT_BF16 && out_dtype == DT_BF16)    \
    return launch_gather_t<bn, __nv_bfloat16, __nv_bfloat16, true>(p, stream);
#define M3D_G_CASE(bn)                                                                                  \
  if (BN == bn) {                                                                                       \
    if (in_dtype == DT_BF16)                                                                            \
      return out_dtype == DT_BF16 ? launch_gather_t<bn, __nv_bfloat16, __nv_bfloat16, false>(p, stream) \
                                  : launch_gather_t<bn, __nv_bfloat16, float, false>(p, stream);        \
    if constexpr (bn <= 128) return launch_gather_t<bn, float, float, false>(p, stream);                \
    return M3D_ERR_UNSUPPORTED;                                                                         \
  }
  if (dcn_fused_supported(p, BN, in_dtype, out_dtype)) return launch_dcn_fused(p, BN, stream);
  M3D_G_STAGED(64)
  M3D_G_STAGED(128)
  M3D_G_STAGED(256)
  M3D_G_CASE(16)
  M3D_G_CASE(32)
  M3D_G_CASE(48)
  M3D_G_CASE(64)
  M3D_G_CASE(128)
  M3D_G_CASE(256)
#undef M3D_G_CASE
#undef M3D_G_STAGED
  return M3D_ERR_UNSUPPORTED;
}

}  // namespace m3d

#ifdef M3D_PROBE
extern "C" int m3d_gather_debug_read(long long* host, int n) {
  return cudaMemcpyFromSymbol(host, m3d::g_gather_dbg, sizeof(long long) * n) == cudaSuccess ? 0 : -1;
}
#endif
