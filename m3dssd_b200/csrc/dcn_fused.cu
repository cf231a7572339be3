// Fused bf16 DCNv2 layer (throughput mode): bilinear offset/mask sampling + output GEMM in one kernel,
// following modulated_deformable_im2col_gpu_kernel + SGEMM of the reference
// (model/DCNv2/src/cuda/dcn_v2_im2col_cuda.cu:18-47,118-180; model/DCNv2/src/dcn_v2_cuda.c:61-97) with no
// `columns` buffer.  Same GEMM view as igemm.cu (M = 128 output pixels, N = BN, k-blocks of 64 channels).
//
// What bounds the layer is the A producer, not the tensor pipe: every A element needs four 2-byte corner reads, i.e.
// 64 KB of gathered 16-byte loads per k-block through the SM's 128 B/clk L1 path (>= 512 clk, against 256 tensor
// clocks for N = 128), plus the blend.  Round 1 blended in fp32 (unpack + FFMA2 + pack: ~228 SASS instructions per
// thread and k-block, 912 issue clocks, the ALU pipe alone 512).  Round 2 (this file):
//   * blend on packed bf16 pairs straight from the loaded vectors (M3D_DCN_BLEND 2: mul/fma.rn.bf16x2 = HFMA2.BF16,
//     16 per 8-channel row; 1: fma.rn.f32.bf16 = FHFMA with fp32 accumulation, 32 per row; 0: the fp32 FFMA2 blend);
//   * 16-byte table entries (one clamped base offset of the 2x2 corner patch + four bf16 weights already multiplied
//     by validity and modulation mask; the other three corners are base + constants), so the table of the NEXT tile
//     fits beside the current one and is built by the otherwise idle epilogue warps while the producers run: the
//     4.8 k-clk bubble between tiles (two block barriers, offset/mask loads, table build) is gone;
//   * the weights of a unit travel with its corner loads through the register ring (no table re-read at blend time).
//
//   warps 0-15 : A producers: per k-block each thread blends 2 rows x 8 channels (4 x 16-byte corner loads per row,
//                issued one k-block ahead of the blend) and writes the swizzled bf16 A tile.
//                After the first k-block of a tile they drain the PREVIOUS tile's accumulator (TMEM -> bias
//                (+ residual) -> LeakyReLU -> bf16 NHWC; warp w: lane quarter w % 4, columns (w / 4) * BN / 4 ...).
//   warp 16    : weight tiles by TMA, one k-block ahead (K walked chunk-major: the 9 taps of a 64-channel chunk
//                re-read one input window, which keeps the corner loads in L1/L2), and tcgen05.mma issue,
//                accumulators in TMEM (two stages).
//   warps 17-19: sample table of the next tile.
#include <cstdlib>

#include "common.cuh"
#include "epilogue.cuh"
#include "igemm.cuh"
#include "ptx.cuh"

#ifndef M3D_DCN_BLEND
#define M3D_DCN_BLEND 2
#endif

namespace m3d {

// bf16 NHWC activation as (C, W, H, N), box {box_c, box_w, box_h, 1}, NO swizzle, zero fill outside (api_conv.cu)
int make_tmap_nhwc_plain(CUtensorMap* map, const void* base, int N, int H, int W, int C, int box_c, int box_w, int box_h);

namespace {

constexpr int kProd = 512;  // producer threads

// Per-k-block timeline probe (tools/probe_dcn_timeline.py): compile with -DM3D_PROBE.  Block 0, first tile, first 24
// k-blocks; stamps in shared memory, copied out at kernel end.
#ifdef M3D_PROBE
__device__ long long g_dcn_dbg[32 * 8];
#define DDBG(kbi, slot) do { if (blockIdx.x == 0 && lane == 0 && (kbi) < 32) s_ddbg[(kbi) * 8 + (slot)] = clock64(); } while (0)
#else
#define DDBG(kbi, slot) do { } while (0)
#endif
// 20 warps = 640 threads: registers are granted per 4 warps, so 20 warps get 96 registers per thread (21-24 warps:
// 80, which spilled ~25 registers of the producers' three-unit load ring; ptxas does not raise a role's budget on
// setmaxnreg here).  Hence: ONE warp both loads the weight tiles (TMA) and issues the MMAs, the producer warps drain
// the previous tile's accumulator themselves (each a 32-row x BN/4-column piece, in the shadow of their own loads),
// and the three remaining warps build the sample table one tile ahead.
constexpr int kBuildThreads = 96;
constexpr int kDcnThreads = kProd + 32 + kBuildThreads;
constexpr int kBK = 64;
constexpr int kTaps = 9;

struct Entry {     // (stored as 3 or 4 words, see DcnCfg::EW)
  uint32_t off;    // byte offset of corner (h0, w0) of the clamped 2x2 patch, channel 0
  uint32_t w01;    // bf16x2: weights of (h0, w0), (h0, w0 + 1)
  uint32_t w23;    // bf16x2: weights of (h0 + 1, w0), (h0 + 1, w0 + 1)
  uint32_t glob;   // HALO mode: 0 = `off` is a byte offset into the staged window, 1 = global byte offset (fallback)
};

// HALO mode: the input window of a (tile, 64-channel chunk) -- the tile's footprint grown by `halo` pixels on every
// side -- is staged in shared memory by TMA (zero fill outside the image), two buffers so that the next chunk's window
// lands while the current one is sampled.  The corner loads then are ld.shared (fixed ~30 clk, no tag lookups, no
// misses; the L1 path hit only 61 % and ran at 71 % of its peak); samples whose 2x2 patch leaves the window fall back
// to global loads, entry by entry.
constexpr int kHaloBytes = 60 * 1024;

template <int BN, int NSTG, bool HALO = false>
struct DcnCfg {
  static constexpr int A_BYTES = kTileM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES = NSTG;
  // table entry = 3 words (offset, 2 x bf16x2 weights), 4 with the window / global flag of HALO mode; with 3 words the
  // whole CTA stays under 100 KB of shared memory, i.e. the 100 KB carve-out: 128 KB of L1 instead of 96 KB for the
  // corner loads (the 16-byte entry of the first version cost exactly that step)
  static constexpr int EW = HALO ? 4 : 3;
  static constexpr int TABLE1 = kTileM * kTaps * EW * 4;  // one tile's table
  static constexpr int TABLE = (2 * TABLE1 + 1023) / 1024 * 1024;
  static constexpr int HALO_OFF = STAGES * STAGE + TABLE;
  static constexpr int BARS_OFF = HALO_OFF + (HALO ? 2 * kHaloBytes : 0);
  static constexpr int SMEM = BARS_OFF + 1024 + 256;
  static constexpr int ACC = BN <= 128 ? 128 : 256;
  static_assert(SMEM <= 227 * 1024, "DCN tile does not fit shared memory");
};

struct Tile {
  int nt, n, p0, q0;
};
__device__ __forceinline__ Tile tile_of(int tile, const ConvGatherParams& p) {
  Tile t;
  t.nt = tile % p.n_tiles;
  int r = tile / p.n_tiles;
  const int tw = r % p.tiles_w;
  r /= p.tiles_w;
  const int th = r % p.tiles_h;
  t.n = r / p.tiles_h;
  t.p0 = th * p.TH;
  t.q0 = tw * p.TW;
  return t;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// One sample-table entry (dcn_v2_im2col_cuda.cu:18-47,151-175).  The reference reads the four corners (h_low, w_low)
// ... (h_low + 1, w_low + 1) and drops those outside the image.  Here the 2x2 patch is CLAMPED into the image
// (h0 = clamp(h_low, 0, H - 2), likewise w0) so that its four addresses are base + constants, and each of its rows /
// columns gets the weight of the reference corner it coincides with (or 0): identical sample value.
template <bool HALO>
__device__ __forceinline__ Entry make_entry(const ConvGatherParams& p, int n, int pp, int qq, int tap, int cs, float o_h,
                                            float o_w, float m, int wy0, int wx0) {
  Entry e;
  e.off = 0, e.w01 = 0, e.w23 = 0, e.glob = 0;
  const int r = tap / 3, sx = tap - r * 3;
  const float hf = static_cast<float>(pp * p.stride - p.pad + r * p.dil) + o_h;
  const float wf = static_cast<float>(qq * p.stride - p.pad + sx * p.dil) + o_w;
  if (p.sigmoid_mask) m = 1.f / (1.f + __expf(-m));
  if (hf > -1.f && wf > -1.f && hf < static_cast<float>(p.H) && wf < static_cast<float>(p.W)) {
    const float hl = floorf(hf), wl = floorf(wf);
    const int h_low = static_cast<int>(hl), w_low = static_cast<int>(wl);
    const float lh = hf - hl, lw = wf - wl, hh = 1.f - lh, hw = 1.f - lw;
    const int h0 = min(max(h_low, 0), p.H - 2), w0 = min(max(w_low, 0), p.W - 2);
    const float wr0 = h0 == h_low ? hh : (h0 == h_low + 1 ? lh : 0.f);
    const float wr1 = h0 == h_low ? lh : (h0 + 1 == h_low ? hh : 0.f);
    const float wc0 = w0 == w_low ? hw : (w0 == w_low + 1 ? lw : 0.f);
    const float wc1 = w0 == w_low ? lw : (w0 + 1 == w_low ? hw : 0.f);
    e.off = static_cast<uint32_t>((n * p.H + h0) * p.W + w0) * static_cast<uint32_t>(cs * 2);
    if constexpr (HALO) {
      const int dy = h0 - wy0, dx = w0 - wx0;  // patch position inside the staged window
      if (dy >= 0 && dx >= 0 && dy + 1 < p.halo_h && dx + 1 < p.halo_w)
        e.off = static_cast<uint32_t>(dy * p.halo_w + dx) * 128u;
      else
        e.glob = 1;
    }
    e.w01 = pack_bf16x2(wr0 * wc0 * m, wr0 * wc1 * m);
    e.w23 = pack_bf16x2(wr1 * wc0 * m, wr1 * wc1 * m);
  }
  return e;
}

// ---- the blend of one 8-channel row: four corner vectors (bf16x2 x 4 each) -> one 16-byte A chunk
__device__ __forceinline__ uint32_t bf2_mul(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t bf2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ float fh_fma(unsigned short a, unsigned short b, float c) {
  float r;
  asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(r) : "h"(a), "h"(b), "f"(c));
  return r;
}

__device__ __forceinline__ uint4 blend_row(const uint4 (&buf)[4], uint32_t w01, uint32_t w23) {
  const uint32_t* q0 = reinterpret_cast<const uint32_t*>(&buf[0]);
  const uint32_t* q1 = reinterpret_cast<const uint32_t*>(&buf[1]);
  const uint32_t* q2 = reinterpret_cast<const uint32_t*>(&buf[2]);
  const uint32_t* q3 = reinterpret_cast<const uint32_t*>(&buf[3]);
  uint32_t o[4];
#if M3D_DCN_BLEND == 2
  // packed bf16: 4 HFMA2.BF16 per channel pair; every partial sum is rounded to bf16 (<= 3 extra roundings of 2^-9)
  const uint32_t wa = __byte_perm(w01, 0, 0x1010), wb = __byte_perm(w01, 0, 0x3232);
  const uint32_t wc = __byte_perm(w23, 0, 0x1010), wd = __byte_perm(w23, 0, 0x3232);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    uint32_t acc = bf2_mul(wa, q0[e]);
    acc = bf2_fma(wb, q1[e], acc);
    acc = bf2_fma(wc, q2[e], acc);
    o[e] = bf2_fma(wd, q3[e], acc);
  }
#elif M3D_DCN_BLEND == 1
  // bf16 operands (weights rounded to bf16), exact products, fp32 accumulation: FHFMA.BF16 with .H0/.H1 selectors
  unsigned short a0, a1, c0, c1;
  asm("mov.b32 {%0, %1}, %2;" : "=h"(a0), "=h"(a1) : "r"(w01));
  asm("mov.b32 {%0, %1}, %2;" : "=h"(c0), "=h"(c1) : "r"(w23));
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    unsigned short l0, h0, l1, h1, l2, h2, l3, h3;
    asm("mov.b32 {%0, %1}, %2;" : "=h"(l0), "=h"(h0) : "r"(q0[e]));
    asm("mov.b32 {%0, %1}, %2;" : "=h"(l1), "=h"(h1) : "r"(q1[e]));
    asm("mov.b32 {%0, %1}, %2;" : "=h"(l2), "=h"(h2) : "r"(q2[e]));
    asm("mov.b32 {%0, %1}, %2;" : "=h"(l3), "=h"(h3) : "r"(q3[e]));
    float lo = fh_fma(l0, a0, 0.f), hi = fh_fma(h0, a0, 0.f);
    lo = fh_fma(l1, a1, lo), hi = fh_fma(h1, a1, hi);
    lo = fh_fma(l2, c0, lo), hi = fh_fma(h2, c0, hi);
    lo = fh_fma(l3, c1, lo), hi = fh_fma(h3, c1, hi);
    o[e] = pack_bf16x2(lo, hi);
  }
#else
  // fp32 blend, two channels per FFMA2 (round-1 arithmetic on the bf16-rounded weights)
  const float fa = __uint_as_float(w01 << 16), fb = __uint_as_float(w01 & 0xffff0000u);
  const float fc = __uint_as_float(w23 << 16), fd = __uint_as_float(w23 & 0xffff0000u);
  const unsigned long long wx = pack_f32x2(fa, fa), wy = pack_f32x2(fb, fb);
  const unsigned long long wz = pack_f32x2(fc, fc), ww = pack_f32x2(fd, fd);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    unsigned long long acc =
        mul_f32x2(wy, pack_f32x2(__uint_as_float(q1[e] << 16), __uint_as_float(q1[e] & 0xffff0000u)));
    acc = fma_f32x2(wx, pack_f32x2(__uint_as_float(q0[e] << 16), __uint_as_float(q0[e] & 0xffff0000u)), acc);
    acc = fma_f32x2(wz, pack_f32x2(__uint_as_float(q2[e] << 16), __uint_as_float(q2[e] & 0xffff0000u)), acc);
    acc = fma_f32x2(ww, pack_f32x2(__uint_as_float(q3[e] << 16), __uint_as_float(q3[e] & 0xffff0000u)), acc);
    o[e] = f32x2_to_bf16x2(acc);
  }
#endif
  return make_uint4(o[0], o[1], o[2], o[3]);
}

// HALF: 64-pixel tiles (A rows 64-127 are never written; their accumulator rows are never read).  The layer is
// bounded by the producers, whose work scales with the rows, so half tiles cost little extra per pixel and the
// device gets filled when there are fewer full tiles than SMs (ida_0.proj_1: 30 tiles -> 72).
template <int BN, int NSTG, bool HALF, bool HALO>
__global__ void __launch_bounds__(kDcnThreads, 1) dcn_fused_kernel(const __grid_constant__ ConvGatherParams p) {
  static_assert(!(HALO && HALF), "the staged window is implemented for full tiles");
#ifdef M3D_PROBE
  __shared__ long long s_ddbg[32 * 8];
  if (threadIdx.x < 32 * 8) s_ddbg[threadIdx.x] = 0;
  int pkb = 0;  // k-blocks seen by this warp role
#endif
  using Cfg = DcnCfg<BN, NSTG, HALO>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int EW = Cfg::EW;
  uint32_t* table = reinterpret_cast<uint32_t*>(smem + STAGES * Cfg::STAGE);  // [2][kTileM * 9][EW]
  uint8_t* halo = smem + Cfg::HALO_OFF;                                 // [2][kHaloBytes] (HALO)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BARS_OFF);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* tbl_full = tempty + 2;
  uint64_t* tbl_empty = tbl_full + 2;
  uint64_t* halo_full = tbl_empty + 2;
  uint64_t* halo_empty = halo_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(halo_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], kProd / 32 + 1);  // one arrival per producer warp + the weight TMA
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], kProd / 32);         // every producer warp has drained its part of the accumulator
      mbar_init(&tbl_full[s], kBuildThreads);    // every table-building thread
      mbar_init(&tbl_empty[s], kProd / 32);      // every producer warp
      mbar_init(&halo_full[s], 1);               // the window TMA
      mbar_init(&halo_empty[s], kProd / 32);     // every producer warp
    }
    fence_barrier_init();
    prefetch_tmap(&p.tmap_b);
    if (HALO) prefetch_tmap(&p.tmap_img);
  }
  if (warp == 16) tmem_alloc<2 * Cfg::ACC>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_sync();

  const int nchunk = p.chunks[0];
  const int total_kb = kTaps * nchunk;

  if (warp < 16) {
    // ------------------------------------------------------------ A producers (+ the deferred epilogue)
    const int pt = threadIdx.x;
    const int j = pt & 7;       // 16-byte chunk (8 channels) of the 64-channel k-block
    const int rbase = pt >> 3;  // rows rbase and rbase + 64
    const int cs = p.in_cstride[0];
    const char* in0 = reinterpret_cast<const char*>(static_cast<const __nv_bfloat16*>(p.in[0]) + p.in_coff[0] + j * 8);
    const uint32_t d1 = static_cast<uint32_t>(cs) * 2u, d2 = static_cast<uint32_t>(p.W) * d1, d3 = d1 + d2;
    // epilogue share of this warp: TMEM lane quarter warp % 4 (rows), column group warp / 4
    constexpr int ECOLS = BN / 4;
    const int quarter = warp & 3, cgrp = warp >> 2;
    const __nv_bfloat16* res = p.res ? static_cast<const __nv_bfloat16*>(p.res) + p.res_coff : nullptr;
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out) + p.out_coff;
    auto epilogue = [&](int etile, int elocal) {  // accumulator of a finished tile -> bias (+ residual) -> LeakyReLU -> bf16
      const Tile t = tile_of(etile, p);
      const int as = elocal & 1;
#ifdef M3D_PROBE
      if (warp == 0 && elocal == 0) DDBG(31, 2);
#endif
      mbar_wait(&tfull[as], (elocal >> 1) & 1);
      tc_fence_after();
#ifdef M3D_PROBE
      if (warp == 0 && elocal == 0) DDBG(31, 3);
#endif
      if (!HALF || quarter < 2)
        epilogue_tile_direct<ECOLS, __nv_bfloat16, __nv_bfloat16>(tmem_base + as * Cfg::ACC + cgrp * ECOLS, quarter, lane,
                                                                  t.n, t.p0, t.q0, p.TW, p.P, p.Q, t.nt * BN + cgrp * ECOLS,
                                                                  p.Cout, p.bias, res, p.res_cstride, out, p.out_cstride,
                                                                  p.slope);
#ifdef M3D_PROBE
      if (warp == 0 && elocal == 0) DDBG(31, 4);
#endif
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
    };
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
      const uint32_t* tb = table + (local & 1) * (kTileM * kTaps * EW);
#ifdef M3D_PROBE
      if (warp == 0 && pkb == 0) DDBG(31, 0);
#endif
      mbar_wait(&tbl_full[local & 1], (local >> 1) & 1);  // built by the table warps during the previous tile
#ifdef M3D_PROBE
      if (warp == 0 && pkb == 0) DDBG(31, 1);
#endif
      // unit = (k-block, row half); K is walked chunk-major: k-block -> (chunk, tap).  The corner loads of a unit are
      // issued two units (one k-block) before its blend, into a ring of three register buffers; the unit's four
      // weights ride along.
      uint4 cv[3][4];
      uint32_t cw[3][2];
      int nx_tap = 0, nx_c = 0;  // k-block of the next unit to issue
      int cur_tap = 0;           // tap of the k-block being blended (M3D_DCN_WREREAD: weights re-read from the table)
      auto issue = [&](int half, uint4 (&buf)[4], uint32_t (&wt)[2]) {
        const uint32_t* ep = &tb[((rbase + 64 * half) * kTaps + nx_tap) * EW];
        uint3 e;  // three 4-byte loads: the 8 lanes of a row read the same words (one wavefront each)
        e.x = ep[0], e.y = ep[1], e.z = ep[2];
        const char* base = in0 + nx_c * (kBK * 2);
        if (HALF || half) {
          if (++nx_tap == kTaps) nx_tap = 0, ++nx_c;
        }
#ifndef M3D_DCN_WREREAD
        wt[0] = e.y, wt[1] = e.z;
#endif
        buf[0] = __ldg(reinterpret_cast<const uint4*>(base + e.x));
        buf[1] = __ldg(reinterpret_cast<const uint4*>(base + (e.x + d1)));
        buf[2] = __ldg(reinterpret_cast<const uint4*>(base + (e.x + d2)));
        buf[3] = __ldg(reinterpret_cast<const uint4*>(base + (e.x + d3)));
      };
      auto finish_kblock = [&]() {
#ifndef M3D_DCN_CONSUMER_FENCE
        fence_proxy_async_smem();
#endif
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
        if (++cur_tap == kTaps) cur_tap = 0;
      };
      auto weights = [&](int half, const uint32_t (&wt)[2], uint32_t& w0, uint32_t& w1) {
#ifdef M3D_DCN_WREREAD
        const uint32_t* ep = &tb[((rbase + 64 * half) * kTaps + cur_tap) * EW];
        w0 = ep[1], w1 = ep[2];
#else
        w0 = wt[0], w1 = wt[1];
#endif
      };
      if constexpr (HALO) {
        // corner loads from the staged window: two register buffers, the next unit's loads are issued before the
        // current unit is blended; nothing is carried across a chunk (= window) boundary
        const uint32_t wp = static_cast<uint32_t>(p.halo_w) * 128u;  // window row pitch
        auto issue_s = [&](int half, int tap, int c, uint32_t hbase, uint4 (&buf)[4], uint32_t (&wt)[2]) {
          const uint4 e = *reinterpret_cast<const uint4*>(&tb[((rbase + 64 * half) * kTaps + tap) * EW]);
          wt[0] = e.y, wt[1] = e.z;
          if (e.w == 0) {
            const uint32_t a = hbase + e.x;
            buf[0] = lds128(a), buf[1] = lds128(a + 128u), buf[2] = lds128(a + wp), buf[3] = lds128(a + wp + 128u);
          } else {  // the 2x2 patch leaves the window: global loads
            const char* base = in0 + c * (kBK * 2);
            buf[0] = __ldg(reinterpret_cast<const uint4*>(base + e.x));
            buf[1] = __ldg(reinterpret_cast<const uint4*>(base + (e.x + d1)));
            buf[2] = __ldg(reinterpret_cast<const uint4*>(base + (e.x + d2)));
            buf[3] = __ldg(reinterpret_cast<const uint4*>(base + (e.x + d3)));
          }
        };
#pragma unroll 1
        for (int c = 0; c < nchunk; ++c) {
          const int u = local * nchunk + c;  // running (tile, chunk) index of this CTA: window buffer u & 1
          const uint32_t hbase = smem_u32(halo) + (u & 1) * kHaloBytes + j * 16;
          mbar_wait(&halo_full[u & 1], (u >> 1) & 1);
          issue_s(0, 0, c, hbase, cv[0], cw[0]);
#pragma unroll 1
          for (int tap = 0; tap < kTaps; ++tap) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* a_tile = smem + stage * Cfg::STAGE;
            issue_s(1, tap, c, hbase, cv[1], cw[1]);
            *reinterpret_cast<uint4*>(a_tile + swizzled_offset<128>(rbase, j)) = blend_row(cv[0], cw[0][0], cw[0][1]);
            if (tap + 1 < kTaps) issue_s(0, tap + 1, c, hbase, cv[0], cw[0]);
            *reinterpret_cast<uint4*>(a_tile + swizzled_offset<128>(rbase + 64, j)) = blend_row(cv[1], cw[1][0], cw[1][1]);
            if (tap + 1 == kTaps) {  // last read of this window: release it before the k-block is published
              __syncwarp();
              if (lane == 0) mbar_arrive(&halo_empty[u & 1]);
            }
            finish_kblock();
            if (c == 0 && tap == 0 && local > 0) epilogue(tile - gridDim.x, local - 1);
          }
        }
      } else if constexpr (!HALF) {
        // one k-block: its halves sit in buffers A / B; the next k-block's halves are issued into NA / NB
        auto kblock = [&](const uint4 (&A)[4], const uint32_t (&WA)[2], const uint4 (&B)[4], const uint32_t (&WB)[2],
                          uint4 (&NA)[4], uint32_t (&NWA)[2], uint4 (&NB)[4], uint32_t (&NWB)[2], bool more) {
#ifdef M3D_PROBE
          if (warp == 0) DDBG(pkb, 0);
#endif
          mbar_wait(&empty[stage], phase ^ 1);
#ifdef M3D_PROBE
          if (warp == 0) DDBG(pkb, 1);
#endif
          uint8_t* a_tile = smem + stage * Cfg::STAGE;
          uint32_t wa0, wa1, wb0, wb1;
          weights(0, WA, wa0, wa1);  // (NB / NWB alias A / WA when the ring wraps: read before issue)
          if (more) issue(0, NA, NWA);
          *reinterpret_cast<uint4*>(a_tile + swizzled_offset<128>(rbase, j)) = blend_row(A, wa0, wa1);
#ifdef M3D_PROBE
          if (warp == 0) DDBG(pkb, 2);
#endif
          weights(1, WB, wb0, wb1);
          if (more) issue(1, NB, NWB);  // NB is A's storage when the ring wraps: A has just been consumed
          *reinterpret_cast<uint4*>(a_tile + swizzled_offset<128>(rbase + 64, j)) = blend_row(B, wb0, wb1);
          finish_kblock();
#ifdef M3D_PROBE
          if (warp == 0) DDBG(pkb, 3);
          ++pkb;
#endif
        };
        issue(0, cv[0], cw[0]);
        issue(1, cv[1], cw[1]);
#pragma unroll 1
        for (int kb = 0; kb < total_kb; kb += 3) {  // total_kb = 9 * nchunk: three k-blocks per trip keep the ring static
          kblock(cv[0], cw[0], cv[1], cw[1], cv[2], cw[2], cv[0], cw[0], true);
          // the previous tile's accumulator is complete by now (its last MMA was issued a k-block ago): drain it while
          // this tile's next loads are in flight
          if (kb == 0 && local > 0) epilogue(tile - gridDim.x, local - 1);
          kblock(cv[2], cw[2], cv[0], cw[0], cv[1], cw[1], cv[2], cw[2], true);
          kblock(cv[1], cw[1], cv[2], cw[2], cv[0], cw[0], cv[1], cw[1], kb + 3 < total_kb);
        }
      } else {
        // one row per thread and k-block; loads run two k-blocks ahead of the blend
        auto kblock1 = [&](const uint4 (&A)[4], const uint32_t (&WA)[2], uint4 (&NA)[4], uint32_t (&NWA)[2], bool more) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint32_t wa0, wa1;
          weights(0, WA, wa0, wa1);
          if (more) issue(0, NA, NWA);
          *reinterpret_cast<uint4*>(smem + stage * Cfg::STAGE + swizzled_offset<128>(rbase, j)) = blend_row(A, wa0, wa1);
          finish_kblock();
        };
        issue(0, cv[0], cw[0]);
        issue(0, cv[1], cw[1]);
#pragma unroll 1
        for (int kb = 0; kb < total_kb; kb += 3) {
          kblock1(cv[0], cw[0], cv[2], cw[2], true);
          if (kb == 0 && local > 0) epilogue(tile - gridDim.x, local - 1);
          kblock1(cv[1], cw[1], cv[0], cw[0], kb + 3 < total_kb);
          kblock1(cv[2], cw[2], cv[1], cw[1], kb + 4 < total_kb);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&tbl_empty[local & 1]);  // this warp has read its last entry of the tile's table
    }
    if (local > 0) epilogue(blockIdx.x + (local - 1) * gridDim.x, local - 1);  // the CTA's last tile
  } else if (warp == 16) {
    // ------------------------------- weight TMA (one k-block ahead) + MMA issue, one warp
    constexpr uint32_t idesc = umma_idesc_bf16(BN);
    int stage = 0, lstage = 0;      // stage of the MMA / of the next weight load
    uint32_t phase = 0, lphase = 0;
    int ltile = blockIdx.x, ltap = 0, lc = 0;  // k-block of the next weight load
    bool lmore = ltile < p.total_tiles;
    auto load_next = [&]() {
      if (!lmore) return;
      const int kblk = ltap * nchunk + lc;  // weights are packed tap-major
      const int nt = ltile % p.n_tiles;
      mbar_wait(&empty[lstage], lphase ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[lstage], Cfg::B_BYTES);
        tma_load_2d(smem + lstage * Cfg::STAGE + Cfg::A_BYTES, &p.tmap_b, &full[lstage], kblk * kBK, nt * BN);
      }
      __syncwarp();
      if (++lstage == STAGES) lstage = 0, lphase ^= 1;
      if (++ltap == kTaps) {
        ltap = 0;
        if (++lc == nchunk) {
          lc = 0;
          ltile += gridDim.x;
          lmore = ltile < p.total_tiles;
        }
      }
    };
    for (int s = 0; s < STAGES - 1; ++s) load_next();  // the weight loads run STAGES - 1 k-blocks ahead of the MMAs
    // HALO: window of the (tile, chunk) unit hu -> buffer hu & 1, requested one unit ahead of the MMAs
    int hu = 0, htile = blockIdx.x, hc = 0;
    auto load_window = [&]() {
      if (!HALO || htile >= p.total_tiles) return;
      const Tile t = tile_of(htile, p);
      mbar_wait(&halo_empty[hu & 1], ((hu >> 1) & 1) ^ 1);  // the producers have released the unit before last
      if (elect_one()) {
        mbar_arrive_expect_tx(&halo_full[hu & 1], static_cast<uint32_t>(p.halo_w * p.halo_h * 128));
        tma_load_4d(halo + (hu & 1) * kHaloBytes, &p.tmap_img, &halo_full[hu & 1], p.in_coff[0] + hc * kBK,
                    t.q0 - p.halo_x, t.p0 - p.halo_y, t.n);
      }
      __syncwarp();
      ++hu;
      if (++hc == nchunk) hc = 0, htile += gridDim.x;
    };
    load_window();
    int local = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + as * Cfg::ACC;
      int ktap = 0;
      for (int kb = 0; kb < total_kb; ++kb) {
        if (ktap == 0) load_window();  // start of a chunk: request the next unit's window (its buffer was released
        if (++ktap == kTaps) ktap = 0;  // when the producers finished the previous chunk, before this k-block's MMA)
        load_next();  // needs MMA(kb - 1) finished (its stage); the A producers wait for the same event
        mbar_wait(&full[stage], phase);
#ifdef M3D_PROBE
        DDBG(pkb, 4);
#endif
#ifdef M3D_DCN_CONSUMER_FENCE
        // EXPERIMENT: the generic-proxy writes of the A tile were released by the producers' mbarrier arrivals and are
        // acquired above; the proxy fence is executed HERE, by the thread that issues the async-proxy reads, instead of
        // by every producer (where it is a MEMBAR.ALL.CTA that also waits for the loads prefetched for the next k-block)
        fence_proxy_async_smem();
#endif
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE);
          const uint64_t da = umma_smem_desc<128>(sa);
          const uint64_t db = umma_smem_desc<128>(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) umma_f16(tmem_acc, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty[stage]);
        }
        __syncwarp();
#ifdef M3D_PROBE
        DDBG(pkb, 5);
        ++pkb;
#endif
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(&tfull[as]);
      __syncwarp();
    }
  } else {
    // ------------------------------------- warps 17-19: sample table of the tile AFTER the one being produced
    const int bt = threadIdx.x - (kProd + 32);  // 0 .. kBuildThreads - 1
    const int tw_shift = 31 - __clz(p.TW);
    const int cs = p.in_cstride[0];
    constexpr int ROWS = HALF ? 64 : kTileM;
    int fill = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++fill) {
      const int b = fill & 1;
      mbar_wait(&tbl_empty[b], ((fill >> 1) & 1) ^ 1);  // the producers are done with the table built two fills ago
      const Tile t = tile_of(tile, p);
      const int wy0 = t.p0 - p.halo_y, wx0 = t.q0 - p.halo_x;  // origin of the tile's staged window (HALO)
      uint32_t* tbw = table + b * (kTileM * kTaps * EW);
      // three entries per trip: their nine offset / mask loads are in flight together
#pragma unroll 1
      for (int e0 = bt; e0 < ROWS * kTaps; e0 += 3 * kBuildThreads) {
        float oh[3], ow[3], mk[3];
        int row[3], tap[3], pp[3], qq[3];
        bool ok[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int e = e0 + i * kBuildThreads;
          row[i] = e / kTaps, tap[i] = e - row[i] * kTaps;
          pp[i] = t.p0 + (row[i] >> tw_shift), qq[i] = t.q0 + (row[i] & (p.TW - 1));
          ok[i] = e < ROWS * kTaps && pp[i] < p.P && qq[i] < p.Q;
          oh[i] = ow[i] = 0.f, mk[i] = 0.f;
          if (ok[i]) {
            const float* om_px = p.om + ((static_cast<long>(t.n) * p.P + pp[i]) * p.Q + qq[i]) * p.om_cstride;
            oh[i] = __ldg(om_px + 2 * tap[i]);
            ow[i] = __ldg(om_px + 2 * tap[i] + 1);
            mk[i] = __ldg(om_px + 2 * kTaps + tap[i]);
          }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int e = e0 + i * kBuildThreads;
          if (e < ROWS * kTaps) {
            Entry en;
            en.off = 0, en.w01 = 0, en.w23 = 0, en.glob = 0;
            if (ok[i]) en = make_entry<HALO>(p, t.n, pp[i], qq[i], tap[i], cs, oh[i], ow[i], mk[i], wy0, wx0);
            tbw[e * EW] = en.off, tbw[e * EW + 1] = en.w01, tbw[e * EW + 2] = en.w23;
            if constexpr (EW == 4) tbw[e * EW + 3] = en.glob;
          }
        }
      }
      mbar_arrive(&tbl_full[b]);  // (release: the entries are visible to the producers that observe the phase)
    }
  }
  tc_fence_before();
  __syncthreads();
#ifdef M3D_PROBE
  if (blockIdx.x == 0 && threadIdx.x < 32 * 8) g_dcn_dbg[threadIdx.x] = s_ddbg[threadIdx.x];
#endif
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc<2 * Cfg::ACC>(tmem_base);
  }
}

template <int BN, int NSTG, bool HALF, bool HALO>
int launch_t(const ConvGatherParams& p, cudaStream_t stream) {
  using Cfg = DcnCfg<BN, NSTG, HALO>;
  auto kern = dcn_fused_kernel<BN, NSTG, HALF, HALO>;
  set_last_kernel("dcn_fused_kernel<%d,%d,%d,%d>", BN, NSTG, int(HALF), int(HALO));
  M3D_ONCE_PER_DEVICE_BEGIN
    M3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    // what the operand ring does not need goes to L1: the 9 taps of a chunk re-read one input window
    const int pct = (Cfg::SMEM + 2048) * 100 / (228 * 1024) + 1;
    M3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct));
  M3D_ONCE_PER_DEVICE_END
  const int sms = persistent_sms();
  int grid = sms;
  if (grid > p.total_tiles) grid = p.total_tiles;
  M3D_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(kDcnThreads), Cfg::SMEM, stream, p));
  return M3D_OK;
}

}  // namespace

// bf16 in / bf16 out 3x3 deformable layers with one input and BN in {128, 256}; anything else (incl. the 1x1
// centre-align layers, which are bounded by their per-tile table build, not the blend) stays on
// conv_gather_kernel (igemm.cu).
bool dcn_fused_supported(const ConvGatherParams& p, int BN, int in_dtype, int out_dtype) {
  if (getenv("M3D_DCN_LEGACY") != nullptr) return false;
  return p.om != nullptr && p.stem_img == nullptr && p.num_inputs == 1 && in_dtype == DT_BF16 &&
         out_dtype == DT_BF16 && (BN == 128 || BN == 256) && p.R == 3 && p.S == 3 && (p.TW & (p.TW - 1)) == 0 &&
         p.H >= 2 && p.W >= 2 && p.om_cstride >= 27 &&
         static_cast<long>(p.N) * p.H * p.W * p.in_cstride[0] * 2 < (1L << 32);  // 32-bit byte offsets in the table
}

int launch_dcn_fused(const ConvGatherParams& p0, int BN, cudaStream_t stream) {
  ConvGatherParams p = p0;
  // 64-pixel tiles when they shorten the schedule: waves x tile cost (measured: a half tile costs ~0.7 of a full
  // one -- the per-k-block barrier / fence work does not shrink), i.e. only when the full tiles cannot fill the device
  const int sms = persistent_sms();
  bool half = false;
  {
    int best_tw = 8, best_th = 8;
    long best = -1;
    const int cands[4][2] = {{8, 8}, {16, 4}, {4, 16}, {32, 2}};
    for (int i = 0; i < 4; ++i) {
      const long tiles = static_cast<long>((p.Q + cands[i][0] - 1) / cands[i][0]) * ((p.P + cands[i][1] - 1) / cands[i][1]);
      if (best < 0 || tiles < best) best = tiles, best_tw = cands[i][0], best_th = cands[i][1];
    }
    const long t_full = p.total_tiles, t_half = best * p.N * p.n_tiles;
    const double cost_full = static_cast<double>((t_full + sms - 1) / sms);
    const double cost_half = 0.72 * static_cast<double>((t_half + sms - 1) / sms);
    const char* e = getenv("M3D_DCN_HALF");  // development override: 0 = never, 1 = always
    half = e != nullptr ? atoi(e) != 0 : cost_half < cost_full;
    if (half && t_half < (1L << 30)) {
      p.TW = best_tw, p.TH = best_th;
      p.tiles_w = (p.Q + best_tw - 1) / best_tw, p.tiles_h = (p.P + best_th - 1) / best_th;
      p.total_tiles = static_cast<int>(t_half);
    } else {
      half = false;
    }
  }
  if (half) return BN == 128 ? launch_t<128, 2, true, false>(p, stream) : launch_t<256, 2, true, false>(p, stream);
  // staged input window (N = 128 layers: two 60 KB window buffers fit beside the operand ring): the largest halo, up
  // to 5 pixels, whose window fits a buffer; M3D_DCN_HALO=0 disables, M3D_DCN_HALO=h forces h
  if (BN == 128) {
    const char* e = getenv("M3D_DCN_HALO");
    int h = e ? atoi(e) : 5;
    while (h > 0 && (p.TW + 2 * h) * (p.TH + 2 * h) * 128 > kHaloBytes) --h;
    if (h >= 2 && p.TW + 2 * h <= 256 && p.TH + 2 * h <= 256) {
      p.halo_x = p.halo_y = h;
      p.halo_w = p.TW + 2 * h, p.halo_h = p.TH + 2 * h;
      const int rc = make_tmap_nhwc_plain(&p.tmap_img, p.in[0], p.N, p.H, p.W, p.in_cstride[0], kBK, p.halo_w, p.halo_h);
      if (rc != M3D_OK) return rc;
      return launch_t<128, 2, false, true>(p, stream);
    }
  }
  return BN == 128 ? launch_t<128, 2, false, false>(p, stream) : launch_t<256, 2, false, false>(p, stream);
}

}  // namespace m3d

#ifdef M3D_PROBE
extern "C" int m3d_dcn_debug_read(long long* host, int n) {
  return cudaMemcpyFromSymbol(host, m3d::g_dcn_dbg, sizeof(long long) * n) == cudaSuccess ? 0 : -1;
}
#endif
