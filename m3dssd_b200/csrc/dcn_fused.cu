// Fused bf16 DCNv2 layer (throughput mode): bilinear offset/mask sampling + output GEMM in one kernel,
// following modulated_deformable_im2col_gpu_kernel + SGEMM of the reference
// (model/DCNv2/src/cuda/dcn_v2_im2col_cuda.cu:18-47,118-180; model/DCNv2/src/dcn_v2_cuda.c:61-97) with no
// `columns` buffer.  Same GEMM view as igemm.cu (M = 128 output pixels, N = BN, k-blocks of 64 channels) but
// sized for what bounds this layer: the fp32 bilinear blend costs ~8.5 CUDA-core instructions per A element
// (unpack + FMA per corner), i.e. >= 544 issue clocks per k-block against 2*BN/... tensor clocks, so the
// kernel spends its warps on the blend:
//
//   warps 0-15 : A producers.  Per tile a shared table holds, for every (row, tap), the four clamped corner
//                offsets and the bilinear weights already multiplied by validity and modulation mask; per
//                k-block each thread blends 2 rows x 8 channels (4 x 16-byte corner loads per row, issued one
//                row ahead of the blend) and writes the swizzled bf16 A tile.
//   warp 16    : weight tiles by TMA (K walked chunk-major: the 9 taps of a 64-channel chunk re-read one
//                input window, which keeps the corner loads in L1/L2).
//   warp 17    : tcgen05.mma issue, accumulators in TMEM (two stages).
//   warps 18-21: epilogue: TMEM -> bias (+ residual) -> LeakyReLU -> bf16 NHWC.
#include <cstdlib>

#include "common.cuh"
#include "epilogue.cuh"
#include "igemm.cuh"
#include "ptx.cuh"

namespace m3d {

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
constexpr int kDcnThreads = kProd + 6 * 32;
constexpr int kBK = 64;

template <int BN, int NSTG>
struct DcnCfg {
  static constexpr int A_BYTES = kTileM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES = NSTG;
  static constexpr int TABLE = kTileM * 9 * 32;
  static constexpr int SMEM = STAGES * STAGE + TABLE + 1024 + 256;
  static constexpr int ACC = BN <= 128 ? 128 : 256;
  static_assert(SMEM <= 227 * 1024, "DCN tile does not fit shared memory");
};

struct Entry {
  int4 off;
  float4 w;
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

// One sample-table entry: clamped corner byte offsets + bilinear weights x validity x modulation mask
// (dcn_v2_im2col_cuda.cu:18-47,151-175).  (pp, qq): output pixel, assumed inside the image.
__device__ __forceinline__ Entry make_entry(const ConvGatherParams& p, int n, int pp, int qq, int tap, int taps, int cs,
                                            float o_h, float o_w, float m) {
  Entry e;
  e.off = make_int4(0, 0, 0, 0);
  e.w = make_float4(0.f, 0.f, 0.f, 0.f);
  const int r = taps == 9 ? tap / 3 : tap / p.S, sx = tap - r * p.S;
  const float hf = static_cast<float>(pp * p.stride - p.pad + r * p.dil) + o_h;
  const float wf = static_cast<float>(qq * p.stride - p.pad + sx * p.dil) + o_w;
  if (p.sigmoid_mask) m = 1.f / (1.f + __expf(-m));
  if (hf > -1.f && wf > -1.f && hf < static_cast<float>(p.H) && wf < static_cast<float>(p.W)) {
    const float hl = floorf(hf), wl = floorf(wf);
    const int h_low = static_cast<int>(hl), w_low = static_cast<int>(wl);
    const int h_high = h_low + 1, w_high = w_low + 1;
    const float lh = hf - hl, lw = wf - wl, hh = 1.f - lh, hw = 1.f - lw;
    const bool hl_ok = h_low >= 0, wl_ok = w_low >= 0, hh_ok = h_high <= p.H - 1, wh_ok = w_high <= p.W - 1;
    const int rl = (n * p.H + (hl_ok ? h_low : 0)) * p.W, rh = (n * p.H + (hh_ok ? h_high : 0)) * p.W;
    const int cl = wl_ok ? w_low : 0, ch = wh_ok ? w_high : 0;
    // byte offsets of the four corner pixels (unsigned 32-bit: one IMAD.WIDE.U32 per load in the main loop)
    e.off = make_int4((rl + cl) * cs * 2, (rl + ch) * cs * 2, (rh + cl) * cs * 2, (rh + ch) * cs * 2);
    e.w = make_float4((hl_ok && wl_ok) ? hh * hw * m : 0.f, (hl_ok && wh_ok) ? hh * lw * m : 0.f,
                      (hh_ok && wl_ok) ? lh * hw * m : 0.f, (hh_ok && wh_ok) ? lh * lw * m : 0.f);
  }
  return e;
}

// HALF: 64-pixel tiles (A rows 64-127 are never written; their accumulator rows are never read).  The layer is
// bounded by the producers' blend, which scales with the rows, so half tiles cost little extra per pixel and the
// device gets filled when there are fewer full tiles than SMs (ida_0.proj_1: 30 tiles -> 72).
template <int BN, int NSTG, bool HALF>
__global__ void __launch_bounds__(kDcnThreads, 1) dcn_fused_kernel(const __grid_constant__ ConvGatherParams p) {
#ifdef M3D_PROBE
  __shared__ long long s_ddbg[32 * 8];
  if (threadIdx.x < 32 * 8) s_ddbg[threadIdx.x] = 0;
  int pkb = 0;  // k-blocks seen by this warp role
#endif
  using Cfg = DcnCfg<BN, NSTG>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Entry* table = reinterpret_cast<Entry*>(smem + STAGES * Cfg::STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE + Cfg::TABLE);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], kProd / 32 + 1);  // one arrival per producer warp + the weight TMA
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 128);
    }
    fence_barrier_init();
    prefetch_tmap(&p.tmap_b);
  }
  if (warp == 17) tmem_alloc<2 * Cfg::ACC>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_sync();

  const int taps = p.R * p.S;
  const int nchunk = p.chunks[0];
  const int total_kb = taps * nchunk;

  if (warp < 16) {
    // ------------------------------------------------------------ A producers
    const int pt = threadIdx.x;
    const int j = pt & 7;       // 16-byte chunk (8 channels) of the 64-channel k-block
    const int rbase = pt >> 3;  // rows rbase and rbase + 64
    const int tw_shift = 31 - __clz(p.TW);
    const __nv_bfloat16* in0 = static_cast<const __nv_bfloat16*>(p.in[0]) + p.in_coff[0];
    const int cs = p.in_cstride[0];
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const Tile t = tile_of(tile, p);
#ifdef M3D_PROBE
      if (warp == 0 && pkb == 0) DDBG(31, 0);
#endif
      named_bar_sync(1, kProd);  // previous tile's table readers are done
      if constexpr (!HALF) {
        // thread pt fills row pt/4, taps (pt%4) + 4k: its offset / mask loads are issued together
        const int trow = pt >> 2;
        const int pp = t.p0 + (trow >> tw_shift), qq = t.q0 + (trow & (p.TW - 1));
        const bool tok = pp < p.P && qq < p.Q;
        const float* om_px = p.om + ((static_cast<long>(t.n) * p.P + pp) * p.Q + qq) * p.om_cstride;
        float o_h[3], o_w[3], o_m[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int tap = (pt & 3) + 4 * k;
          o_h[k] = o_w[k] = 0.f, o_m[k] = 1.f;
          if (tok && tap < taps) {
            o_h[k] = __ldg(om_px + 2 * tap);
            o_w[k] = __ldg(om_px + 2 * tap + 1);
            o_m[k] = __ldg(om_px + 2 * taps + tap);
          }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int tap = (pt & 3) + 4 * k;
          if (tap >= taps) break;
          Entry e;
          e.off = make_int4(0, 0, 0, 0);
          e.w = make_float4(0.f, 0.f, 0.f, 0.f);
          if (tok) e = make_entry(p, t.n, pp, qq, tap, taps, cs, o_h[k], o_w[k], o_m[k]);
          table[trow * taps + tap] = e;
        }
      } else {
        // 64 rows x 9 taps: thread pt fills (row pt/8, tap pt%8); threads with pt%8 == 0 also fill tap 8
        const int trow = pt >> 3;
        const int pp = t.p0 + (trow >> tw_shift), qq = t.q0 + (trow & (p.TW - 1));
        const bool tok = pp < p.P && qq < p.Q;
        const float* om_px = p.om + ((static_cast<long>(t.n) * p.P + pp) * p.Q + qq) * p.om_cstride;
        float o_h[2], o_w[2], o_m[2];
        int tp[2] = {pt & 7, 8};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          o_h[k] = o_w[k] = 0.f, o_m[k] = 1.f;
          if (tok && tp[k] < taps && (k == 0 || (pt & 7) == 0)) {
            o_h[k] = __ldg(om_px + 2 * tp[k]);
            o_w[k] = __ldg(om_px + 2 * tp[k] + 1);
            o_m[k] = __ldg(om_px + 2 * taps + tp[k]);
          }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (tp[k] >= taps || (k == 1 && (pt & 7) != 0)) continue;
          Entry e;
          e.off = make_int4(0, 0, 0, 0);
          e.w = make_float4(0.f, 0.f, 0.f, 0.f);
          if (tok) e = make_entry(p, t.n, pp, qq, tp[k], taps, cs, o_h[k], o_w[k], o_m[k]);
          table[trow * taps + tp[k]] = e;
        }
      }
      named_bar_sync(1, kProd);
#ifdef M3D_PROBE
      if (warp == 0 && pkb == 0) DDBG(31, 1);
#endif

      // unit = (k-block, row half); K is walked chunk-major: k-block -> (chunk, tap).  The corner loads of a unit
      // are issued two units (one k-block) before its blend, into a ring of three register buffers; the bilinear
      // weights are re-read from the table at blend time so the ring fits the 88-register budget of 704 threads.
      uint4 cv[3][4];
      int nx_tap = 0, nx_c = 0;  // k-block of the next unit to issue
      auto issue = [&](int half, uint4 (&buf)[4]) {
        const Entry& e = table[(rbase + 64 * half) * taps + nx_tap];
        const int4 o = e.off;
        const char* base = reinterpret_cast<const char*>(in0 + nx_c * kBK + j * 8);
        if (half) {
          if (++nx_tap == taps) nx_tap = 0, ++nx_c;
        }
        buf[0] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<uint32_t>(o.x)));
        buf[1] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<uint32_t>(o.y)));
        buf[2] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<uint32_t>(o.z)));
        buf[3] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<uint32_t>(o.w)));
      };
      auto blend = [&](int half, int tap, const uint4 (&buf)[4], uint8_t* a_tile) {
        const float4 w4 = table[(rbase + 64 * half) * taps + tap].w;
        const uint32_t* q0 = reinterpret_cast<const uint32_t*>(&buf[0]);
        const uint32_t* q1 = reinterpret_cast<const uint32_t*>(&buf[1]);
        const uint32_t* q2 = reinterpret_cast<const uint32_t*>(&buf[2]);
        const uint32_t* q3 = reinterpret_cast<const uint32_t*>(&buf[3]);
        // same arithmetic as the scalar fp32 blend of conv_gather_kernel, two channels per FFMA2
        const unsigned long long wx = pack_f32x2(w4.x, w4.x), wy = pack_f32x2(w4.y, w4.y);
        const unsigned long long wz = pack_f32x2(w4.z, w4.z), ww = pack_f32x2(w4.w, w4.w);
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          // (operation order of the scalar expression as nvcc contracts it: y*q1 first, then fma x, z, w)
          unsigned long long acc =
              mul_f32x2(wy, pack_f32x2(__uint_as_float(q1[e] << 16), __uint_as_float(q1[e] & 0xffff0000u)));
          acc = fma_f32x2(wx, pack_f32x2(__uint_as_float(q0[e] << 16), __uint_as_float(q0[e] & 0xffff0000u)), acc);
          acc = fma_f32x2(wz, pack_f32x2(__uint_as_float(q2[e] << 16), __uint_as_float(q2[e] & 0xffff0000u)), acc);
          acc = fma_f32x2(ww, pack_f32x2(__uint_as_float(q3[e] << 16), __uint_as_float(q3[e] & 0xffff0000u)), acc);
          w[e] = f32x2_to_bf16x2(acc);
        }
        *reinterpret_cast<uint4*>(a_tile + swizzled_offset<128>(rbase + 64 * half, j)) = make_uint4(w[0], w[1], w[2], w[3]);
      };
      int cur_tap = 0;  // tap of the k-block being blended
      auto finish_kblock = [&]() {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
        if (++cur_tap == taps) cur_tap = 0;
      };
      if constexpr (!HALF) {
        // one k-block: its halves sit in buffers A / B; the next k-block's halves are issued into NA / NB
        auto kblock = [&](const uint4 (&A)[4], const uint4 (&B)[4], uint4 (&NA)[4], uint4 (&NB)[4], bool more) {
#ifdef M3D_PROBE
          if (warp == 0) DDBG(pkb, 0);
#endif
          mbar_wait(&empty[stage], phase ^ 1);
#ifdef M3D_PROBE
          if (warp == 0) DDBG(pkb, 1);
#endif
          uint8_t* a_tile = smem + stage * Cfg::STAGE;
          if (more) issue(0, NA);
          blend(0, cur_tap, A, a_tile);
#ifdef M3D_PROBE
          if (warp == 0) DDBG(pkb, 2);
#endif
          if (more) issue(1, NB);  // NB is A's storage when the ring wraps: A has just been consumed
          blend(1, cur_tap, B, a_tile);
          finish_kblock();
#ifdef M3D_PROBE
          if (warp == 0) DDBG(pkb, 3);
          ++pkb;
#endif
        };
        issue(0, cv[0]);
        issue(1, cv[1]);
#pragma unroll 1
        for (int kb = 0; kb < total_kb; kb += 3) {  // total_kb = 9 * nchunk: three k-blocks per trip keep the ring static
          kblock(cv[0], cv[1], cv[2], cv[0], true);
          kblock(cv[2], cv[0], cv[1], cv[2], true);
          kblock(cv[1], cv[2], cv[0], cv[1], kb + 3 < total_kb);
        }
      } else {
        // one row per thread and k-block; loads run two k-blocks ahead of the blend
        auto issue1 = [&](uint4 (&buf)[4]) {
          const Entry& e = table[rbase * taps + nx_tap];
          const int4 o = e.off;
          const char* base = reinterpret_cast<const char*>(in0 + nx_c * kBK + j * 8);
          if (++nx_tap == taps) nx_tap = 0, ++nx_c;
          buf[0] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<uint32_t>(o.x)));
          buf[1] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<uint32_t>(o.y)));
          buf[2] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<uint32_t>(o.z)));
          buf[3] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<uint32_t>(o.w)));
        };
        auto kblock1 = [&](const uint4 (&A)[4], uint4 (&NA)[4], bool more) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (more) issue1(NA);
          blend(0, cur_tap, A, smem + stage * Cfg::STAGE);
          finish_kblock();
        };
        issue1(cv[0]);
        issue1(cv[1]);
#pragma unroll 1
        for (int kb = 0; kb < total_kb; kb += 3) {
          kblock1(cv[0], cv[2], true);
          kblock1(cv[1], cv[0], kb + 3 < total_kb);
          kblock1(cv[2], cv[1], kb + 4 < total_kb);
        }
      }
    }
  } else if (warp == 16) {
    // ---------------------------------------------------- weight TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const Tile t = tile_of(tile, p);
      int tap = 0, c = 0;
      for (int kb = 0; kb < total_kb; ++kb) {
        const int kblk = tap * nchunk + c;  // weights are packed tap-major
        if (++tap == taps) tap = 0, ++c;
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[stage], Cfg::B_BYTES);
          tma_load_2d(smem + stage * Cfg::STAGE + Cfg::A_BYTES, &p.tmap_b, &full[stage], kblk * kBK, t.nt * BN);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 17) {
    // ------------------------------------------------------------ MMA issuer
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
#ifdef M3D_PROBE
        DDBG(pkb, 4);
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
    // --------------------------------------------------------------- epilogue
    const int quarter = warp & 3;
    int local = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
      const Tile t = tile_of(tile, p);
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const __nv_bfloat16* res = p.res ? static_cast<const __nv_bfloat16*>(p.res) + p.res_coff : nullptr;
      __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out) + p.out_coff;
      if (!HALF || quarter < 2)
        epilogue_tile_direct<BN, __nv_bfloat16, __nv_bfloat16>(tmem_base + as * Cfg::ACC, quarter, lane, t.n, t.p0, t.q0,
                                                               p.TW, p.P, p.Q, t.nt * BN, p.Cout, p.bias, res,
                                                               p.res_cstride, out, p.out_cstride, p.slope);
      tc_fence_before();
      mbar_arrive(&tempty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
#ifdef M3D_PROBE
  if (blockIdx.x == 0 && threadIdx.x < 32 * 8) g_dcn_dbg[threadIdx.x] = s_ddbg[threadIdx.x];
#endif
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc<2 * Cfg::ACC>(tmem_base);
  }
}

template <int BN, int NSTG, bool HALF>
int launch_t(const ConvGatherParams& p, cudaStream_t stream) {
  using Cfg = DcnCfg<BN, NSTG>;
  auto kern = dcn_fused_kernel<BN, NSTG, HALF>;
  set_last_kernel("dcn_fused_kernel<%d,%d,%d>", BN, NSTG, int(HALF));
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
         out_dtype == DT_BF16 && (BN == 128 || BN == 256) && p.R * p.S == 9 && (p.TW & (p.TW - 1)) == 0 &&
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
  const char* e = getenv("M3D_DCN_STAGES");
  const int nstg = e ? atoi(e) : 2;
  if (half) return BN == 128 ? launch_t<128, 2, true>(p, stream) : launch_t<256, 2, true>(p, stream);
  if (nstg == 3) return BN == 128 ? launch_t<128, 3, false>(p, stream) : launch_t<256, 3, false>(p, stream);
  return BN == 128 ? launch_t<128, 2, false>(p, stream) : launch_t<256, 2, false>(p, stream);
}

}  // namespace m3d

#ifdef M3D_PROBE
extern "C" int m3d_dcn_debug_read(long long* host, int n) {
  return cudaMemcpyFromSymbol(host, m3d::g_dcn_dbg, sizeof(long long) * n) == cudaSuccess ? 0 : -1;
}
#endif
