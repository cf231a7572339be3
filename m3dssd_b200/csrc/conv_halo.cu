// 3x3 stride-1 convolution with a vertically shared A window.
//
// The plain implicit-GEMM kernel (igemm.cu) fetches one 128-pixel A tile per tap: 9 x 16 KB per 64-channel chunk,
// and the SM's ~64 B/clk L2 port -- not the tensor pipe -- bounds every layer whose N is small (A bytes per
// tensor clock = 8192 / N).  Here a pipeline stage is (kernel column s, channel chunk): ONE TMA box of
// (TH + 2) x TW pixels serves the three taps r = 0, 1, 2 of that column, because with TW a multiple of 8 the
// tile for tap r is the same shared-memory image shifted by r * TW rows = a whole number of 1024-byte
// swizzle atoms: the UMMA descriptor just starts r * TW * 128 bytes later.  A traffic drops 3 * TH / (TH + 2)
// = 2.4x (TH = 8); the stage's three weight k-blocks arrive as one 4-D TMA box.  12 MMAs per stage also
// amortise the barrier hand-shakes.  Same warp roles / TMEM double buffering / epilogues as conv_tma_kernel.
//
// Resident-weight variants (BRES: Cin <= 128 and one N tile -- level0 / level2, the offset/mask convs) go one step
// further: the tile is TH x TW = 16 x 8, ONE TMA box of (TH+2) x (TW+2) = 18 x 10 pixels per 64-channel chunk holds
// every tap, and tap (r, s) of tile row g starts at window row (g + r) * 10 + s: an A descriptor with start
// (r * 10 + s) * 128 bytes and 1280 bytes between 8-row groups (umma_smem_desc_sbo: the swizzle is a function of the
// absolute address, so neither needs 1024-byte alignment).  A bytes per tile and chunk: 3 x 20 KB -> 22.5 KB.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "epilogue.cuh"
#include "igemm.cuh"
#include "ptx.cuh"

namespace m3d {

namespace {

__host__ __device__ constexpr int halo_acc_cols(int bn) { return bn <= 32 ? 32 : (bn <= 64 ? 64 : (bn <= 128 ? 128 : 256)); }

// BRES: the whole weight matrix (9 taps x BN x 64 channels, Cin = 64 layers) is loaded once per CTA and stays in
// shared memory; stages then carry only the 20 KB A window (level0 / level2: 50+ tiles per CTA re-used 72 KB of
// weights per tile through the L2 port).
template <int BN, bool STAGED, bool BRES>
struct HaloCfg {
  static constexpr int A_ROWS = 160;                 // (TH + 2) * TW with TH * TW = 128, TH = 8... see host
  static constexpr int A_BYTES = 20 * 1024;          // host guarantees (TH + 2) * TW * 128 <= A_BYTES
  static constexpr int B_BYTES = BN * 128;           // one tap
  // BRES (Cin = 64): a stage holds the windows of all three kernel columns -> one stage (36 MMAs) per tile
  static constexpr int WIN_BYTES = 23 * 1024;        // BRES: (16 + 2) x (8 + 2) pixels x 128 B = 23 040, rounded up
  static constexpr int STAGE = BRES ? WIN_BYTES : A_BYTES + 3 * B_BYTES;
  static constexpr int RES_BYTES = BRES ? 72 * 1024 : 0;  // 9 * nchunk taps of [BN][64]: BN = 64 x 1 chunk, BN = 32 x <= 2
  static constexpr int EXTRA = (STAGED ? (2 * kSlabBytes + 1024) : 0) + RES_BYTES;
  static constexpr int BUDGET = 225 * 1024 + 512 - EXTRA;
  static constexpr int FIT = BUDGET / STAGE;
  static constexpr int STAGES = FIT >= 6 ? 6 : FIT;
  static constexpr int SMEM = STAGES * STAGE + EXTRA + 1024 + 256;
  static constexpr int ACC = halo_acc_cols(BN);
  static_assert(STAGES >= 2, "halo conv tile does not fit shared memory");
  static_assert(B_BYTES % 1024 == 0, "weight tap tiles must keep the 1024-byte swizzle alignment");
};

struct HTile {
  int nt, n, p0, q0;
};
__device__ __forceinline__ HTile htile(const TileWalk& w, const ConvTmaParams& p) {
  HTile t;
  t.nt = w.nt, t.n = w.n, t.p0 = w.th * p.TH, t.q0 = w.tw * p.TW;
  return t;
}
#define HALO_WALK_INIT TileWalk tw_; tw_.init(p.total_tiles, p.n_tiles, p.tiles_w, p.tiles_h, p.N)
#define HALO_WALK_NEXT tw_.next(p.n_tiles, p.tiles_w, p.tiles_h, p.N)

template <int BN, typename OutT, bool STAGED, bool BRES>
__global__ void __launch_bounds__(STAGED ? 320 : 192, 1) conv_halo_kernel(const __grid_constant__ ConvTmaParams p) {
  using Cfg = HaloCfg<BN, STAGED, BRES>;
  constexpr int NW = STAGED ? 8 : 4;  // epilogue warps (see conv_tma_kernel)
  // one-slab tiles: the two 4-warp groups take alternate tiles whole (epilogue.cuh, "grouped", SPLIT = false) -- the
  // serial drain of a 64-column tile (~2.8 k clocks: barriers, residual TMA, store) was longer than its 36 MMAs
  constexpr bool ALT = STAGED && BN == 64;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* b_res = smem + STAGES * Cfg::STAGE;               // BRES: [s][r][BN][64] resident weights
  uint8_t* stage_out = b_res + Cfg::RES_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE + Cfg::EXTRA);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* res_bar = tempty + 2;
  uint64_t* bres_bar = res_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], ALT ? 4 : NW);  // one arrival per epilogue warp that drains this accumulator stage
      mbar_init(&res_bar[s], 1);
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
    prefetch_tmap(&p.tmap_a[0]);
    prefetch_tmap(&p.tmap_b);
    if (STAGED) prefetch_tmap(&p.tmap_out);
  }
  if (warp == 1) tmem_alloc<2 * Cfg::ACC>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_sync();

  const int nchunk = p.chunks[0];
  const int n_stages = BRES ? nchunk : 3 * nchunk;  // (s, chunk) pairs; resident weights: one stage per chunk
  const uint32_t a_bytes = static_cast<uint32_t>((p.TH + 2) * (BRES ? p.TW + 2 : p.TW) * 128);
  const uint32_t tap_shift = static_cast<uint32_t>((BRES ? p.TW + 2 : p.TW) * 128) >> 4;  // descriptor units (16 B) per kernel row

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    if (BRES && elect_one()) {  // n_tiles == 1: the whole weight matrix, once, as [s * nchunk + c][r][BN][64]
      mbar_arrive_expect_tx(bres_bar, 9 * nchunk * Cfg::B_BYTES);
      for (int i = 0; i < 3 * nchunk; ++i) tma_load_4d(b_res + i * 3 * Cfg::B_BYTES, &p.tmap_b, bres_bar, 0, 0, i, 0);
    }
    __syncwarp();
    HALO_WALK_INIT;
    for (int tile = tw_.first; tile < tw_.last; ++tile, HALO_WALK_NEXT) {
      const HTile t = htile(tw_, p);
      const int brow = t.nt * BN;
      int s = 0, c = 0;
      for (int st = 0; st < n_stages; ++st) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * Cfg::STAGE;
          if constexpr (BRES) {  // the whole (TH+2) x (TW+2) window of this chunk
            mbar_arrive_expect_tx(&full[stage], a_bytes);
            tma_load_4d(sa, &p.tmap_a[0], &full[stage], p.a_coff[0] + st * 64, t.q0 - 1, t.p0 - 1, t.n);
          } else {
            mbar_arrive_expect_tx(&full[stage], ((p.dbg & 1) ? 0 : a_bytes) + 3 * Cfg::B_BYTES);
            if (!(p.dbg & 1)) tma_load_4d(sa, &p.tmap_a[0], &full[stage], p.a_coff[0] + c * 64, t.q0 - 1 + s, t.p0 - 1, t.n);
            tma_load_4d(sa + Cfg::A_BYTES, &p.tmap_b, &full[stage], 0, brow, s * nchunk + c, 0);
          }
        }
        __syncwarp();
        if (++c == nchunk) c = 0, ++s;
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
    if (BRES) mbar_wait(bres_bar, 0);
    const unsigned long long kzero = BRES ? p.kzero : 0ull;
    const uint64_t db_res = umma_smem_desc<128>(smem_u32(b_res));                        // BRES: resident weights
    const uint32_t sc_stride = static_cast<uint32_t>(nchunk) * 3 * (Cfg::B_BYTES >> 4);  // one kernel column of them
    HALO_WALK_INIT;
    for (int tile = tw_.first; tile < tw_.last; ++tile, ++local) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + as * Cfg::ACC;
      for (int st = 0; st < n_stages; ++st) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        // Descriptor bases are computed HERE, in warp-uniform code, so they live in uniform registers and the elected
        // thread spends ~3 instructions per MMA (two 64-bit immediate adds + the MMA).  Computed inside the elected
        // branch they cost ~11 (IMAD / IADD3 / R2UR per operand): a lone thread then needs ~70 clocks per MMA while
        // an N = 64 MMA occupies the tensor pipe for 32 -- level0 / level2 ran at 2.5 k clocks per 36-MMA tile.
        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE);
        if constexpr (BRES) {
          constexpr uint32_t kTap = ((8 + 2) * 128) >> 4;  // descriptor units per kernel row of the 18 x 10 window
          constexpr uint32_t kB16 = Cfg::B_BYTES >> 4;
          const uint64_t da0 = umma_smem_desc_sbo(sa, kTap << 4);
          const uint64_t db_a = db_res + static_cast<uint32_t>(st) * (3 * kB16);  // kernel column 0 of this chunk
          const uint64_t db_b = db_a + sc_stride, db_c = db_b + sc_stride;        // columns 1, 2
          if (kzero != 0) {
            // structurally sparse weights (the space-to-depth rewrite of level0: 16 live k-steps of 36): a k-step
            // whose 16 weight columns are zero in every row is not issued.  Host: one chunk, never all 36.
            if (elect_one()) {
              uint32_t acc = 0;
#pragma unroll
              for (int sc = 0; sc < 3; ++sc) {
                const uint64_t db = sc == 0 ? db_a : (sc == 1 ? db_b : db_c);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    if (!((kzero >> ((r * 3 + sc) * 4 + k)) & 1ull)) {
                      umma_f16(tmem_acc, da0 + (sc * 8 + r * kTap + 2 * k), db + (r * kB16 + 2 * k), idesc, acc);
                      acc = 1;
                    }
                  }
                }
              }
              umma_commit(&empty[stage]);
            }
          } else if (elect_one()) {
#pragma unroll
            for (int sc = 0; sc < 3; ++sc) {
              // tile row g of tap (r, sc) = window rows (g + r) * (TW + 2) + sc ...: 8-row groups (TW + 2) * 128 B apart
              const uint64_t db = sc == 0 ? db_a : (sc == 1 ? db_b : db_c);
#pragma unroll
              for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16(tmem_acc, da0 + (sc * 8 + r * kTap + 2 * k), db + (r * kB16 + 2 * k), idesc, (st | sc | r | k) != 0);
              }
            }
            umma_commit(&empty[stage]);
          }
        } else {
          const uint64_t da_0 = umma_smem_desc<128>(sa);
          const uint64_t da_1 = da_0 + tap_shift, da_2 = da_1 + tap_shift;
          const uint64_t db = da_0 + (Cfg::A_BYTES >> 4);
          const bool skip = (p.dbg & 8) != 0;  // development probe: no MMAs
          if (elect_one()) {
            if (!skip) {
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                const uint64_t da = r == 0 ? da_0 : (r == 1 ? da_1 : da_2);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16(tmem_acc, da + 2 * k, db + (r * (Cfg::B_BYTES >> 4) + 2 * k), idesc, (st | r | k) != 0);
              }
            }
            umma_commit(&empty[stage]);
          }
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
  } else if constexpr (ALT) {
    const int quarter = warp & 3;
    const int ep_tid = threadIdx.x - 64;
    const int group = ep_tid >> 7, gtid = ep_tid & 127;
    float* bias_s = reinterpret_cast<float*>(stage_out + 2 * kSlabBytes);  // 256 floats (host: n_tiles * BN <= 256)
    for (int i = ep_tid; i < p.n_tiles * BN; i += 32 * NW) bias_s[i] = (p.bias != nullptr && i < p.Cout) ? __ldg(p.bias + i) : 0.f;
    named_bar_sync(kEpiBarrier + 2, 32 * NW);
    GroupedEpilogue st;
    st.init(stage_out + group * kSlabBytes, bias_s, &res_bar[group]);
    const void* tmap_res = p.res ? &p.tmap_res : nullptr;
    HALO_WALK_INIT;
    int tile = tw_.first;
    if (group == 1) {
      ++tile;
      HALO_WALK_NEXT;
    }
    auto epi_tile = [&]() {
      const HTile t = htile(tw_, p);
      return EpiTile{t.n, t.p0, t.q0, t.nt * BN};
    };
    EpiTile cur = epi_tile();
    if (tile < tw_.last) epilogue_grouped_begin<BN, false>(st, gtid, group, cur, tmap_res, p.res_coff);
    for (int local = group; tile < tw_.last; tile += 2, local += 2) {
      HALO_WALK_NEXT;
      HALO_WALK_NEXT;
      const EpiTile nxt = epi_tile();  // this group's next tile (coordinates only meaningful when it exists)
      const uint32_t aphase = (local >> 1) & 1;
      mbar_wait(&tfull[group], aphase);
      tc_fence_after();
      if (p.dbg & 4) {  // development probe: skip the drain
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[group]);
        cur = nxt;
        continue;
      }
      epilogue_tile_grouped<BN, false>(st, tmem_base + group * Cfg::ACC, quarter, lane, gtid, group, cur,
                                       tile + 2 < tw_.last ? &nxt : nullptr, &p.tmap_out, p.out_coff, tmap_res, p.res_coff,
                                       p.slope, [&]() {
                                         tc_fence_before();
                                         __syncwarp();
                                         if (lane == 0) mbar_arrive(&tempty[group]);
                                       });
      cur = nxt;
    }
    if (gtid == 0) tma_store_wait_all();
  }  else {
    const int quarter = warp & 3;
    const int ep_tid = threadIdx.x - 64;
    StagedEpilogue st;
    if constexpr (STAGED) st.init(stage_out, reinterpret_cast<float*>(stage_out + 2 * kSlabBytes), res_bar);
    int local = 0;
    HALO_WALK_INIT;
    for (int tile = tw_.first; tile < tw_.last; ++tile, ++local, HALO_WALK_NEXT) {
      const HTile t = htile(tw_, p);
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      if (p.dbg & 4) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);
        continue;
      }
      if constexpr (STAGED) {
        const int col0 = t.nt * BN;
        epilogue_tile_staged<BN, NW>(st, tmem_base + as * Cfg::ACC, quarter, lane, ep_tid, t.n, t.p0, t.q0, &p.tmap_out,
                                 p.out_coff + col0, p.res ? &p.tmap_res : nullptr, p.res_coff + col0,
                                 p.bias ? p.bias + col0 : nullptr, p.Cout - col0, p.slope, [&]() {
                                   tc_fence_before();
                                   __syncwarp();
                                   if (lane == 0) mbar_arrive(&tempty[as]);
                                 });
      } else {
        const __nv_bfloat16* res = p.res ? static_cast<const __nv_bfloat16*>(p.res) + p.res_coff : nullptr;
        OutT* out = static_cast<OutT*>(p.out) + p.out_coff;
        epilogue_tile_direct<BN, OutT, __nv_bfloat16>(tmem_base + as * Cfg::ACC, quarter, lane, t.n, t.p0, t.q0, p.TW,
                                                      p.P, p.Q, t.nt * BN, p.Cout, p.bias, res, p.res_cstride, out,
                                                      p.out_cstride, p.slope);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);
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

template <int BN, typename OutT, bool STAGED, bool BRES = false>
int launch_t(const ConvTmaParams& p, cudaStream_t stream) {
  using Cfg = HaloCfg<BN, STAGED, BRES>;
  auto kern = conv_halo_kernel<BN, OutT, STAGED, BRES>;
  set_last_kernel("conv_halo_kernel<%d,%s,%d,%d>", BN, sizeof(OutT) == 4 ? "f32" : "bf16", int(STAGED), int(BRES));
  M3D_ONCE_PER_DEVICE_BEGIN
    M3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  M3D_ONCE_PER_DEVICE_END
  const int sms = persistent_sms();
  const int per_sm = (2 * (Cfg::SMEM + 1024) <= 227 * 1024 && 4 * Cfg::ACC <= 512) ? 2 : 1;
  int grid = sms * per_sm;
  if (grid > p.total_tiles) grid = p.total_tiles;
  ConvTmaParams q = p;
  q.dbg = getenv("M3D_DBG") ? atoi(getenv("M3D_DBG")) : 0;
  M3D_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(STAGED ? 320 : 192), Cfg::SMEM, stream, q));
  return M3D_OK;
}

}  // namespace

// Packed bf16 weights [rows][K], K = (r, s, chunk, 64) as (64, rows, 3 * nchunk, 3): box {64, bn, 1, 3} brings the
// three kernel rows of one (s, chunk) as consecutive [bn][64] tiles.
int make_tmap_b_halo(CUtensorMap* map, const void* base, long rows, int nchunk, int bn);

// Will launch_conv_halo pick a resident-weight (unified-window, 16 x 8 pixel tiles) variant for this layer?
bool conv_halo_unified(int BN, int out_dtype, bool staged, int nchunk, int n_tiles) {
  static const bool bres = !(getenv("M3D_HALO_BRES") && atoi(getenv("M3D_HALO_BRES")) == 0);
  if (!bres || n_tiles != 1) return false;
  if (staged) return BN == 64 && nchunk == 1;
  return BN == 32 && out_dtype != DT_BF16 && nchunk <= 2;
}

bool conv_halo_supported(int BN, int out_dtype, bool staged) {
  if (staged) return out_dtype == DT_BF16 && (BN == 64 || BN == 128);
  return BN == 32 || BN == 64;
}

int launch_conv_halo(const ConvTmaParams& p, int BN, int out_dtype, bool staged, cudaStream_t stream) {
  // resident-weight variants: the host (api_conv.cu) asked conv_halo_unified() too and built 16 x 8 tiles + window map
  const bool bres = conv_halo_unified(BN, out_dtype, staged, p.chunks[0], p.n_tiles);
  if (bres && (p.TW != 8 || p.TH != 16)) return M3D_ERR_UNSUPPORTED;
  if (staged) {
    if (BN == 64 && p.n_tiles * 64 > 256) return M3D_ERR_UNSUPPORTED;  // bias area of the alternate-tile epilogue
    if (BN == 64 && bres) return launch_t<64, __nv_bfloat16, true, true>(p, stream);
    if (BN == 64) return launch_t<64, __nv_bfloat16, true>(p, stream);
    if (BN == 128) return launch_t<128, __nv_bfloat16, true>(p, stream);
    return M3D_ERR_UNSUPPORTED;
  }
  if (BN == 32 && bres) return launch_t<32, float, false, true>(p, stream);  // offset/mask convs of the Cin <= 128 DCNs
  if (BN == 32)
    return out_dtype == DT_BF16 ? launch_t<32, __nv_bfloat16, false>(p, stream) : launch_t<32, float, false>(p, stream);
  if (BN == 64)
    return out_dtype == DT_BF16 ? launch_t<64, __nv_bfloat16, false>(p, stream) : launch_t<64, float, false>(p, stream);
  return M3D_ERR_UNSUPPORTED;
}

}  // namespace m3d
