// Fused three-layer 1x1 regression heads (model/M3d_inference_align.py:66-210, 236-277):
//
//   out_g = W3_g . lrelu(W2_g . lrelu(W1_g . x + b1_g) + b2_g) + b3_g        g = 0..G-1 heads sharing the input x
//
// (BatchNorm folded into W1/b1 and W2/b2 on the host.)  The reference runs each head as conv1x1 + BN + LeakyReLU,
// conv1x1 + BN + LeakyReLU, conv1x1: nine kernels and four round trips of a [B,256,48,160] tensor through HBM per
// head.  Here one persistent kernel keeps the 256-channel intermediates on chip:
//
//   work item = (128-pixel tile, head).  x tile [128 x Cx] bf16 stays in shared memory for the G heads of a tile.
//   GEMM1  acc1[128 x 256] (TMEM cols 0-255)   = X . W1^T       K = Cx
//   E1     acc1 -> +b1, LeakyReLU, bf16 -> Y (4 swizzled 64-channel slabs in shared memory)
//   GEMM2  acc2[128 x 256] (TMEM cols 256-511) = Y . W2^T       K = 256: k-block j only needs slab j of Y, so it
//                                                               starts as soon as E1 has written that slab
//   E2     acc2 -> +b2, LeakyReLU, bf16 -> Y (free again: GEMM2 has completed)
//   GEMM3  acc3[128 x R3]  (TMEM cols 256-...) = Y . W3^T       k-block j after slab j of E2
//   E3     acc3 -> +b3 -> fp32 global (the reference's [B, H, W, 11*A] head buffer slot)
//
// Warp roles: warp 0 streams X tiles and the weights (W1, W2, W3 of every item, 32 KB ring slots) by TMA; warp 1
// issues tcgen05.mma (GEMM1 of the next item is issued between GEMM2 and GEMM3 of the current one, so it runs
// under E2); warps 2-9 are the epilogue (two warps per TMEM lane quarter, each takes half of the columns).
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstring>
#include <cstdlib>

#include "common.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace m3d {

int make_tmap_nhwc(CUtensorMap* map, const void* base, int N, int H, int W, int C, int bk, int tw, int th, int stride);
int make_tmap_nhwc_f32(CUtensorMap* map, const void* base, int N, int H, int W, int C, int box_c, int tw, int th,
                       bool swizzle128);
int make_tmap_b3d(CUtensorMap* map, const void* base, long rows, long cols, int bk, int box_rows, int ksub);
void pick_tile(int P, int Q, int max_tw, int* TW, int* TH);

// Phase timeline probe (tools/probe_heads.py): compile with -DM3D_PROBE
#ifdef M3D_PROBE
// (stamps go to shared memory and are copied out at the end: a global store per stamp makes the next
// fence.proxy.async -- a MEMBAR -- of the stamping warp wait an L2 round trip, which delays exactly the slab hand-offs
// the probe is looking at)
__device__ long long g_head_dbg[1024];
#define HDBG(slot) do { if (blockIdx.x == 0 && lane == 0 && li < 6) s_head_dbg[li * 32 + (slot)] = clock64(); } while (0)
#else
#define HDBG(slot) do { } while (0)
#endif
constexpr int kHeadMid = 256;
constexpr int kHeadWSlots = 4;
constexpr int kHeadSlotBytes = kHeadMid * 128;  // one k-block of a 256-row weight matrix
constexpr int kHeadEpiWarps = 16;
constexpr int kHeadThreads = 64 + 32 * kHeadEpiWarps;
constexpr int kHeadSmem = 2 * 16384 /*X*/ + 4 * 16384 /*Y*/ + kHeadWSlots * kHeadSlotBytes + 1024 + 256;

struct alignas(64) HeadMlpParams {
  CUtensorMap tmap_x;   // (C, W, H, N) box {64, TW, TH, 1}
  CUtensorMap tmap_w1;  // (64, G*256, K1) box {64, 256, 1}
  CUtensorMap tmap_w2;  // (64, G*256, 4)  box {64, 256, 1}
  CUtensorMap tmap_w3;  // (64, G*R3, 4)   box {64, R3, 4}
  CUtensorMap tmap_out; // fp32 (C, W, H, N) box {A, TW, TH, 1}: one head's A channels of a pixel tile
  const float *b1, *b2, *b3;
  float* out;
  int out_cstride, out_coff;
  int x_coff, K1;
  int G, A, R3;
  int N, P, Q, TW, TH, tiles_w, tiles_h;
  int total_items;
  float slope;
};

struct HeadTile {
  int n, p0, q0;
};
__device__ __forceinline__ HeadTile head_tile(int tile, const HeadMlpParams& p) {
  HeadTile t;
  const int tw = tile % p.tiles_w;
  int r = tile / p.tiles_w;
  const int th = r % p.tiles_h;
  t.n = r / p.tiles_h;
  t.p0 = th * p.TH;
  t.q0 = tw * p.TW;
  return t;
}

template <int R3>
__global__ void __launch_bounds__(kHeadThreads, 1) head_mlp_kernel(const __grid_constant__ HeadMlpParams p) {
  extern __shared__ uint8_t smem_raw[];
#ifdef M3D_PROBE
  __shared__ long long s_head_dbg[6 * 32];
  if (threadIdx.x < 6 * 32) s_head_dbg[threadIdx.x] = 0;
#endif
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sx = smem;                    // K1 k-blocks of X
  uint8_t* sy = smem + 2 * 16384;        // 4 slabs of Y
  uint8_t* sw = smem + 6 * 16384;        // weight ring
  // [128][A] fp32 output staging (the TMA store's box, up to 24 KB) over slabs 2-3 of Y, free between GEMM3 and the
  // moment the next E1 reaches slab 2: the store is not waited for in E3, only (via so_free) before that slab is
  // rewritten
  float* so = reinterpret_cast<float*>(sy + 2 * 16384);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * 16384 + kHeadWSlots * kHeadSlotBytes);
  uint64_t* w_full = bars;                      // [kHeadWSlots]
  uint64_t* w_empty = bars + kHeadWSlots;       // [kHeadWSlots]
  uint64_t* x_full = bars + 2 * kHeadWSlots;    // TMA
  uint64_t* x_empty = x_full + 1;               // MMA commit (GEMM1 of an item has read X)
  uint64_t* acc1_full = x_full + 2;
  uint64_t* acc1_empty = x_full + 3;            // 8 epilogue warps
  uint64_t* acc2_full = x_full + 4;
  uint64_t* acc2_empty = x_full + 5;            // after E3
  uint64_t* acc3_full = x_full + 6;
  uint64_t* y1_ready = x_full + 7;              // [4], 8 arrivals each
  uint64_t* y2_ready = x_full + 11;             // [4]
  uint64_t* so_free = x_full + 15;              // the output TMA store of an item has read its staging (TMA warp)
  uint64_t* so_ready = x_full + 16;             // E3 has staged an item's output (16 epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_full + 17);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kHeadWSlots; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    mbar_init(acc1_full, 1);
    mbar_init(acc1_empty, kHeadEpiWarps);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, kHeadEpiWarps);
    mbar_init(acc3_full, 1);
    mbar_init(so_free, 1);
    mbar_init(so_ready, kHeadEpiWarps);
    for (int s = 0; s < 4; ++s) {
      mbar_init(&y1_ready[s], kHeadEpiWarps);
      mbar_init(&y2_ready[s], kHeadEpiWarps);
    }
    fence_barrier_init();
    prefetch_tmap(&p.tmap_x);
    prefetch_tmap(&p.tmap_w1);
    prefetch_tmap(&p.tmap_w2);
    prefetch_tmap(&p.tmap_w3);
    prefetch_tmap(&p.tmap_out);
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_sync();

  // contiguous, balanced share of the (tile, head) items: consecutive items of a CTA mostly share the x tile
  const int start = static_cast<int>(static_cast<long>(blockIdx.x) * p.total_items / gridDim.x);
  const int end = static_cast<int>(static_cast<long>(blockIdx.x + 1) * p.total_items / gridDim.x);
  const int K1 = p.K1;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int slot = 0;
    uint32_t wphase = 0;
    int li = 0, nload = 0;
    auto load_w = [&](const CUtensorMap* tm, int row, int kb, uint32_t bytes) {
      mbar_wait_sleep(&w_empty[slot], wphase ^ 1);
      if (nload >= 0 && nload < 8) HDBG(24 + nload);
      ++nload;
      if (elect_one()) {
        mbar_arrive_expect_tx(&w_full[slot], bytes);
        tma_load_3d(sw + slot * kHeadSlotBytes, tm, &w_full[slot], 0, row, kb);
      }
      __syncwarp();
      if (++slot == kHeadWSlots) slot = 0, wphase ^= 1;
    };
    auto load_x = [&](int tile) {
      const HeadTile t = head_tile(tile, p);
      if (elect_one()) {
        mbar_arrive_expect_tx(x_full, static_cast<uint32_t>(K1) * 16384u);
        for (int kb = 0; kb < K1; ++kb) tma_load_4d(sx + kb * 16384, &p.tmap_x, x_full, p.x_coff + kb * 64, t.q0, t.p0, t.n);
      }
      __syncwarp();
    };
    // The output tile of an item goes out as ONE TMA store issued here (this warp is mostly idle and always the same
    // thread issues, bulk groups being per thread): wait for E3's staging, store, and once the store has read the
    // staging release it to the E1 of the next item.
    auto store_out = [&](int oit, uint32_t parity) {
      mbar_wait_sleep(so_ready, parity);
      if (lane == 0) {
        const int otile = oit / p.G, og = oit - otile * p.G;
        const HeadTile t = head_tile(otile, p);
        tma_store_4d(&p.tmap_out, so, p.out_coff + og * p.A, t.q0, t.p0, t.n);
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(so_free);
      }
      __syncwarp();
    };
    for (int it = start; it < end; ++it, ++li) {
      const int tile = it / p.G, g = it - tile * p.G;
      nload = li == 0 ? -2 : 0;  // probe: loads of this item in order W2[0..3], W1'[0..1], W3
      if (li == 0) {
        load_x(tile);
        for (int kb = 0; kb < K1; ++kb) load_w(&p.tmap_w1, g * kHeadMid, kb, kHeadSlotBytes);
      }
      for (int kb = 0; kb < 4; ++kb) load_w(&p.tmap_w2, g * kHeadMid, kb, kHeadSlotBytes);
      if (li > 0) store_out(it - 1, (li - 1) & 1);  // the previous item's E3 is staging about now
      if (it + 1 < end) {
        const int ntile = (it + 1) / p.G, ng = (it + 1) - ntile * p.G;
        mbar_wait_sleep(x_empty, li & 1);  // GEMM1 of this item has read X
        if (ntile != tile) load_x(ntile);
        for (int kb = 0; kb < K1; ++kb) load_w(&p.tmap_w1, ng * kHeadMid, kb, kHeadSlotBytes);
      }
      load_w(&p.tmap_w3, g * R3, 0, 4u * R3 * 128u);
    }
    if (end > start) {
      store_out(end - 1, (end - 1 - start) & 1);
      if (lane == 0) tma_store_wait_all();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_mid = umma_idesc_bf16(kHeadMid);
    constexpr uint32_t idesc_out = umma_idesc_bf16(R3);
    const uint32_t acc1 = tmem_base, acc2 = tmem_base + kHeadMid, acc3 = tmem_base + kHeadMid;
    int slot = 0;
    uint32_t wphase = 0;
    uint32_t xcount = 0;
    auto next_slot = [&]() {
      if (++slot == kHeadWSlots) slot = 0, wphase ^= 1;
    };
    auto gemm1 = [&](bool new_x) {
      if (new_x) {
        mbar_wait_sleep(x_full, xcount & 1);
        ++xcount;
      }
      tc_fence_after();
      for (int kb = 0; kb < K1; ++kb) {
        mbar_wait_sleep(&w_full[slot], wphase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = umma_smem_desc<128>(smem_u32(sx + kb * 16384));
          const uint64_t db = umma_smem_desc<128>(smem_u32(sw + slot * kHeadSlotBytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(acc1, da + 2 * k, db + 2 * k, idesc_mid, (kb | k) != 0);
          umma_commit(&w_empty[slot]);
        }
        __syncwarp();
        next_slot();
      }
      if (elect_one()) {
        umma_commit(x_empty);
        umma_commit(acc1_full);
      }
      __syncwarp();
    };
    for (int it = start, li = 0; it < end; ++it, ++li) {
      const uint32_t par = li & 1;
      const int tile = it / p.G;
      if (li == 0) gemm1(true);
      HDBG(0);
      // GEMM2: acc2 = Y1 . W2^T
      mbar_wait_sleep(acc2_empty, par ^ 1);  // E3 of the previous item has drained acc2 / acc3
      tc_fence_after();
      for (int kb = 0; kb < 4; ++kb) {
        mbar_wait_sleep(&w_full[slot], wphase);
        HDBG(16 + kb);
        mbar_wait_sleep(&y1_ready[kb], par);
        HDBG(20 + kb);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = umma_smem_desc<128>(smem_u32(sy + kb * 16384));
          const uint64_t db = umma_smem_desc<128>(smem_u32(sw + slot * kHeadSlotBytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(acc2, da + 2 * k, db + 2 * k, idesc_mid, (kb | k) != 0);
          umma_commit(&w_empty[slot]);
        }
        __syncwarp();
        next_slot();
      }
      if (elect_one()) umma_commit(acc2_full);
      __syncwarp();
      HDBG(1);
      // GEMM1 of the next item runs under E2 of this one
      if (it + 1 < end) {
        mbar_wait_sleep(acc1_empty, par);  // E1 of this item has drained acc1
        gemm1((it + 1) / p.G != tile);
      }
      HDBG(2);
      // GEMM3: acc3 = Y2 . W3^T
      mbar_wait_sleep(&w_full[slot], wphase);
      for (int kb = 0; kb < 4; ++kb) {
        mbar_wait_sleep(&y2_ready[kb], par);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = umma_smem_desc<128>(smem_u32(sy + kb * 16384));
          const uint64_t db = umma_smem_desc<128>(smem_u32(sw + slot * kHeadSlotBytes + kb * (R3 * 128)));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(acc3, da + 2 * k, db + 2 * k, idesc_out, (kb | k) != 0);
        }
        __syncwarp();
      }
      if (elect_one()) {
        umma_commit(&w_empty[slot]);
        umma_commit(acc3_full);
      }
      __syncwarp();
      HDBG(3);
      next_slot();
    }
  } else {
    // ------------------------------------------------------------------ epilogue (16 warps)
    const int quarter = warp & 3;          // TMEM lane quarter this warp may read
    const int part = (warp - 2) >> 2;      // which 16 of a slab's 64 columns
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const unsigned long long slope2 = pack_f32x2(p.slope, p.slope);
    // acc -> +bias, LeakyReLU, bf16 -> Y: every warp converts 16 of the 64 columns of each slab, slab 0 first, so
    // GEMM2 / GEMM3 start on slab 0 while the rest is drained; the TMEM load of slab s+1 is in flight while slab s
    // is converted.  (The bias vectors of the item were prefetched into L1 during the previous item: an L2 round
    // trip per slab used to be exposed here.)
    // Each warp keeps the 4 x 16 bias values of the coming drain in lanes 0-15 of four registers, loaded one epilogue
    // phase ahead and broadcast with shuffles, instead of four read-only-path loads per slab in the drain itself
    // (measured round 2: headsA/C 75 -> 72 us, headsB 45.5 -> 43.5 us; the slab cadence is latency-bound).
    float breg[4];
    auto load_bias_regs = [&](const float* bias) {
#pragma unroll
      for (int s = 0; s < 4; ++s) breg[s] = __ldg(bias + s * 64 + part * 16 + (lane & 15));
    };
    if (start < end) load_bias_regs(p.b1 + (start % p.G) * kHeadMid);
    auto drain_to_y = [&](uint32_t acc, const float* bias, uint64_t* ready, uint64_t* acc_empty, bool so_wait,
                          uint32_t so_par) {
      uint32_t a[2][16];
      tmem_ld16(acc + lane_off + part * 16, a[0]);
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        // E1 only: slabs 2-3 hold the output staging of the previous item until its TMA store has read it
        if (s == 2 && so_wait) mbar_wait(so_free, so_par);
        tmem_ld_wait();
        if (s < 3) tmem_ld16(acc + lane_off + (s + 1) * 64 + part * 16, a[(s + 1) & 1]);
        uint8_t* slab = sy + s * 16384;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          // the warp's 16 bias values of this slab live in lanes 0-15 of breg[s] (loaded an epilogue phase ahead)
          float4 b0, b1;
          b0.x = __shfl_sync(0xffffffffu, breg[s], j * 8 + 0), b0.y = __shfl_sync(0xffffffffu, breg[s], j * 8 + 1);
          b0.z = __shfl_sync(0xffffffffu, breg[s], j * 8 + 2), b0.w = __shfl_sync(0xffffffffu, breg[s], j * 8 + 3);
          b1.x = __shfl_sync(0xffffffffu, breg[s], j * 8 + 4), b1.y = __shfl_sync(0xffffffffu, breg[s], j * 8 + 5);
          b1.z = __shfl_sync(0xffffffffu, breg[s], j * 8 + 6), b1.w = __shfl_sync(0xffffffffu, breg[s], j * 8 + 7);
          // two channels per instruction (add.f32x2 / mul.f32x2 round like their scalar forms)
          const unsigned long long bb[4] = {pack_f32x2(b0.x, b0.y), pack_f32x2(b0.z, b0.w), pack_f32x2(b1.x, b1.y),
                                            pack_f32x2(b1.z, b1.w)};
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const unsigned long long v = add_f32x2(
                pack_f32x2(__uint_as_float(a[s & 1][j * 8 + 2 * e]), __uint_as_float(a[s & 1][j * 8 + 2 * e + 1])), bb[e]);
            const unsigned long long t = mul_f32x2(v, slope2);
            float v0, v1, t0, t1;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(v0), "=f"(v1) : "l"(v));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(t));
            __nv_bfloat162 o = __floats2bfloat162_rn(fmaxf(v0, t0), fmaxf(v1, t1));
            w[e] = *reinterpret_cast<uint32_t*>(&o);
          }
          *reinterpret_cast<uint4*>(slab + swizzled_offset<128>(row, part * 2 + j)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[s]);
      }
      if (acc_empty != nullptr) {  // every TMEM load of this warp has completed (last wait::ld above)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
      }
    };
    // one 128-byte line of the bias vectors of head `hg` per thread -> L1
    const int etid = threadIdx.x - 64;
    auto prefetch_bias = [&](int hg) {
      const float* a = nullptr;
      if (etid < 8) a = p.b1 + hg * kHeadMid + etid * 32;
      else if (etid < 16) a = p.b2 + hg * kHeadMid + (etid - 8) * 32;
      else if (etid < 16 + (R3 + 31) / 32) a = p.b3 + hg * R3 + (etid - 16) * 32;
      if (a != nullptr) asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
    };
    if (start < end) prefetch_bias(start % p.G);
    for (int it = start, li = 0; it < end; ++it, ++li) {
      const uint32_t par = li & 1;
      const int tile = it / p.G, g = it - tile * p.G;
      // E1
      if (warp == 2) HDBG(8);
      mbar_wait(acc1_full, par);
      if (warp == 2) HDBG(9);
      tc_fence_after();
      drain_to_y(tmem_base, p.b1 + g * kHeadMid, y1_ready, acc1_empty, li > 0, (li - 1) & 1);
      load_bias_regs(p.b2 + g * kHeadMid);  // in flight while GEMM2 finishes
      // E2 (acc2_full also means GEMM2 has finished reading Y)
      if (it + 1 < end) prefetch_bias((it + 1) % p.G);
      if (warp == 2) HDBG(10);
      mbar_wait(acc2_full, par);
      if (warp == 2) HDBG(11);
      tc_fence_after();
      drain_to_y(tmem_base + kHeadMid, p.b2 + g * kHeadMid, y2_ready, nullptr, false, 0);
      if (it + 1 < end) load_bias_regs(p.b1 + ((it + 1) % p.G) * kHeadMid);  // in flight during GEMM3 / E3
      // E3: acc3 + b3 -> fp32 staging [128][R3] -> coalesced 16-byte stores (A*4 contiguous bytes per pixel)
      const int A = p.A;
      float4 b3v[4];
      if (part * 16 < R3) {
#pragma unroll
        for (int c = 0; c < 4; ++c) b3v[c] = __ldg(reinterpret_cast<const float4*>(p.b3 + g * R3 + part * 16 + 4 * c));
      }
      if (warp == 2) HDBG(12);
      mbar_wait(acc3_full, par);
      if (warp == 2) HDBG(13);
      tc_fence_after();
      if (part * 16 < A) {
        uint32_t a[16];
        tmem_ld16(tmem_base + kHeadMid + lane_off + part * 16, a);
        tmem_ld_wait();
        float* dst = so + row * A + part * 16;  // 4 * A-byte rows: A = 36 -> the 16-byte stores of a quarter warp hit distinct banks
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
          if (part * 16 + c < A) {
            const float4 b = b3v[c >> 2];
            *reinterpret_cast<float4*>(dst + c) = make_float4(__uint_as_float(a[c]) + b.x, __uint_as_float(a[c + 1]) + b.y,
                                                              __uint_as_float(a[c + 2]) + b.z, __uint_as_float(a[c + 3]) + b.w);
          }
        }
        fence_proxy_async_smem();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(acc2_empty);
        mbar_arrive(so_ready);  // the TMA warp stores the tile
      }
      if (warp == 2) HDBG(14);
    }
  }
  tc_fence_before();
  __syncthreads();
#ifdef M3D_PROBE
  if (blockIdx.x == 0 && threadIdx.x < 6 * 32) g_head_dbg[threadIdx.x] = s_head_dbg[threadIdx.x];
#endif
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace m3d

using namespace m3d;

// x: bf16 NHWC [N,H,W,x_cstride], channels [x_coff, x_coff + Cx).  w1: bf16 [G*256][Cx], w2: bf16 [G*256][256],
// w3: bf16 [G*rows3][256] (rows >= A of each head are zero padding), b1/b2: fp32 [G*256], b3: fp32 [G*rows3].
// out: fp32 NHWC [N,H,W,out_cstride]; head g writes channels [out_coff + g*A, out_coff + (g+1)*A).
extern "C" int m3d_head_mlp(const void* x, int N, int H, int W, int x_cstride, int x_coff, int Cx, const void* w1,
                            const float* b1, const void* w2, const float* b2, const void* w3, const float* b3, int G,
                            int A, int rows3, float* out, int out_cstride, int out_coff, float slope,
                            m3d_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3D_REQUIRE(x && w1 && b1 && w2 && b2 && w3 && b3 && out, "NULL pointer");
  M3D_REQUIRE(N >= 1 && H >= 1 && W >= 1 && G >= 1, "bad geometry");
  M3D_REQUIRE(Cx == 64 || Cx == 128, "head input channels must be 64 or 128 (got %d)", Cx);
  M3D_REQUIRE(x_cstride % 8 == 0 && x_coff % 8 == 0, "x channel stride/offset must keep 16-byte alignment");
  M3D_REQUIRE(rows3 == 48 && A >= 1 && A <= rows3, "head_mlp: rows3 must be 48 and A <= 48 (got rows3=%d A=%d)", rows3, A);
  M3D_REQUIRE(out_cstride % 4 == 0 && out_coff % 4 == 0 && A % 4 == 0, "output channels must keep 16-byte alignment");
  int TW = 16, TH = 8;
  pick_tile(H, W, 256, &TW, &TH);
  HeadMlpParams p;
  memset(&p, 0, sizeof(p));
  int rc = make_tmap_nhwc(&p.tmap_x, x, N, H, W, x_cstride, 64, TW, TH, 1);
  if (rc != M3D_OK) return rc;
  rc = make_tmap_b3d(&p.tmap_w1, w1, static_cast<long>(G) * kHeadMid, Cx, 64, kHeadMid, 1);
  if (rc != M3D_OK) return rc;
  rc = make_tmap_b3d(&p.tmap_w2, w2, static_cast<long>(G) * kHeadMid, kHeadMid, 64, kHeadMid, 1);
  if (rc != M3D_OK) return rc;
  rc = make_tmap_b3d(&p.tmap_w3, w3, static_cast<long>(G) * rows3, kHeadMid, 64, rows3, 4);
  if (rc != M3D_OK) return rc;
  rc = make_tmap_nhwc_f32(&p.tmap_out, out, N, H, W, out_cstride, A, TW, TH, false);
  if (rc != M3D_OK) return rc;
  p.b1 = b1, p.b2 = b2, p.b3 = b3;
  p.out = out, p.out_cstride = out_cstride, p.out_coff = out_coff;
  p.x_coff = x_coff, p.K1 = Cx / 64;
  p.G = G, p.A = A, p.R3 = rows3;
  p.N = N, p.P = H, p.Q = W, p.TW = TW, p.TH = TH;
  p.tiles_w = (W + TW - 1) / TW, p.tiles_h = (H + TH - 1) / TH;
  const long items = static_cast<long>(p.tiles_w) * p.tiles_h * N * G;
  M3D_REQUIRE(items < (1L << 30), "too many work items");
  p.total_items = static_cast<int>(items);
  p.slope = slope;
  auto kern = head_mlp_kernel<48>;
  M3D_ONCE_PER_DEVICE_BEGIN
    M3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeadSmem));
  M3D_ONCE_PER_DEVICE_END
  const int sms = persistent_sms();
  int grid = sms;
  if (grid > p.total_items) grid = p.total_items;
  M3D_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(kHeadThreads), kHeadSmem, stream, p));
  return M3D_OK;
}

#ifdef M3D_PROBE
extern "C" int m3d_head_debug_read(long long* host, int n) {
  return cudaMemcpyFromSymbol(host, m3d::g_head_dbg, sizeof(long long) * n) == cudaSuccess ? 0 : -1;
}
#endif
