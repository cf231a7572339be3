// Stem of the bf16 trunk: the 7x7 / pad 3 convolution of the fp32 NCHW image (pose_dla_dcn.py:336-340, base_layer)
// computed in 2x2 space-to-depth form: output pixel (Y, X) holds the 2x2 block of full-resolution pixels x 16
// channels = 64 "channels", its receptive field is the 8x8 window at (2Y-3, 2X-3) of the 3 image channels, K = 192.
// Weights are packed by ops.pack_stem_s2d ([64][192] bf16, zero where a tap falls outside a sub-pixel's 7x7).
//
// A dedicated kernel because what bounds this layer is neither HBM nor the tensor pipe but the CUDA-core work of
// building the A operand and draining the accumulator (timeline probe of the shared gather kernel: 8 producer warps
// needed ~880 clk per k-block and 4 epilogue warps 2700 clk per tile, the image TMA a further ~780 clk because it
// was issued only one tile ahead):
//   warps 0-11  : A producers.  A tile = all three k-blocks (48 KB): ONE hand-shake per tile, 8 window rows of 8
//                 floats per thread (5 ld.shared each, issued together), fp32 -> bf16, swizzled st.shared.
//   warp 12     : TMA: the weights once (resident, 24 KB), image tiles three tiles ahead (3 buffers).
//   warp 13     : tcgen05.mma, 12 per tile, two TMEM accumulators.
//   warps 14-21 : staged epilogue (bias + LeakyReLU -> bf16 -> swizzled staging -> TMA store), 8 warps.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "epilogue.cuh"
#include "igemm.cuh"
#include "ptx.cuh"

namespace m3d {

int make_tmap_2d(CUtensorMap* map, const void* base, long rows, long cols, int box_cols, int box_rows);
int make_tmap_nhwc(CUtensorMap* map, const void* base, int N, int H, int W, int C, int bk, int tw, int th, int stride);
int make_tmap_img(CUtensorMap* map, const void* base, int N, int H, int W, int box_w, int box_h);
void pick_tile(int P, int Q, int max_tw, int* TW, int* TH);

namespace {

constexpr int kStemProd = 384;                      // producer threads (12 warps)
constexpr int kStemEpiWarps = 8;
constexpr int kStemThreads = kStemProd + 64 + 32 * kStemEpiWarps;  // 704
constexpr int kStemABytes = 3 * kTileM * 128;       // three k-blocks of [128][64] bf16
constexpr int kStemAStages = 2;
constexpr int kStemBBytes = 3 * 64 * 128;           // [3][64 rows][64] bf16
constexpr int kStemImgBufs = 3;
constexpr int kStemImgBytes = 12 * 1024;            // >= 3 * (2TH+6) * ld * 4 (host checks)
constexpr int kStemSmem = kStemAStages * kStemABytes + kStemBBytes + kStemImgBufs * kStemImgBytes + 2 * kSlabBytes + 1024 +
                          512 + 1024;

struct alignas(64) StemParams {
  CUtensorMap tmap_img;  // fp32 (W, H, 3, N) box {ld, 2TH+6, 3, 1}
  CUtensorMap tmap_b;    // bf16 [64][192] box {64, 64}
  CUtensorMap tmap_out;  // bf16 (64, Q, P, N) box {64, TW, TH, 1}
  const float* bias;     // [64]
  int N, P, Q, TW, TH, tiles_w, tiles_h, total_tiles;
  int ld;  // row pitch (floats) of the image tile in shared memory
  float slope;
};

__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&b);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

struct STile {
  int n, p0, q0;
};
__device__ __forceinline__ STile stile(int tile, const StemParams& p) {
  STile t;
  const int tw = tile % p.tiles_w;
  const int r = tile / p.tiles_w;
  const int th = r % p.tiles_h;
  t.n = r / p.tiles_h;
  t.p0 = th * p.TH;
  t.q0 = tw * p.TW;
  return t;
}

__global__ void __launch_bounds__(kStemThreads, 1) stem_s2d_kernel(const __grid_constant__ StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sa = smem;                                              // A stages
  uint8_t* sb = smem + kStemAStages * kStemABytes;                 // resident weights
  uint8_t* simg = sb + kStemBBytes;                                // image buffers
  uint8_t* stage_out = simg + kStemImgBufs * kStemImgBytes;        // 2 staging slabs + bias
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + 2 * kSlabBytes + 1024);
  uint64_t* a_full = bars;                      // [2] 12 producer warps
  uint64_t* a_empty = bars + 2;                 // [2] MMA commit
  uint64_t* img_full = bars + 4;                // [3] TMA
  uint64_t* img_empty = bars + 7;               // [3] 12 producer warps
  uint64_t* tfull = bars + 10;                  // [2]
  uint64_t* tempty = bars + 12;                 // [2] 8 epilogue warps
  uint64_t* b_full = bars + 14;
  uint64_t* res_bar = bars + 15;                // [2] (staged epilogue interface; no residual here)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], kStemProd / 32);
      mbar_init(&a_empty[s], 1);
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], kStemEpiWarps);
      mbar_init(&res_bar[s], 1);
    }
    for (int s = 0; s < kStemImgBufs; ++s) {
      mbar_init(&img_full[s], 1);
      mbar_init(&img_empty[s], kStemProd / 32);
    }
    mbar_init(b_full, 1);
    fence_barrier_init();
    prefetch_tmap(&p.tmap_img);
    prefetch_tmap(&p.tmap_b);
    prefetch_tmap(&p.tmap_out);
  }
  if (warp == 13) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_sync();

  const int th2 = 2 * p.TH + 6, ld = p.ld;
  const uint32_t img_bytes = static_cast<uint32_t>(3 * th2 * ld * 4);

  if (warp < 12) {
    // ------------------------------------------------------------ A producers
    const int pt = threadIdx.x;
    const int tw_shift = 31 - __clz(p.TW);
    int local = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
      const int ib = local % kStemImgBufs, st = local & 1;
      mbar_wait(&img_full[ib], (local / kStemImgBufs) & 1);
      mbar_wait(&a_empty[st], ((local >> 1) & 1) ^ 1);
      const uint32_t s_img = smem_u32(simg + ib * kStemImgBytes);
      uint8_t* a_tile = sa + st * kStemABytes;
      // 3 k-blocks x 128 rows x 8 window rows = 3072 (row, j) cells, 8 per thread; cell (kb, row, j) = the 8 floats at
      // image row 2Y + j, columns 2X + 1 .. 2X + 8 of the tile (TMA origin (2q0 - 4, 2p0 - 3)) of channel kb
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = pt + kStemProd * i;
        const int kb = idx >> 10, row = (idx >> 3) & 127, j = idx & 7;
        const uint32_t src =
            s_img + static_cast<uint32_t>((kb * th2 + 2 * (row >> tw_shift) + j) * ld + 2 * (row & (p.TW - 1))) * 4u;
        float v[8];
        asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(v[0]) : "r"(src));
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+8];" : "=f"(v[1]), "=f"(v[2]) : "r"(src));
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+16];" : "=f"(v[3]), "=f"(v[4]) : "r"(src));
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+24];" : "=f"(v[5]), "=f"(v[6]) : "r"(src));
        asm volatile("ld.shared.f32 %0, [%1+32];" : "=f"(v[7]) : "r"(src));
        *reinterpret_cast<uint4*>(a_tile + kb * (kTileM * 128) + swizzled_offset<128>(row, j)) = pack8(v);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&a_full[st]);
        mbar_arrive(&img_empty[ib]);  // (the loads above have completed: their values were consumed)
      }
    }
  } else if (warp == 12) {
    // ------------------------------------------------------------ TMA: weights once, image tiles ahead
    if (elect_one()) {
      mbar_arrive_expect_tx(b_full, kStemBBytes);
      for (int kb = 0; kb < 3; ++kb) tma_load_2d(sb + kb * (64 * 128), &p.tmap_b, b_full, kb * 64, 0);
    }
    __syncwarp();
    int local = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
      const int ib = local % kStemImgBufs;
      mbar_wait_sleep(&img_empty[ib], ((local / kStemImgBufs) & 1) ^ 1);
      if (elect_one()) {
        const STile t = stile(tile, p);
        mbar_arrive_expect_tx(&img_full[ib], img_bytes);
        // x origin 2*q0 - 4 keeps the innermost coordinate 16-byte aligned (window columns start at +1)
        tma_load_4d(simg + ib * kStemImgBytes, &p.tmap_img, &img_full[ib], 2 * t.q0 - 4, 2 * t.p0 - 3, 0, t.n);
      }
      __syncwarp();
    }
  } else if (warp == 13) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(64);
    mbar_wait(b_full, 0);
    int local = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
      const int as = local & 1;
      const uint32_t ph = (local >> 1) & 1;
      mbar_wait(&tempty[as], ph ^ 1);
      mbar_wait(&a_full[as], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = umma_smem_desc<128>(smem_u32(sa + as * kStemABytes));
        const uint64_t db = umma_smem_desc<128>(smem_u32(sb));
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + as * 64, da + kb * ((kTileM * 128) >> 4) + 2 * k, db + kb * ((64 * 128) >> 4) + 2 * k, idesc,
                     (kb | k) != 0);
        }
        umma_commit(&a_empty[as]);
        umma_commit(&tfull[as]);
      }
      __syncwarp();
    }
  } else {
    // --------------------------------------------------------------- epilogue (8 warps)
    const int quarter = warp & 3;
    const int ep_tid = threadIdx.x - (kStemProd + 64);
    StagedEpilogue st;
    st.init(stage_out, reinterpret_cast<float*>(stage_out + 2 * kSlabBytes), res_bar);
    int local = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
      const STile t = stile(tile, p);
      const int as = local & 1;
      mbar_wait(&tfull[as], (local >> 1) & 1);
      tc_fence_after();
      epilogue_tile_staged<64, kStemEpiWarps>(st, tmem_base + as * 64, quarter, lane, ep_tid, t.n, t.p0, t.q0, &p.tmap_out, 0,
                                              nullptr, 0, p.bias, 64, p.slope, [&]() {
                                                tc_fence_before();
                                                __syncwarp();
                                                if (lane == 0) mbar_arrive(&tempty[as]);
                                              });
    }
    if (ep_tid == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

}  // namespace

// Returns M3D_ERR_UNSUPPORTED when the tile's image window does not fit the buffers (the caller falls back to the
// shared gather kernel).
int launch_stem_s2d(const float* image, const void* weight, const float* bias, void* out, int N, int H, int W, float slope,
                    cudaStream_t stream) {
  const int P = H / 2, Q = W / 2;
  int TW = 16, TH = 8;
  pick_tile(P, Q, 64, &TW, &TH);
  // Row pitch of the image tile: >= 2TW + 8 and = 12 or 20 (mod 32) floats, so that the 8 window rows x 2 pixels a
  // half warp reads with one 64-bit ld.shared fall into 16 distinct bank pairs (pitch 40 was a 2-way conflict and the
  // kernel is bound by shared-memory wavefronts: every image float is read ~16 times).
  int ld = 2 * TW + 8;
  while (ld % 32 != 12 && ld % 32 != 20) ld += 4;
  if (3 * (2 * TH + 6) * ld * 4 > kStemImgBytes || (TW & (TW - 1)) != 0) return M3D_ERR_UNSUPPORTED;
  StemParams p;
  memset(&p, 0, sizeof(p));
  int rc = make_tmap_2d(&p.tmap_b, weight, 64, 192, 64, 64);
  if (rc != M3D_OK) return rc;
  rc = make_tmap_nhwc(&p.tmap_out, out, N, P, Q, 64, 64, TW, TH, 1);
  if (rc != M3D_OK) return rc;
  rc = make_tmap_img(&p.tmap_img, image, N, H, W, ld, 2 * TH + 6);
  if (rc != M3D_OK) return rc;
  p.bias = bias;
  p.ld = ld;
  p.N = N, p.P = P, p.Q = Q, p.TW = TW, p.TH = TH;
  p.tiles_w = (Q + TW - 1) / TW, p.tiles_h = (P + TH - 1) / TH;
  const long tiles = static_cast<long>(p.tiles_w) * p.tiles_h * N;
  M3D_REQUIRE(tiles < (1L << 30), "too many tiles");
  p.total_tiles = static_cast<int>(tiles);
  p.slope = slope;
  M3D_ONCE_PER_DEVICE_BEGIN
    M3D_CUDA_OK(cudaFuncSetAttribute(stem_s2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemSmem));
  M3D_ONCE_PER_DEVICE_END
  int grid = persistent_sms();
  if (grid > p.total_tiles) grid = p.total_tiles;
  M3D_CUDA_OK(launch_pdl(stem_s2d_kernel, dim3(grid), dim3(kStemThreads), kStemSmem, stream, p));
  return M3D_OK;
}

}  // namespace m3d
