// Accumulator drains shared by the conv kernels (TMEM -> bias/residual/LeakyReLU -> global).
//
//  * epilogue_tile_staged: bf16 outputs whose channel count is a multiple of 64.  The 128 x BN tile
//    leaves through shared memory in 64-channel slabs (one swizzled 128-byte row per pixel) written by
//    TMA stores; the residual slab arrives in the same buffer by TMA load.  One bulk transaction per
//    slab instead of 32 scattered 16-byte stores per warp instruction; TMA clips tiles that overhang
//    the image.
//  * epilogue_tile_direct: everything else (fp32 outputs, odd channel counts: 27-ch offsets, 36-ch
//    heads): each thread owns one pixel and stores its channels itself.
#pragma once
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace m3d {

constexpr int kEpiThreads = 128;
constexpr int kEpiBarrier = 2;         // named barrier id of the 4 epilogue warps
constexpr int kSlabBytes = 128 * 128;  // 128 pixels x 64 bf16 channels

// LeakyReLU; slope in [0,1] (1 = identity): max(v, slope*v) is the same value in two instructions
__device__ __forceinline__ float lrelu(float v, float slope) { return fmaxf(v, v * slope); }

// ---------------------------------------------------------------- direct
template <typename OutT, typename ResT>
__device__ __forceinline__ void epilogue_chunk16(const uint32_t (&acc)[16], const float* __restrict__ bias,
                                                 const ResT* __restrict__ res, OutT* __restrict__ out, int nvalid,
                                                 float slope) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(acc[i]);
  const bool full = nvalid >= 16;
  if (bias != nullptr) {
    if (full && (reinterpret_cast<uintptr_t>(bias) & 15) == 0) {
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + h);
        v[4 * h] += b.x, v[4 * h + 1] += b.y, v[4 * h + 2] += b.z, v[4 * h + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i < nvalid) v[i] += __ldg(bias + i);
    }
  }
  if (res != nullptr) {
    if (full && (reinterpret_cast<uintptr_t>(res) & 15) == 0) {
      if constexpr (sizeof(ResT) == 2) {
        const uint4* r4 = reinterpret_cast<const uint4*>(res);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint4 u = __ldg(r4 + h);
          v[h * 8 + 0] += __uint_as_float(u.x << 16), v[h * 8 + 1] += __uint_as_float(u.x & 0xffff0000u);
          v[h * 8 + 2] += __uint_as_float(u.y << 16), v[h * 8 + 3] += __uint_as_float(u.y & 0xffff0000u);
          v[h * 8 + 4] += __uint_as_float(u.z << 16), v[h * 8 + 5] += __uint_as_float(u.z & 0xffff0000u);
          v[h * 8 + 6] += __uint_as_float(u.w << 16), v[h * 8 + 7] += __uint_as_float(u.w & 0xffff0000u);
        }
      } else {
        const float4* r4 = reinterpret_cast<const float4*>(res);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float4 u = __ldg(r4 + h);
          v[4 * h] += u.x, v[4 * h + 1] += u.y, v[4 * h + 2] += u.z, v[4 * h + 3] += u.w;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i < nvalid) v[i] += static_cast<float>(res[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = lrelu(v[i], slope);
  if (full && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
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
#pragma unroll
    for (int i = 0; i < 16; ++i)  // compile-time indices: stays in registers
      if (i < nvalid) out[i] = static_cast<OutT>(v[i]);
  }
}

template <int BN, typename OutT, typename ResT>
__device__ __forceinline__ void epilogue_tile_direct(uint32_t tmem_acc, int quarter, int lane, int n, int p0, int q0,
                                                     int TW, int P, int Q, int col_base, int cout, const float* bias,
                                                     const ResT* res, int res_cstride, OutT* out, int out_cstride,
                                                     float slope) {
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
    const int nvalid = cout - col;
    if (pix_ok && nvalid > 0) {
      epilogue_chunk16<OutT, ResT>(acc, bias ? bias + col : nullptr, res ? res + pix * res_cstride + col : nullptr,
                                   out + pix * out_cstride + col, nvalid, slope);
    }
  }
}

// ---------------------------------------------------------------- staged
struct StagedEpilogue {
  // (no arrays indexed at run time here: they would live in local memory and put LDL / STL into the slab loop)
  uint8_t* stage;      // two 16 KB swizzled staging buffers (1024-byte aligned), buffer b at stage + b * kSlabBytes
  float* bias_s;       // BN floats
  uint64_t* res_bar;   // [2] residual-landed barriers
  uint32_t res_uses0, res_uses1;
  uint32_t slab_count;  // running slab index: selects the buffer
  __device__ __forceinline__ void init(uint8_t* stage_, float* bias_smem, uint64_t* bars) {
    stage = stage_;
    bias_s = bias_smem, res_bar = bars;
    res_uses0 = res_uses1 = 0;
    slab_count = 0;
  }
  __device__ __forceinline__ uint8_t* slab(int b) const { return stage + b * kSlabBytes; }
};

// One 128 x BN bf16 tile.  `ep_tid` in [0, 32*NW) numbers the epilogue threads (NW = 4 or 8 warps; with 8 the
// two warps of a TMEM lane quarter each take 32 of a slab's 64 columns: `half`); thread 0 issues every
// TMA operation (bulk groups are per thread).  c_out / c_res: first channel of the tile inside the
// output / residual buffers.  `on_tmem_drained` is called once the accumulator has been read.
template <int BN, int NW, typename F>
__device__ __forceinline__ void epilogue_tile_staged(StagedEpilogue& st, uint32_t tmem_acc, int quarter, int lane,
                                                     int ep_tid, int n, int p0, int q0, const void* tmap_out, int c_out,
                                                     const void* tmap_res, int c_res, const float* bias, int nbias,
                                                     float slope, F on_tmem_drained) {
  static_assert(BN % 64 == 0, "staged epilogue works on 64-channel slabs");
  static_assert(NW == 4 || NW == 8, "4 or 8 epilogue warps");
  constexpr int NSLAB = BN / 64;
  constexpr int NT = 32 * NW;
  constexpr int CPT = 64 / (NW / 4);  // columns of a slab per thread: 64 or 32
  const int half = NW == 8 ? (ep_tid >> 7) : 0;
  const int row = quarter * 32 + lane;
  const bool leader = ep_tid == 0;
  const bool has_res = tmap_res != nullptr;

  named_bar_sync(kEpiBarrier, NT);  // previous tile no longer reads bias_s
  for (int i = ep_tid; i < BN; i += NT) st.bias_s[i] = (bias != nullptr && i < nbias) ? __ldg(bias + i) : 0.f;
  if (has_res && leader) {
    const int b = st.slab_count & 1;
    tma_store_wait_read<1>();  // the store that last used buffer b (two slabs ago) has drained it
    mbar_arrive_expect_tx(&st.res_bar[b], kSlabBytes);
    tma_load_4d(st.slab(b), tmap_res, &st.res_bar[b], c_res, q0, p0, n);
  }
  named_bar_sync(kEpiBarrier, NT);  // bias_s visible

#pragma unroll 1
  for (int s = 0; s < NSLAB; ++s) {
    const int b = st.slab_count & 1;
    uint8_t* buf = st.slab(b);
    const uint32_t buf_s = smem_u32(st.stage) + static_cast<uint32_t>(b) * kSlabBytes;  // shared-space address of buf
    if (has_res) {
      if (leader && s + 1 < NSLAB) {  // prefetch the next residual slab into the other buffer
        tma_store_wait_read<0>();
        mbar_arrive_expect_tx(&st.res_bar[b ^ 1], kSlabBytes);
        tma_load_4d(st.slab(b ^ 1), tmap_res, &st.res_bar[b ^ 1], c_res + (s + 1) * 64, q0, p0, n);
      }
      mbar_wait(&st.res_bar[b], (b ? st.res_uses1 : st.res_uses0) & 1);
      if (b) ++st.res_uses1; else ++st.res_uses0;
    } else {
      if (leader) tma_store_wait_read<1>();
      named_bar_sync(kEpiBarrier, NT);  // buffer b is free for everybody
    }
    uint32_t a[CPT];
    const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + s * 64 + half * 32;
    if constexpr (CPT == 64) {
      tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(a));
      tmem_ld32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(a + 32));
    } else {
      tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(a));
    }
    tmem_ld_wait();
    if (s == NSLAB - 1) on_tmem_drained();
#pragma unroll
    for (int j = 0; j < CPT / 8; ++j) {  // chunks of 8 channels
      const int chunk = half * 4 + j;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(a[j * 8 + e]);
      const float4 b0 = *reinterpret_cast<const float4*>(st.bias_s + s * 64 + chunk * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(st.bias_s + s * 64 + chunk * 8 + 4);
      v[0] += b0.x, v[1] += b0.y, v[2] += b0.z, v[3] += b0.w, v[4] += b1.x, v[5] += b1.y, v[6] += b1.z, v[7] += b1.w;
      const uint32_t cell = buf_s + swizzled_offset<128>(row, chunk);
      if (has_res) {
        const uint4 u = lds128(cell);
        v[0] += __uint_as_float(u.x << 16), v[1] += __uint_as_float(u.x & 0xffff0000u);
        v[2] += __uint_as_float(u.y << 16), v[3] += __uint_as_float(u.y & 0xffff0000u);
        v[4] += __uint_as_float(u.z << 16), v[5] += __uint_as_float(u.z & 0xffff0000u);
        v[6] += __uint_as_float(u.w << 16), v[7] += __uint_as_float(u.w & 0xffff0000u);
      }
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        __nv_bfloat162 t = __floats2bfloat162_rn(lrelu(v[2 * e], slope), lrelu(v[2 * e + 1], slope));
        w[e] = *reinterpret_cast<uint32_t*>(&t);
      }
      sts128(cell, make_uint4(w[0], w[1], w[2], w[3]));
    }
    fence_proxy_async_smem();
    named_bar_sync(kEpiBarrier, NT);  // slab complete
    if (leader) {
      tma_store_4d(tmap_out, buf, c_out + s * 64, q0, p0, n);
      tma_store_commit();
    }
    st.slab_count++;
  }
}

// ---------------------------------------------------------------- grouped
// Same job as epilogue_tile_staged with the serial chain taken apart (timeline probe, round 2: the staged drain of a
// 128 x 128 tile held its accumulator for ~5.5 k clocks and took 6.8 k in all -- bias __ldg, residual TMA issued only
// after the accumulator was complete, two slabs in lock step -- against 4.6 k clocks of MMAs per tile).  Here
//  * the 8 epilogue warps are two independent groups of 4; group g owns columns [g*BN/2, (g+1)*BN/2) of the tile, one
//    16 KB slab buffer, one `ready` mbarrier and one named barrier: no cross-group synchronisation;
//  * a thread pulls all 64 columns of its row of a slab out of TMEM at once, so a 128-column accumulator is released
//    a few hundred clocks after it completed;
//  * the residual slab of the NEXT tile is TMA-loaded into the group's buffer as soon as the store of the current
//    one has read it, i.e. a whole tile time before it is needed; without a residual the leader just arrives;
//  * the layer's biases are copied to shared memory once per kernel.
// SPLIT = true: the two groups share every tile (columns halved; BN = 128, 256: conv_halo2.cu).  SPLIT = false: the
// groups take alternate tiles whole (BN = 64, one slab per tile: conv_halo.cu) -- group g then always drains
// accumulator stage g, and `next` is the group's next tile, two tiles ahead.
struct EpiTile {
  int n, p0, q0, col0;  // image, first tile row / column, first output channel of the BN-wide tile
};
#ifndef EPI_STAMP  // timeline probe hook (conv_halo2.cu defines it under -DM3D_PROBE)
#define EPI_STAMP(slot) do { } while (0)
#endif

struct GroupedEpilogue {
  uint8_t* buf;         // this group's slab (1024-byte aligned)
  const float* bias_s;  // every bias of the layer (zero padded to n_tiles * BN)
  uint64_t* ready;      // residual landed / buffer free
  uint32_t uses;
  __device__ __forceinline__ void init(uint8_t* buf_, const float* bias_smem, uint64_t* ready_) {
    buf = buf_, bias_s = bias_smem, ready = ready_;
    uses = 0;
  }
};

// column offset (inside the BN tile) of slab k of group g
template <int BN, bool SPLIT = true>
__device__ __forceinline__ int grouped_col(int group, int k) {
  return (SPLIT ? group * (BN / 2) : 0) + k * 64;
}

// Before the first tile: the group leader (gtid == 0) fetches the first residual slab / declares the buffer free.
template <int BN, bool SPLIT = true>
__device__ __forceinline__ void epilogue_grouped_begin(GroupedEpilogue& st, int gtid, int group, const EpiTile& t,
                                                       const void* tmap_res, int res_coff) {
  if (gtid != 0) return;
  if (tmap_res != nullptr) {
    mbar_arrive_expect_tx(st.ready, kSlabBytes);
    tma_load_4d(st.buf, tmap_res, st.ready, res_coff + t.col0 + grouped_col<BN, SPLIT>(group, 0), t.q0, t.p0, t.n);
  } else {
    mbar_arrive(st.ready);
  }
}

template <int BN, bool SPLIT = true, typename F>
__device__ __forceinline__ void epilogue_tile_grouped(GroupedEpilogue& st, uint32_t tmem_acc, int quarter, int lane,
                                                      int gtid, int group, const EpiTile& t, const EpiTile* next,
                                                      const void* tmap_out, int out_coff, const void* tmap_res,
                                                      int res_coff, float slope, F on_tmem_drained) {
  static_assert(BN % (SPLIT ? 128 : 64) == 0, "groups work on whole 64-channel slabs");
  constexpr int NS = SPLIT ? BN / 128 : BN / 64;  // slabs per group and tile
  const int row = quarter * 32 + lane;
  const bool has_res = tmap_res != nullptr;
  const uint32_t buf_s = smem_u32(st.buf);
#pragma unroll 1
  for (int k = 0; k < NS; ++k) {
    const int c = grouped_col<BN, SPLIT>(group, k);
    uint32_t a[64];
    const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + c;
    tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(a));
    tmem_ld32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(a + 32));
    tmem_ld_wait();
    EPI_STAMP(0);
    if (k == NS - 1) on_tmem_drained();
    mbar_wait(st.ready, st.uses & 1);
    EPI_STAMP(1);
    ++st.uses;
    const float* bias = st.bias_s + t.col0 + c;
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // chunks of 8 channels
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(a[j * 8 + e]);
      const float4 b0 = *reinterpret_cast<const float4*>(bias + j * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(bias + j * 8 + 4);
      v[0] += b0.x, v[1] += b0.y, v[2] += b0.z, v[3] += b0.w, v[4] += b1.x, v[5] += b1.y, v[6] += b1.z, v[7] += b1.w;
      const uint32_t cell = buf_s + swizzled_offset<128>(row, j);
      if (has_res) {
        const uint4 u = lds128(cell);
        v[0] += __uint_as_float(u.x << 16), v[1] += __uint_as_float(u.x & 0xffff0000u);
        v[2] += __uint_as_float(u.y << 16), v[3] += __uint_as_float(u.y & 0xffff0000u);
        v[4] += __uint_as_float(u.z << 16), v[5] += __uint_as_float(u.z & 0xffff0000u);
        v[6] += __uint_as_float(u.w << 16), v[7] += __uint_as_float(u.w & 0xffff0000u);
      }
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        __nv_bfloat162 h = __floats2bfloat162_rn(lrelu(v[2 * e], slope), lrelu(v[2 * e + 1], slope));
        w[e] = *reinterpret_cast<uint32_t*>(&h);
      }
      sts128(cell, make_uint4(w[0], w[1], w[2], w[3]));
    }
    EPI_STAMP(2);
    fence_proxy_async_smem();
    EPI_STAMP(3);
    named_bar_sync(kEpiBarrier + group, 128);  // slab complete
    EPI_STAMP(4);
    if (gtid == 0) {
      tma_store_4d(tmap_out, st.buf, out_coff + t.col0 + c, t.q0, t.p0, t.n);
      tma_store_commit();
      const bool same = k + 1 < NS;
      if (same || next != nullptr) {
        tma_store_wait_read<0>();  // the buffer has left
        EPI_STAMP(5);
        if (has_res) {
          const EpiTile& u = same ? t : *next;
          mbar_arrive_expect_tx(st.ready, kSlabBytes);
          tma_load_4d(st.buf, tmap_res, st.ready, res_coff + u.col0 + grouped_col<BN, SPLIT>(group, same ? k + 1 : 0), u.q0,
                      u.p0, u.n);
        } else {
          mbar_arrive(st.ready);
        }
      }
    }
  }
}

// One 128 x BN fp32 tile (bf16 activations, fp32 output: the class logits).  Same staging as above with 32-column
// slabs (128 rows x 128 bytes, swizzled, one TMA store per slab): a thread writing its row's 576 bytes straight to
// global memory touches 32 different lines per warp instruction, which made cls.l3 LSU-bound.  No residual.  Slabs
// past `nvalid` columns are skipped; the last one may be partial only if it ends at the tensor's channel extent
// (the TMA store clips there) -- the host checks.
template <int BN, int NW, typename F>
__device__ __forceinline__ void epilogue_tile_staged_f32(StagedEpilogue& st, uint32_t tmem_acc, int quarter, int lane,
                                                         int ep_tid, int n, int p0, int q0, const void* tmap_out,
                                                         int c_out, const float* bias, int nvalid, float slope,
                                                         F on_tmem_drained) {
  static_assert(NW == 8, "8 epilogue warps");
  constexpr int NT = 32 * NW;
  const int half = ep_tid >> 7;
  const int row = quarter * 32 + lane;
  const bool leader = ep_tid == 0;
  const int nslab = (min(nvalid, BN) + 31) >> 5;

  named_bar_sync(kEpiBarrier, NT);  // previous tile no longer reads bias_s
  for (int i = ep_tid; i < BN; i += NT) st.bias_s[i] = (bias != nullptr && i < nvalid) ? __ldg(bias + i) : 0.f;
  named_bar_sync(kEpiBarrier, NT);  // bias_s visible

#pragma unroll 1
  for (int s = 0; s < nslab; ++s) {
    const int b = st.slab_count & 1;
    const uint32_t buf_s = smem_u32(st.stage) + static_cast<uint32_t>(b) * kSlabBytes;
    uint32_t a[16];
    tmem_ld16(tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + s * 32 + half * 16, a);
    if (leader) tma_store_wait_read<1>();  // the store that last used buffer b (two slabs ago) has drained it
    named_bar_sync(kEpiBarrier, NT);       // buffer b is free for everybody
    tmem_ld_wait();
    if (s == nslab - 1) on_tmem_drained();
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // chunks of 4 channels
      const float4 bb = *reinterpret_cast<const float4*>(st.bias_s + s * 32 + half * 16 + j * 4);
      const float v0 = __uint_as_float(a[4 * j]) + bb.x, v1 = __uint_as_float(a[4 * j + 1]) + bb.y;
      const float v2 = __uint_as_float(a[4 * j + 2]) + bb.z, v3 = __uint_as_float(a[4 * j + 3]) + bb.w;
      sts128(buf_s + swizzled_offset<128>(row, half * 4 + j),
             make_uint4(__float_as_uint(lrelu(v0, slope)), __float_as_uint(lrelu(v1, slope)),
                        __float_as_uint(lrelu(v2, slope)), __float_as_uint(lrelu(v3, slope))));
    }
    fence_proxy_async_smem();
    named_bar_sync(kEpiBarrier, NT);  // slab complete
    if (leader) {
      tma_store_4d(tmap_out, st.slab(b), c_out + s * 32, q0, p0, n);
      tma_store_commit();
    }
    st.slab_count++;
  }
}

}  // namespace m3d
