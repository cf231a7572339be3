// 3x3 stride-1 convolution on CTA pairs (tcgen05 cta_group::2), vertical-halo A windows as in conv_halo.cu.
//
// The single-CTA kernels of the 128/256/512-channel layers are bounded by the SM's ~64 B/clk L2 port: an
// M = 128 tile needs 64 bytes of weights per tensor clock whatever N is.  A CTA pair computes M = 256 (each CTA
// its own 128-pixel tile, its own accumulator in its own TMEM) x N = BN with ONE tcgen05.mma.cta_group::2 issued by
// the leader; each CTA loads only HALF of the weight rows (BN/2) of a stage and the tensor cores read both halves,
// so the weight bytes through each SM's port halve:
//
//   stage (kernel column s, 64-channel chunk):  A window (TH+2) x TW x 64ch = 20 KB  +  3 taps x BN/2 x 128 B
//   BN = 256: 68 KB per 1536 tensor clocks = 44 B/clk      BN = 128: 44 KB per 768 clocks = 57 B/clk
//
// Protocol (same shared-memory offsets in both CTAs):
//   full[s]   leader's barrier only: the leader's producer arms it with the bytes of BOTH CTAs; both producers'
//             TMA loads (.cta_group::2, barrier operand = mapa address of the leader's copy) complete on it
//   empty[s]  each CTA's own: the leader's tcgen05.commit multicasts the arrival to both CTAs
//   tfull[a]  each CTA's own (multicast commit): accumulator stage a is complete, both epilogues drain their TMEM
//   tempty[a] leader's only: 2 x NW arrivals, the peer's epilogue warps arrive remotely (mapa)
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#ifdef M3D_PROBE
// epilogue stamps of the leader CTA's first epilogue thread: rows 32.. (8 slots per tile, two rows of the table)
namespace m3d { __device__ long long g_h2_epi[64]; __device__ int g_h2_epi_row; }
#define EPI_STAMP(slot) do { if (blockIdx.x == 0 && threadIdx.x == 64 && m3d::g_h2_epi_row < 7) m3d::g_h2_epi[m3d::g_h2_epi_row * 8 + (slot)] = clock64(); } while (0)
#endif
#include "epilogue.cuh"
#include "igemm.cuh"
#include "ptx.cuh"

namespace m3d {

int make_tmap_b_halo(CUtensorMap* map, const void* base, long rows, int nchunk, int bn);

namespace {

constexpr int kNW = 8;                        // epilogue warps per CTA
constexpr int kThreads2 = 64 + 32 * kNW;

template <int BN>
struct Halo2Cfg {
  static constexpr int A_BYTES = 20 * 1024;
  static constexpr int BH_BYTES = (BN / 2) * 128;  // this CTA's half of one tap's weight tile
  static constexpr int STAGE = A_BYTES + 3 * BH_BYTES;
  static constexpr int BIAS_FLOATS = 1024;  // every bias of the layer lives in shared memory (host gate: Cout <= 1024)
  static constexpr int EXTRA = 2 * kSlabBytes + BIAS_FLOATS * 4;
  static constexpr int BUDGET = 225 * 1024 + 512 - EXTRA;
  static constexpr int FIT = BUDGET / STAGE;
  static constexpr int STAGES = FIT >= 4 ? 4 : FIT;
  static constexpr int SMEM = STAGES * STAGE + EXTRA + 1024 + 256;
  static constexpr int ACC = BN <= 128 ? 128 : 256;
  static_assert(STAGES >= 2, "pair conv tile does not fit shared memory");
};

// ---- cta_group::2 PTX wrappers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of the same shared-memory offset in the leader (rank 0) CTA
__device__ __forceinline__ uint32_t leader_addr(const void* p) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(smem_u32(p)));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once the issued MMAs have completed
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// TMA loads into OWN shared memory whose completion is signalled on the LEADER's barrier (mapa address)
__device__ __forceinline__ void tma2_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_addr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// arrive on the leader CTA's copy of `bar` (works from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(leader_addr(bar))
               : "memory");
}
// instruction descriptor: bf16 x bf16 -> fp32, K-major operands, M = 256 (the pair), N = n
__host__ __device__ constexpr uint32_t umma2_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);
}

struct H2Tile {
  int nt, n, p0, q0;
};

// Timeline probe (tools/probe_halo2.py): compile with -DM3D_PROBE.  Cluster 0, leader CTA; rows = pipeline stages
// (first 40) / items; stamps in shared memory, copied out at kernel end.
#ifdef M3D_PROBE
__device__ long long g_h2_dbg[64 * 4];
#define H2DBG(row, slot) do { if (blockIdx.x == 0 && lane == 0 && (row) < 64) s_dbg[(row) * 4 + (slot)] = clock64(); } while (0)
#else
#define H2DBG(row, slot) do { } while (0)
#endif

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
    conv_halo2_kernel(const __grid_constant__ ConvTmaParams p) {
  using Cfg = Halo2Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
#ifdef M3D_PROBE
  __shared__ long long s_dbg[64 * 4];
  if (blockIdx.x == 0 && threadIdx.x == 64) g_h2_epi_row = 0;
  if (threadIdx.x < 64 * 4) s_dbg[threadIdx.x] = 0;
  int prow = 0;
  const long long t_entry = clock64();
#endif
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_out = smem + STAGES * Cfg::STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE + Cfg::EXTRA);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* ready = tempty + 2;  // per epilogue group: residual slab landed / slab buffer free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ready + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);   // the leader's producer (expect_tx of both CTAs' bytes)
      mbar_init(&empty[s], 1);  // multicast commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);          // multicast commit
      mbar_init(&tempty[s], 2 * kNW);   // one arrival per epilogue warp of both CTAs
      mbar_init(&ready[s], 1);
    }
    fence_barrier_init();
    prefetch_tmap(&p.tmap_a[0]);
    prefetch_tmap(&p.tmap_b);
    prefetch_tmap(&p.tmap_out);
  }
  if (warp == 1) tmem_alloc2<2 * Cfg::ACC>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before anything signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
#ifdef M3D_PROBE
  const long long t_setup = clock64();
#endif
  grid_dep_sync();
#ifdef M3D_PROBE
  if (blockIdx.x == 0 && threadIdx.x == 0) s_dbg[47 * 4] = t_entry, s_dbg[47 * 4 + 1] = t_setup, s_dbg[47 * 4 + 2] = clock64();
#endif

  const int nchunk = p.chunks[0];
  const int n_stages = 3 * nchunk;
  const uint32_t a_bytes = static_cast<uint32_t>((p.TH + 2) * p.TW * 128);
  const uint32_t tap_shift = static_cast<uint32_t>(p.TW * 128) >> 4;

  // work items: (pair of consecutive 128-pixel tiles of the flattened (image, row, column) list, N tile): the two
  // CTAs share the weights, not the window, so the tiles need not be neighbours.  Contiguous range per cluster.
  const int m_tiles = p.tiles_w * p.tiles_h * p.N;  // even (host)
  const int items = (m_tiles / 2) * p.n_tiles;
  const int ncl = gridDim.x / 2, cl = blockIdx.x / 2;
  const int first = static_cast<int>(static_cast<long>(cl) * items / ncl);
  const int last = static_cast<int>(static_cast<long>(cl + 1) * items / ncl);
  auto item_tile = [&](int it) {
    H2Tile t;
    t.nt = it % p.n_tiles;
    int r = 2 * (it / p.n_tiles) + static_cast<int>(rank);  // this CTA's M tile
    const int tw = r % p.tiles_w;
    r /= p.tiles_w;
    const int th = r % p.tiles_h;
    t.n = r / p.tiles_h;
    t.p0 = th * p.TH;
    t.q0 = tw * p.TW;
    return t;
  };

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int it = first; it < last; ++it) {
      const H2Tile t = item_tile(it);
      const int brow = t.nt * BN + static_cast<int>(rank) * (BN / 2);
      int s = 0, c = 0;
      for (int st = 0; st < n_stages; ++st) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * Cfg::STAGE;
          if (leader) mbar_arrive_expect_tx(&full[stage], 2 * (a_bytes + 3 * Cfg::BH_BYTES));
          tma2_load_4d(sa, &p.tmap_a[0], &full[stage], p.a_coff[0] + c * 64, t.q0 - 1 + s, t.p0 - 1, t.n);
          tma2_load_4d(sa + Cfg::A_BYTES, &p.tmap_b, &full[stage], 0, brow, s * nchunk + c, 0);
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
    if (leader) {
      constexpr uint32_t idesc = umma2_idesc_bf16(BN);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int it = first; it < last; ++it, ++local) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        H2DBG(48 + local, 0);
        mbar_wait(&tempty[as], aphase ^ 1);
        H2DBG(48 + local, 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + as * Cfg::ACC;
        for (int st = 0; st < n_stages; ++st) {
          H2DBG(prow, 0);
          mbar_wait(&full[stage], phase);
          H2DBG(prow, 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE);
            const uint64_t da = umma_smem_desc<128>(sa);
            const uint64_t db = umma_smem_desc<128>(sa + Cfg::A_BYTES);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma2_f16(tmem_acc, da + r * tap_shift + 2 * k, db + r * (Cfg::BH_BYTES >> 4) + 2 * k, idesc,
                          (st | r | k) != 0);
            }
            umma2_commit_mc(&empty[stage]);
          }
          __syncwarp();
#ifdef M3D_PROBE
          H2DBG(prow, 2);
          ++prow;
#endif
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma2_commit_mc(&tfull[as]);
        __syncwarp();
      }
    }
  } else {
    // two independent groups of 4 warps, each draining half of the tile's columns (epilogue.cuh, "grouped")
    const int quarter = warp & 3;
    const int ep_tid = threadIdx.x - 64;
    const int group = ep_tid >> 7, gtid = ep_tid & 127;
    float* bias_s = reinterpret_cast<float*>(stage_out + 2 * kSlabBytes);
    const int nb = p.n_tiles * BN;
    for (int i = ep_tid; i < nb; i += 32 * kNW) bias_s[i] = (p.bias != nullptr && i < p.Cout) ? __ldg(p.bias + i) : 0.f;
    named_bar_sync(kEpiBarrier + 2, 32 * kNW);
    GroupedEpilogue st;
    st.init(stage_out + group * kSlabBytes, bias_s, &ready[group]);
    const void* tmap_res = p.res ? &p.tmap_res : nullptr;
    auto epi_tile = [&](int it) {
      const H2Tile t = item_tile(it);
      return EpiTile{t.n, t.p0, t.q0, t.nt * BN};
    };
    EpiTile cur = epi_tile(first);
    if (first < last) epilogue_grouped_begin<BN>(st, gtid, group, cur, tmap_res, p.res_coff);
    int local = 0;
    for (int it = first; it < last; ++it, ++local) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      EpiTile nxt = cur;
      if (it + 1 < last) nxt = epi_tile(it + 1);
      if (warp == 2) H2DBG(48 + local, 2);
      mbar_wait(&tfull[as], aphase);
      if (warp == 2) H2DBG(48 + local, 3);
      tc_fence_after();
      epilogue_tile_grouped<BN>(st, tmem_base + as * Cfg::ACC, quarter, lane, gtid, group, cur,
                                it + 1 < last ? &nxt : nullptr, &p.tmap_out, p.out_coff, tmap_res, p.res_coff, p.slope,
                                [&]() {
                                  tc_fence_before();
                                  __syncwarp();
                                  if (lane == 0) mbar_arrive_leader(&tempty[as]);
                                });
      cur = nxt;
#ifdef M3D_PROBE
      EPI_STAMP(6);
      if (blockIdx.x == 0 && threadIdx.x == 64) ++g_h2_epi_row;
#endif
    }
    if (gtid == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
#ifdef M3D_PROBE
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) s_dbg[47 * 4 + 3] = clock64();
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < 64 * 4) g_h2_dbg[threadIdx.x] = s_dbg[threadIdx.x];
#endif
  cluster_sync_all();  // the peer may still be signalling / reading this CTA's shared memory and TMEM
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2<2 * Cfg::ACC>(tmem_base);
  }
}

template <int BN>
int launch_t(const ConvTmaParams& p, cudaStream_t stream) {
  using Cfg = Halo2Cfg<BN>;
  auto kern = conv_halo2_kernel<BN>;
  set_last_kernel("conv_halo2_kernel<%d>", BN);
  M3D_ONCE_PER_DEVICE_BEGIN
    M3D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  M3D_ONCE_PER_DEVICE_END
  const int items = (p.tiles_w * p.tiles_h * p.N / 2) * p.n_tiles;
  // a CTA pair needs both SMs of a TPC: every CTA running on a reserved SM may strand its sibling
  int clusters = (persistent_sms() - reserved_sms()) / 2;
  if (clusters > items) clusters = items;
  // __cluster_dims__ on the kernel fixes the cluster shape; PDL attribute as everywhere else
  M3D_CUDA_OK(launch_pdl(kern, dim3(2 * clusters), dim3(kThreads2), Cfg::SMEM, stream, p));
  return M3D_OK;
}

}  // namespace

// staged bf16 3x3 stride-1 convs with 128-channel N tiles and an even number of 128-pixel tiles
bool conv_halo2_supported(int BN, long m_tiles, int cout) {
  if (getenv("M3D_NO_PAIR") != nullptr) return false;
  return BN == 128 && m_tiles % 2 == 0 && (cout + BN - 1) / BN * BN <= Halo2Cfg<128>::BIAS_FLOATS;
}

int launch_conv_halo2(const ConvTmaParams& p, int BN, cudaStream_t stream) {
  if (BN != 128) return M3D_ERR_UNSUPPORTED;
  return launch_t<128>(p, stream);
}

}  // namespace m3d

#ifdef M3D_PROBE
extern "C" int m3d_halo2_debug_read(long long* host, int n) {
  return cudaMemcpyFromSymbol(host, m3d::g_h2_dbg, sizeof(long long) * n) == cudaSuccess ? 0 : -1;
}
extern "C" int m3d_halo2_epi_read(long long* host) {
  return cudaMemcpyFromSymbol(host, m3d::g_h2_epi, sizeof(long long) * 64) == cudaSuccess ? 0 : -1;
}
#endif
