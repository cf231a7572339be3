// m3d_conv2d_nhwc: validate the descriptor, pick tile shapes, encode the TMA
// tensor maps and launch the tcgen05 implicit-GEMM kernel.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "common.cuh"
#include "igemm.cuh"

namespace m3d {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

static thread_local char g_last_kernel[128] = "";

void set_last_kernel(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_kernel, sizeof(g_last_kernel), fmt, ap);
  va_end(ap);
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
  }
  return fn;
}

static CUtensorMapSwizzle swizzle_for(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// 2-D bf16 matrix [rows][cols] (cols contiguous), box {box_cols, box_rows}.
int make_tmap_2d(CUtensorMap* map, const void* base, long rows, long cols, int box_cols, int box_rows) {
  auto enc = get_encode();
  M3D_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(box_cols * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  M3D_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d rows=%ld cols=%ld box=%dx%d) failed: %d", rows, cols,
              box_cols, box_rows, static_cast<int>(r));
  return M3D_OK;
}

// Packed bf16 weights [rows][K] as (bk, rows, K/bk): box {bk, box_rows, ksub} lands as ksub consecutive
// [box_rows][bk] k-block tiles.
int make_tmap_b3d(CUtensorMap* map, const void* base, long rows, long cols, int bk, int box_rows, int ksub) {
  auto enc = get_encode();
  M3D_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  M3D_REQUIRE(cols % bk == 0, "weight K=%ld is not a multiple of the k-block %d", cols, bk);
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(bk), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(cols / bk)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(cols) * 2, static_cast<cuuint64_t>(bk) * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(bk), static_cast<cuuint32_t>(box_rows), static_cast<cuuint32_t>(ksub)};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(bk * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  M3D_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights rows=%ld cols=%ld box=%dx%dx%d) failed: %d", rows, cols,
              bk, box_rows, ksub, static_cast<int>(r));
  return M3D_OK;
}

// Packed bf16 weights [rows][K] of a 3x3 conv, K = ((r*3+s)*nchunk + c)*64 + ch, as (64, rows, 3*nchunk, 3):
// box {64, bn, 1, 3} brings the three kernel rows r of one (s, chunk) as consecutive [bn][64] tiles (conv_halo.cu).
int make_tmap_b_halo(CUtensorMap* map, const void* base, long rows, int nchunk, int bn) {
  auto enc = get_encode();
  M3D_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  const cuuint64_t K = static_cast<cuuint64_t>(9) * nchunk * 64;
  cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(3 * nchunk), 3};
  cuuint64_t strides[3] = {K * 2, 128, static_cast<cuuint64_t>(3 * nchunk) * 128};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(bn), 1, 3};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  M3D_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(halo weights rows=%ld nchunk=%d bn=%d) failed: %d", rows, nchunk,
              bn, static_cast<int>(r));
  return M3D_OK;
}

// NHWC bf16 activation as (bk, W, H, N, C/bk): box {bk, tw*stride, th*stride, 1, ksub} lands as ksub
// consecutive [th*tw][bk] channel-chunk tiles of the same window.
int make_tmap_nhwc5(CUtensorMap* map, const void* base, int N, int H, int W, int C, int bk, int tw, int th, int stride,
                    int ksub) {
  auto enc = get_encode();
  M3D_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[5] = {static_cast<cuuint64_t>(bk), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(C / bk)};
  cuuint64_t strides[4] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2,
                           static_cast<cuuint64_t>(H) * W * C * 2, static_cast<cuuint64_t>(bk) * 2};
  cuuint32_t box[5] = {static_cast<cuuint32_t>(bk), static_cast<cuuint32_t>(tw * stride),
                       static_cast<cuuint32_t>(th * stride), 1, static_cast<cuuint32_t>(ksub)};
  cuuint32_t es[5] = {1, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(bk * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  M3D_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(nhwc5 N=%d H=%d W=%d C=%d box=%dx%dx%dx%d) failed: %d", N, H, W,
              C, bk, tw * stride, th * stride, ksub, static_cast<int>(r));
  return M3D_OK;
}

// NHWC bf16 activation [N][H][W][C] as (C, W, H, N); box {bk, tw*stride, th*stride, 1}
// traversed with element strides {1, stride, stride, 1}.
int make_tmap_nhwc(CUtensorMap* map, const void* base, int N, int H, int W, int C, int bk, int tw, int th, int stride) {
  auto enc = get_encode();
  M3D_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2,
                           static_cast<cuuint64_t>(H) * W * C * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(bk), static_cast<cuuint32_t>(tw * stride),
                       static_cast<cuuint32_t>(th * stride), 1};
  cuuint32_t es[4] = {1, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(bk * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  M3D_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(nhwc N=%d H=%d W=%d C=%d box=%dx%dx%d) failed: %d", N, H, W, C,
              bk, tw * stride, th * stride, static_cast<int>(r));
  return M3D_OK;
}

// bf16 NHWC activation as (C, W, H, N); box {box_c, box_w, box_h, 1}, no swizzle, zero fill outside the tensor:
// the staged input window of the fused DCN kernel (plain ld.shared addressing: pixel-major, 2 * box_c bytes per pixel).
int make_tmap_nhwc_plain(CUtensorMap* map, const void* base, int N, int H, int W, int C, int box_c, int box_w, int box_h) {
  auto enc = get_encode();
  M3D_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2,
                           static_cast<cuuint64_t>(H) * W * C * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  M3D_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(nhwc plain N=%d H=%d W=%d C=%d box=%dx%dx%d) failed: %d", N, H, W,
              C, box_c, box_w, box_h, static_cast<int>(r));
  return M3D_OK;
}

// fp32 NHWC tensor as (C, W, H, N); box {box_c, tw, th, 1}, no swizzle (the fused heads' TMA store: box rows are the
// box_c * 4 contiguous bytes of one pixel, clipped at the image edge).
int make_tmap_nhwc_f32(CUtensorMap* map, const void* base, int N, int H, int W, int C, int box_c, int tw, int th,
                       bool swizzle128) {
  auto enc = get_encode();
  M3D_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 4, static_cast<cuuint64_t>(W) * C * 4,
                           static_cast<cuuint64_t>(H) * W * C * 4};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(tw), static_cast<cuuint32_t>(th), 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  M3D_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(nhwc f32 N=%d H=%d W=%d C=%d box=%dx%dx%d) failed: %d", N, H, W, C,
              box_c, tw, th, static_cast<int>(r));
  return M3D_OK;
}

// fp32 NCHW 3-channel image as (W, H, 3, N); box {box_w, box_h, 3, 1}, no swizzle, zero fill outside.
int make_tmap_img(CUtensorMap* map, const void* base, int N, int H, int W, int box_w, int box_h) {
  auto enc = get_encode();
  M3D_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), 3, static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(W) * 4, static_cast<cuuint64_t>(H) * W * 4,
                           static_cast<cuuint64_t>(H) * W * 12};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 3, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  M3D_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(image %dx%dx%d box %dx%d) failed: %d", N, H, W, box_w, box_h,
              static_cast<int>(r));
  return M3D_OK;
}

// Output tile = TH x TW pixels with TH * TW = 128: pick the shape wasting the
// fewest padded pixels (ties go to the squarest, which shares the most halo).
void pick_tile(int P, int Q, int max_tw, int* TW, int* TH) {
  const int cands[5] = {16, 8, 32, 64, 128};
  long best = -1;
  for (int i = 0; i < 5; ++i) {
    const int tw = cands[i], th = kTileM / tw;
    if (tw > max_tw) continue;
    const long tiles = static_cast<long>((Q + tw - 1) / tw) * ((P + th - 1) / th);
    if (best < 0 || tiles < best) {
      best = tiles;
      *TW = tw;
      *TH = th;
    }
  }
}

static int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}

int launch_conv_simt_f32(const m3d_conv_desc* d, int P, int Q, cudaStream_t stream);
bool conv_halo_supported(int BN, int out_dtype, bool staged);
bool conv_halo2_supported(int BN, long m_tiles, int cout);
int launch_conv_halo2(const ConvTmaParams& p, int BN, cudaStream_t stream);
int launch_conv_halo(const ConvTmaParams& p, int BN, int out_dtype, bool staged, cudaStream_t stream);
bool conv_halo_unified(int BN, int out_dtype, bool staged, int nchunk, int n_tiles);

static int pick_bn(int cout, int bk, long m_tiles, bool gather, bool split) {
  if (bk == 16) return cout <= 16 ? 16 : 32;
  if (bk == 32) return cout <= 32 ? 32 : 64;
  if (cout <= 16 && gather) return 16;
  if (cout <= 32) return 32;
  if (cout <= 48) return 48;
  if (cout <= 64) return 64;
  if (cout <= 128) return 128;
  // A deformable gather is paid once per N tile: always take the widest tile
  // (the 3-part fp32 mode only fits 128 columns of operands in shared memory).
  if (gather) return split ? 128 : 256;
  // Plain conv: 256-wide tiles halve the A traffic, but only when there are
  // enough tiles left to occupy every SM.
  return (m_tiles * ((cout + 255) / 256) >= sm_count()) ? 256 : 128;
}

}  // namespace m3d

using namespace m3d;

namespace m3d {
int launch_stem_s2d(const float* image, const void* weight, const float* bias, void* out, int N, int H, int W, float slope,
                    cudaStream_t stream);
}

extern "C" int m3d_stem_conv7x7_s2d(const float* image, const void* weight, const float* bias, void* out, int N, int H,
                                    int W, float slope, m3d_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3D_REQUIRE(image && weight && bias && out, "NULL pointer");
  M3D_REQUIRE(H % 2 == 0 && W % 4 == 0 && H >= 2 && W >= 4, "space-to-depth stem needs H %% 2 == 0, W %% 4 == 0 (got %dx%d)", H, W);
  if (getenv("M3D_STEM_LEGACY") == nullptr) {  // dedicated kernel (stem.cu); the shared gather kernel otherwise
    const int rc = m3d::launch_stem_s2d(image, weight, bias, out, N, H, W, slope, stream);
    if (rc != M3D_ERR_UNSUPPORTED) return rc;
  }
  const int P = H / 2, Q = W / 2;
  int TW = 16, TH = 8;
  pick_tile(P, Q, 64, &TW, &TH);  // image tile (2TW+8) x (2TH+6) x 3 fp32 must fit the producer scratch
  ConvGatherParams p;
  memset(&p, 0, sizeof(p));
  int rc = make_tmap_2d(&p.tmap_b, weight, 64, 192, 64, 64);
  if (rc != M3D_OK) return rc;
  rc = make_tmap_nhwc(&p.tmap_out, out, N, P, Q, 64, 64, TW, TH, 1);
  if (rc != M3D_OK) return rc;
  rc = make_tmap_img(&p.tmap_img, image, N, H, W, 2 * TW + 8, 2 * TH + 6);
  if (rc != M3D_OK) return rc;
  p.num_inputs = 1;
  p.chunks[0] = 3;  // one k-block per image channel
  p.stem_img = image;
  p.H = H, p.W = W;
  p.R = 1, p.S = 1, p.stride = 1, p.pad = 0, p.dil = 1;
  p.N = N, p.P = P, p.Q = Q;
  p.TW = TW, p.TH = TH, p.tiles_w = (Q + TW - 1) / TW, p.tiles_h = (P + TH - 1) / TH;
  p.Cout = 64, p.n_tiles = 1;
  p.out = out, p.out_cstride = 64;
  p.bias = bias;
  p.slope = slope;
  p.total_tiles = p.tiles_w * p.tiles_h * N;
  return launch_conv_gather(p, 64, DT_BF16, DT_BF16, true, stream);
}

namespace m3d {
static int g_sm_limit = 0;
int persistent_sms() {
  const int all = sm_count();
  return (g_sm_limit > 0 && g_sm_limit < all) ? g_sm_limit : all;
}
int reserved_sms() { return sm_count() - persistent_sms(); }
}  // namespace m3d

namespace m3d {
static bool g_pdl_on = true;
bool pdl_switch() { return g_pdl_on; }
}  // namespace m3d

extern "C" int m3d_set_pdl(int on) {
  m3d::g_pdl_on = on != 0;
  return M3D_OK;
}

extern "C" int m3d_set_sm_limit(int sms) {
  M3D_REQUIRE(sms >= 0, "sm limit must be >= 0 (0 = the whole device)");
  m3d::g_sm_limit = sms;
  return M3D_OK;
}

extern "C" const char* m3d_last_error(void) { return g_last_error; }
extern "C" const char* m3d_last_kernel(void) { return g_last_kernel; }
extern "C" int m3d_version(void) { return 100; }
extern "C" size_t m3d_conv_desc_size(void) { return sizeof(m3d_conv_desc); }

extern "C" int m3d_conv2d_nhwc(const m3d_conv_desc* d, m3d_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  M3D_REQUIRE(d != nullptr, "desc is NULL");
  M3D_REQUIRE(d->num_inputs >= 1 && d->num_inputs <= M3D_MAX_CONCAT, "num_inputs=%d out of range", d->num_inputs);
  M3D_REQUIRE(d->R >= 1 && d->S >= 1 && d->stride >= 1 && d->dil >= 1 && d->pad >= 0, "bad kernel geometry");
  M3D_REQUIRE(d->N >= 1 && d->H >= 1 && d->W >= 1 && d->Cout >= 1, "bad tensor geometry");
  M3D_REQUIRE((d->weight != nullptr || d->weight_f32 != nullptr) && d->out != nullptr, "weight/out is NULL");
  const int groups = d->groups < 1 ? 1 : d->groups;
  const int P = d->out_h > 0 ? d->out_h : (d->H + 2 * d->pad - (d->dil * (d->R - 1) + 1)) / d->stride + 1;  // dcn_v2_cuda.c:40-41
  const int Q = d->out_w > 0 ? d->out_w : (d->W + 2 * d->pad - (d->dil * (d->S - 1) + 1)) / d->stride + 1;
  M3D_REQUIRE(P >= 1 && Q >= 1, "empty output %dx%d", P, Q);
  const bool split = d->act_dtype == M3D_F32;
  const bool gather = split || d->om != nullptr || d->force_gather;
  if (split) {
    M3D_REQUIRE(d->out_dtype == M3D_F32, "fp32 activations need fp32 output");
    if (d->weight_f32 != nullptr) {
      M3D_REQUIRE(groups == 1, "fp32 path does not batch groups");
      return launch_conv_simt_f32(d, P, Q, stream);
    }
    M3D_REQUIRE(d->weight != nullptr && d->weight_mid != nullptr && d->weight_lo != nullptr,
                "fp32 activations need weight_f32 (reference accuracy) or weight + weight_mid + weight_lo (bf16x3)");
  }
  long ktot = 0;
  int bk = 64;
  for (int i = 0; i < d->num_inputs; ++i) {
    M3D_REQUIRE(d->in[i] != nullptr, "input %d is NULL", i);
    M3D_REQUIRE(d->in_c[i] % 16 == 0 && d->in_c[i] > 0, "input %d: %d channels (need a multiple of 16)", i, d->in_c[i]);
    if (d->in_c[i] % 64 != 0) {
      const int need = d->in_c[i] % 32 == 0 ? 32 : 16;
      if (need < bk) bk = need;
    }
    ktot += static_cast<long>(d->R) * d->S * d->in_c[i];
  }
  if (gather) {
    M3D_REQUIRE(bk == 64, "gather/DCN path needs input channels in multiples of 64");
    M3D_REQUIRE(groups == 1, "gather/DCN path does not batch groups");
    M3D_REQUIRE(d->R * d->S <= 9 || d->om == nullptr, "deformable kernels above 3x3 unsupported");
    if (!split) {
      M3D_REQUIRE(d->num_inputs == 1 && d->R * d->S <= 9, "bf16 gather path: one input, at most 9 taps");
    }
  }
  int TW = 16, TH = 8;
  pick_tile(P, Q, 256 / d->stride, &TW, &TH);
  const int tiles_w = (Q + TW - 1) / TW, tiles_h = (P + TH - 1) / TH;
  const long m_tiles = static_cast<long>(tiles_w) * tiles_h * d->N;
  int BN = pick_bn(d->Cout, bk, m_tiles * groups, gather, split);
  // (A 256-wide CTA-pair tile for the level-4 256 -> 256 3x3 convs was built and measured in round 2: 27 us per layer
  // against 25 us for conv_tma_kernel<256,64,1> on 120 CTAs -- 60 pair items cannot fill 70 clusters -- and removed.)
  const int n_tiles = (d->Cout + BN - 1) / BN;
  const long total_tiles = m_tiles * n_tiles * groups;
  M3D_REQUIRE(total_tiles < (1L << 30), "too many tiles");

  // bf16 tiles of whole 64-channel slabs leave through shared memory + TMA stores
  const bool staged = d->act_dtype == M3D_BF16 && d->out_dtype == M3D_BF16 && (bk == 64 || (bk == 32 && BN == 64 && !gather)) &&
                      BN % 64 == 0 &&
                      d->Cout % 64 == 0 && d->out_cstride % 8 == 0 && d->out_coff % 8 == 0 && d->out_goff % 8 == 0 &&
                      (d->res == nullptr || (d->res_cstride % 8 == 0 && d->res_coff % 8 == 0 && d->res_goff % 8 == 0));
  // fp32 output of a bf16 1x1 layer with many channels (the class logits): 32-column slabs through the same staging.
  // A partial last slab is clipped by the TMA store only at the tensor's channel extent.
  const bool staged_f32 = !gather && d->act_dtype == M3D_BF16 && d->out_dtype == M3D_F32 && bk == 64 && BN == 256 &&
                          groups == 1 && d->res == nullptr && d->Cout > 64 && d->out_cstride % 4 == 0 &&
                          d->out_coff % 4 == 0 && (d->Cout % 32 == 0 || d->out_coff + d->Cout == d->out_cstride) &&
                          !(getenv("M3D_F32_STAGED") && atoi(getenv("M3D_F32_STAGED")) == 0);
  if (!gather) {
    ConvTmaParams p;
    memset(&p, 0, sizeof(p));
    // 3x3 stride-1 convs: one (TH+2)-row window per kernel column serves its three taps (conv_halo.cu)
    const bool halo = d->R == 3 && d->S == 3 && d->stride == 1 && d->dil == 1 && d->pad == 1 && d->num_inputs == 1 &&
                      groups == 1 && bk == 64 && d->out_h <= 0 && d->out_w <= 0 && TW <= 16 &&
                      d->in_cstride[0] % 8 == 0 && d->in_coff[0] % 8 == 0 &&
                      conv_halo_supported(BN, d->out_dtype == M3D_F32 ? DT_F32 : DT_BF16, staged) &&
                      getenv("M3D_NO_HALO") == nullptr;
    // resident-weight variants read all nine taps from one (TH+2) x (TW+2) window: 16 x 8 pixel tiles
    const bool uni = halo && conv_halo_unified(BN, d->out_dtype == M3D_F32 ? DT_F32 : DT_BF16, staged, d->in_c[0] / 64, n_tiles);
    if (uni) TW = 8, TH = 16;
    const int tiles_w = (Q + TW - 1) / TW, tiles_h = (P + TH - 1) / TH;  // (shadow the estimates above)
    const long m_tiles = static_cast<long>(tiles_w) * tiles_h * d->N;
    const long total_tiles = m_tiles * n_tiles * groups;
    if (staged_f32) {
      int rc2 = make_tmap_nhwc_f32(&p.tmap_out, d->out, d->N, P, Q, d->out_cstride, 32, TW, TH, true);
      if (rc2 != M3D_OK) return rc2;
    }
    if (staged) {
      int rc2 = make_tmap_nhwc(&p.tmap_out, d->out, d->N, P, Q, d->out_cstride, 64, TW, TH, 1);
      if (rc2 != M3D_OK) return rc2;
      if (d->res != nullptr) {
        rc2 = make_tmap_nhwc(&p.tmap_res, d->res, d->N, P, Q, d->res_cstride, 64, TW, TH, 1);
        if (rc2 != M3D_OK) return rc2;
      }
    }
    if (halo) {
      const bool pair = staged && conv_halo2_supported(BN, m_tiles, d->Cout);  // CTA pairs: each CTA loads half of the weight rows
      int rc = make_tmap_nhwc(&p.tmap_a[0], d->in[0], d->N, d->H, d->W, d->in_cstride[0], 64, uni ? TW + 2 : TW, TH + 2, 1);
      if (rc != M3D_OK) return rc;
      rc = make_tmap_b_halo(&p.tmap_b, d->weight, d->weight_rows, d->in_c[0] / 64, pair ? BN / 2 : BN);
      if (rc != M3D_OK) return rc;
      p.num_inputs = 1;
      p.chunks[0] = d->in_c[0] / 64;
      p.a_coff[0] = d->in_coff[0];
      p.R = 3, p.S = 3, p.stride = 1, p.pad = 1, p.dil = 1;
      p.N = d->N, p.P = P, p.Q = Q;
      p.TW = TW, p.TH = TH, p.tiles_w = tiles_w, p.tiles_h = tiles_h;
      p.Cout = d->Cout, p.n_tiles = n_tiles, p.groups = 1;
      p.out = d->out, p.out_cstride = d->out_cstride, p.out_coff = d->out_coff;
      p.bias = d->bias;
      p.res = d->res, p.res_cstride = d->res_cstride, p.res_coff = d->res_coff;
      p.slope = d->slope;
      p.total_tiles = static_cast<int>(total_tiles);
      if (uni && d->in_c[0] == 64 && getenv("M3D_NO_KSKIP") == nullptr) {  // K = 9 taps x 64 = 36 slices of 16
        const unsigned long long all = (1ull << 36) - 1;
        p.kzero = d->k16_zero[0] & all;
        if (p.kzero == all) p.kzero = 0;  // nothing left to initialise the accumulator with: run it dense
      }
      if (pair) return launch_conv_halo2(p, BN, stream);
      return launch_conv_halo(p, BN, d->out_dtype == M3D_F32 ? DT_F32 : DT_BF16, staged, stream);
    }
    // k-blocks per pipeline stage: enough tensor-pipe clocks per stage (N/2 per k16) to cover the
    // issue cost of the stage's barrier hand-shakes and TMA instructions
    const long total_kb = ktot / bk;
    int ksub = 1;
    if (staged && bk == 64) {
      for (int k = (BN >= 256 ? 1 : (BN >= 128 ? 2 : 4)); k > 1; --k)
        if (total_kb % k == 0) {
          ksub = k;
          break;
        }
      const char* e = getenv("M3D_KSUB");  // development override
      if (e != nullptr && atoi(e) >= 1 && total_kb % atoi(e) == 0 && (atoi(e) <= 2 || BN == 64)) ksub = atoi(e);
    }
    bool wide = ksub > 1;
    for (int i = 0; i < d->num_inputs; ++i) {
      M3D_REQUIRE(d->in_cstride[i] % 8 == 0, "input %d: channel stride %d not a multiple of 8", i, d->in_cstride[i]);
      if ((d->in_c[i] / bk) % ksub != 0 || d->in_cstride[i] % bk != 0 || d->in_coff[i] % bk != 0 ||
          d->in_goff[i] % bk != 0)
        wide = false;
    }
    for (int i = 0; i < d->num_inputs; ++i) {
      int rc = wide ? make_tmap_nhwc5(&p.tmap_a[i], d->in[i], d->N, d->H, d->W, d->in_cstride[i], bk, TW, TH, d->stride,
                                      ksub)
                    : make_tmap_nhwc(&p.tmap_a[i], d->in[i], d->N, d->H, d->W, d->in_cstride[i], bk, TW, TH, d->stride);
      if (rc != M3D_OK) return rc;
      p.chunks[i] = d->in_c[i] / bk;
      p.a_coff[i] = d->in_coff[i];
      p.a_goff[i] = d->in_goff[i];
    }
    p.a_wide = wide ? 1 : 0;
    int rc = make_tmap_b3d(&p.tmap_b, d->weight, d->weight_rows, ktot, bk, BN, ksub);
    if (rc != M3D_OK) return rc;
    p.num_inputs = d->num_inputs;
    p.R = d->R, p.S = d->S, p.stride = d->stride, p.pad = d->pad, p.dil = d->dil;
    p.N = d->N, p.P = P, p.Q = Q;
    p.TW = TW, p.TH = TH, p.tiles_w = tiles_w, p.tiles_h = tiles_h;
    p.Cout = d->Cout, p.n_tiles = n_tiles, p.groups = groups, p.b_goff = d->weight_goff;
    p.out = d->out, p.out_cstride = d->out_cstride, p.out_coff = d->out_coff, p.out_goff = d->out_goff;
    p.bias = d->bias, p.bias_goff = d->bias_goff;
    p.res = d->res, p.res_cstride = d->res_cstride, p.res_coff = d->res_coff, p.res_goff = d->res_goff;
    p.slope = d->slope;
    p.total_tiles = static_cast<int>(total_tiles);
    rc = launch_conv_tma(p, BN, bk, ksub, d->out_dtype, staged || staged_f32, stream);
    if (rc == M3D_ERR_UNSUPPORTED) set_last_error("no TMA conv kernel for BN=%d BK=%d", BN, bk);
    return rc;
  }

  ConvGatherParams p;
  memset(&p, 0, sizeof(p));
  int rc = make_tmap_2d(&p.tmap_b, d->weight, d->weight_rows, ktot, 64, BN);
  if (rc != M3D_OK) return rc;
  if (staged) {
    rc = make_tmap_nhwc(&p.tmap_out, d->out, d->N, P, Q, d->out_cstride, 64, TW, TH, 1);
    if (rc != M3D_OK) return rc;
    if (d->res != nullptr) {
      rc = make_tmap_nhwc(&p.tmap_res, d->res, d->N, P, Q, d->res_cstride, 64, TW, TH, 1);
      if (rc != M3D_OK) return rc;
    }
  }
  if (split) {
    rc = make_tmap_2d(&p.tmap_b_mid, d->weight_mid, d->weight_rows, ktot, 64, BN);
    if (rc != M3D_OK) return rc;
    rc = make_tmap_2d(&p.tmap_b_lo, d->weight_lo, d->weight_rows, ktot, 64, BN);
    if (rc != M3D_OK) return rc;
  }
  p.num_inputs = d->num_inputs;
  for (int i = 0; i < d->num_inputs; ++i) {
    p.in[i] = d->in[i];
    p.in_cstride[i] = d->in_cstride[i];
    p.in_coff[i] = d->in_coff[i];
    p.chunks[i] = d->in_c[i] / 64;
    M3D_REQUIRE((d->in_cstride[i] % (split ? 4 : 8)) == 0 && (d->in_coff[i] % (split ? 4 : 8)) == 0,
                "input %d: channel stride/offset must keep 16-byte alignment", i);
  }
  p.H = d->H, p.W = d->W;
  p.R = d->R, p.S = d->S, p.stride = d->stride, p.pad = d->pad, p.dil = d->dil;
  p.N = d->N, p.P = P, p.Q = Q;
  p.TW = TW, p.TH = TH, p.tiles_w = tiles_w, p.tiles_h = tiles_h;
  p.Cout = d->Cout, p.n_tiles = n_tiles;
  p.om = d->om, p.om_cstride = d->om_cstride, p.sigmoid_mask = d->sigmoid_mask;
  if (d->om != nullptr) M3D_REQUIRE(d->om_cstride >= 3 * d->R * d->S, "om_cstride %d < 3*R*S", d->om_cstride);
  p.out = d->out, p.out_cstride = d->out_cstride, p.out_coff = d->out_coff;
  p.bias = d->bias;
  p.res = d->res, p.res_cstride = d->res_cstride, p.res_coff = d->res_coff;
  p.slope = d->slope;
  p.total_tiles = static_cast<int>(total_tiles);
  rc = launch_conv_gather(p, BN, d->act_dtype == M3D_F32 ? DT_F32 : DT_BF16, d->out_dtype == M3D_F32 ? DT_F32 : DT_BF16,
                          staged, stream);
  if (rc == M3D_ERR_UNSUPPORTED) set_last_error("no gather conv kernel for BN=%d", BN);
  return rc;
}
