// Anchor <-> ground-truth target assignment on the device: the reference computes it per image
// with numpy on the data-loader's CPU workers (lib/rpn_util.py:430-532 compute_targets, lib/dataloader.py:1014-1144
// _targets: an [M, G] IoU matrix, argmax both ways, index-set arithmetic with np.unique / np.setdiff1d) and ships five
// [M]-sized arrays per image to the GPU.  Here one batch is three small launches over the M = A*H*W anchors of every
// image; the anchors' boxes are never materialised (rebuilt from (anchor, row, column) as locate_anchors does).
//
//   pass 1  per anchor: IoU with every valid box -> best overlap + its box (first maximum, like np.argmax);
//           per box: atomicMax of the IoU bit pattern (overlaps are >= 0, so the order of the doubles is the order
//           of their bits)
//   pass 2  per anchor: every box whose best overlap this anchor attains -> atomicMin of the anchor index
//           (np.argmax over the anchor axis returns the FIRST maximum)
//   pass 3  per anchor: foreground = overlap >= fg_thresh or best anchor of a box whose best overlap >= best_thresh;
//           ignored = covered by an ignore region (intersection / anchor area >= ign_thresh); background = overlap in
//           [bg_lo, bg_hi) and neither of the others; labels, 2D / 3D regression targets, normalisation.
//
// Arithmetic follows numpy's promotion rules in the reference: anchor boxes are float32 (their widths, centres and
// areas are float32 expressions), ground-truth boxes float64, everything that mixes the two float64, results rounded
// to float32 when stored, (x - mean) / std in float32.  Labels are bit-exact, targets equal to the last bit.
#include <cstdint>

#include "common.cuh"

namespace m3d {
namespace {

struct TargetParams {
  const double* gts_val;   // [B, Gmax, 4] x1, y1, x2, y2
  const double* gts_3d;    // [B, Gmax, 7] cx, cy, z, w, h, l, rotY
  const int* box_lbls;     // [B, Gmax] class index >= 1
  const int* n_val;        // [B]
  const double* gts_ign;   // [B, Imax, 4]
  const int* n_ign;        // [B]
  const float* anchors;    // [A, 9]
  int B, Gmax, Imax, A, H, W;
  double stride;
  double fg_thresh, ign_thresh, bg_lo, bg_hi, best_thresh;
  float means[11], stds[11];
  // workspace
  unsigned long long* best_bits;  // [B, Gmax] IoU bit pattern of every box's best anchor
  int* best_idx;                  // [B, Gmax] first anchor attaining it
  // outputs
  unsigned char *labels_fg, *labels_bg, *labels_ign;  // [B, M]
  long long* labels;                                    // [B, M]: 0 background, class, 3000 ignored
  float* bbox_2d;                                       // [B, M, 4]
  float* bbox_3d;                                       // [B, M, 7]
  unsigned char* any_val;                               // [B]
};

struct Roi {
  float x1, y1, x2, y2;
  int a;
};

// locate_anchors (lib/rpn_util.py:1345-1386): float64 shift + float32 anchor, stored as float32
__device__ __forceinline__ Roi make_roi(const TargetParams& p, int m) {
  const int HW = p.H * p.W;
  const int a = m / HW, h = (m / p.W) % p.H, w = m % p.W;
  const float* an = p.anchors + a * 9;
  Roi r;
  r.x1 = static_cast<float>(static_cast<double>(w) * p.stride + static_cast<double>(an[0]));
  r.y1 = static_cast<float>(static_cast<double>(h) * p.stride + static_cast<double>(an[1]));
  r.x2 = static_cast<float>(static_cast<double>(w) * p.stride + static_cast<double>(an[2]));
  r.y2 = static_cast<float>(static_cast<double>(h) * p.stride + static_cast<double>(an[3]));
  r.a = a;
  return r;
}

// intersect (lib/core.py:266-282), numpy branch: clip(min(x2) - max(x1), 0) per axis, no +1
__device__ __forceinline__ double inter_area(const Roi& r, const double* g) {
  const double iw = fmin(static_cast<double>(r.x2), g[2]) - fmax(static_cast<double>(r.x1), g[0]);
  const double ih = fmin(static_cast<double>(r.y2), g[3]) - fmax(static_cast<double>(r.y1), g[1]);
  return fmax(iw, 0.0) * fmax(ih, 0.0);
}
// iou (lib/core.py:341-372), numpy 'combinations': the anchor's area is a float32 product
__device__ __forceinline__ double iou_val(const Roi& r, const double* g) {
  const double inter = inter_area(r, g);
  const float area_a = (r.x2 - r.x1) * (r.y2 - r.y1);
  const double area_b = (g[2] - g[0]) * (g[3] - g[1]);
  return inter / (static_cast<double>(area_a) + area_b - inter);
}

__global__ void targets_init_kernel(const TargetParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p.B * p.Gmax) {
    p.best_bits[i] = 0ull;
    p.best_idx[i] = 0x7fffffff;
  }
  if (i < p.B) p.any_val[i] = p.n_val[i] > 0 ? 1 : 0;
}

__global__ void __launch_bounds__(256) targets_pass1_kernel(const TargetParams p) {
  const int b = blockIdx.y, M = p.A * p.H * p.W;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int G = p.n_val[b];
  if (m >= M || G <= 0) return;
  const Roi r = make_roi(p, m);
  for (int g = 0; g < G; ++g) {
    const double v = iou_val(r, p.gts_val + (static_cast<long>(b) * p.Gmax + g) * 4);
    const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(v));
    if (v > 0.0) atomicMax(&p.best_bits[b * p.Gmax + g], bits);  // (a box no anchor touches keeps 0: argmax = anchor 0)
  }
}

__global__ void __launch_bounds__(256) targets_pass2_kernel(const TargetParams p) {
  const int b = blockIdx.y, M = p.A * p.H * p.W;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int G = p.n_val[b];
  if (m >= M || G <= 0) return;
  const Roi r = make_roi(p, m);
  for (int g = 0; g < G; ++g) {
    const double v = iou_val(r, p.gts_val + (static_cast<long>(b) * p.Gmax + g) * 4);
    if (static_cast<unsigned long long>(__double_as_longlong(v)) == p.best_bits[b * p.Gmax + g]) atomicMin(&p.best_idx[b * p.Gmax + g], m);
  }
}

__global__ void __launch_bounds__(256) targets_pass3_kernel(const TargetParams p) {
  const int b = blockIdx.y, M = p.A * p.H * p.W;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const long row = static_cast<long>(b) * M + m;
  const int G = p.n_val[b], I = p.n_ign[b];
  float t2[4] = {0.f, 0.f, 0.f, 0.f}, t3[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int code = -1;  // transforms[:, 4]: > 0 class of a foreground anchor, -1 background, 0 ignored
  if (G > 0) {    // (_targets calls compute_targets only when the image has a valid box; else everything is background)
    const Roi r = make_roi(p, m);
    double ols_max = 0.0;
    int target = 0;
    bool is_best = false;
    for (int g = 0; g < G; ++g) {
      const double v = iou_val(r, p.gts_val + (static_cast<long>(b) * p.Gmax + g) * 4);
      if (g == 0 || v > ols_max) ols_max = v, target = g;
      if (p.best_idx[b * p.Gmax + g] == m && __longlong_as_double(static_cast<long long>(p.best_bits[b * p.Gmax + g])) >= p.best_thresh)
        is_best = true;
    }
    double ign_max = 0.0;  // iou_ign (lib/core.py:402-430): intersection over the ANCHOR's area
    const float area_a = (r.x2 - r.x1) * (r.y2 - r.y1);
    for (int g = 0; g < I; ++g) {
      const double v = inter_area(r, p.gts_ign + (static_cast<long>(b) * p.Imax + g) * 4) / static_cast<double>(area_a);
      if (g == 0 || v > ign_max) ign_max = v;
    }
    const bool fg = ols_max >= p.fg_thresh || is_best;
    const bool ign = ign_max >= p.ign_thresh;
    const bool bg = ols_max >= p.bg_lo && ols_max < p.bg_hi && !ign && !fg;
    code = fg ? p.box_lbls[b * p.Gmax + target] : (bg ? -1 : 0);
    if (fg) {
      // bbox_transform (lib/rpn_util.py:1101-1134): anchor side in float32, box side in float64
      const double* gt = p.gts_val + (static_cast<long>(b) * p.Gmax + target) * 4;
      const float ew = r.x2 - r.x1 + 1.0f, eh = r.y2 - r.y1 + 1.0f;
      const float ecx = r.x1 + 0.5f * (ew - 1.0f), ecy = r.y1 + 0.5f * (eh - 1.0f);
      const double gw = gt[2] - gt[0] + 1.0, gh = gt[3] - gt[1] + 1.0;
      const double gcx = gt[0] + 0.5 * (gw - 1.0), gcy = gt[1] + 0.5 * (gh - 1.0);
      t2[0] = static_cast<float>((gcx - static_cast<double>(ecx)) / static_cast<double>(ew));
      t2[1] = static_cast<float>((gcy - static_cast<double>(ecy)) / static_cast<double>(eh));
      t2[2] = static_cast<float>(log(gw / static_cast<double>(ew)));
      t2[3] = static_cast<float>(log(gh / static_cast<double>(eh)));
      // bbox_transform_3d (lib/rpn_util.py:1059-1098) against the anchor's 3D priors anchors[a, 4:9]
      const double* g3 = p.gts_3d + (static_cast<long>(b) * p.Gmax + target) * 7;
      const float* an = p.anchors + r.a * 9;
      t3[0] = static_cast<float>((g3[0] - static_cast<double>(ecx)) / static_cast<double>(ew));
      t3[1] = static_cast<float>((g3[1] - static_cast<double>(ecy)) / static_cast<double>(eh));
      t3[2] = static_cast<float>(g3[2] - static_cast<double>(an[4]));
      t3[3] = static_cast<float>(log(g3[3] / static_cast<double>(an[5])));
      t3[4] = static_cast<float>(log(g3[4] / static_cast<double>(an[6])));
      t3[5] = static_cast<float>(log(g3[5] / static_cast<double>(an[7])));
      t3[6] = static_cast<float>(g3[6] - static_cast<double>(an[8]));
    }
    // lib/dataloader.py:1109-1113: (t - mean) / std, two float32 operations, on EVERY row
#pragma unroll
    for (int j = 0; j < 4; ++j) t2[j] = __fdiv_rn(__fsub_rn(t2[j], p.means[j]), p.stds[j]);
#pragma unroll
    for (int j = 0; j < 7; ++j) t3[j] = __fdiv_rn(__fsub_rn(t3[j], p.means[4 + j]), p.stds[4 + j]);
  }
  p.labels_fg[row] = code > 0;
  p.labels_bg[row] = code < 0;
  p.labels_ign[row] = code == 0;
  p.labels[row] = code > 0 ? code : (code == 0 ? 3000 : 0);
  *reinterpret_cast<float4*>(p.bbox_2d + row * 4) = make_float4(t2[0], t2[1], t2[2], t2[3]);
#pragma unroll
  for (int j = 0; j < 7; ++j) p.bbox_3d[row * 7 + j] = t3[j];
}

}  // namespace
}  // namespace m3d

using namespace m3d;

extern "C" size_t m3d_compute_targets_workspace(int batch, int max_gts) {
  return static_cast<size_t>(batch) * (max_gts > 0 ? max_gts : 1) * (sizeof(unsigned long long) + sizeof(int)) + 16;
}

extern "C" int m3d_compute_targets(const double* gts_val, const double* gts_3d, const int* box_lbls, const int* n_val,
                                   int max_gts, const double* gts_ign, const int* n_ign, int max_ign, const float* anchors,
                                   int batch, int A, int H, int W, float feat_stride, double fg_thresh, double ign_thresh,
                                   double bg_thresh_lo, double bg_thresh_hi, double best_thresh, const float* means11,
                                   const float* stds11, unsigned char* labels_fg, unsigned char* labels_bg,
                                   unsigned char* labels_ign, long long* labels, float* bbox_2d, float* bbox_3d,
                                   unsigned char* any_val, void* workspace, size_t workspace_bytes, m3d_stream_t stream) {
  M3D_REQUIRE(n_val && n_ign && anchors && means11 && stds11 && labels_fg && labels_bg && labels_ign && labels && bbox_2d &&
                  bbox_3d && any_val,
              "NULL pointer");
  M3D_REQUIRE(batch >= 1 && A >= 1 && H >= 1 && W >= 1 && max_gts >= 0 && max_ign >= 0, "bad geometry");
  M3D_REQUIRE(max_gts == 0 || (gts_val && gts_3d && box_lbls), "NULL ground-truth arrays");
  M3D_REQUIRE(max_ign == 0 || gts_ign, "NULL ignore-region array");
  M3D_REQUIRE(static_cast<long>(A) * H * W < (1L << 31), "too many anchors");
  if (workspace == nullptr || workspace_bytes < m3d_compute_targets_workspace(batch, max_gts) ||
      (reinterpret_cast<uintptr_t>(workspace) & 7) != 0) {
    set_last_error("m3d_compute_targets: 8-byte aligned workspace of %zu bytes needed, got %zu",
                   m3d_compute_targets_workspace(batch, max_gts), workspace_bytes);
    return M3D_ERR_WORKSPACE;
  }
  TargetParams p;
  p.gts_val = gts_val, p.gts_3d = gts_3d, p.box_lbls = box_lbls, p.n_val = n_val, p.gts_ign = gts_ign, p.n_ign = n_ign;
  p.anchors = anchors;
  p.B = batch, p.Gmax = max_gts > 0 ? max_gts : 1, p.Imax = max_ign > 0 ? max_ign : 1, p.A = A, p.H = H, p.W = W;
  p.stride = static_cast<double>(feat_stride);
  p.fg_thresh = fg_thresh, p.ign_thresh = ign_thresh, p.bg_lo = bg_thresh_lo, p.bg_hi = bg_thresh_hi, p.best_thresh = best_thresh;
  for (int i = 0; i < 11; ++i) p.means[i] = means11[i], p.stds[i] = stds11[i];
  p.best_bits = static_cast<unsigned long long*>(workspace);
  p.best_idx = reinterpret_cast<int*>(p.best_bits + static_cast<size_t>(batch) * p.Gmax);
  p.labels_fg = labels_fg, p.labels_bg = labels_bg, p.labels_ign = labels_ign, p.labels = labels;
  p.bbox_2d = bbox_2d, p.bbox_3d = bbox_3d, p.any_val = any_val;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int M = A * H * W;
  const dim3 grid((M + 255) / 256, batch);
  targets_init_kernel<<<(batch * p.Gmax + batch + 255) / 256, 256, 0, st>>>(p);
  M3D_CUDA_OK(cudaGetLastError());
  if (max_gts > 0) {
    targets_pass1_kernel<<<grid, 256, 0, st>>>(p);
    M3D_CUDA_OK(cudaGetLastError());
    targets_pass2_kernel<<<grid, 256, 0, st>>>(p);
    M3D_CUDA_OK(cudaGetLastError());
  }
  targets_pass3_kernel<<<grid, 256, 0, st>>>(p);
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}
