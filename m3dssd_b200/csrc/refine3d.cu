// Post-NMS 3D refinement: the per-box loop of test_kitti_3d
// (lib/rpn_util.py:1801-1852) with hill_climb (:652-708), test_projection (:2015-2050), project_3d (:921-970) and
// convertAlpha2Rot / convertRot2Alpha (lib/util.py:516-535).  The reference runs it box by box in Python / numpy
// float64 (30-60 projections per box); here one thread per kept box does the same float64 arithmetic on the device,
// straight from the [B, max_out, 14] rows the NMS gather leaves in HBM, and writes the 14 numbers of the KITTI
// result line (class index, alpha, x1, y1, x2, y2, h3d, w3d, l3d, x3d, y3d, z3d, ry3d, score).
#include <cuda_runtime.h>

#include "common.cuh"

namespace m3d {
namespace {

constexpr double kPi = 3.141592653589793;

struct Box3 {
  double w3d, h3d, l3d;
};

// project_3d + the extent / validity part of test_projection: returns ol = -L1 distance between the 2D box and the
// extent of the projected 3D box; *invalid = any corner at or behind the camera.
__device__ double test_projection(const double* __restrict__ p2, const double* __restrict__ p2i, double bx, double by,
                                  double bw, double bh, double cx, double cy, double z, const Box3& b, double rot,
                                  bool* invalid) {
  const double x2 = bx + bw - 1, y2 = by + bh - 1;
  const double v[4] = {cx * z, cy * z, z, 1.0};
  double c3[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) c3[r] = ((p2i[r * 4] * v[0] + p2i[r * 4 + 1] * v[1]) + p2i[r * 4 + 2] * v[2]) + p2i[r * 4 + 3] * v[3];
  const double c = cos(rot), s = sin(rot);
  const double xs[8] = {0, b.l3d, b.l3d, b.l3d, b.l3d, 0, 0, 0};
  const double ys[8] = {0, 0, b.h3d, b.h3d, 0, 0, b.h3d, b.h3d};
  const double zs[8] = {0, 0, 0, b.w3d, b.w3d, b.w3d, b.w3d, 0};
  double xmin = 1e300, ymin = 1e300, xmax = -1e300, ymax = -1e300;
  bool inv = false;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double xc = xs[k] - b.l3d / 2, yc = ys[k] - b.h3d / 2, zc = zs[k] - b.w3d / 2;
    // R = [[c, 0, s], [0, 1, 0], [-s, 0, c]]  (numpy dot: sum in index order, the zero terms are exact)
    const double X = ((c * xc + 0.0 * yc) + s * zc) + c3[0];
    const double Y = ((0.0 * xc + 1.0 * yc) + 0.0 * zc) + c3[1];
    const double Z = ((-s * xc + 0.0 * yc) + c * zc) + c3[2];
    inv = inv || (Z <= 0.0);
    const double u = ((p2[0] * X + p2[1] * Y) + p2[2] * Z) + p2[3] * 1.0;
    const double w = ((p2[4] * X + p2[5] * Y) + p2[6] * Z) + p2[7] * 1.0;
    const double q = ((p2[8] * X + p2[9] * Y) + p2[10] * Z) + p2[11] * 1.0;
    const double px = u / q, py = w / q;
    xmin = fmin(xmin, px), xmax = fmax(xmax, px);
    ymin = fmin(ymin, py), ymax = fmax(ymax, py);
  }
  *invalid = inv;
  return -(((fabs(bx - xmin) + fabs(by - ymin)) + fabs(x2 - xmax)) + fabs(y2 - ymax));
}

__device__ double wrap_pi(double a) {
  while (a > kPi) a -= kPi * 2;
  while (a < -kPi) a += kPi * 2;
  return a;
}

__global__ void refine3d_kernel(const float* __restrict__ kept, const int* __restrict__ num_keep, int B, int max_out,
                                int row_len, const double* __restrict__ p2s, const double* __restrict__ p2invs,
                                float score_thresh, int hill_climbing, double step_r_init, double r_lim,
                                double* __restrict__ out, int* __restrict__ valid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * max_out) return;
  const int n = i / max_out, r = i - n * max_out;
  double* o = out + static_cast<long>(i) * 14;
  const float* box = kept + static_cast<long>(i) * row_len;
  const bool live = r < num_keep[n] && box[4] >= score_thresh;
  valid[i] = live ? 1 : 0;
  if (!live) {
    for (int k = 0; k < 14; ++k) o[k] = 0.0;
    return;
  }
  const double* p2 = p2s + n * 16;
  const double* p2i = p2invs + n * 16;
  const double x1 = box[0], y1 = box[1], x2 = box[2], y2 = box[3];
  const double bw = x2 - x1 + 1, bh = y2 - y1 + 1;
  const double x3d = box[6], y3d = box[7];
  double z3d = box[8];
  Box3 b = {box[9], box[10], box[11]};
  double ry = box[12];
  auto back_project = [&](double z, double* c3) {
    const double v[4] = {x3d * z, y3d * z, 1 * z, 1.0};
    for (int k = 0; k < 3; ++k) c3[k] = ((p2i[k * 4] * v[0] + p2i[k * 4 + 1] * v[1]) + p2i[k * 4 + 2] * v[2]) + p2i[k * 4 + 3] * v[3];
  };
  double c3[3];
  back_project(z3d, c3);
  ry = wrap_pi(ry + atan2(-c3[2], c3[0]) + 0.5 * kPi);  // convertAlpha2Rot
  if (hill_climbing) {
    // step_z_init = 0 at the reference's call site: only the rotation is searched
    double step_r = step_r_init;
    bool invalid;
    double ol_best = test_projection(p2, p2i, x1, y1, bw, bh, x3d, y3d, z3d, b, ry, &invalid);
    if (!invalid) {
      while (step_r > r_lim) {
        bool inv_neg, inv_pos;
        const double ol_neg = test_projection(p2, p2i, x1, y1, bw, bh, x3d, y3d, z3d, b, ry - step_r, &inv_neg);
        const double ol_pos = test_projection(p2, p2i, x1, y1, bw, bh, x3d, y3d, z3d, b, ry + step_r, &inv_pos);
        if ((ol_pos - ol_best) <= 0.0 && (ol_neg - ol_best) <= 0.0) {
          step_r = step_r * 0.5;
        } else if ((ol_pos - ol_best) > 0.0 && ol_pos > ol_neg && !inv_pos) {
          ry += step_r;
          ol_best = ol_pos;
        } else if ((ol_neg - ol_best) > 0.0 && !inv_neg) {
          ry -= step_r;
          ol_best = ol_neg;
        } else {
          step_r = step_r * 0.5;
        }
      }
      ry = wrap_pi(ry);
    }
  }
  back_project(z3d, c3);
  const double alpha = wrap_pi(ry - atan2(-c3[2], c3[0]) - 0.5 * kPi);  // convertRot2Alpha
  o[0] = static_cast<double>(box[5]) - 1.0;
  o[1] = alpha;
  o[2] = x1, o[3] = y1, o[4] = x2, o[5] = y2;
  o[6] = b.h3d, o[7] = b.w3d, o[8] = b.l3d;
  o[9] = c3[0], o[10] = c3[1] + b.h3d / 2, o[11] = c3[2];
  o[12] = ry;
  o[13] = box[4];
}

}  // namespace
}  // namespace m3d

using namespace m3d;

// kept: fp32 [B, max_out, row_len >= 13] rows (x1, y1, x2, y2, score, cls, x3d, y3d, z3d, w3d, h3d, l3d, alpha, ..) as
// m3d_gather_kept leaves them; num_keep [B]; p2 / p2_inv: float64 [B, 16] row-major 4x4 (device).  out: float64
// [B, max_out, 14] KITTI-line values, valid[B, max_out] = 1 where the row passed the score cut.
extern "C" int m3d_refine_3d(const float* kept, const int* num_keep, int B, int max_out, int row_len, const double* p2,
                             const double* p2_inv, float score_thresh, int hill_climbing, double step_r_init,
                             double r_lim, double* out, int* valid, m3d_stream_t stream) {
  M3D_REQUIRE(kept && num_keep && p2 && p2_inv && out && valid, "NULL pointer");
  M3D_REQUIRE(B >= 1 && max_out >= 1 && row_len >= 13, "bad geometry");
  M3D_REQUIRE(!hill_climbing || r_lim > 0.0, "hill climbing needs r_lim > 0 (the reference's loop would not end)");
  const int total = B * max_out;
  refine3d_kernel<<<(total + 63) / 64, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      kept, num_keep, B, max_out, row_len, p2, p2_inv, score_thresh, hill_climbing, step_r_init, r_lim, out, valid);
  M3D_CUDA_OK(cudaGetLastError());
  return M3D_OK;
}
