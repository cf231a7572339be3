/* TEST INFRASTRUCTURE ONLY -- CPU oracle for m3dssd_b200.
 *
 * Restates, in plain C, the reference's native arithmetic on the hot path so
 * that the CUDA kernels can be checked against it.  Nothing under oracle/ is
 * linked, imported or executed by the product (m3dssd_b200/): only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may call it, and only as the checker or the timed CPU baseline.
 *
 * Pinning: validated against (a) the reference's own known-answer test
 * check_zero_offset (model/DCNv2/test.py:32-65), (b) torchvision.ops.
 * deform_conv2d on random inputs, (c) the reference's unmodified CUDA kernels
 * compiled from /root/reference into oracle/_ref and run on the GPU box
 * (tests/test_ref_kernels_gpu.py), (d) the reference's lib/nms/py_cpu_nms.py
 * run in the build container (fixtures in tests/golden/).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define REAL float
#define FN(x) x##_f32
#define FLOOR floorf
#define FABS fabsf
#include "dcn_body.inc"
#undef REAL
#undef FN
#undef FLOOR
#undef FABS

#define REAL double
#define FN(x) x##_f64
#define FLOOR floor
#define FABS fabs
#include "dcn_body.inc"
#undef REAL
#undef FN
#undef FLOOR
#undef FABS

/* ------------------------------------------------------------------ NMS
 * lib/nms/nms_kernel.cu.  devIoU (:24-32) in the operation order nvcc's default
 * -fmad=true gives the reference kernel (SASS of the unmodified source built for
 * sm_100a: Sa = FMUL; Sa + Sb = FFMA(wb, hb, Sa); interS = FMUL; FADD; IEEE
 * division), so the threshold test `IoU > thresh` (:71) is bit-identical.
 * Compile this file with -ffp-contract=off: the only fused op is the fmaf below. */
static float dev_iou(const float* a, const float* b) {
  float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  float width = fmaxf(right - left + 1.f, 0.f), height = fmaxf(bottom - top + 1.f, 0.f);
  float interS = width * height;
  float Sa = (a[2] - a[0] + 1.f) * (a[3] - a[1] + 1.f);
  float SaSb = fmaf(b[2] - b[0] + 1.f, b[3] - b[1] + 1.f, Sa);
  return interS / (SaSb - interS);
}

/* _nms (:91-144): boxes are already sorted by score (descending); box i
 * suppresses every later box j with IoU(i, j) > thresh; greedy sweep (:124-139). */
void m3d_oracle_nms(int* keep_out, int* num_out, const float* boxes, int n, int dim, float thresh) {
  unsigned char* removed = (unsigned char*)calloc((size_t)(n > 0 ? n : 1), 1);
  int kept = 0;
  for (int i = 0; i < n; ++i) {
    if (removed[i]) continue;
    keep_out[kept++] = i;
    for (int j = i + 1; j < n; ++j)
      if (!removed[j] && dev_iou(boxes + (size_t)i * dim, boxes + (size_t)j * dim) > thresh) removed[j] = 1;
  }
  *num_out = kept;
  free(removed);
}

/* The 64-bit suppression bitmask the reference kernel emits (:61-77), for
 * checking the device bitmask itself: mask[i * col_blocks + cb] bit k set iff
 * j = cb*64 + k > i and IoU(i, j) > thresh. */
void m3d_oracle_nms_mask(unsigned long long* mask, const float* boxes, int n, int dim, float thresh) {
  const int col_blocks = (n + 63) / 64;
  memset(mask, 0, sizeof(unsigned long long) * (size_t)n * col_blocks);
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j)
      if (dev_iou(boxes + (size_t)i * dim, boxes + (size_t)j * dim) > thresh)
        mask[(size_t)i * col_blocks + j / 64] |= 1ULL << (j % 64);
}
