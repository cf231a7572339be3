// shape_align offsets of one pixel (model/module/feturealign_mgpu.py:119-136, 160-172), shared by the stand-alone
// builder, the softmax kernel's tail (elementwise.cu) and the fused class-head epilogue (igemm.cu):
// tap (i, j) of the 3x3 DCNv2 moves by ((ah/stride/3 - 1)(i - 1), (aw/stride/3 - 1)(j - 1)) for the top-1 anchor, zeroed
// where fg <= thresh; modulation mask = fg.  om: [npix, 27] = 18 offsets (dh, dw per tap) + 9 masks.
#pragma once

namespace m3d {

__device__ __forceinline__ void shape_align_om_pixel(float fg, int a, const float* __restrict__ anchors, int anchor_ld,
                                                     float feat_stride, float thresh, float* __restrict__ om, long i) {
  const float hard = fg > thresh ? 1.f : 0.f;
  const float aw = anchors[a * anchor_ld + 2] - anchors[a * anchor_ld + 0];
  const float ah = anchors[a * anchor_ld + 3] - anchors[a * anchor_ld + 1];
  const float hstep = ah / feat_stride / 3.f - 1.f;
  const float wstep = aw / feat_stride / 3.f - 1.f;
  float* o = om + i * 27;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int ti = t / 3, tj = t % 3;
    o[2 * t] = hstep * (static_cast<float>(ti) - 1.5f + 0.5f) * hard;
    o[2 * t + 1] = wstep * (static_cast<float>(tj) - 1.5f + 0.5f) * hard;
    o[18 + t] = fg;
  }
}

}  // namespace m3d
