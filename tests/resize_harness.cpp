// Host harness around csrc/resize_math.h (the arithmetic of resize_u8_to_f32_kernel): lets the CPU tests check the
// kernel's per-pixel math against the cv2-pinned oracle without a GPU.  Test infrastructure only -- never part of
// libb200det.so.
#include <stddef.h>
#include <stdint.h>

#include "../object_detection_tracking_b200/csrc/resize_math.h"

extern "C" void resize_harness(const uint8_t* src, int sh, int sw, float* dst, int dh, int dw) {
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x) {
      const b2::ResizeTap tx = b2::resize_tap_x(x, sw, dw), ty = b2::resize_tap_y(y, sh, dh);
      const uint8_t* r0 = src + static_cast<size_t>(ty.i0) * sw * 3;
      const uint8_t* r1 = src + static_cast<size_t>(ty.i1) * sw * 3;
      for (int c = 0; c < 3; ++c)
        dst[(static_cast<size_t>(y) * dw + x) * 3 + c] =
            b2::resize_sample(r0[tx.i0 * 3 + c], r0[tx.i1 * 3 + c], r1[tx.i0 * 3 + c], r1[tx.i1 * 3 + c], tx, ty);
    }
}
