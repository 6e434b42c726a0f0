/* Host build of dfki_quad_b200/csrc/wbc_scene.cu (kernel K6) for logic checks against the numpy scene assembly - test infrastructure. */
#include "emu_preamble.h"
#define B200QP_HOST_EMULATION 1
#include "../../dfki_quad_b200/csrc/wbc_scene.cu"

extern "C" void wbc_scene_host(int n, const b200qp::WbcSceneConfig *C, const double *scene_in, double *H, double *g, double *A, double *lo,
                               double *hi, double *feet) {
  for (int p = 0; p < n; ++p)
    b200qp::wbc_scene_kernel(1, *C, scene_in + (size_t)p * 72, H + (size_t)p * 900, g + (size_t)p * 30, A + (size_t)p * 1380, lo + (size_t)p * 46,
                             hi + (size_t)p * 46, feet ? feet + (size_t)p * 24 : nullptr);
}
