/* Host build of dfki_quad_b200/csrc/sequencer_full.cu (kernel K5b) for logic checks against oracle/_ref - test infrastructure. */
#include "emu_preamble.h"
#define B200QP_HOST_EMULATION 1
#include "../../dfki_quad_b200/csrc/sequencer_full.cu"

extern "C" void seqfull_host_step(int n, const b200qp::SeqModel *M, const b200qp::SeqOptions *O, const double *seq_in,
                                  const uint8_t *measured, double *state, double *rec, uint8_t *contact, int in_stride) {
  for (int p = 0; p < n; ++p)
    b200qp::sequencer_full_kernel(1, *M, *O, seq_in + (size_t)p * 64, measured ? measured + 4 * p : nullptr, state + (size_t)p * 192,
                                  rec + (size_t)p * in_stride, contact + (size_t)p * M->N * 4, in_stride);
}
