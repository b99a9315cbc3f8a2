// Internal interface of the tcgen05 GEMM / implicit-conv kernel.
#pragma once
#include "common.cuh"

namespace lc {

// Implicit 3x3 convolution A-operand addressing (disabled for plain GEMMs).
struct ConvLoad {
  int enabled = 0;
  int H = 0, W = 0, n_frames = 0;
  int Wt = 0, Nt = 0;   // tile = Nt frames x 1 image row x Wt columns (Wt*Nt <= 128)
  int tiles_x = 0;      // ceil(W / Wt)
  int kb_per_tap = 0;   // Cp / 64
  int a_bytes = 0;      // bytes one A box delivers (64*2*Wt*Nt)
  int c_real = 0;       // un-padded input channels (algorithmic FLOP count only)
};

int conv3x3_bf16(const void* xpad, int n_frames, int H, int W, int Cp, const void* wmat, int C_out,
                 const EpiParams& epi, cudaStream_t stream, int c_real = 0);

}  // namespace lc
