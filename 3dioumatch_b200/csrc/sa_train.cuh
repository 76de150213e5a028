// sa_train.cuh -- operand descriptions shared by the training-mode kernels (sa_train.cu, sa_train_dw.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

// dz_l = scale * (g - S1/R - xhat * S2/R), built on the fly from the saved raw conv output z and the upstream gradient
struct DzSrc {
  const float *z;        // (R, C) raw conv output of layer l
  const float *g;        // (R, C) gradient w.r.t. the BN output with the ReLU mask applied (hidden layers) or NULL (top layer)
  const float *gout_pm;  // top layer: (G, C) grad_out, point-major
  const int32_t *arg_pm; // top layer: (G, C) arg-max slot
  const float *mean, *invstd, *scale, *shift, *S1, *S2;  // (C)
  int C, ns;
  float inv_R;
};

// a_{l-1}: the first layer's input rows as they are, or relu(scale z + shift) of the previous layer
struct ActSrc {
  const float *rows;                // (R, ld)
  const float *scale, *shift;       // NULL: rows are used as they are (layer 1)
  const float *mean, *invstd;       // for xhat of the previous layer (bwd_da epilogue)
  int C, ld;
};


struct DwParams {
  long long R;
  int cin, cout, vec_act;
  DzSrc dz;
  ActSrc act;
  const float *coef_b, *coef_c;  // dz = dz.scale * g + coef_b + coef_c * z  (dz_coeff_kernel)
  float *partial;                // [splits][cout][cin]
};
int dw_tc_launch(DwParams &p, int splits, cudaStream_t stream);

}  // namespace b200
