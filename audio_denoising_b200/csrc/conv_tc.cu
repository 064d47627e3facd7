// conv_tc.cu -- tcgen05 / TMEM implicit-GEMM variants of the GRUUNet2 encoder / decoder convolutions.
// (placeholder until the tensor-core path lands: conv_mode 1/2 report B2D_ERR_UNSUPPORTED)
#include "kernels.cuh"

namespace b2d {

int model_pack_tc(b2d_model* m) {
  m->d_tc = nullptr;
  m->tc_bytes = 0;
  return B2D_OK;
}

int model_forward_tc(const b2d_model*, const float*, size_t, float*, float*, float*, float*, int, cudaStream_t) {
  return fail(B2D_ERR_UNSUPPORTED, "tcgen05 encoder path not built into this library yet (use conv_mode=0)");
}

int model_decode_tc(const b2d_model*, const float*, const float*, const float*, const float*, const float*, size_t, float*,
                    float*, int, float, int, cudaStream_t) {
  return fail(B2D_ERR_UNSUPPORTED, "tcgen05 decoder path not built into this library yet (use conv_mode=0)");
}

}  // namespace b2d
