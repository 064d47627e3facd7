// model_layout.cuh -- GRUUNet2 shipped configuration constants and the packed parameter blob layout shared by the
// CUDA-core kernels (model.cu) and the warp-level tensor-core kernels (unet_mma.cu).
#pragma once

namespace b2d {

// shipped configuration (all three checkpoints): hidden 17, 4 levels, 4 compressed bins
constexpr int H = 17;
constexpr int HP = 20;        // hidden padded to a multiple of 4 (float4 weight rows)
constexpr int H3 = 51;
constexpr int H3P = 52;
constexpr int LEVELS = 4;
constexpr int BINS = 4;
constexpr int NMEL = BINS << LEVELS;  // 64

// per-frame activation sizes
constexpr int D0 = H * 32, D1 = H * 16, D2 = H * 8, GX = H3 * 4, HS = H * 4;

// ---- packed parameter blob layout (floats) --------------------------------------------------------
// encoder layer l : W[ci][k][coP]  ,  PB[j][coP]
// recurrent       : W[ci][k][gate][c] (c padded to HP) , PB[gate][c][j]
// decoder layer i : W[ci][k][coP]  ,  PB[o][coP]
struct Packed {
  int enc_w[LEVELS], enc_pb[LEVELS];
  int rec_w, rec_pb;
  int dec_w[LEVELS], dec_pb[LEVELS];
  int total;
};
__host__ __device__ inline Packed packed_layout() {
  Packed p{};
  int o = 0;
  const int cin[LEVELS] = {1, H, H, H};
  const int cop[LEVELS] = {HP, HP, HP, H3P};
  const int lout[LEVELS] = {32, 16, 8, 4};
  for (int l = 0; l < LEVELS; ++l) {
    p.enc_w[l] = o; o += cin[l] * 3 * cop[l];
    p.enc_pb[l] = o; o += lout[l] * cop[l];
  }
  p.rec_w = o; o += H * 3 * 3 * HP;
  p.rec_pb = o; o += 3 * H * 4;
  const int dcin[LEVELS] = {H, 2 * H, 2 * H, 2 * H};
  const int dcop[LEVELS] = {HP, HP, HP, 4};
  const int dlout[LEVELS] = {8, 16, 32, 64};
  for (int i = 0; i < LEVELS; ++i) {
    p.dec_w[i] = o; o += dcin[i] * 3 * dcop[i];
    p.dec_pb[i] = o; o += dlout[i] * dcop[i];
  }
  p.total = o;
  return p;
}

}  // namespace b2d
