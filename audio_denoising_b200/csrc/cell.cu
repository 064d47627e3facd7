// cell.cu -- the conv-GRU U-Net cell of the reference's model family for ANY configuration:
//   arch GRUUNET2  gruunet2.py:71-306 / gruunet.py (same maths): Gaussian position channels appended at every encoder and
//                  decoder layer;
//   arch MOMO3     momo3.py:191-324: two input channels (x_t, x_t - x_{t-1}), Gaussian position channels at the encoder input
//                  only, none in the decoder; shipped weights saves/MOMO3-4d4ea0: hidden (16,16,16), paddings (1,0,1), 3 bins.
// Arbitrary hidden sizes, kernel sizes, strides and paddings per level, odd lengths, padding 0.  The tuned kernels of
// model.cu / unet_mma.cu serve the shipped GRUUNet2 configuration (hidden 17 x 4, k3 s2 p1); everything else runs here:
// the same three phases (encoder time-parallel over all B*T frames -> recurrence, one CTA per clip -> decoder
// time-parallel), fp32 FMA, one CTA per frame with the frame's activations in shared memory, the Gaussian channels folded
// into per-position biases at pack time (they do not depend on the input).
#include <math.h>
#include <string.h>
#include <vector>

#include "kernels.cuh"

namespace b2d {

constexpr int kCellMaxLayers = 8;

struct CellLayer {
  int transposed, cin, cout, k, s, p, lin, lout, relu;
  int w_off;   // floats into the blob: Conv1d [cout][cin][k] / ConvTranspose1d [cin][cout][k] (data channels only)
  int pb_off;  // [lout][cout] bias + folded Gaussian channels
  int skip_c, skip_off;  // decoder: channels / offset (floats, within a frame's skip record) of the skip concatenated after this layer
  int out_off;           // encoder: where this layer's output goes in the frame's skip record (-1: it is the gate tensor)
};
struct CellDesc {  // passed by value to the kernels
  int arch, in_ch, n_mels, levels, H, bins, max_act;
  int skip_stride;  // floats per frame of encoder outputs kept for the decoder
  CellLayer enc[kCellMaxLayers], dec[kCellMaxLayers];
  int rec_w_off, rec_pb_off;  // recurrent conv [3H][H][3], [bins][3H]
};

}  // namespace b2d

struct b2d_cell {
  b2d_cell_config cfg;
  b2d::CellDesc d;
  float* d_blob;
  int device;
};

namespace b2d {

// out[co][j] = act(pb[j][co] + sum_ci sum_kk W[co][ci][kk] in[ci][s j - p + kk])
__device__ __forceinline__ void cell_conv(const CellLayer& L, const float* __restrict__ blob, const float* in, float* out) {
  const float* W = blob + L.w_off;
  const float* pb = blob + L.pb_off;
  for (int idx = threadIdx.x; idx < L.cout * L.lout; idx += blockDim.x) {
    const int co = idx / L.lout, j = idx - co * L.lout;
    float acc = pb[j * L.cout + co];
    for (int ci = 0; ci < L.cin; ++ci) {
      const float* w = W + ((size_t)co * L.cin + ci) * L.k;
      const float* x = in + ci * L.lin;
      for (int kk = 0; kk < L.k; ++kk) {
        const int q = L.s * j - L.p + kk;
        if (q >= 0 && q < L.lin) acc = fmaf(w[kk], x[q], acc);
      }
    }
    out[co * L.lout + j] = L.relu ? fmaxf(acc, 0.f) : acc;
  }
}
// ConvTranspose1d: out[co][o] = act(pb[o][co] + sum_ci sum_kk W[ci][co][kk] in[ci][i]),  o = s i - p + kk
__device__ __forceinline__ void cell_convt(const CellLayer& L, const float* __restrict__ blob, const float* in, float* out) {
  const float* W = blob + L.w_off;
  const float* pb = blob + L.pb_off;
  for (int idx = threadIdx.x; idx < L.cout * L.lout; idx += blockDim.x) {
    const int co = idx / L.lout, o = idx - co * L.lout;
    float acc = pb[o * L.cout + co];
    for (int kk = 0; kk < L.k; ++kk) {
      const int num = o + L.p - kk;
      if (num < 0 || num % L.s) continue;
      const int i = num / L.s;
      if (i >= L.lin) continue;
      for (int ci = 0; ci < L.cin; ++ci) acc = fmaf(W[((size_t)ci * L.cout + co) * L.k + kk], in[ci * L.lin + i], acc);
    }
    out[co * L.lout + o] = L.relu ? fmaxf(acc, 0.f) : acc;
  }
}

// encoder: one CTA per frame.  x [B, T, n_mels]; prev [B, n_mels] or null (MOMO3: the frame before the first one).
__global__ void __launch_bounds__(128) cell_encoder_kernel(const CellDesc d, const float* __restrict__ blob, const float* __restrict__ x,
                                                            const float* __restrict__ prev, int T, float* __restrict__ skips,
                                                            float* __restrict__ gx) {
  extern __shared__ float sm[];
  float* a = sm;
  float* b = sm + d.max_act;
  const size_t f = blockIdx.x;
  const int t = (int)(f % T);
  const float* xf = x + f * d.n_mels;
  for (int i = threadIdx.x; i < d.n_mels; i += blockDim.x) {
    const float v = xf[i];
    a[i] = v;
    if (d.in_ch == 2) {  // momo3.py:277-283: delta to the previous frame; the first frame is its own predecessor unless `prev` is given
      const float pv = (t > 0) ? xf[i - d.n_mels] : (prev ? prev[(f / T) * d.n_mels + i] : v);
      a[d.n_mels + i] = v - pv;
    }
  }
  __syncthreads();
  for (int l = 0; l < d.levels; ++l) {
    const CellLayer& L = d.enc[l];
    cell_conv(L, blob, a, b);
    __syncthreads();
    float* dst = (L.out_off >= 0) ? skips + f * d.skip_stride + L.out_off : gx + f * (size_t)(L.cout * L.lout);
    for (int i = threadIdx.x; i < L.cout * L.lout; i += blockDim.x) dst[i] = b[i];
    float* tmp = a; a = b; b = tmp;
    __syncthreads();
  }
}

// recurrence (gruunet2.py:231-240 / momo3.py:228-243): one CTA per clip, thread (g, c, j) owns one gate pre-activation
__global__ void __launch_bounds__(256) cell_recurrence_kernel(const CellDesc d, const float* __restrict__ blob, const float* __restrict__ gx,
                                                               float* __restrict__ hx, float* __restrict__ hseq, int T) {
  extern __shared__ float sm[];
  const int H = d.H, bins = d.bins, HB = H * bins;
  float* h = sm;             // [H][bins]
  float* pre = sm + HB;      // [3H][bins] relu(conv(h))
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < HB; i += blockDim.x) h[i] = hx[(size_t)b * HB + i];
  __syncthreads();
  const float* W = blob + d.rec_w_off;
  const float* pb = blob + d.rec_pb_off;
  for (int t = 0; t < T; ++t) {
    for (int idx = threadIdx.x; idx < 3 * HB; idx += blockDim.x) {
      const int co = idx / bins, j = idx - co * bins;
      float acc = pb[j * 3 * H + co];
      for (int ci = 0; ci < H; ++ci) {
        const float* w = W + ((size_t)co * H + ci) * 3;
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
          const int q = j - 1 + kk;
          if (q >= 0 && q < bins) acc = fmaf(w[kk], h[ci * bins + q], acc);
        }
      }
      pre[idx] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    const float* g = gx + ((size_t)b * T + t) * 3 * HB;
    for (int idx = threadIdx.x; idx < HB; idx += blockDim.x) {
      const float ir = g[idx], ii = g[HB + idx], in = g[2 * HB + idx];  // chunk(3, dim=1): reset, update, new
      const float hr = pre[idx], hi = pre[HB + idx], hn = pre[2 * HB + idx];
      const float z = 1.0f / (1.0f + expf(-(ii + hi)));
      const float r = 1.0f / (1.0f + expf(-(ir + hr)));
      const float nw = tanhf(in + r * hn);
      const float hv = nw + z * (h[idx] - nw);
      hseq[((size_t)b * T + t) * HB + idx] = hv;
      pre[idx] = hv;  // parked: h is still being read by nobody (all pre-activations are done), but keep the barrier order simple
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < HB; idx += blockDim.x) h[idx] = pre[idx];
    __syncthreads();
  }
  for (int i = threadIdx.x; i < HB; i += blockDim.x) hx[(size_t)b * HB + i] = h[i];
}

// decoder: one CTA per frame
__global__ void __launch_bounds__(128) cell_decoder_kernel(const CellDesc d, const float* __restrict__ blob, const float* __restrict__ hseq,
                                                            const float* __restrict__ skips, float* __restrict__ out) {
  extern __shared__ float sm[];
  float* a = sm;
  float* b = sm + d.max_act;
  const size_t f = blockIdx.x;
  for (int i = threadIdx.x; i < d.H * d.bins; i += blockDim.x) a[i] = hseq[f * (size_t)(d.H * d.bins) + i];
  __syncthreads();
  for (int l = 0; l < d.levels; ++l) {
    const CellLayer& L = d.dec[l];
    cell_convt(L, blob, a, b);
    if (L.skip_c > 0) {  // cat(relu(up), skip): upsampled channels first, the (already ReLU'd) encoder output second
      const float* s = skips + f * d.skip_stride + L.skip_off;
      for (int i = threadIdx.x; i < L.skip_c * L.lout; i += blockDim.x) b[L.cout * L.lout + i] = s[i];
    }
    __syncthreads();
    float* tmp = a; a = b; b = tmp;
  }
  for (int i = threadIdx.x; i < d.n_mels; i += blockDim.x) out[f * d.n_mels + i] = a[i];  // last layer: 1 channel x n_mels (squeeze(-2))
}

static void cell_smear(const float* off, int G, int nbins, std::vector<double>& s) {
  // gruunet2.py:54-68 / momo3.py:54-68: exp(coeff * (p - o)^2), p = linspace(0, 1, nbins), coeff = -0.5 / (o[1] - o[0])^2
  const double gap = (double)(float)(off[1] - off[0]);
  const double coeff = -0.5 / (gap * gap);
  s.assign((size_t)G * nbins, 0.0);
  for (int p = 0; p < nbins; ++p) {
    const double pos = (nbins > 1) ? (double)(float)((double)p / (double)(nbins - 1)) : 0.0;
    for (int g = 0; g < G; ++g) {
      const double dd = (double)(float)(pos - (double)off[g]);
      s[(size_t)g * nbins + p] = exp((double)(float)coeff * (double)(float)(dd * dd));
    }
  }
}

}  // namespace b2d

using namespace b2d;

extern "C" {

int b2d_cell_create(const b2d_cell_config* c, const float* const* hp, int n_params, const float* const* offs, b2d_cell** out) {
  B2D_REQUIRE(out && c && hp && offs, B2D_ERR_BAD_ARG, "NULL argument");
  *out = nullptr;
  B2D_REQUIRE(c->arch == B2D_ARCH_GRUUNET2 || c->arch == B2D_ARCH_MOMO3, B2D_ERR_BAD_ARG, "unknown arch %d", c->arch);
  const int Lv = c->levels, G = c->num_gaussians;
  B2D_REQUIRE(Lv >= 1 && Lv <= kCellMaxLayers, B2D_ERR_UNSUPPORTED, "levels must be in [1, %d] (got %d)", kCellMaxLayers, Lv);
  B2D_REQUIRE(G >= 2 && G <= 64, B2D_ERR_UNSUPPORTED, "num_gaussians must be in [2, 64] (got %d)", G);
  B2D_REQUIRE(c->n_mels >= 1 && c->n_mels <= 4096 && c->num_compressed_bins >= 1, B2D_ERR_BAD_ARG, "bad n_mels / num_compressed_bins");
  B2D_REQUIRE(n_params == 4 * Lv + 2, B2D_ERR_BAD_ARG, "expected %d parameter tensors, got %d", 4 * Lv + 2, n_params);
  for (int i = 0; i < n_params; ++i) B2D_REQUIRE(hp[i], B2D_ERR_BAD_ARG, "parameter %d is NULL", i);
  B2D_REQUIRE(offs[0] && offs[1] && (c->arch == B2D_ARCH_MOMO3 || offs[2]), B2D_ERR_BAD_ARG, "gs.offset pointer is NULL");
  const bool momo = c->arch == B2D_ARCH_MOMO3;
  b2d_cell* m = new b2d_cell();
  memset(m, 0, sizeof(*m));
  m->cfg = *c;
  CellDesc& d = m->d;
  d.arch = c->arch; d.in_ch = momo ? 2 : 1; d.n_mels = c->n_mels; d.levels = Lv; d.H = c->hidden[Lv - 1]; d.bins = c->num_compressed_bins;
  // ---- lengths down the encoder (gruunet2.py:127-157: floor((L + 2 p - k) / s) + 1) ----
  int len[kCellMaxLayers + 1];
  len[0] = c->n_mels;
  for (int l = 0; l < Lv; ++l) {
    B2D_REQUIRE(c->hidden[l] >= 1 && c->kernel[l] >= 1 && c->stride[l] >= 1 && c->padding[l] >= 0, B2D_ERR_BAD_ARG, "bad level %d", l);
    const int num = len[l] + 2 * c->padding[l] - c->kernel[l];
    if (num < 0) { delete m; return fail(B2D_ERR_BAD_ARG, "n_mels = %d is too short for level %d of this configuration", c->n_mels, l); }
    len[l + 1] = num / c->stride[l] + 1;
  }
  if (len[Lv] != c->num_compressed_bins) {
    const int got = len[Lv];
    delete m;
    return fail(B2D_ERR_BAD_ARG, "n_mels = %d compresses to %d bins, the model has num_compressed_bins = %d", c->n_mels, got, c->num_compressed_bins);
  }
  std::vector<float> blob;
  std::vector<double> S;
  int max_act = 2 * c->n_mels;
  // ---- encoder (input_gate): weights [cout][cin + G?][k] ----
  int skip_stride = 0;
  for (int l = 0; l < Lv; ++l) {
    CellLayer& L = d.enc[l];
    const bool gauss = !momo || l == 0;
    L.transposed = 0; L.cin = (l == 0) ? d.in_ch : c->hidden[l - 1]; L.cout = (l == Lv - 1) ? 3 * c->hidden[l] : c->hidden[l];
    L.k = c->kernel[l]; L.s = c->stride[l]; L.p = c->padding[l]; L.lin = len[l]; L.lout = len[l + 1]; L.relu = 1;
    const int CT = L.cin + (gauss ? G : 0);
    const float* W = hp[2 * l];
    const float* bias = hp[2 * l + 1];
    L.w_off = (int)blob.size();
    for (int co = 0; co < L.cout; ++co)
      for (int ci = 0; ci < L.cin; ++ci)
        for (int kk = 0; kk < L.k; ++kk) blob.push_back(W[((size_t)co * CT + ci) * L.k + kk]);
    if (gauss) cell_smear(offs[0], G, L.lin, S);
    L.pb_off = (int)blob.size();
    for (int j = 0; j < L.lout; ++j)
      for (int co = 0; co < L.cout; ++co) {
        double acc = bias[co];
        if (gauss)
          for (int g = 0; g < G; ++g)
            for (int kk = 0; kk < L.k; ++kk) {
              const int q = L.s * j - L.p + kk;
              if (q >= 0 && q < L.lin) acc += (double)W[((size_t)co * CT + L.cin + g) * L.k + kk] * S[(size_t)g * L.lin + q];
            }
        blob.push_back((float)acc);
      }
    if (l < Lv - 1) { L.out_off = skip_stride; skip_stride += L.cout * L.lout; } else L.out_off = -1;
    if (L.cout * L.lout > max_act) max_act = L.cout * L.lout;
  }
  d.skip_stride = skip_stride > 0 ? skip_stride : 1;
  // ---- recurrent conv (reset_gate): [3H][H + G][3], stride 1, padding 1 ----
  {
    const int H = d.H, bins = d.bins, CT = H + G;
    const float* W = hp[2 * Lv];
    const float* bias = hp[2 * Lv + 1];
    d.rec_w_off = (int)blob.size();
    for (int co = 0; co < 3 * H; ++co)
      for (int ci = 0; ci < H; ++ci)
        for (int kk = 0; kk < 3; ++kk) blob.push_back(W[((size_t)co * CT + ci) * 3 + kk]);
    cell_smear(offs[1], G, bins, S);
    d.rec_pb_off = (int)blob.size();
    for (int j = 0; j < bins; ++j)
      for (int co = 0; co < 3 * H; ++co) {
        double acc = bias[co];
        for (int g = 0; g < G; ++g)
          for (int kk = 0; kk < 3; ++kk) {
            const int q = j - 1 + kk;
            if (q >= 0 && q < bins) acc += (double)W[((size_t)co * CT + H + g) * 3 + kk] * S[(size_t)g * bins + q];
          }
        blob.push_back((float)acc);
      }
  }
  // ---- decoder (output_gate): level i undoes encoder level Lv-1-i; ConvTranspose weights [cin + G?][cout][k] ----
  int cur_c = d.H, cur_l = d.bins;
  for (int i = 0; i < Lv; ++i) {
    CellLayer& L = d.dec[i];
    const int e = Lv - 1 - i;  // the encoder level whose input length this layer restores
    const bool gauss = !momo;
    L.transposed = 1; L.cin = cur_c; L.cout = (e == 0) ? 1 : c->hidden[e - 1];
    L.k = c->kernel[e]; L.s = c->stride[e]; L.p = c->padding[e]; L.lin = cur_l; L.lout = len[e]; L.relu = (i < Lv - 1);
    const int natural = (L.lin - 1) * L.s - 2 * L.p + L.k;  // ConvTranspose1d(output_size=...): output_padding = lout - natural in [0, stride)
    if (L.lout < natural || L.lout - natural > (L.s > 1 ? L.s - 1 : 0)) {
      delete m;
      return fail(B2D_ERR_BAD_ARG, "decoder level %d cannot produce length %d from %d (ConvTranspose1d output_size rule)", i, len[e], cur_l);
    }
    const int CT = L.cin + (gauss ? G : 0);
    const float* W = hp[2 * Lv + 2 + 2 * i];
    const float* bias = hp[2 * Lv + 3 + 2 * i];
    L.w_off = (int)blob.size();
    for (int ci = 0; ci < L.cin; ++ci)
      for (int co = 0; co < L.cout; ++co)
        for (int kk = 0; kk < L.k; ++kk) blob.push_back(W[((size_t)ci * L.cout + co) * L.k + kk]);
    if (gauss) cell_smear(offs[2], G, L.lin, S);
    L.pb_off = (int)blob.size();
    for (int o = 0; o < L.lout; ++o)
      for (int co = 0; co < L.cout; ++co) {
        double acc = bias[co];
        if (gauss)
          for (int g = 0; g < G; ++g)
            for (int kk = 0; kk < L.k; ++kk) {
              const int num = o + L.p - kk;
              if (num < 0 || num % L.s) continue;
              const int q = num / L.s;
              if (q < L.lin) acc += (double)W[((size_t)(L.cin + g) * L.cout + co) * L.k + kk] * S[(size_t)g * L.lin + q];
            }
        blob.push_back((float)acc);
      }
    if (i < Lv - 1) {  // cat(relu(up), skip = encoder output of level e-1)
      L.skip_c = d.enc[e - 1].cout; L.skip_off = d.enc[e - 1].out_off;
      cur_c = L.cout + L.skip_c;
    } else {
      L.skip_c = 0; L.skip_off = 0;
      cur_c = L.cout;
    }
    cur_l = L.lout;
    if (cur_c * cur_l > max_act) max_act = cur_c * cur_l;
  }
  if (d.H * d.bins > max_act) max_act = d.H * d.bins;
  d.max_act = (max_act + 3) & ~3;
  if ((size_t)2 * d.max_act * sizeof(float) > 200 * 1024) { delete m; return fail(B2D_ERR_UNSUPPORTED, "activations of one frame exceed shared memory"); }
  B2D_CUDA(cudaGetDevice(&m->device));
  if (cudaMalloc(&m->d_blob, blob.size() * sizeof(float)) != cudaSuccess ||
      cudaMemcpy(m->d_blob, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(m->d_blob);
    delete m;
    return fail(B2D_ERR_CUDA, "uploading the packed cell failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  *out = m;
  return B2D_OK;
}

void b2d_cell_destroy(b2d_cell* m) {
  if (!m) return;
  cudaFree(m->d_blob);
  delete m;
}

size_t b2d_cell_workspace_bytes(const b2d_cell* m, int B, int T) {
  if (!m || B < 1 || T < 1) return 0;
  const size_t nf = (size_t)B * T, HB = (size_t)m->d.H * m->d.bins;
  return align_up(nf * m->d.skip_stride * 4, 256) + align_up(nf * 3 * HB * 4, 256) + align_up(nf * HB * 4, 256);
}

int b2d_cell_forward(const b2d_cell* m, const float* x, const float* prev, float* hx, float* out, int B, int T, void* ws,
                     size_t ws_bytes, void* stream) {
  B2D_REQUIRE(m && x && hx && out && ws, B2D_ERR_BAD_ARG, "NULL pointer");
  B2D_REQUIRE(B >= 1 && T >= 1 && (long long)B * T < (1ll << 31), B2D_ERR_BAD_ARG, "bad batch / sequence length");
  B2D_REQUIRE(ws_bytes >= b2d_cell_workspace_bytes(m, B, T), B2D_ERR_WORKSPACE, "cell workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const CellDesc& d = m->d;
  const size_t nf = (size_t)B * T, HB = (size_t)d.H * d.bins;
  unsigned char* base = static_cast<unsigned char*>(ws);
  float* skips = reinterpret_cast<float*>(base); base += align_up(nf * d.skip_stride * 4, 256);
  float* gx = reinterpret_cast<float*>(base); base += align_up(nf * 3 * HB * 4, 256);
  float* hseq = reinterpret_cast<float*>(base);
  const size_t smem = (size_t)2 * d.max_act * sizeof(float);
  B2D_SMEM_OPT_IN(smem, cell_encoder_kernel);
  cell_encoder_kernel<<<(unsigned)nf, 128, smem, st>>>(d, m->d_blob, x, prev, T, skips, gx);
  B2D_LAUNCH_CHECK("cell_encoder_kernel");
  const size_t rsmem = (size_t)4 * HB * sizeof(float);
  B2D_REQUIRE(rsmem <= 200 * 1024, B2D_ERR_UNSUPPORTED, "hidden state exceeds shared memory");
  {
    const size_t smem = rsmem;
    B2D_SMEM_OPT_IN(smem, cell_recurrence_kernel);
  }
  cell_recurrence_kernel<<<B, 256, rsmem, st>>>(d, m->d_blob, gx, hx, hseq, T);
  B2D_LAUNCH_CHECK("cell_recurrence_kernel");
  {
    B2D_SMEM_OPT_IN(smem, cell_decoder_kernel);
  }
  cell_decoder_kernel<<<(unsigned)nf, 128, smem, st>>>(d, m->d_blob, hseq, skips, out);
  B2D_LAUNCH_CHECK("cell_decoder_kernel");
  return B2D_OK;
}

}  // extern "C"
