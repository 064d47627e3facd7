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
  // backward: the layer's parameters in the flat gradient vector (parameters() order, torch layouts incl. the Gaussian channels)
  int gw_off, gb_off, ct, gauss, smear_off;  // ct = cin + Gaussian channels; smear_off: S[g][lin] table in the blob
};
struct CellDesc {  // passed by value to the kernels
  int arch, in_ch, n_mels, levels, H, bins, max_act;
  int skip_stride;  // floats per frame of encoder outputs kept for the decoder
  CellLayer enc[kCellMaxLayers], dec[kCellMaxLayers];
  int rec_w_off, rec_pb_off;  // recurrent conv [3H][H][3], [bins][3H]
  int rec_gw_off, rec_gb_off, rec_smear_off, G;
  int blob_floats, n_param_floats;
  int dec_act_total, enc_act_total;  // floats of all layer inputs (+ last output) of one frame, for the recomputing backward kernels
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

// =================================================================================================
// fp32 backward of the cell (fine-tuning through the B200 module; server.py:86-142 builds
// AdamW(self.inner.parameters()) around it).  Same three phases in reverse: decoder (time-parallel) -> recurrence (BPTT, one
// CTA per clip) -> encoder (time-parallel).  Each kernel recomputes its forward activations of a frame in shared memory from
// what the forward pass left in the workspace (skips, gate_x, hidden sequence), back-propagates, and accumulates weight /
// position-bias gradients in shared memory over all the frames the CTA walks, with one atomicAdd per element per CTA at the
// end.  A last kernel turns (d weights, d position-bias) into the torch-layout gradients, Gaussian channels and bias included.
// =================================================================================================
__device__ __forceinline__ void cell_zero(float* p, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = 0.f;
}
__device__ __forceinline__ void cell_flush(float* __restrict__ g, const float* acc, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = acc[i];
    if (v != 0.f) atomicAdd(g + i, v);
  }
}

// d_in[ci][q] = sum_co sum_kk W[co][ci][kk] dpre[co][j], q = s j - p + kk ;  dW[co][ci][kk] += in[ci][q] dpre[co][j] ;  dpb[j][co] += dpre[co][j]
__device__ __forceinline__ void cell_conv_bwd(const CellLayer& L, const float* __restrict__ blob, const float* in, const float* dpre,
                                              float* din, float* gW, float* gpb) {
  const float* W = blob + L.w_off;
  for (int idx = threadIdx.x; idx < L.cin * L.lin; idx += blockDim.x) {
    const int ci = idx / L.lin, q = idx - ci * L.lin;
    float acc = 0.f;
    for (int kk = 0; kk < L.k; ++kk) {
      const int num = q + L.p - kk;
      if (num < 0 || num % L.s) continue;
      const int j = num / L.s;
      if (j >= L.lout) continue;
      for (int co = 0; co < L.cout; ++co) acc = fmaf(W[((size_t)co * L.cin + ci) * L.k + kk], dpre[co * L.lout + j], acc);
    }
    din[idx] = acc;
  }
  for (int idx = threadIdx.x; idx < L.cout * L.cin * L.k; idx += blockDim.x) {
    const int kk = idx % L.k, ci = (idx / L.k) % L.cin, co = idx / (L.k * L.cin);
    float acc = 0.f;
    for (int j = 0; j < L.lout; ++j) {
      const int q = L.s * j - L.p + kk;
      if (q >= 0 && q < L.lin) acc = fmaf(in[ci * L.lin + q], dpre[co * L.lout + j], acc);
    }
    gW[idx] += acc;
  }
  for (int idx = threadIdx.x; idx < L.cout * L.lout; idx += blockDim.x) {
    const int co = idx / L.lout, j = idx - co * L.lout;
    gpb[j * L.cout + co] += dpre[idx];
  }
}
// ConvTranspose1d: d_in[ci][i] = sum_co sum_kk W[ci][co][kk] dpre[co][o], o = s i - p + kk ;  dW[ci][co][kk] += in[ci][i] dpre[co][o]
__device__ __forceinline__ void cell_convt_bwd(const CellLayer& L, const float* __restrict__ blob, const float* in, const float* dpre,
                                               float* din, float* gW, float* gpb) {
  const float* W = blob + L.w_off;
  for (int idx = threadIdx.x; idx < L.cin * L.lin; idx += blockDim.x) {
    const int ci = idx / L.lin, i = idx - ci * L.lin;
    float acc = 0.f;
    for (int kk = 0; kk < L.k; ++kk) {
      const int o = L.s * i - L.p + kk;
      if (o < 0 || o >= L.lout) continue;
      for (int co = 0; co < L.cout; ++co) acc = fmaf(W[((size_t)ci * L.cout + co) * L.k + kk], dpre[co * L.lout + o], acc);
    }
    din[idx] = acc;
  }
  for (int idx = threadIdx.x; idx < L.cin * L.cout * L.k; idx += blockDim.x) {
    const int kk = idx % L.k, co = (idx / L.k) % L.cout, ci = idx / (L.k * L.cout);
    float acc = 0.f;
    for (int i = 0; i < L.lin; ++i) {
      const int o = L.s * i - L.p + kk;
      if (o >= 0 && o < L.lout) acc = fmaf(in[ci * L.lin + i], dpre[co * L.lout + o], acc);
    }
    gW[idx] += acc;
  }
  for (int idx = threadIdx.x; idx < L.cout * L.lout; idx += blockDim.x) {
    const int co = idx / L.lout, o = idx - co * L.lout;
    gpb[o * L.cout + co] += dpre[idx];
  }
}

// shared memory: acts[dec_act_total] | ga[max_act] | gb[max_act] | gacc[decoder part of the blob]
__global__ void __launch_bounds__(128) cell_decoder_bwd_kernel(const CellDesc d, const float* __restrict__ blob, const float* __restrict__ hseq,
                                                                const float* __restrict__ skips, const float* __restrict__ gout,
                                                                float* __restrict__ d_hseq, float* __restrict__ d_skips,
                                                                float* __restrict__ gblob, size_t nframes, int acc_lo, int acc_n) {
  extern __shared__ float sm[];
  float* acts = sm;
  float* ga = acts + d.dec_act_total;
  float* gb = ga + d.max_act;
  float* gacc = gb + d.max_act;  // mirrors blob[acc_lo, acc_lo + acc_n)
  cell_zero(gacc, acc_n);
  __syncthreads();
  for (size_t f = blockIdx.x; f < nframes; f += gridDim.x) {
    // ---- forward recompute, keeping every layer's input ----
    float* a = acts;
    for (int i = threadIdx.x; i < d.H * d.bins; i += blockDim.x) a[i] = hseq[f * (size_t)(d.H * d.bins) + i];
    __syncthreads();
    for (int l = 0; l < d.levels; ++l) {
      const CellLayer& L = d.dec[l];
      float* nx = a + L.cin * L.lin;
      cell_convt(L, blob, a, nx);
      if (L.skip_c > 0) {
        const float* s = skips + f * d.skip_stride + L.skip_off;
        for (int i = threadIdx.x; i < L.skip_c * L.lout; i += blockDim.x) nx[L.cout * L.lout + i] = s[i];
      }
      __syncthreads();
      a = nx;
    }
    // ---- backward ----
    for (int i = threadIdx.x; i < d.n_mels; i += blockDim.x) ga[i] = gout[f * d.n_mels + i];
    __syncthreads();
    for (int l = d.levels - 1; l >= 0; --l) {
      const CellLayer& L = d.dec[l];
      const float* outp = a;               // this layer's output (first cout channels of the next layer's input)
      float* inp = a - L.cin * L.lin;      // this layer's input
      if (L.relu)
        for (int i = threadIdx.x; i < L.cout * L.lout; i += blockDim.x) if (!(outp[i] > 0.f)) ga[i] = 0.f;
      __syncthreads();
      cell_convt_bwd(L, blob, inp, ga, gb, gacc + (L.w_off - acc_lo), gacc + (L.pb_off - acc_lo));
      __syncthreads();
      if (l > 0) {  // input = cat(relu(up of layer l-1), skip): split the gradient
        const CellLayer& P = d.dec[l - 1];
        float* ds = d_skips + f * d.skip_stride + P.skip_off;
        for (int i = threadIdx.x; i < P.skip_c * P.lout; i += blockDim.x) ds[i] = gb[P.cout * P.lout + i];
        for (int i = threadIdx.x; i < P.cout * P.lout; i += blockDim.x) ga[i] = gb[i];
      } else {
        for (int i = threadIdx.x; i < d.H * d.bins; i += blockDim.x) d_hseq[f * (size_t)(d.H * d.bins) + i] = gb[i];
      }
      __syncthreads();
      a = inp;
    }
  }
  cell_flush(gblob + acc_lo, gacc, acc_n);
}

// BPTT.  shared memory: h[HB] | pre[3HB] | dpre[3HB] | dh[HB] | dhp[HB] | gW[3H*H*3] | gpb[bins*3H]
__global__ void __launch_bounds__(256) cell_recurrence_bwd_kernel(const CellDesc d, const float* __restrict__ blob, const float* __restrict__ gx,
                                                                   const float* __restrict__ hx0, const float* __restrict__ hseq,
                                                                   const float* __restrict__ d_hseq, const float* __restrict__ d_hT,
                                                                   float* __restrict__ d_gx, float* __restrict__ d_h0, float* __restrict__ gblob,
                                                                   int T) {
  extern __shared__ float sm[];
  const int H = d.H, bins = d.bins, HB = H * bins;
  float* h = sm;
  float* pre = h + HB;
  float* dpre = pre + 3 * HB;
  float* dh = dpre + 3 * HB;
  float* dhp = dh + HB;
  float* gW = dhp + HB;
  float* gpb = gW + 3 * H * H * 3;
  const int b = blockIdx.x;
  const float* W = blob + d.rec_w_off;
  const float* pb = blob + d.rec_pb_off;
  cell_zero(gW, 3 * H * H * 3 + bins * 3 * H);
  for (int i = threadIdx.x; i < HB; i += blockDim.x) dh[i] = d_hT ? d_hT[(size_t)b * HB + i] : 0.f;
  __syncthreads();
  for (int t = T - 1; t >= 0; --t) {
    const float* hprev = (t > 0) ? hseq + ((size_t)b * T + t - 1) * HB : hx0 + (size_t)b * HB;
    for (int i = threadIdx.x; i < HB; i += blockDim.x) {
      h[i] = hprev[i];
      dh[i] += d_hseq[((size_t)b * T + t) * HB + i];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 3 * HB; idx += blockDim.x) {  // recompute relu(conv(h_{t-1}))
      const int co = idx / bins, j = idx - co * bins;
      float acc = pb[j * 3 * H + co];
      for (int ci = 0; ci < H; ++ci)
        for (int kk = 0; kk < 3; ++kk) {
          const int q = j - 1 + kk;
          if (q >= 0 && q < bins) acc = fmaf(W[((size_t)co * H + ci) * 3 + kk], h[ci * bins + q], acc);
        }
      pre[idx] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    const float* g = gx + ((size_t)b * T + t) * 3 * HB;
    float* dg = d_gx + ((size_t)b * T + t) * 3 * HB;
    for (int idx = threadIdx.x; idx < HB; idx += blockDim.x) {
      const float hr = pre[idx], hi = pre[HB + idx], hn = pre[2 * HB + idx];
      const float z = 1.0f / (1.0f + expf(-(g[HB + idx] + hi)));
      const float r = 1.0f / (1.0f + expf(-(g[idx] + hr)));
      const float n = tanhf(g[2 * HB + idx] + r * hn);
      const float dht = dh[idx];
      const float dn = dht * (1.0f - z), dz = dht * (h[idx] - n);
      const float dan = dn * (1.0f - n * n), daz = dz * z * (1.0f - z);
      const float dar = dan * hn * r * (1.0f - r);
      dg[idx] = dar; dg[HB + idx] = daz; dg[2 * HB + idx] = dan;   // chunk(3, 1): reset, update, new
      dpre[idx] = hr > 0.f ? dar : 0.f;
      dpre[HB + idx] = hi > 0.f ? daz : 0.f;
      dpre[2 * HB + idx] = hn > 0.f ? dan * r : 0.f;
      dhp[idx] = dht * z;  // direct path h_t = n + z (h_{t-1} - n)
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < HB; idx += blockDim.x) {  // conv input gradient
      const int ci = idx / bins, q = idx - ci * bins;
      float acc = dhp[idx];
      for (int kk = 0; kk < 3; ++kk) {
        const int j = q + 1 - kk;
        if (j < 0 || j >= bins) continue;
        for (int co = 0; co < 3 * H; ++co) acc = fmaf(W[((size_t)co * H + ci) * 3 + kk], dpre[co * bins + j], acc);
      }
      dh[idx] = acc;  // gradient w.r.t. h_{t-1}, carried to the next (earlier) step
    }
    for (int idx = threadIdx.x; idx < 3 * H * H * 3; idx += blockDim.x) {
      const int kk = idx % 3, ci = (idx / 3) % H, co = idx / (3 * H);
      float acc = 0.f;
      for (int j = 0; j < bins; ++j) {
        const int q = j - 1 + kk;
        if (q >= 0 && q < bins) acc = fmaf(h[ci * bins + q], dpre[co * bins + j], acc);
      }
      gW[idx] += acc;
    }
    for (int idx = threadIdx.x; idx < 3 * HB; idx += blockDim.x) {
      const int co = idx / bins, j = idx - co * bins;
      gpb[j * 3 * H + co] += dpre[idx];
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < HB; i += blockDim.x) d_h0[(size_t)b * HB + i] = dh[i];
  cell_flush(gblob + d.rec_w_off, gW, 3 * H * H * 3);
  cell_flush(gblob + d.rec_pb_off, gpb, bins * 3 * H);
}

// shared memory: acts[enc_act_total] | ga[max_act] | gb[max_act] | gacc[encoder part of the blob]
__global__ void __launch_bounds__(128) cell_encoder_bwd_kernel(const CellDesc d, const float* __restrict__ blob, const float* __restrict__ x,
                                                                const float* __restrict__ prev, int T, const float* __restrict__ d_skips,
                                                                const float* __restrict__ d_gx, float* __restrict__ d_x,
                                                                float* __restrict__ gblob, size_t nframes, int acc_lo, int acc_n) {
  extern __shared__ float sm[];
  float* acts = sm;
  float* ga = acts + d.enc_act_total;
  float* gb = ga + d.max_act;
  float* gacc = gb + d.max_act;
  cell_zero(gacc, acc_n);
  __syncthreads();
  for (size_t f = blockIdx.x; f < nframes; f += gridDim.x) {
    const int t = (int)(f % T);
    const float* xf = x + f * d.n_mels;
    float* a = acts;
    for (int i = threadIdx.x; i < d.n_mels; i += blockDim.x) {
      const float v = xf[i];
      a[i] = v;
      if (d.in_ch == 2) {
        const float pv = (t > 0) ? xf[i - d.n_mels] : (prev ? prev[(f / T) * d.n_mels + i] : v);
        a[d.n_mels + i] = v - pv;
      }
    }
    __syncthreads();
    for (int l = 0; l < d.levels; ++l) {
      const CellLayer& L = d.enc[l];
      float* nx = a + L.cin * L.lin;
      cell_conv(L, blob, a, nx);
      __syncthreads();
      a = nx;
    }
    // a = output of the last layer (the gate tensor)
    const CellLayer& LL = d.enc[d.levels - 1];
    for (int i = threadIdx.x; i < LL.cout * LL.lout; i += blockDim.x) ga[i] = d_gx[f * (size_t)(LL.cout * LL.lout) + i];
    __syncthreads();
    for (int l = d.levels - 1; l >= 0; --l) {
      const CellLayer& L = d.enc[l];
      const float* outp = a;
      float* inp = a - L.cin * L.lin;
      for (int i = threadIdx.x; i < L.cout * L.lout; i += blockDim.x) if (!(outp[i] > 0.f)) ga[i] = 0.f;  // every encoder layer has a ReLU
      __syncthreads();
      cell_conv_bwd(L, blob, inp, ga, gb, gacc + (L.w_off - acc_lo), gacc + (L.pb_off - acc_lo));
      __syncthreads();
      if (l > 0) {  // the input is encoder output l-1, which also fed the decoder as a skip
        const CellLayer& P = d.enc[l - 1];
        const float* ds = d_skips + f * d.skip_stride + P.out_off;
        for (int i = threadIdx.x; i < P.cout * P.lout; i += blockDim.x) ga[i] = gb[i] + ds[i];
      } else {
        // x_t enters as channel 0 and (MOMO3) through the delta channel; the previous frame is detached (momo3.py:278, 287)
        for (int i = threadIdx.x; i < d.n_mels; i += blockDim.x) d_x[f * d.n_mels + i] = gb[i] + (d.in_ch == 2 ? gb[d.n_mels + i] : 0.f);
      }
      __syncthreads();
      a = inp;
    }
  }
  cell_flush(gblob + acc_lo, gacc, acc_n);
}

// (d data-channel weights, d position-bias) -> gradients in the torch parameter layouts (Gaussian-channel weights and bias included)
__device__ __forceinline__ void cell_expand_layer(const CellLayer& L, int G, const float* __restrict__ blob, const float* __restrict__ gblob,
                                                  float* __restrict__ gparams, int tid, int nthreads) {
  const float* gW = gblob + L.w_off;
  const float* gpb = gblob + L.pb_off;
  const float* S = blob + L.smear_off;
  const int CT = L.ct;
  for (int idx = tid; idx < L.cout * CT * L.k; idx += nthreads) {
    int co, c, kk;
    if (!L.transposed) { kk = idx % L.k; c = (idx / L.k) % CT; co = idx / (L.k * CT); }     // [cout][CT][k]
    else               { kk = idx % L.k; co = (idx / L.k) % L.cout; c = idx / (L.k * L.cout); }  // [CT][cout][k]
    float v;
    if (c < L.cin) {
      v = !L.transposed ? gW[((size_t)co * L.cin + c) * L.k + kk] : gW[((size_t)c * L.cout + co) * L.k + kk];
    } else {
      const int g = c - L.cin;
      v = 0.f;
      for (int o = 0; o < L.lout; ++o) {
        int q;
        if (!L.transposed) { q = L.s * o - L.p + kk; if (q < 0 || q >= L.lin) continue; }
        else { const int num = o + L.p - kk; if (num < 0 || num % L.s) continue; q = num / L.s; if (q >= L.lin) continue; }
        v = fmaf(gpb[o * L.cout + co], S[(size_t)g * L.lin + q], v);
      }
    }
    gparams[L.gw_off + idx] = v;
  }
  for (int co = tid; co < L.cout; co += nthreads) {
    float v = 0.f;
    for (int o = 0; o < L.lout; ++o) v += gpb[o * L.cout + co];
    gparams[L.gb_off + co] = v;
  }
}
__global__ void __launch_bounds__(256) cell_expand_grads_kernel(const CellDesc d, const float* __restrict__ blob, const float* __restrict__ gblob,
                                                                 float* __restrict__ gparams) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  for (int l = 0; l < d.levels; ++l) {
    cell_expand_layer(d.enc[l], d.G, blob, gblob, gparams, tid, nt);
    cell_expand_layer(d.dec[l], d.G, blob, gblob, gparams, tid, nt);
  }
  // recurrent conv [3H][H + G][3], stride 1, padding 1, length bins
  const int H = d.H, bins = d.bins, CT = H + d.G;
  const float* gW = gblob + d.rec_w_off;
  const float* gpb = gblob + d.rec_pb_off;
  const float* S = blob + d.rec_smear_off;
  for (int idx = tid; idx < 3 * H * CT * 3; idx += nt) {
    const int kk = idx % 3, c = (idx / 3) % CT, co = idx / (3 * CT);
    float v = 0.f;
    if (c < H) v = gW[((size_t)co * H + c) * 3 + kk];
    else
      for (int j = 0; j < bins; ++j) {
        const int q = j - 1 + kk;
        if (q >= 0 && q < bins) v = fmaf(gpb[j * 3 * H + co], S[(size_t)(c - H) * bins + q], v);
      }
    gparams[d.rec_gw_off + idx] = v;
  }
  for (int co = tid; co < 3 * H; co += nt) {
    float v = 0.f;
    for (int j = 0; j < bins; ++j) v += gpb[j * 3 * H + co];
    gparams[d.rec_gb_off + co] = v;
  }
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
  int gpos = 0;  // running offset into the flat parameter-gradient vector (parameters() order)
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
    L.ct = CT; L.gauss = gauss ? 1 : 0;
    L.gw_off = gpos; gpos += L.cout * CT * L.k;
    L.gb_off = gpos; gpos += L.cout;
    L.smear_off = (int)blob.size();
    if (gauss) for (double v : S) blob.push_back((float)v);
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
    d.rec_gw_off = gpos; gpos += 3 * H * CT * 3;
    d.rec_gb_off = gpos; gpos += 3 * H;
    d.rec_smear_off = (int)blob.size();
    for (double v : S) blob.push_back((float)v);
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
    L.ct = CT; L.gauss = gauss ? 1 : 0;
    L.gw_off = gpos; gpos += CT * L.cout * L.k;
    L.gb_off = gpos; gpos += L.cout;
    L.smear_off = (int)blob.size();
    if (gauss) for (double v : S) blob.push_back((float)v);
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
  d.G = G; d.blob_floats = (int)blob.size(); d.n_param_floats = gpos;
  d.enc_act_total = d.in_ch * d.n_mels;
  for (int l = 0; l < Lv; ++l) d.enc_act_total += d.enc[l].cout * d.enc[l].lout;
  d.dec_act_total = d.H * d.bins;
  for (int i = 0; i < Lv; ++i) d.dec_act_total += (d.dec[i].cout + d.dec[i].skip_c) * d.dec[i].lout;
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

int b2d_cell_num_param_floats(const b2d_cell* m) { return m ? m->d.n_param_floats : B2D_ERR_BAD_ARG; }

// scratch of the backward pass: gradient blob | d_hseq | d_skips | d_gx
size_t b2d_cell_backward_workspace_bytes(const b2d_cell* m, int B, int T) {
  if (!m || B < 1 || T < 1) return 0;
  const size_t nf = (size_t)B * T, HB = (size_t)m->d.H * m->d.bins;
  return align_up((size_t)m->d.blob_floats * 4, 256) + align_up(nf * HB * 4, 256) + align_up(nf * m->d.skip_stride * 4, 256) +
         align_up(nf * 3 * HB * 4, 256);
}

int b2d_cell_backward(const b2d_cell* m, const float* x, const float* prev, const float* hx_in, const void* forward_workspace,
                      const float* grad_out, const float* grad_hx_out, float* grad_x, float* grad_hx_in, float* grad_params, int B, int T,
                      void* workspace, size_t workspace_bytes, void* stream) {
  B2D_REQUIRE(m && x && hx_in && forward_workspace && grad_out && grad_x && grad_hx_in && grad_params && workspace, B2D_ERR_BAD_ARG, "NULL pointer");
  B2D_REQUIRE(B >= 1 && T >= 1 && (long long)B * T < (1ll << 31), B2D_ERR_BAD_ARG, "bad batch / sequence length");
  B2D_REQUIRE(workspace_bytes >= b2d_cell_backward_workspace_bytes(m, B, T), B2D_ERR_WORKSPACE, "cell backward workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const CellDesc& d = m->d;
  const size_t nf = (size_t)B * T, HB = (size_t)d.H * d.bins;
  // what the forward pass left behind (b2d_cell_forward's workspace layout)
  const unsigned char* fb = static_cast<const unsigned char*>(forward_workspace);
  const float* skips = reinterpret_cast<const float*>(fb); fb += align_up(nf * d.skip_stride * 4, 256);
  const float* gx = reinterpret_cast<const float*>(fb); fb += align_up(nf * 3 * HB * 4, 256);
  const float* hseq = reinterpret_cast<const float*>(fb);
  unsigned char* base = static_cast<unsigned char*>(workspace);
  float* gblob = reinterpret_cast<float*>(base); base += align_up((size_t)d.blob_floats * 4, 256);
  float* d_hseq = reinterpret_cast<float*>(base); base += align_up(nf * HB * 4, 256);
  float* d_skips = reinterpret_cast<float*>(base); base += align_up(nf * d.skip_stride * 4, 256);
  float* d_gx = reinterpret_cast<float*>(base);
  B2D_CUDA(cudaMemsetAsync(gblob, 0, (size_t)d.blob_floats * 4, st));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
  const int grid = (int)(nf < (size_t)sms * 4 ? nf : (size_t)sms * 4);
  {  // decoder: accumulators mirror blob[dec[0].w_off, end)
    const int lo = d.dec[0].w_off, n = d.blob_floats - lo;
    const size_t smem = sizeof(float) * (size_t)(d.dec_act_total + 2 * d.max_act + n);
    B2D_REQUIRE(smem <= 220 * 1024, B2D_ERR_UNSUPPORTED, "decoder backward does not fit shared memory for this configuration");
    B2D_SMEM_OPT_IN(smem, cell_decoder_bwd_kernel);
    cell_decoder_bwd_kernel<<<grid, 128, smem, st>>>(d, m->d_blob, hseq, skips, grad_out, d_hseq, d_skips, gblob, nf, lo, n);
    B2D_LAUNCH_CHECK("cell_decoder_bwd_kernel");
  }
  {
    const size_t smem = sizeof(float) * (size_t)(9 * HB + 3 * d.H * d.H * 3 + d.bins * 3 * d.H);
    B2D_REQUIRE(smem <= 220 * 1024, B2D_ERR_UNSUPPORTED, "recurrence backward does not fit shared memory for this configuration");
    B2D_SMEM_OPT_IN(smem, cell_recurrence_bwd_kernel);
    cell_recurrence_bwd_kernel<<<B, 256, smem, st>>>(d, m->d_blob, gx, hx_in, hseq, d_hseq, grad_hx_out, d_gx, grad_hx_in, gblob, T);
    B2D_LAUNCH_CHECK("cell_recurrence_bwd_kernel");
  }
  {  // encoder: accumulators mirror blob[0, rec_w_off)
    const int lo = 0, n = d.rec_w_off;
    const size_t smem = sizeof(float) * (size_t)(d.enc_act_total + 2 * d.max_act + n);
    B2D_REQUIRE(smem <= 220 * 1024, B2D_ERR_UNSUPPORTED, "encoder backward does not fit shared memory for this configuration");
    B2D_SMEM_OPT_IN(smem, cell_encoder_bwd_kernel);
    cell_encoder_bwd_kernel<<<grid, 128, smem, st>>>(d, m->d_blob, x, prev, T, d_skips, d_gx, grad_x, gblob, nf, lo, n);
    B2D_LAUNCH_CHECK("cell_encoder_bwd_kernel");
  }
  cell_expand_grads_kernel<<<16, 256, 0, st>>>(d, m->d_blob, gblob, grad_params);
  B2D_LAUNCH_CHECK("cell_expand_grads_kernel");
  return B2D_OK;
}

}  // extern "C"
