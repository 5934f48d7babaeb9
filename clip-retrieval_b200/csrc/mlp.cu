// Dense fp32 MLP head on the reconstructed rows of a search (the last post-filter):
// the H14 NSFW detector the reference runs on the CPU per request — h14_nsfw_model.py:15-34
// (nn.Sequential of 7 Linear layers 1024-1024-2048-1024-256-128-16-1, ReLU between the first five,
// Dropout = identity in eval) called from KnnService.get_unsafe_items (clip_back.py:315-319).
//
// y = relu?(x · W^T + b) per layer, fp32 operands and fp32 accumulation like the reference's torch CPU
// path (the 0.5 threshold on the logit must not move with a lower-precision product).  k <= a few
// thousand rows x <= 2048 columns: 2 * 5.5 MFLOP per row, latency-bound; a 64x64x16 shared-memory tiled
// FMA kernel per layer, ping-pong activations inside the handle.
#include "common.cuh"
#include <mutex>
#include <vector>

struct b200_mlp {
  int device = 0;
  std::vector<int> dims;          // [L + 1]
  std::vector<uint8_t> relu;      // [L]
  std::vector<float*> w, b;       // device, W [out, in] row-major, b [out]
  std::vector<bool> loaded;
  float* act[2] = {nullptr, nullptr};
  size_t act_rows = 0;
  std::mutex mu;
};

namespace b200 {

constexpr int MLP_BM = 64, MLP_BN = 64, MLP_BK = 16;

// Y[n, N] = act(X[n, K] · W[N, K]^T + bias)
__global__ void __launch_bounds__(256)
mlp_layer_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ bias,
                 float* __restrict__ Y, int n, int N, int K, int relu) {
  __shared__ float sX[MLP_BK][MLP_BM + 4];
  __shared__ float sW[MLP_BK][MLP_BN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads, 4 x 4 outputs each
  const int m0 = blockIdx.y * MLP_BM, n0 = blockIdx.x * MLP_BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += MLP_BK) {
    // 64 rows x 16 columns of each operand: 1024 elements, 4 per thread
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int idx = threadIdx.x + e * 256;
      const int r = idx >> 4, c = idx & 15;
      const int gm = m0 + r, gn = n0 + r, gk = k0 + c;
      sX[c][r] = (gm < n && gk < K) ? X[(int64_t)gm * K + gk] : 0.f;
      sW[c][r] = (gn < N && gk < K) ? W[(int64_t)gn * K + gk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < MLP_BK; kk++) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = sX[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; j++) bb[j] = sW[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j] + (bias ? bias[gn] : 0.f);
      if (relu) v = fmaxf(v, 0.f);
      Y[(int64_t)gm * N + gn] = v;
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_mlp_create(int n_layers, const int32_t* dims, const uint8_t* relu, int device, b200_mlp** out) {
  B200_CHECK(out && dims && relu && n_layers >= 1 && n_layers <= 64, B200_ERR_INVALID, "mlp_create: bad argument");
  for (int i = 0; i <= n_layers; i++)
    B200_CHECK(dims[i] >= 1 && dims[i] <= 65536, B200_ERR_INVALID, "mlp_create: dims[%d]=%d", i, dims[i]);
  int ndev = 0;
  B200_CUDA(cudaGetDeviceCount(&ndev));
  B200_CHECK(device >= 0 && device < ndev, B200_ERR_INVALID, "mlp_create: device %d of %d", device, ndev);
  DeviceGuard g(device);
  b200_mlp* m = new (std::nothrow) b200_mlp();
  B200_CHECK(m != nullptr, B200_ERR_OOM, "mlp_create: host allocation failed");
  m->device = device;
  m->dims.assign(dims, dims + n_layers + 1);
  m->relu.assign(relu, relu + n_layers);
  m->w.assign(n_layers, nullptr);
  m->b.assign(n_layers, nullptr);
  m->loaded.assign(n_layers, false);
  for (int l = 0; l < n_layers; l++) {
    if (cudaMalloc((void**)&m->w[l], (size_t)dims[l] * dims[l + 1] * 4) != cudaSuccess ||
        cudaMalloc((void**)&m->b[l], (size_t)dims[l + 1] * 4) != cudaSuccess) {
      set_error("mlp_create: device allocation failed");
      b200_mlp_destroy(m);
      return B200_ERR_OOM;
    }
  }
  *out = m;
  return B200_OK;
}

int b200_mlp_destroy(b200_mlp* m) {
  if (!m) return B200_OK;
  DeviceGuard g(m->device);
  cudaDeviceSynchronize();
  for (float* p : m->w) if (p) cudaFree(p);
  for (float* p : m->b) if (p) cudaFree(p);
  for (float* p : m->act) if (p) cudaFree(p);
  delete m;
  return B200_OK;
}

int b200_mlp_load_layer(b200_mlp* m, int layer, const float* h_weight, const float* h_bias) {
  B200_CHECK(m && h_weight && h_bias && layer >= 0 && layer < (int)m->w.size(), B200_ERR_INVALID, "mlp_load_layer: bad argument");
  DeviceGuard g(m->device);
  B200_CUDA(cudaMemcpy(m->w[layer], h_weight, (size_t)m->dims[layer] * m->dims[layer + 1] * 4, cudaMemcpyHostToDevice));
  B200_CUDA(cudaMemcpy(m->b[layer], h_bias, (size_t)m->dims[layer + 1] * 4, cudaMemcpyHostToDevice));
  m->loaded[layer] = true;
  return B200_OK;
}

int b200_mlp_forward_device(b200_mlp* m, const float* d_x, int n, float* d_y, void* stream) {
  B200_CHECK(m && n >= 0, B200_ERR_INVALID, "mlp_forward: bad argument");
  if (n == 0) return B200_OK;
  B200_CHECK(d_x && d_y, B200_ERR_INVALID, "mlp_forward: null buffer");
  for (size_t l = 0; l < m->loaded.size(); l++)
    B200_CHECK(m->loaded[l], B200_ERR_STATE, "mlp_forward: layer %zu has no weights", l);
  std::lock_guard<std::mutex> lock(m->mu);
  DeviceGuard g(m->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int L = (int)m->w.size();
  int widest = 0;
  for (int l = 1; l < L; l++) widest = std::max(widest, m->dims[l]);
  if (L > 1 && m->act_rows < (size_t)n) {
    B200_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 2; i++) {
      if (m->act[i]) B200_CUDA(cudaFree(m->act[i]));
      m->act[i] = nullptr;
      B200_CUDA(cudaMalloc((void**)&m->act[i], (size_t)n * widest * 4));
    }
    m->act_rows = (size_t)n;
  }
  const float* in = d_x;
  for (int l = 0; l < L; l++) {
    float* o = l == L - 1 ? d_y : m->act[l & 1];
    const int K = m->dims[l], N = m->dims[l + 1];
    dim3 grid((N + MLP_BN - 1) / MLP_BN, (n + MLP_BM - 1) / MLP_BM);
    mlp_layer_kernel<<<grid, 256, 0, st>>>(in, m->w[l], m->b[l], o, n, N, K, m->relu[l]);
    B200_LAUNCH_OK();
    in = o;
  }
  return B200_OK;
}

}  // extern "C"
