// Post-filters of the search path on the device.
//
// After `index.search_and_reconstruct` the reference filters the k reconstructed rows on the host
// (clip_retrieval/clip_back.py:290-324):
//   * dedup  — `get_non_uniques`: FAISS IndexFlatIP over the k rows, `range_search(rows, 0.94)`,
//              connected components of the "inner product > threshold" graph; every member of a
//              component except its first (lowest-index) one is dropped (clip_back.py:270-311);
//   * violence detector — `np.einsum("ij,kj->ik", rows, prompts)`, argmax over the prompts == 1
//              (clip_back.py:321-324).
// The rows are already in HBM when the search returns (`d_R` of b200_index_search_device), k <= 4096
// (the UI asks for 3000, front/src/clip-front.js:45), so this is two tiny kernels:
//   adjacency_kernel   one warp per row i: fp32 dot products with every row j, one bit per pair
//                      (2·k²·d flops — 14 GFLOP at k=3000, d=768; latency-bound);
//   components_kernel  one CTA: min-label propagation with pointer jumping over the bit matrix in
//                      shared memory until no label changes; label[i] != i  <=>  row i is dropped.
// NaN rows (the padding past the last result, clip_back.py:370-378) never compare greater than the
// threshold, so they are isolated nodes and are kept, as in the reference.
#include "common.cuh"

using namespace b200;

namespace {

constexpr int PF_MAX_K = 4096;

// adj[i][w] bit b = (dot(E_i, E_j) > thr), j = 32 w + b.  grid = k blocks of 8 warps; warp = column word stripe.
__global__ void __launch_bounds__(256)
adjacency_kernel(const float* __restrict__ E, int k, int d, float thr, uint32_t* __restrict__ adj, int words) {
  extern __shared__ float s_row[];   // row i
  const int i = blockIdx.x;
  for (int c = threadIdx.x; c < d; c += blockDim.x) s_row[c] = E[(int64_t)i * d + c];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int w = warp; w < words; w += nwarps) {
    uint32_t bits = 0;
    for (int b = 0; b < 32; b++) {
      const int j = w * 32 + b;
      float a = 0.f;
      if (j < k) {
        const float* r = E + (int64_t)j * d;
        for (int c = lane; c < d; c += 32) a = fmaf(s_row[c], r[c], a);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (j < k && a > thr) bits |= 1u << b;
    }
    if (lane == 0) adj[(int64_t)i * words + w] = bits;
  }
}

// One CTA.  label[i] = min index reachable from i.  Each sweep: label[i] = min over neighbours j of
// label[j], then pointer jumping label[i] = label[label[i]]; stop when a sweep changes nothing.
__global__ void __launch_bounds__(1024)
components_kernel(const uint32_t* __restrict__ adj, int k, int words, int32_t* __restrict__ label_out,
                  uint8_t* __restrict__ drop) {
  __shared__ int32_t label[PF_MAX_K];
  __shared__ int changed;
  for (int i = threadIdx.x; i < k; i += blockDim.x) label[i] = i;
  __syncthreads();
  while (true) {
    if (threadIdx.x == 0) changed = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
      int best = label[i];
      for (int w = 0; w < words; w++) {
        // the graph is made symmetric here: an edge exists if either direction passed the threshold
        uint32_t bits = adj[(int64_t)i * words + w];
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          best = min(best, label[w * 32 + b]);
        }
      }
      if (best < label[i]) { atomicMin(&label[i], best); changed = 1; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
      int l = label[i];
      while (label[l] < l) l = label[l];
      if (l < label[i]) { label[i] = l; changed = 1; }
    }
    __syncthreads();
    // push labels back along edges (j adjacent to i gets min too): handles one-directional bits
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
      const int li = label[i];
      for (int w = 0; w < words; w++) {
        uint32_t bits = adj[(int64_t)i * words + w];
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          const int j = w * 32 + b;
          if (li < label[j]) { atomicMin(&label[j], li); changed = 1; }
        }
      }
    }
    __syncthreads();
    const int again = changed;
    __syncthreads();
    if (!again) break;
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    if (label_out) label_out[i] = label[i];
    drop[i] = label[i] != i ? 1 : 0;
  }
}

// flag[i] = (argmax_p dot(E_i, P_p) == target); first maximum wins, as np.argmax.
__global__ void prompt_argmax_kernel(const float* __restrict__ E, int k, int d, const float* __restrict__ P, int np_,
                                     int target, uint8_t* __restrict__ flag) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= k) return;
  float best = -INFINITY;
  int arg = 0;
  for (int p = 0; p < np_; p++) {
    float a = 0.f;
    for (int c = lane; c < d; c += 32) a = fmaf(E[(int64_t)i * d + c], P[(int64_t)p * d + c], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (p == 0 || a > best) { best = a; arg = p; }
  }
  if (lane == 0) flag[i] = arg == target ? 1 : 0;
}

}  // namespace

extern "C" {

int b200_dedup_device(const float* d_rows, int k, int d, float threshold, uint8_t* d_drop, int32_t* d_labels,
                      void* d_workspace, size_t workspace_bytes, int device, void* stream) {
  B200_CHECK(k >= 0 && k <= PF_MAX_K, B200_ERR_UNSUPPORTED, "dedup: k=%d (max %d rows)", k, PF_MAX_K);
  if (k == 0) return B200_OK;
  B200_CHECK(d_rows && d_drop && d_workspace && d >= 1, B200_ERR_INVALID, "dedup: bad argument");
  const int words = (k + 31) / 32;
  B200_CHECK(workspace_bytes >= (size_t)k * words * 4, B200_ERR_INVALID, "dedup: workspace of %zu bytes, need %zu",
             workspace_bytes, (size_t)k * words * 4);
  DeviceGuard g(device);
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* adj = (uint32_t*)d_workspace;
  const size_t smem = (size_t)d * 4;
  if (smem > 48 * 1024)
    B200_CUDA(cudaFuncSetAttribute(adjacency_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  adjacency_kernel<<<k, 256, smem, st>>>(d_rows, k, d, threshold, adj, words);
  B200_LAUNCH_OK();
  components_kernel<<<1, 1024, 0, st>>>(adj, k, words, d_labels, d_drop);
  B200_LAUNCH_OK();
  return B200_OK;
}

int b200_prompt_argmax_device(const float* d_rows, int k, int d, const float* d_prompts, int n_prompts, int target,
                              uint8_t* d_flag, int device, void* stream) {
  if (k == 0) return B200_OK;
  B200_CHECK(d_rows && d_prompts && d_flag && k > 0 && d >= 1 && n_prompts >= 1, B200_ERR_INVALID,
             "prompt_argmax: bad argument");
  DeviceGuard g(device);
  prompt_argmax_kernel<<<(k + 7) / 8, 256, 0, (cudaStream_t)stream>>>(d_rows, k, d, d_prompts, n_prompts, target, d_flag);
  B200_LAUNCH_OK();
  return B200_OK;
}

}  // extern "C"
