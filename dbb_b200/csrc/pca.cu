// "Next" row N4: PCA of the signature matrix, pca.py:56-75 (PCA.pcaMatrix) + :113-126 (Center).
//
// Reference: A -= A.mean(0); optionally A /= A.std(0) (std 0 -> 1); U, d, Vt = svd(A); variance = d^2 / sum(d^2);
// pc() = U[:, :npc] * d[:npc] (= A . Vt[:npc].T).  The only dense contraction is the W x W Gram matrix of the centred
// data (W = 136 columns): with N rows it costs 2 N W^2 flops against N W 8 bytes of input, i.e. 34 flops per byte, and
// the whole job is ~2 GFLOP for 50 000 scaffolds -- done here in FP64 on the CUDA cores (same precision as the
// reference; a tensor-core version would have to be tf32 and buys microseconds).  The W x W symmetric
// eigenproblem is solved on the host (numpy), the scores are one more pass on the GPU.
//   pass 1: column sums                      -> mean
//   pass 2: column sums of squared deviations -> std (population, as numpy's .std())      [only if scaling]
//   pass 3: Gram matrix of (A - mean) / std
//   pass 4: scores = ((A - mean) / std) . V[:, :npc]
#include "common.cuh"
#include <algorithm>

namespace {

constexpr int SLAB = 32;                 // rows staged per iteration

// out[c] += sum over rows of f(A[r][c]); f = identity (mean == nullptr) or squared deviation from mean
__global__ void __launch_bounds__(256) col_reduce_kernel(const double* __restrict__ A, int64_t n, int w, const double* __restrict__ mean,
                                                         double* __restrict__ out) {
  const int c = threadIdx.x;
  if (c >= w) return;
  double acc = 0.0;
  const double m = mean ? mean[c] : 0.0;
  for (int64_t r = blockIdx.x; r < n; r += gridDim.x) {
    const double v = A[r * w + c] - m;
    acc += mean ? v * v : v;
  }
  atomicAdd(out + c, acc);
}

// gram[i][j] += sum_r x[r][i] x[r][j], x = (A - mean) * inv_std; one CTA per group of row slabs, thread t owns the
// pairs (i, j) = t, t + blockDim, ... of the w x w matrix (row-major); slabs are staged (centred, scaled) in shared memory
__global__ void __launch_bounds__(256) gram_kernel(const double* __restrict__ A, int64_t n, int w, const double* __restrict__ mean,
                                                   const double* __restrict__ inv_std, double* __restrict__ gram) {
  extern __shared__ double xs[];                       // [SLAB][w]
  const int total = w * w;
  constexpr int MAXP = 80;                             // pairs per thread: ceil(144 * 144 / 256) = 81 > 136^2 / 256 = 72.25
  double acc[MAXP];
  #pragma unroll
  for (int p = 0; p < MAXP; ++p) acc[p] = 0.0;
  for (int64_t r0 = (int64_t)blockIdx.x * SLAB; r0 < n; r0 += (int64_t)gridDim.x * SLAB) {
    const int rows = (int)min((int64_t)SLAB, n - r0);
    __syncthreads();
    for (int e = threadIdx.x; e < rows * w; e += blockDim.x) {
      const int c = e % w;
      xs[e] = (A[r0 * w + e] - mean[c]) * inv_std[c];
    }
    __syncthreads();
    #pragma unroll
    for (int p = 0; p < MAXP; ++p) {
      const int idx = threadIdx.x + p * 256;
      if (idx < total) {
        const int i = idx / w, j = idx - i * w;
        double s = 0.0;
        for (int r = 0; r < rows; ++r) s += xs[r * w + i] * xs[r * w + j];
        acc[p] += s;
      }
    }
  }
  #pragma unroll
  for (int p = 0; p < MAXP; ++p) {
    const int idx = threadIdx.x + p * 256;
    if (idx < total) atomicAdd(gram + idx, acc[p]);
  }
}

// scores[r][k] = sum_c x[r][c] V[c][k]; one warp per row, lanes over columns, k looped
__global__ void __launch_bounds__(256) project_kernel(const double* __restrict__ A, int64_t n, int w, const double* __restrict__ mean,
                                                      const double* __restrict__ inv_std, const double* __restrict__ V, int npc,
                                                      double* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n) return;
  double x[8];
  #pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int c = lane + 32 * q;
    x[q] = c < w ? (A[r * w + c] - mean[c]) * inv_std[c] : 0.0;
  }
  for (int k = 0; k < npc; ++k) {
    double s = 0.0;
    #pragma unroll
    for (int q = 0; q < 8; ++q) { const int c = lane + 32 * q; if (c < w) s += x[q] * V[(int64_t)c * npc + k]; }
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) scores[r * npc + k] = s;
  }
}

}  // namespace

// Host-buffer entry points (the data is a TSV-sized matrix; everything is copied once).
DBB_EXPORT int dbb_pca_gram_host(const double* A, int64_t n, int32_t w, int32_t scale, double* mean, double* std_out, double* gram) {
  DBB_REQUIRE(A && mean && std_out && gram, "NULL argument");
  DBB_REQUIRE(n > 0 && w > 0 && w <= 143, "bad shape (1..143 columns: 80 matrix entries per thread of the Gram kernel)");
  cudaStream_t st;
  DBB_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  double *dA = nullptr, *dm = nullptr, *ds = nullptr, *dg = nullptr;
  const size_t bytesA = (size_t)n * w * sizeof(double);
  int rc = DBB_OK;
  cudaError_t e = cudaSuccess;
  do {
    if ((e = cudaMalloc(&dA, bytesA)) != cudaSuccess || (e = cudaMalloc(&dm, w * 8)) != cudaSuccess ||
        (e = cudaMalloc(&ds, w * 8)) != cudaSuccess || (e = cudaMalloc(&dg, (size_t)w * w * 8)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(dA, A, bytesA, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    cudaMemsetAsync(dm, 0, w * 8, st); cudaMemsetAsync(ds, 0, w * 8, st); cudaMemsetAsync(dg, 0, (size_t)w * w * 8, st);
    const unsigned grid = (unsigned)std::min<int64_t>(n, 148 * 4);
    col_reduce_kernel<<<grid, 256, 0, st>>>(dA, n, w, nullptr, dm);
    if ((e = cudaMemcpyAsync(mean, dm, w * 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
    for (int c = 0; c < w; ++c) mean[c] /= (double)n;                              // A.mean(axis=0)
    if ((e = cudaMemcpyAsync(dm, mean, w * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    for (int c = 0; c < w; ++c) std_out[c] = 1.0;
    if (scale) {
      col_reduce_kernel<<<grid, 256, 0, st>>>(dA, n, w, dm, ds);
      if ((e = cudaMemcpyAsync(std_out, ds, w * 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
      if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
      for (int c = 0; c < w; ++c) { const double s = sqrt(std_out[c] / (double)n); std_out[c] = s != 0.0 ? s : 1.0; }   // pca.py:121-122
    }
    double inv[256];
    for (int c = 0; c < w; ++c) inv[c] = 1.0 / std_out[c];
    if ((e = cudaMemcpyAsync(ds, inv, w * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    const size_t smem = (size_t)SLAB * w * sizeof(double);
    const unsigned ggrid = (unsigned)std::min<int64_t>((n + SLAB - 1) / SLAB, 148 * 2);
    gram_kernel<<<ggrid, 256, smem, st>>>(dA, n, w, dm, ds, dg);
    if ((e = cudaGetLastError()) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(gram, dg, (size_t)w * w * 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    e = cudaStreamSynchronize(st);
  } while (0);
  if (e != cudaSuccess) { dbb::set_error("dbb_pca_gram_host failed: %s", cudaGetErrorString(e)); rc = DBB_ERR_CUDA; }
  cudaFree(dA); cudaFree(dm); cudaFree(ds); cudaFree(dg);
  cudaStreamDestroy(st);
  return rc;
}

DBB_EXPORT int dbb_pca_project_host(const double* A, int64_t n, int32_t w, const double* mean, const double* std_in, const double* V,
                                    int32_t npc, double* scores) {
  DBB_REQUIRE(A && mean && std_in && V && scores, "NULL argument");
  DBB_REQUIRE(n > 0 && w > 0 && w <= 256 && npc > 0 && npc <= w, "bad shape");
  cudaStream_t st;
  DBB_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  double *dA = nullptr, *dm = nullptr, *ds = nullptr, *dV = nullptr, *dS = nullptr;
  int rc = DBB_OK;
  cudaError_t e = cudaSuccess;
  do {
    if ((e = cudaMalloc(&dA, (size_t)n * w * 8)) != cudaSuccess || (e = cudaMalloc(&dm, w * 8)) != cudaSuccess ||
        (e = cudaMalloc(&ds, w * 8)) != cudaSuccess || (e = cudaMalloc(&dV, (size_t)w * npc * 8)) != cudaSuccess ||
        (e = cudaMalloc(&dS, (size_t)n * npc * 8)) != cudaSuccess) break;
    double inv[256];
    for (int c = 0; c < w; ++c) inv[c] = 1.0 / std_in[c];
    if ((e = cudaMemcpyAsync(dA, A, (size_t)n * w * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(dm, mean, w * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(ds, inv, w * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(dV, V, (size_t)w * npc * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    const int64_t grid = (n * 32 + 255) / 256;
    project_kernel<<<(unsigned)grid, 256, 0, st>>>(dA, n, w, dm, ds, dV, npc, dS);
    if ((e = cudaGetLastError()) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(scores, dS, (size_t)n * npc * 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    e = cudaStreamSynchronize(st);
  } while (0);
  if (e != cudaSuccess) { dbb::set_error("dbb_pca_project_host failed: %s", cudaGetErrorString(e)); rc = DBB_ERR_CUDA; }
  cudaFree(dA); cudaFree(dm); cudaFree(ds); cudaFree(dV); cudaFree(dS);
  cudaStreamDestroy(st);
  return rc;
}
