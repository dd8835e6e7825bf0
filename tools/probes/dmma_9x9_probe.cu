// dmma_9x9_probe.cu -- measurement behind the "tensor cores for the SPARTACUS matrix exponential?" decision.
//
// The matrix exponential of solver_spartacus_sw/lw (radiation_matrix.F90:805-903 expm, :145 mat_x_mat) is a chain of 9x9 (SW)
// or 6x6 (LW) fp64 products per (layer, g-point).  The only fp64 tensor instruction is DMMA m8n8k4 (mma.sync.aligned.m8n8k4.f64);
// a 9x9 operand has to be padded to 16x16 (four 8x8 output tiles x four k-steps = 16 DMMA per product, 18 % useful flops).
// This probe times a batch of chained products  X <- A * X  (the shape of the repeated squaring / Pade products) three ways:
//   fma_thread   one thread per matrix, 729 FMAs in registers / local memory        (the current layer kernels)
//   fma_9lanes   nine lanes per matrix: lane j holds column j, A broadcast from shared memory   (the planned redesign)
//   dmma_pad16   one warp per matrix, operands padded to 16x16 in shared memory, 16 DMMA per product
// and prints the time per product, the fp64 rate on useful flops and the largest element difference from the first variant.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_9x9_probe dmma_9x9_probe.cu && ./dmma_9x9_probe
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#define N 9
#define CHAIN 8          // products per matrix (an expm does 6-15)

__global__ void fma_thread(const double* __restrict__ A, double* __restrict__ X, int nmat) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nmat) return;
  double a[N * N], x[N * N], y[N * N];
  for (int i = 0; i < N * N; ++i) { a[i] = A[(size_t)i * nmat + m]; x[i] = X[(size_t)i * nmat + m]; }
  for (int it = 0; it < CHAIN; ++it) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) s = fma(a[i * N + k], x[k * N + j], s);
        y[i * N + j] = s;
      }
    for (int i = 0; i < N * N; ++i) x[i] = y[i];
  }
  for (int i = 0; i < N * N; ++i) X[(size_t)i * nmat + m] = x[i];
}

// 3 matrices per warp (27 lanes); lane j of a matrix keeps column j of X in registers, A sits in shared memory (row-major)
__global__ void fma_9lanes(const double* __restrict__ A, double* __restrict__ X, int nmat) {
  __shared__ double sA[8][3][N * N];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / N, j = lane - sub * N;
  const int m = (blockIdx.x * 8 + warp) * 3 + sub;
  const bool on = sub < 3 && m < nmat;
  double x[N], y[N];
  if (on) {
    for (int i = 0; i < N; ++i) { sA[warp][sub][i * N + j] = A[(size_t)(i * N + j) * nmat + m]; x[i] = X[(size_t)(i * N + j) * nmat + m]; }
  }
  __syncwarp();
  if (on) {
    const double* a = sA[warp][sub];
    for (int it = 0; it < CHAIN; ++it) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) s = fma(a[i * N + k], x[k], s);
        y[i] = s;
      }
#pragma unroll
      for (int i = 0; i < N; ++i) x[i] = y[i];
    }
    for (int i = 0; i < N; ++i) X[(size_t)(i * N + j) * nmat + m] = x[i];
  }
}

// one warp per matrix; A and X padded to 16x16 in shared memory; C = A * X by 2x2 output tiles x 4 k-steps of DMMA m8n8k4
__global__ void dmma_pad16(const double* __restrict__ A, double* __restrict__ X, int nmat) {
  __shared__ double sA[8][16 * 16], sX[8][16 * 16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + warp;
  if (m >= nmat) return;
  double* a = sA[warp]; double* x = sX[warp];
  for (int i = lane; i < 256; i += 32) { a[i] = 0.0; x[i] = 0.0; }
  __syncwarp();
  for (int i = lane; i < N * N; i += 32) { const int r = i / N, c = i - r * N; a[r * 16 + c] = A[(size_t)i * nmat + m]; x[r * 16 + c] = X[(size_t)i * nmat + m]; }
  __syncwarp();
  const int g = lane >> 2, t = lane & 3;
  for (int it = 0; it < CHAIN; ++it) {
    double c[2][2][2];
#pragma unroll
    for (int ti = 0; ti < 2; ++ti)
#pragma unroll
      for (int tj = 0; tj < 2; ++tj) {
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const double af = a[(8 * ti + g) * 16 + 4 * ks + t];      // A fragment: row g, column t of the 8x4 tile
          const double bf = x[(4 * ks + t) * 16 + 8 * tj + g];      // B fragment: row t, column g of the 4x8 tile
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(af), "d"(bf));
        }
        c[ti][tj][0] = d0; c[ti][tj][1] = d1;
      }
    __syncwarp();
#pragma unroll
    for (int ti = 0; ti < 2; ++ti)
#pragma unroll
      for (int tj = 0; tj < 2; ++tj) { x[(8 * ti + g) * 16 + 8 * tj + 2 * t] = c[ti][tj][0]; x[(8 * ti + g) * 16 + 8 * tj + 2 * t + 1] = c[ti][tj][1]; }
    __syncwarp();
  }
  for (int i = lane; i < N * N; i += 32) { const int r = i / N, cc = i - r * N; X[(size_t)i * nmat + m] = x[r * 16 + cc]; }
}

int main() {
  const int nmat = 148 * 2048;
  const size_t n = (size_t)nmat * N * N;
  double *hA = (double*)malloc(n * 8), *hX = (double*)malloc(n * 8), *h0 = (double*)malloc(n * 8), *h1 = (double*)malloc(n * 8);
  srand(1);
  for (size_t i = 0; i < n; ++i) { hA[i] = 0.3 * (rand() / (double)RAND_MAX - 0.5); hX[i] = rand() / (double)RAND_MAX - 0.5; }
  double *dA, *dX;
  cudaMalloc(&dA, n * 8); cudaMalloc(&dX, n * 8);
  cudaMemcpy(dA, hA, n * 8, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[3] = {"fma_thread", "fma_9lanes", "dmma_pad16"};
  for (int v = 0; v < 3; ++v) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaMemcpy(dX, hX, n * 8, cudaMemcpyHostToDevice);
      cudaEventRecord(e0);
      if (v == 0) fma_thread<<<(nmat + 127) / 128, 128>>>(dA, dX, nmat);
      else if (v == 1) fma_9lanes<<<(nmat + 23) / 24, 256>>>(dA, dX, nmat);
      else dmma_pad16<<<(nmat + 7) / 8, 256>>>(dA, dX, nmat);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    cudaMemcpy(v == 0 ? h0 : h1, dX, n * 8, cudaMemcpyDeviceToHost);
    double diff = 0.0, big = 0.0;
    if (v) for (size_t i = 0; i < n; ++i) { diff = fmax(diff, fabs(h1[i] - h0[i])); big = fmax(big, fabs(h0[i])); }
    const double useful = 2.0 * N * N * N * CHAIN * (double)nmat;
    printf("%-11s %8.3f ms  %7.2f ns per 9x9 product per SM-slot  %6.2f TFLOP/s on useful flops  max |diff vs fma_thread| %.3e (max |x| %.3e)  %s\n", names[v], best,
           best * 1e6 / ((double)nmat * CHAIN), useful / (best * 1e-3) / 1e12, diff, big, cudaGetErrorString(err));
  }
  return 0;
}
