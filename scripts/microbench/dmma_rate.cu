// Microbenchmark for the north-star's tensor-core question: what is the FP64 tensor (DMMA) rate of B200 against the
// FP64 vector pipe (DFMA), and do the two overlap when a kernel issues both?
//   variant 0: DFMA only (8 independent chains per thread, one multiplicand uniform)
//   variant 1: DMMA only (mma.sync.aligned.m8n8k4.row.col.f64: 8x8x4 = 256 FMA per warp instruction, 8 accumulator
//              tiles per warp)
//   variant 2: both in every warp, interleaved 1 DMMA : 8 DFMA (equal FMA counts on both pipes per trip)
//   variant 3: half of the warps DFMA only, the other half DMMA only
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_rate dmma_rate.cu ; run on a GPU box; under ncu
// the metrics sm__inst_executed_pipe_fp64 / sm__pipe_fp64_cycles_active / sm__inst_executed_pipe_tensor* show which
// pipe the DMMA occupies.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int VARIANT>
__global__ void __launch_bounds__(256) probe(double* out, int iters, double x, double y) {
  double acc[8], t0[8], t1[8];
  const double yv = y * (double)(threadIdx.x + 1);
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[j] = 1.0 + 1e-3 * (threadIdx.x + j); t0[j] = 1e-3 * j; t1[j] = 2e-3 * j; }
  const bool vec_warp = (VARIANT == 0) || (VARIANT == 2) || (VARIANT == 3 && ((threadIdx.x >> 5) & 1) == 0);
  const bool ten_warp = (VARIANT == 1) || (VARIANT == 2) || (VARIANT == 3 && ((threadIdx.x >> 5) & 1) == 1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (ten_warp) dmma(t0[r], t1[r], a, b);               // 256 FMA per warp
      if (vec_warp) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fma(acc[j], x, yv);     // 8 x 32 = 256 FMA per warp
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += acc[j] + t0[j] + t1[j];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int VARIANT>
void run(const char* name, double* d, int blocks) {
  const int iters = 2048;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    probe<VARIANT><<<blocks, 256>>>(d, iters, 0.999999, 1e-7);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double warps = (double)blocks * 8;
  const double per_trip = 8.0 * 256.0;                       // FMA per warp and trip on each active pipe
  double vec = (VARIANT == 0 || VARIANT == 2) ? warps : (VARIANT == 3 ? warps / 2 : 0);
  double ten = (VARIANT == 1 || VARIANT == 2) ? warps : (VARIANT == 3 ? warps / 2 : 0);
  printf("%-44s %8.3f ms   vector %6.2f TFLOP/s   tensor %6.2f TFLOP/s   total %6.2f\n", name, best,
         2 * vec * per_trip * iters / best / 1e9, 2 * ten * per_trip * iters / best / 1e9,
         2 * (vec + ten) * per_trip * iters / best / 1e9);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8;
  double* d; cudaMalloc(&d, (size_t)blocks * 256 * 8);
  run<0>("DFMA only", d, blocks);
  run<1>("DMMA m8n8k4 only", d, blocks);
  run<2>("DFMA + DMMA interleaved in every warp", d, blocks);
  run<3>("DFMA warps next to DMMA warps", d, blocks);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
