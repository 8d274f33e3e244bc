// Microbenchmark: does a non-FP64 instruction issued between FP64 instructions cost FP64 throughput on B200?
// Each variant runs 8 independent DFMA chains per thread plus `MIX` extra instructions of another class per
// DFMA.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue fp64_issue.cu ; run on a GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND, int MIX>
__global__ void __launch_bounds__(256) probe(double* out, int* iout, int iters, double x, double y, int seed) {
  double acc[8];
  int ia[8];
  float fa[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[j] = 1.0 + 1e-3 * (threadIdx.x + j); ia[j] = seed + threadIdx.x * (j + 1); fa[j] = 1.0f + j; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] = fma(acc[j], x, y);
#pragma unroll
        for (int m = 0; m < MIX; ++m) {
          if (KIND == 1) ia[j] = (ia[j] ^ (ia[(j + 1) & 7] >> 3)) + 0x9e3779b9;      // LOP3 / IADD (ALU pipe)
          if (KIND == 2) fa[j] = fmaf(fa[j], 0.999f, 0.5f);                         // FFMA (FMA pipe)
          if (KIND == 3) ia[j] = ia[j] * 3 + ia[(j + 3) & 7];                         // IMAD (FMA pipe)
        }
      }
    }
  }
  double s = 0.0; int is = 0; float fs = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { s += acc[j]; is += ia[j]; fs += fa[j]; }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s + fs;
  iout[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = is;
}

template <int KIND, int MIX>
void run(const char* name, double* d, int* di, int blocks) {
  const int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    probe<KIND, MIX><<<blocks, 256>>>(d, di, iters, 0.999999, 1e-7, 12345);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  double dfma = (double)blocks * 256 * iters * 64;
  printf("%-28s %8.3f ms  DFMA rate %6.2f TFLOP/s  (extra instr per DFMA: %d)\n", name, best, 2 * dfma / best / 1e9, MIX);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8;
  double* d; int* di;
  cudaMalloc(&d, (size_t)blocks * 256 * 8); cudaMalloc(&di, (size_t)blocks * 256 * 4);
  run<0, 0>("DFMA only", d, di, blocks);
  run<1, 1>("DFMA + 1 ALU (LOP3/IADD)", d, di, blocks);
  run<1, 2>("DFMA + 2 ALU", d, di, blocks);
  run<2, 1>("DFMA + 1 FFMA", d, di, blocks);
  run<2, 2>("DFMA + 2 FFMA", d, di, blocks);
  run<3, 1>("DFMA + 1 IMAD", d, di, blocks);
  run<3, 2>("DFMA + 2 IMAD", d, di, blocks);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
