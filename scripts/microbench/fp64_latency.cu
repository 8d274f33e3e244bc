// Microbenchmark 2: FP64 pipe utilisation vs resident warps per SM for dependent chains (ILP = CHAINS per
// thread), optionally with a dependent ALU-pipe select after every DFMA.
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS, int SEL>
__global__ void probe(double* out, int iters, double x, double y, double z) {
  double acc[CHAINS];
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) acc[j] = 1.0 + 1e-3 * (threadIdx.x + j);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
#pragma unroll
      for (int j = 0; j < CHAINS; ++j) {
        acc[j] = fma(acc[j], x, y);
        if (SEL) acc[j] = (acc[j] > z) ? acc[j] : -acc[j];      // DSETP + select dependent on the DFMA
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) s += acc[j];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS, int SEL>
void run(double* d, int sms) {
  for (int warps = 4; warps <= 32; warps += (warps < 16 ? 4 : 8)) {
    const int iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      probe<CHAINS, SEL><<<sms, warps * 32>>>(d, iters, 0.999999, 1e-7, -5.0);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    double dfma = (double)sms * warps * 32 * iters * 16 * CHAINS;
    printf("chains=%d sel=%d warps/SM=%2d  %7.3f ms  %6.2f TFLOP/s (DFMA only)\n", CHAINS, SEL, warps, best, 2 * dfma / best / 1e9);
  }
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* d; cudaMalloc(&d, (size_t)sms * 1024 * 8);
  run<1, 0>(d, sms); run<2, 0>(d, sms); run<4, 0>(d, sms);
  run<1, 1>(d, sms); run<2, 1>(d, sms);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
