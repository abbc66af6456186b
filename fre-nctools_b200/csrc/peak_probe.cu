// Measured FP64 roofline denominator.  MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only;
// the exchange-grid clip kernel is bound by the FP64 pipe, so bench.py measures the DFMA rate of
// this GPU live (dependent-chain-free register DFMAs, every SM saturated) and reports roofline
// fractions against it.
#include <cuda_runtime.h>
#include "xgrid_plan.h"

namespace xgb {
constexpr int kPeakAcc = 8;
constexpr int kPeakIters = 4096;

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, double a, double b)
{
  double acc[kPeakAcc];
#pragma unroll
  for (int k = 0; k < kPeakAcc; ++k) acc[k] = (double)(threadIdx.x + k);
  for (int it = 0; it < kPeakIters; ++it) {
#pragma unroll
    for (int k = 0; k < kPeakAcc; ++k) acc[k] = fma(acc[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < kPeakAcc; ++k) s += acc[k];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true: keeps the chain alive
}
}  // namespace xgb

// returns 0 and the best-of-5 DFMA throughput in TFLOP/s (2 flops per DFMA)
extern "C" int xgb_fp64_peak_tflops(int device, double* tflops)
{
  using namespace xgb;
  if (cudaSetDevice(device) != cudaSuccess) { xgb_set_error("cudaSetDevice failed"); return 1; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  double* out = nullptr;
  if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) { xgb_set_error("cudaMalloc failed"); return 1; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    dfma_peak_kernel<<<blocks, threads>>>(out, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * kPeakAcc * kPeakIters * (double)blocks * threads;
    if (rep > 0 && ms > 0) { const double t = fl / (ms * 1e-3) * 1e-12; if (t > best) best = t; }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  if (cudaGetLastError() != cudaSuccess) { xgb_set_error("dfma peak probe failed"); return 1; }
  *tflops = best;
  return 0;
}
