// Measurement utility (used by bench.py only): sustained FP32 FMA throughput of the device, the
// denominator of the CUDA-core roofline.  MEASURED_PEAKS.json carries HBM and bf16 tensor peaks but no
// FP32 CUDA-core figure, and this path deliberately uses no tensor cores (matrices <= 32x32).
#include "common.cuh"

namespace {
__global__ void __launch_bounds__(256) k_fma_peak(float *out, int iters, float b, float c) {
  float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
      a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
}  // namespace

extern "C" int tfmpc_measure_fp32_peak(double *tflops, double *ms_out) {
  if (!tflops) return tfmpc_set_error(TFMPC_E_INVALID, "null argument");
  int dev = 0, sms = 148;
  CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * 8, threads = 256, iters = 4096;
  float *buf = nullptr;
  CUDA_TRY(cudaMalloc((void **)&buf, (size_t)blocks * threads * sizeof(float)));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  double best = 0, best_ms = 0;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0, 0);
    k_fma_peak<<<blocks, threads>>>(buf, iters, 1.0000001f, 1e-7f);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 64.0 * iters * (double)blocks * threads;
    double tf = fl / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) { best = tf; best_ms = ms; }
  }
  tfmpc_count_launch(6);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf);
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) return tfmpc_set_error(TFMPC_E_CUDA, "fp32 peak kernel: %s", cudaGetErrorString(le));
  *tflops = best;
  if (ms_out) *ms_out = best_ms;
  return TFMPC_OK;
}
