// FP64 issue-rate microbenchmark for sm_100a: DFMA throughput per SM sub-partition as a function of resident warps,
// independent chains per thread (ILP) and operand kinds.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -o fp64_issue fp64_issue.cu ; run on the GPU box.  Prints DFMA per cycle per SMSP (peak 0.5 if the pipe retires one
// warp instruction every 2 cycles).
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int MODE>
__global__ void k(double *out, const double *in, int iters) {
  double a[ILP], b[ILP], c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { a[i] = in[threadIdx.x + i]; b[i] = in[64 + threadIdx.x + i]; c[i] = in[128 + threadIdx.x + i]; }
  const double k0 = in[0], k1 = in[1];
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        if (MODE == 0) a[i] = fma(a[i], 1.0000001, 1e-9);            // R, imm, imm
        if (MODE == 1) a[i] = fma(a[i], k0, k1);                     // R, R(shared), R(shared)
        if (MODE == 2) a[i] = fma(a[i], b[i], c[i]);                 // R, R, R all distinct
        if (MODE == 3) a[i] = fma(b[i], c[i], a[i]);                 // accumulate form
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += a[i];
  if (s == 123.456) out[0] = s;
}

template <int ILP, int MODE>
void run(double *out, double *in, int sms, double mhz) {
  for (int wps = 1; wps <= 8; wps *= 2) {      // warps per SMSP
    const int threads = 128 * wps > 1024 ? 1024 : 128 * wps;
    const int blocks_per_sm = (128 * wps) / threads;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    k<ILP, MODE><<<sms * blocks_per_sm, threads>>>(out, in, 10);
    cudaEventRecord(e0);
    k<ILP, MODE><<<sms * blocks_per_sm, threads>>>(out, in, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * mhz * 1e6;
    const double dfma_per_smsp = (double)iters * 8 * ILP * wps;       // warp instructions per SMSP
    printf("mode %d ilp %2d warps/SMSP %d : %.3f DFMA/cycle/SMSP (%.2f ms)\n", MODE, ILP, wps, dfma_per_smsp / cycles, ms);
  }
}

int main() {
  int dev = 0, sms, khz;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  double mhz = khz / 1000.0;
  printf("SMs %d clock %.0f MHz\n", sms, mhz);
  double *out, *in;
  cudaMalloc(&out, 8);
  cudaMalloc(&in, 4096 * 8);
  cudaMemset(in, 0, 4096 * 8);
  run<1, 0>(out, in, sms, mhz); run<2, 0>(out, in, sms, mhz); run<4, 0>(out, in, sms, mhz); run<8, 0>(out, in, sms, mhz);
  run<4, 1>(out, in, sms, mhz); run<8, 1>(out, in, sms, mhz);
  run<1, 2>(out, in, sms, mhz); run<2, 2>(out, in, sms, mhz); run<4, 2>(out, in, sms, mhz); run<8, 2>(out, in, sms, mhz); run<10, 2>(out, in, sms, mhz);
  run<4, 3>(out, in, sms, mhz); run<8, 3>(out, in, sms, mhz);
  return 0;
}
