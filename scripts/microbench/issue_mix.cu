// Micro-benchmark: does a non-FP64 instruction issue in the shadow of a DFMA's second dispatch cycle on sm_100a?
// Each thread runs NF independent DFMA chains and NI independent integer (IMAD) / FP32 (FFMA) chains per loop trip.
// Prints cycles per loop trip per SM sub-partition at full occupancy for several (NF, NI) mixes.  If the cost is
// 2 NF + NI the integer work adds to the FP64 work (issue-bound model of profiles/r01_v12_summary.md); if it is
// max(2 NF, NF + NI) the other pipes run in the FP64 pipe's shadow.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_mix issue_mix.cu && ./issue_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NI, int KIND>
__global__ void __launch_bounds__(256) mix_kernel(int iters, double seed, int iseed, double* sink, long long* cycles) {
  double a[NF > 0 ? NF : 1];
  int b[NI > 0 ? NI : 1];
  float f[NI > 0 ? NI : 1];
#pragma unroll
  for (int i = 0; i < NF; ++i) a[i] = seed + i;
#pragma unroll
  for (int i = 0; i < NI; ++i) { b[i] = iseed + i + (int)threadIdx.x; f[i] = (float)seed + i + threadIdx.x; }  // per-thread values: keeps the chains off the uniform datapath
  const int tsel = threadIdx.x & 1 ? 0x55555555 : 0x33333333;
  const double m = 0.9999999, c = 1e-9;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < (NF > NI ? NF : NI); ++i) {
        if (i < NF) a[i] = fma(a[i], m, c);
        if (i < NI) {
          if (KIND == 0) b[i] = b[i] * tsel + 12345;            // IMAD
          else if (KIND == 1) f[i] = fmaf(f[i], 0.999f, 1e-3f);  // FFMA
          else if (KIND == 2) b[i] = (b[i] ^ tsel) & (b[i] | 0x0f0f0f0f);  // LOP3 (one three-input op)
          else if (KIND == 3) b[i] = b[i] + tsel + i;           // IADD3
          else b[i] = (b[i] > tsel) ? (b[i] - tsel) : (b[i] + 77);  // ISETP + SEL / VIADD
        }
      }
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NF; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < NI; ++i) s += b[i] + f[i];
  if (s == 123.456) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int NF, int NI, int KIND>
void run(const char* kind, double* sink, long long* cyc) {
  const int iters = 4096;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  // 8 warps per sub-partition: 4 blocks of 256 threads per SM
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  mix_kernel<NF, NI, KIND><<<sms * 4, 256>>>(16, 1.0, 3, sink, cyc);
  cudaEventRecord(e0);
  mix_kernel<NF, NI, KIND><<<sms * 4, 256>>>(iters, 1.0, 3, sink, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  // per sub-partition: 8 warps, each iters*4 trips of (NF + NI) instructions
  const double trips = (double)iters * 4 * 8;
  printf("NF=%d NI=%d %-5s  %.3f ms  cycles/trip/subpartition = %.2f   (2NF+NI = %d, max(2NF, NF+NI) = %d)\n", NF, NI, kind, ms,
         (double)h / trips, 2 * NF + NI, (2 * NF > NF + NI ? 2 * NF : NF + NI));
}

int main() {
  double* sink; long long* cyc;
  cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
  run<8, 0, 0>("-", sink, cyc);
  run<0, 8, 0>("imad", sink, cyc);
  run<8, 4, 0>("imad", sink, cyc);
  run<8, 8, 0>("imad", sink, cyc);
  run<8, 16, 0>("imad", sink, cyc);
  run<8, 4, 1>("ffma", sink, cyc);
  run<8, 8, 1>("ffma", sink, cyc);
  run<8, 16, 1>("ffma", sink, cyc);
  run<0, 8, 2>("lop3", sink, cyc);
  run<8, 8, 2>("lop3", sink, cyc);
  run<8, 16, 2>("lop3", sink, cyc);
  run<0, 8, 3>("iadd3", sink, cyc);
  run<8, 8, 3>("iadd3", sink, cyc);
  run<8, 16, 3>("iadd3", sink, cyc);
  run<0, 8, 4>("setp+sel", sink, cyc);
  run<8, 8, 4>("setp+sel", sink, cyc);
  run<4, 8, 0>("imad", sink, cyc);
  run<4, 8, 1>("ffma", sink, cyc);
  return 0;
}
