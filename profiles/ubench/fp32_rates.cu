// Micro-benchmark: issue / pipe rate of scalar and packed (two-wide) fp32 instructions on sm_100a.
// For every instruction kind, every thread runs ILP independent dependency chains; the kernel reports
// warp-instructions per cycle per SM (clock64 on one SM, all SMs loaded).  Build:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp32_rates fp32_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

template <int KIND, int ILP>
__global__ void k(float2* out, long long* cyc, float2 s, float2 c) {
  float2 v[ILP], u[ILP], w[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    v[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    u[i] = make_float2(1.0f + 1e-6f * (i + threadIdx.x), 1.0f - 1e-6f * i * s.x);
    w[i] = make_float2(1e-3f * i * c.x, 1e-3f * (i + 1) * c.y);
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (KIND == 0) { v[i].x = fmaf(v[i].x, s.x, c.x); }                                   // FFMA
      if (KIND == 1) { v[i].x = v[i].x + c.x; }                                             // FADD
      if (KIND == 2) { v[i] = __ffma2_rn(v[i], s, c); }                                     // FFMA2
      if (KIND == 3) { v[i] = __fadd2_rn(v[i], c); }                                        // FADD2
      if (KIND == 4) { v[i] = __fmul2_rn(v[i], s); }                                        // FMUL2
      if (KIND == 5) { v[i].x = fmaf(v[i].x, s.x, c.x); v[i].y = fmaf(v[i].y, s.y, c.y); }  // 2 x FFMA
      if (KIND == 6) { v[i].x = v[i].x * s.x; }                                             // FMUL
      if (KIND == 7) { v[i].x = fmaf(v[i].x, 1.0001f, 0.5f); }                              // FFMA imm
      if (KIND == 8) { v[i] = __ffma2_rn(v[i], u[i], w[i]); }                               // FFMA2, 3 distinct register pairs
      if (KIND == 9) { if (i & 1) v[i] = __ffma2_rn(v[i], s, c); else v[i].x = fmaf(v[i].x, s.x, c.x); }   // FFMA2 + FFMA alternating
      if (KIND == 10) { if (i & 1) v[i] = __fadd2_rn(v[i], c); else v[i].x = fmaf(v[i].x, s.x, c.x); }     // FADD2 + FFMA alternating
      if (KIND == 13) { if ((i & 3) == 0) v[i] = __ffma2_rn(v[i], s, c); else v[i].x = fmaf(v[i].x, s.x, c.x); }   // 1 FFMA2 : 3 FFMA
      if (KIND == 14) { if ((i & 3) < 2) v[i] = __ffma2_rn(v[i], s, c); else v[i].x = fmaf(v[i].x, s.x, c.x); }    // 2 FFMA2 : 2 FFMA (grouped)
      if (KIND == 15) { if ((i & 7) == 0) v[i] = __ffma2_rn(v[i], s, c); else v[i].x = fmaf(v[i].x, s.x, c.x); }   // 1 FFMA2 : 7 FFMA
      if (KIND == 16) { if (i & 1) v[i] = __ffma2_rn(v[i], s, c); else v[i].x = v[i].x * s.x; }   // FFMA2 + FMUL alternating
      if (KIND == 11) { v[i].x = fmaf(v[i].x, u[i].x, w[i].x); }                            // FFMA, 3 distinct registers
      if (KIND == 12) { v[i] = __fadd2_rn(v[i], w[i]); }                                    // FADD2, 2 distinct pairs
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < ILP; ++i) { acc.x += v[i].x; acc.y += v[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND, int ILP>
void run(const char* name, int threads, int ctas_per_sm, int per_iter) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * ctas_per_sm;
  float2* out; long long* cyc;
  cudaMalloc(&out, sizeof(float2) * grid * threads);
  cudaMalloc(&cyc, sizeof(long long) * grid);
  k<KIND, ILP><<<grid, threads>>>(out, cyc, make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f));
  k<KIND, ILP><<<grid, threads>>>(out, cyc, make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f));
  cudaDeviceSynchronize();
  long long h[2048];
  cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
  const double winst = (double)ITERS * ILP * per_iter * (threads / 32) * ctas_per_sm;
  printf("%-10s ILP=%d threads=%4d ctas/SM=%d  warp-inst/cycle/SM = %.3f   (%s)\n", name, ILP, threads, ctas_per_sm,
         winst / avg, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {256, 512, 1024}) {
    run<0, 8>("FFMA", threads, 1, 1);
    run<7, 8>("FFMA.imm", threads, 1, 1);
    run<1, 8>("FADD", threads, 1, 1);
    run<6, 8>("FMUL", threads, 1, 1);
    run<2, 8>("FFMA2", threads, 1, 1);
    run<3, 8>("FADD2", threads, 1, 1);
    run<4, 8>("FMUL2", threads, 1, 1);
    run<5, 8>("2xFFMA", threads, 1, 2);
  }
  run<8, 8>("FFMA2.3r", 512, 1, 1);
  run<11, 8>("FFMA.3r", 512, 1, 1);
  run<12, 8>("FADD2.2r", 512, 1, 1);
  run<9, 8>("FFMA2+FFMA", 512, 1, 1);
  run<10, 8>("FADD2+FFMA", 512, 1, 1);
  run<13, 8>("1xFFMA2:3xFFMA", 512, 1, 1);
  run<14, 8>("2xFFMA2:2xFFMA", 512, 1, 1);
  run<15, 8>("1xFFMA2:7xFFMA", 512, 1, 1);
  run<16, 8>("FFMA2+FMUL", 512, 1, 1);
  run<9, 8>("FFMA2+FFMA", 256, 1, 1);
  run<9, 8>("FFMA2+FFMA", 1024, 1, 1);
  run<0, 4>("FFMA", 512, 1, 1);
  run<2, 4>("FFMA2", 512, 1, 1);
  run<0, 16>("FFMA", 512, 1, 1);
  run<2, 16>("FFMA2", 512, 1, 1);
  return 0;
}
