// Developer check (GPU): accuracy of rmv_f64.cuh on the device against the CUDA math library, and of the reciprocal seed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -I reinmav-gym_b200/csrc tools/f64_check.cu -o build/f64_check
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "rmv_f64.cuh"

__device__ double u01(uint64_t& s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return (double)(s >> 11) * (1.0 / 9007199254740992.0); }

__global__ void check(double* out, int iters) {
  uint64_t s = 0x9e3779b97f4a7c15ull * (blockIdx.x * blockDim.x + threadIdx.x + 1);
  double e_seed = 0, e_div = 0, e_atan2 = 0, e_asin = 0, e_atan2_rel = 0;
  for (int i = 0; i < iters; ++i) {
    const double d = exp2(600.0 * (u01(s) - 0.5)) * (1.0 + u01(s)), n = (u01(s) - 0.5) * exp2(40.0 * (u01(s) - 0.5));
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    e_seed = fmax(e_seed, fabs(fma(-d, r, 1.0)));
    const double q = rmv::div_newton(n, d), qe = n / d;
    e_div = fmax(e_div, fabs(q - qe) / fabs(qe));
    const double ang = 6.283185307179586 * (u01(s) - 0.5), rad = exp2(60.0 * (u01(s) - 0.5));
    double sn, cs;
    sincos(ang, &sn, &cs);
    const double y = rad * sn, x = rad * cs;
    const double a = rmv::atan2_f64(y, x), ae = atan2(y, x);
    e_atan2 = fmax(e_atan2, fabs(a - ae));
    e_atan2_rel = fmax(e_atan2_rel, fabs(a - ae) / fabs(ae));
    const double v = 2.0 * u01(s) - 1.0;
    e_asin = fmax(e_asin, fabs(rmv::asin_f64(v) - asin(v)) / fabs(asin(v)));
  }
  double* o = out + 5 * (blockIdx.x * blockDim.x + threadIdx.x);
  o[0] = e_seed; o[1] = e_div; o[2] = e_atan2; o[3] = e_asin; o[4] = e_atan2_rel;
}

int main() {
  const int blocks = 296, threads = 128, iters = 20000, n = blocks * threads;
  double* d; cudaMalloc(&d, sizeof(double) * 5 * n);
  check<<<blocks, threads>>>(d, iters);
  double* h = new double[5 * n];
  if (cudaMemcpy(h, d, sizeof(double) * 5 * n, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error\n"); return 1; }
  double m[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i) for (int k = 0; k < 5; ++k) m[k] = h[5 * i + k] > m[k] ? h[5 * i + k] : m[k];
  printf("samples %lld\n", (long long)n * iters);
  printf("reciprocal seed  max |1 - d r|          %.3e  (2^%.1f)\n", m[0], log2(m[0]));
  printf("div_newton       max rel err vs n / d    %.3e\n", m[1]);
  printf("atan2_f64        max abs err vs atan2    %.3e   max rel %.3e\n", m[2], m[4]);
  printf("asin_f64         max rel err vs asin     %.3e\n", m[3]);
  return 0;
}
