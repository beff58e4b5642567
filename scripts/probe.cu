// latency / throughput probes for the FP64 pipe, MUFU.RCP64H and shared memory on the target GPU
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat_dfma(double* out, long long* cyc, int n, double a, double b)
{
  double x = threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; i++) x = fma(x, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lat_dadd(double* out, long long* cyc, int n, double a)
{
  double x = threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; i++) x = __dadd_rn(x, a);
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lat_rcp(double* out, long long* cyc, int n)
{
  double x = 1.5 + threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; i++) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y; }
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lat_lds(double* out, long long* cyc, int n)
{
  __shared__ int s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 37 + 11) & 1023;
  __syncthreads();
  int j = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; i++) j = s[j];
  long long t1 = clock64();
  out[threadIdx.x] = j; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// throughput with k independent chains per thread, many warps
template <int K>
__global__ void k_tp_dfma(double* out, int n, double a, double b)
{
  double x[K];
#pragma unroll
  for (int k = 0; k < K; k++) x[k] = threadIdx.x + k;
  for (int i = 0; i < n; i++)
  {
#pragma unroll
    for (int k = 0; k < K; k++) x[k] = fma(x[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < K; k++) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// rcp seed accuracy
__global__ void k_rcp_acc(double* maxerr0, double* maxerr1, double* maxerr2)
{
  double e0 = 0, e1 = 0, e2 = 0;
  for (int i = 0; i < 4096; i++)
  {
    double x = 0.3 + (threadIdx.x * 4096 + i) * (8.0 / (1024 * 4096));
    double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double ex = 1.0 / x;
    e0 = fmax(e0, fabs(y - ex) / ex);
    double e = fma(-x, y, 1.0); y = fma(y, e, y);
    e1 = fmax(e1, fabs(y - ex) / ex);
    e = fma(-x, y, 1.0); y = fma(y, e, y);
    e2 = fmax(e2, fabs(y - ex) / ex);
  }
  maxerr0[threadIdx.x] = e0; maxerr1[threadIdx.x] = e1; maxerr2[threadIdx.x] = e2;
}
int main()
{
  double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 64);
  long long h; const int n = 4096;
  k_lat_dfma<<<1, 32>>>(out, cyc, n, 1.0000001, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("DFMA dependent latency  %.2f cyc\n", (double)h / n);
  k_lat_dadd<<<1, 32>>>(out, cyc, n, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("DADD dependent latency  %.2f cyc\n", (double)h / n);
  k_lat_rcp<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("RCP64H dependent latency %.2f cyc\n", (double)h / n);
  k_lat_lds<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("LDS dependent latency   %.2f cyc\n", (double)h / n);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 1; warps <= 16; warps *= 2)
  {
    float ms;
    const int blocks = 148 * 4;   // one block per SM sub-partition share: blocks of `warps` warps, 4 blocks per SM
    k_tp_dfma<1><<<blocks, 32 * warps>>>(out, 20000, 1.0000001, 1e-9); cudaEventRecord(e0); k_tp_dfma<1><<<blocks, 32 * warps>>>(out, 20000, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("warps/SM %3d  K=1: %.2f TFLOP/s", warps * 4, 2.0 * 20000 * blocks * 32 * warps / ms / 1e9);
    k_tp_dfma<4><<<blocks, 32 * warps>>>(out, 20000, 1.0000001, 1e-9); cudaEventRecord(e0); k_tp_dfma<4><<<blocks, 32 * warps>>>(out, 20000, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("   K=4: %.2f TFLOP/s\n", 2.0 * 4 * 20000 * blocks * 32 * warps / ms / 1e9);
  }
  double *m0, *m1, *m2; cudaMalloc(&m0, 8192); cudaMalloc(&m1, 8192); cudaMalloc(&m2, 8192);
  k_rcp_acc<<<1, 1024>>>(m0, m1, m2);
  double h0[1024], h1[1024], h2[1024]; cudaMemcpy(h0, m0, 8192, cudaMemcpyDeviceToHost); cudaMemcpy(h1, m1, 8192, cudaMemcpyDeviceToHost); cudaMemcpy(h2, m2, 8192, cudaMemcpyDeviceToHost);
  double a = 0, b = 0, c = 0; for (int i = 0; i < 1024; i++) { a = fmax(a, h0[i]); b = fmax(b, h1[i]); c = fmax(c, h2[i]); }
  printf("rcp.approx.ftz.f64 max rel err: seed %.3e, 1 Newton %.3e, 2 Newton %.3e\n", a, b, c);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
