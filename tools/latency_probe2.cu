// latency_probe2.cu -- where the microseconds of a one-launch small-system call go: a kernel that reads
// `nin` doubles from mapped host memory, writes `nout` doubles back, fences and raises a flag, with
// %globaltimer stamps after every stage; variants: inputs copied to device memory first, outputs
// without the fence, results to device memory + one cudaMemcpyAsync.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/latency_probe2 tools/latency_probe2.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <algorithm>
#include <chrono>
#include <vector>

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void k_io(const double* in, double* out, volatile unsigned* flag, int nin, int nout, int fences,
                     unsigned long long* stamps) {
  __shared__ double s[2048];
  const unsigned long long t0 = gtime();
  for (int t = threadIdx.x; t < nin; t += blockDim.x) s[t] = in[t];
  __syncthreads();
  const unsigned long long t1 = gtime();
  double acc = 0.0;
  for (int t = 0; t < nin; ++t) acc += s[t] * (double)(threadIdx.x + 1);
  for (int t = threadIdx.x; t < nout; t += blockDim.x) out[t] = acc + s[t % nin];
  const unsigned long long t2 = gtime();
  if (fences >= 2) __threadfence_system();
  __syncthreads();
  const unsigned long long t3 = gtime();
  if (threadIdx.x == 0) {
    out[nout] = acc;
    if (fences >= 1) __threadfence_system();
    const unsigned long long t4 = gtime();
    *flag = 0u;
    stamps[0] = t0;
    stamps[1] = t1;
    stamps[2] = t2;
    stamps[3] = t3;
    stamps[4] = t4;
  }
}

int main() {
  double *h_in, *h_out;
  unsigned* h_flag;
  cudaHostAlloc(&h_in, 1 << 16, cudaHostAllocMapped);
  cudaHostAlloc(&h_out, 1 << 16, cudaHostAllocMapped);
  cudaHostAlloc(&h_flag, 64, cudaHostAllocMapped);
  for (int i = 0; i < 4096; ++i) h_in[i] = 1.0 + i;
  double *m_in, *m_out, *d_in, *d_out;
  unsigned* m_flag;
  unsigned long long *d_st, h_st[8];
  cudaHostGetDevicePointer(&m_in, h_in, 0);
  cudaHostGetDevicePointer(&m_out, h_out, 0);
  cudaHostGetDevicePointer(&m_flag, h_flag, 0);
  cudaMalloc(&d_in, 1 << 16);
  cudaMalloc(&d_out, 1 << 16);
  cudaMalloc(&d_st, 64);
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  const int reps = 3000;
  const int sizes[][2] = {{39 + 9, 39}, {12 + 9 + 2, 12}, {654 + 9 + 109, 654}, {1500 + 9, 1500}};
  for (auto& sz : sizes) {
    const int nin = sz[0], nout = sz[1];
    // mode 0: mapped in, mapped out, 2 fences | 1: mapped in/out, 1 fence (thread 0 only, after the barrier)
    // mode 2: H2D copy of the inputs, mapped out | 3: H2D copy in, device out + D2H copy, stream sync
    for (int mode = 0; mode < 4; ++mode) {
      for (int threads : {32, 128}) {
        std::vector<double> us(reps);
        double stage[4] = {0, 0, 0, 0};
        for (int r = -200; r < reps; ++r) {
          auto t0 = std::chrono::steady_clock::now();
          *(volatile unsigned*)h_flag = 0xffffffffu;
          const double* in = m_in;
          double* out = m_out;
          if (mode >= 2) {
            cudaMemcpyAsync(d_in, h_in, 8 * nin, cudaMemcpyHostToDevice, st);
            in = d_in;
          }
          if (mode == 3) out = d_out;
          k_io<<<1, threads, 0, st>>>(in, out, m_flag, nin, nout, mode == 0 ? 2 : 1, d_st);
          if (mode == 3) {
            cudaMemcpyAsync(h_out, d_out, 8 * (nout + 1), cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
          } else {
            while (*(volatile unsigned*)h_flag == 0xffffffffu) __builtin_ia32_pause();
          }
          if (r >= 0) us[r] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e6;
          if (r >= 0 && r < 64) {
            cudaStreamSynchronize(st);
            cudaMemcpy(h_st, d_st, 40, cudaMemcpyDeviceToHost);
            for (int k = 0; k < 4; ++k) stage[k] += (double)(h_st[k + 1] - h_st[k]) / 64.0;
          }
        }
        std::sort(us.begin(), us.end());
        printf("{\"nin\": %d, \"nout\": %d, \"mode\": %d, \"threads\": %d, \"us_median\": %.2f, \"us_min\": %.2f, "
               "\"kernel_ns\": {\"loads\": %.0f, \"compute_stores\": %.0f, \"fence_all\": %.0f, \"fence_t0\": %.0f}}\n",
               nin, nout, mode, threads, us[reps / 2], us[0], stage[0], stage[1], stage[2], stage[3]);
      }
    }
  }
  return 0;
}
