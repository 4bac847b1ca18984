// latency_probe.cu -- what one tiny kernel round trip costs on this box (the floor of any single small
// system call): launch + completion seen by stream synchronisation or by polling a mapped flag, with
// 0 / 1 / 2 dependent reads of mapped host memory inside the kernel.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o latency_probe tools/latency_probe.cu && ./latency_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <chrono>

__global__ void k_probe(const volatile long long* in, volatile unsigned* flag, int reads) {
  long long v = 1;
  if (reads >= 1) v = in[0];
  if (reads >= 2) v = in[v];
  __threadfence_system();
  *flag = (unsigned)v;
}

int main() {
  long long* h_in;
  unsigned* h_flag;
  cudaHostAlloc(&h_in, 4096, cudaHostAllocMapped);
  cudaHostAlloc(&h_flag, 64, cudaHostAllocMapped);
  h_in[0] = 1;
  h_in[1] = 7;
  long long* d_in;
  unsigned* d_flag;
  cudaHostGetDevicePointer(&d_in, h_in, 0);
  cudaHostGetDevicePointer(&d_flag, h_flag, 0);
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  const int reps = 5000;
  for (int reads = 0; reads <= 2; ++reads) {
    for (int mode = 0; mode < 2; ++mode) {
      for (int w = 0; w < 200; ++w) {
        k_probe<<<1, 32, 0, st>>>(d_in, d_flag, reads);
        cudaStreamSynchronize(st);
      }
      auto t0 = std::chrono::steady_clock::now();
      for (int r = 0; r < reps; ++r) {
        *(volatile unsigned*)h_flag = 0xffffffffu;
        k_probe<<<1, 32, 0, st>>>(d_in, d_flag, reads);
        if (mode == 0) {
          cudaStreamSynchronize(st);
        } else {
          while (*(volatile unsigned*)h_flag == 0xffffffffu) __builtin_ia32_pause();
        }
      }
      double us = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps * 1e6;
      printf("{\"what\": \"latency_probe\", \"mapped_reads\": %d, \"wait\": \"%s\", \"us\": %.2f}\n", reads,
             mode == 0 ? "cudaStreamSynchronize" : "poll mapped flag", us);
    }
  }
  return 0;
}
