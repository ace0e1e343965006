// Dev microbenchmark: achievable HBM read bandwidth of 1-D bulk async copies (cp.async.bulk)
// as a function of row size, rows per stage, ring depth and CTAs per SM.  Consumers only wait
// and release.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bw tools/tma_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../restrictive-hierarchical-semantic-segmentation_b200/csrc/pipeline.cuh"
using namespace rhseg;

__global__ void __launch_bounds__(160) k_tma(const char* src, size_t pitch, int row_bytes, int rows, int ns, long units,
                                             size_t span, int misalign, float* sink, int lanes_issue) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = (uint64_t*)smem;
  unsigned char* ring = smem + 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < ns; ++i) { mbar_init(smem_u32(&bars[i]), 1); mbar_init(smem_u32(&bars[ns + i]), 4); }
    mbar_fence_init();
  }
  __syncthreads();
  const long u0 = units * blockIdx.x / gridDim.x, u1 = units * (blockIdx.x + 1) / gridDim.x;
  const int stage_bytes = rows * (row_bytes + 16);
  if (warp == 4) {
    const uint64_t pol = l2_evict_first_policy();
    long i = 0;
    for (long u = u0; u < u1; ++u, ++i) {
      const int slot = i % ns; const uint32_t it = i / ns;
      mbar_wait(smem_u32(&bars[ns + slot]), (it & 1) ^ 1);
      // unit u covers `rows` rows starting at column block (u % colblocks) of row group (u / colblocks)
      const size_t colblocks = span / row_bytes;
      const size_t rg = u / colblocks, cb = u % colblocks;
      const int bytes = row_bytes + (misalign ? 16 : 0);
      if (lane == 0) mbar_arrive_expect_tx(smem_u32(&bars[slot]), bytes * rows);
      __syncwarp();
      if (lanes_issue) {
        if (lane < rows)
          bulk_g2s(smem_u32(ring + (size_t)slot * stage_bytes + lane * (row_bytes + 16)), src + (rg * rows + lane) * pitch + cb * row_bytes, bytes, smem_u32(&bars[slot]), pol);
      } else if (lane == 0) {
        for (int r = 0; r < rows; ++r)
          bulk_g2s(smem_u32(ring + (size_t)slot * stage_bytes + r * (row_bytes + 16)), src + (rg * rows + r) * pitch + cb * row_bytes, bytes, smem_u32(&bars[slot]), pol);
      }
    }
    return;
  }
  long i = 0; float acc = 0.f;
  for (long u = u0; u < u1; ++u, ++i) {
    const int slot = i % ns; const uint32_t it = i / ns;
    mbar_wait(smem_u32(&bars[slot]), it & 1);
    acc += ((float*)(ring + (size_t)slot * stage_bytes))[tid];
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bars[ns + slot]));
  }
  if (acc == 123.456f) sink[0] = acc;
}

int main() {
  const size_t total = 1ull << 30;  // 1 GiB
  char* buf; cudaMalloc(&buf, total + (1 << 20)); cudaMemset(buf, 1, total + (1 << 20));
  float* sink; cudaMalloc(&sink, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("row_bytes rows ns cta/sm misalign lanes  ->  GB/s\n");
  int cfgs[][6] = {
      {2048, 16, 3, 2, 0, 1}, {2048, 16, 3, 2, 1, 1}, {2048, 8, 4, 3, 0, 1}, {2048, 8, 4, 3, 1, 1}, {1024, 16, 4, 3, 0, 1},
      {1024, 16, 4, 3, 1, 1}, {512, 16, 4, 6, 0, 1}, {4096, 8, 3, 2, 0, 1}, {8192, 4, 3, 2, 0, 1}, {16384, 2, 3, 2, 0, 1},
      {2048, 16, 3, 2, 0, 0}, {2048, 8, 4, 4, 0, 1}, {2048, 16, 6, 1, 0, 1}, {4096, 16, 3, 1, 0, 1}, {2048, 8, 3, 4, 1, 1},
  };
  for (auto& c : cfgs) {
    const int row_bytes = c[0], rows = c[1], ns = c[2], per_sm = c[3], mis = c[4], lanes = c[5];
    const size_t span = 96000 / row_bytes * row_bytes;  // bytes used per row (HRNet-like plane of ~96 KB)
    const size_t pitch = mis ? 96100 : 96256;           // 96100: 4-byte aligned planes (16B-aligned start emulated below)
    const size_t nrows = total / pitch / rows * rows;
    const long units = (long)(nrows / rows) * (span / row_bytes);
    const size_t smem = 128 + (size_t)ns * rows * (row_bytes + 16);
    const int grid = 148 * per_sm;
    const char* src = buf;  // misalign: rows start at multiples of 96100 = 4 mod 16 -> we round each down to 16B
    // for the misaligned case emulate "aligned-down span": pitch rounded to 16B multiple keeps alignment legal
    const size_t use_pitch = mis ? 96096 + 16 : pitch;  // 96112: 16B-aligned but not 32/128B-aligned rows
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k_tma<<<grid, 160, smem>>>(src, use_pitch, row_bytes, rows, ns, units, span, mis, sink, lanes);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    const double bytes = (double)units * rows * (row_bytes + (mis ? 16 : 0));
    printf("%6d %4d %2d %2d %d %d  ->  %8.1f GB/s  (%.3f ms, smem %zu KB) %s\n", row_bytes, rows, ns, per_sm, mis, lanes,
           bytes / ms / 1e6, ms, smem / 1024, err == cudaSuccess ? "" : cudaGetErrorString(err));
  }
  return 0;
}
