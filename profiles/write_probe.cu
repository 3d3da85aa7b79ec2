// write_probe.cu -- dev microbenchmark: store bandwidth of ONE CTA per SM writing 335 MB (the act[0] tensor of config 2)
// the way the leaf GEMM's epilogue does (8 or 16 warps, one 128-byte line per warp store, lines 512 B apart inside a
// 640 KB tile block) versus contiguous lines, st.global.cs versus default, and TMA bulk stores from shared memory.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// mode 0: strided lines (column stride 512 B), mode 1: contiguous lines; cs: streaming stores
__global__ void __launch_bounds__(1024, 1) st_kernel(float* out, size_t n_lines, int mode, int cs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const size_t per_cta = n_lines / gridDim.x;
  const size_t base = (size_t)blockIdx.x * per_cta;
  const float v = (float)lane;
  for (size_t l = warp; l < per_cta; l += nw) {
    size_t line = base + l;
    if (mode == 0) {   // permute inside blocks of 1280 x 4 lines: (column c, quarter q) -> line c * 4 + q, visited q-major
      const size_t blk = line / 5120, r = line % 5120;
      line = blk * 5120 + (r % 1280) * 4 + r / 1280;
    }
    float* p = out + line * 32 + lane;
    if (cs) __stcs(p, v); else *p = v;
  }
}

// TMA bulk store: each warp fills a `chunk`-byte staging buffer in shared memory (STS), one lane issues cp.async.bulk
__global__ void __launch_bounds__(1024, 1) tma_st_kernel(float* out, size_t bytes, int chunk) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float* stg = reinterpret_cast<float*>(sm + (size_t)warp * chunk);
  const size_t n_chunks = bytes / chunk, per_cta = n_chunks / gridDim.x;
  for (size_t c = warp; c < per_cta; c += nw) {
    for (int i = lane; i < chunk / 4; i += 32) stg[i] = (float)i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<unsigned char*>(out) + ((size_t)blockIdx.x * per_cta + c) * chunk),
                   "r"((uint32_t)__cvta_generic_to_shared(stg)), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
float time_ms(F f, int n = 20) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaEventRecord(a));
  for (int i = 0; i < n; ++i) f();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / n;
}

int main() {
  const size_t bytes = (size_t)1280 * 65536 * 4;   // 335 MB
  float* out; CK(cudaMalloc(&out, bytes));
  int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  const size_t n_lines = bytes / 128;
  // per-SM limits: only 16 CTAs write (a ninth of the buffer), so neither HBM nor L2 is the bound
  for (int threads : {256, 512})
    for (int mode : {0, 1}) {
      float ms = time_ms([&] { st_kernel<<<16, threads>>>(out, n_lines / 9, mode, 1); });
      printf("WPROBE st 16 CTAs warps %2d %s : %.4f ms = %.1f B/clk/SM\n", threads / 32, mode ? "contiguous" : "strided   ", ms, bytes / 9.0 / ms / 1e6 / 16 / 1.9);
    }
  CK(cudaFuncSetAttribute(tma_st_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int chunk : {512, 2048, 8192}) {
    float ms = time_ms([&] { tma_st_kernel<<<16, 256, (size_t)8 * chunk>>>(out, bytes / 9, chunk); });
    printf("WPROBE tma st 16 CTAs warps 8 chunk %5d : %.4f ms = %.1f B/clk/SM\n", chunk, ms, bytes / 9.0 / ms / 1e6 / 16 / 1.9);
  }
  for (int threads : {128, 256, 512, 1024})
    for (int mode : {0, 1})
      for (int cs : {1, 0}) {
        float ms = time_ms([&] { st_kernel<<<nsm, threads>>>(out, n_lines, mode, cs); });
        printf("WPROBE st     warps %2d %s %s : %.4f ms %.0f GB/s = %.1f B/clk/SM @1.9GHz\n", threads / 32, mode ? "contiguous" : "strided   ", cs ? "cs" : "  ",
               ms, bytes / ms / 1e6, bytes / ms / 1e6 / nsm / 1.9);
      }
  CK(cudaFuncSetAttribute(tma_st_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int threads : {128, 256})
    for (int chunk : {512, 2048, 8192}) {
      float ms = time_ms([&] { tma_st_kernel<<<nsm, threads, (size_t)(threads / 32) * chunk>>>(out, bytes, chunk); });
      printf("WPROBE tma st warps %2d chunk %5d : %.4f ms %.0f GB/s = %.1f B/clk/SM\n", threads / 32, chunk, ms, bytes / ms / 1e6, bytes / ms / 1e6 / nsm / 1.9);
    }
  CK(cudaDeviceSynchronize());
  return 0;
}
