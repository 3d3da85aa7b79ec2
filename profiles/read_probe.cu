// read_probe.cu -- dev microbenchmark: how fast can ONE persistent CTA per SM stream a 205 MB fp32 buffer out of HBM,
// per load path?  (bulk TMA into a stage ring | cp.async 16 B | ld.global.nc.v4 into registers.)  Decides the load path
// of csrc/ratspn_leaf_stream.cu.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/read_probe profiles/read_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}

// mode 3: 2-D tensor TMA, box [rows x 32 floats] of the (65536, 784) matrix, 128B swizzle (what a K-block stage of the leaf
//         GEMM needs);  `extra` > 0 adds that many 4 KB bulk copies of an L2-resident buffer per stage (the weight images)
// mode 0: bulk TMA, one thread issues chunk-byte copies into an S-deep ring, warp 1 consumes (touches one word, releases)
// mode 1: cp.async 16 B by `nload` threads into the ring, completion by cp.async.mbarrier.arrive
__global__ void __launch_bounds__(1024, 1) ring_kernel(const unsigned char* x, size_t bytes, int chunk, int S, int mode, int nload, float* out,
                                                       const __grid_constant__ CUtensorMap map, int extra, const unsigned char* w) {
  extern __shared__ __align__(1024) unsigned char sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)S * chunk);
  uint64_t* empty = full + 16;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full + s, (mode == 0 || mode == 3) ? 1 : nload); mbar_init(empty + s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t nchunks = mode == 3 ? (size_t)(65536 / (chunk / 128)) * 25 : bytes / chunk;
  const size_t mine = (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int t = threadIdx.x;
  const int rows = chunk / 128, KBn = 25;
  if ((mode == 0 || mode == 3) ? (t == 0) : (t < nload)) {
    for (size_t g = 0; g < mine; ++g) {
      const int s = (int)(g % S);
      const uint32_t ph = (uint32_t)(g / S) & 1u;
      const unsigned char* src = x + (blockIdx.x + g * gridDim.x) * (size_t)chunk;
      mbar_wait(empty + s, ph ^ 1u);
      if (mode == 3) {
        const size_t gg = blockIdx.x + g * gridDim.x;   // stage index: tile = gg / KBn, kb = gg % KBn  (tile-major like the GEMM)
        mbar_expect_tx(full + s, chunk + extra * 4096);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(sm + (size_t)s * chunk)),
                     "l"(&map), "r"((int)(gg % KBn) * 32), "r"((int)(gg / KBn) * rows), "r"(smem_u32(full + s)) : "memory");
        for (int e = 0; e < extra; ++e)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + 200 * 1024 + e * 4096)), "l"(w + ((gg % KBn) * 2 + e) * 4096), "r"(4096), "r"(smem_u32(full + s)) : "memory");
      } else if (mode == 0) {
        mbar_expect_tx(full + s, chunk);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + (size_t)s * chunk)), "l"(src), "r"(chunk), "r"(smem_u32(full + s)) : "memory");
      } else {
        for (int o = t * 16; o < chunk; o += nload * 16)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sm + (size_t)s * chunk + o)), "l"(src + o) : "memory");
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(full + s)) : "memory");
      }
    }
    if (mode == 1) asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (t >= 992) {   // last warp: consumer
    float acc = 0.f;
    for (size_t g = 0; g < mine; ++g) {
      const int s = (int)(g % S);
      const uint32_t ph = (uint32_t)(g / S) & 1u;
      mbar_wait(full + s, ph);
      acc += *reinterpret_cast<const float*>(sm + (size_t)s * chunk + (t - 992) * 4);
      __syncwarp();
      if (t == 992) mbar_arrive(empty + s);
    }
    if (acc == 123.456f) out[0] = acc;
  }
}

// mode 2: ld.global.nc.v4 into registers, U loads in flight per thread, persistent grid-stride over 16-byte words
template <int U>
__global__ void ldg_kernel(const float4* __restrict__ x, size_t n16, float* out) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n16; i += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(x + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  for (; i < n16; i += stride) { const float4 v = __ldg(x + i); acc += v.x + v.y + v.z + v.w; }
  if (acc == 123.456f) out[0] = acc;
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

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const size_t bytes = (size_t)65536 * 784 * 4;
  unsigned char* x; float* out;
  CK(cudaMalloc(&x, bytes)); CK(cudaMalloc(&out, 4));
  CK(cudaMemset(x, 1, bytes));
  int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  unsigned char* w; CK(cudaMalloc(&w, 25 * 2 * 4096)); CK(cudaMemset(w, 0, 25 * 2 * 4096));
  EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr));
  CUtensorMap dummy;
  auto make_map = [&](int rows, CUtensorMapL2promotion promo) {
    CUtensorMap m;
    cuuint64_t dims[2] = {784, 65536}, strides[1] = {784 * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)rows}, es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
  };
  dummy = make_map(128, CU_TENSOR_MAP_L2_PROMOTION_NONE);
  CK(cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  for (int chunk : {8192, 16384, 32768})
    for (int S : {2, 4, 6, 12}) {
      if ((size_t)S * chunk > 200 * 1024) continue;
      const size_t smem = (size_t)S * chunk + 512;
      float ms = time_ms([&] { ring_kernel<<<nsm, 1024, smem>>>(x, bytes, chunk, S, 0, 0, out, dummy, 0, w); });
      printf("PROBE bulk   chunk %6d S %2d : %.4f ms %.0f GB/s\n", chunk, S, ms, bytes / ms / 1e6);
      for (int nload : {128, 512}) {
        ms = time_ms([&] { ring_kernel<<<nsm, 1024, smem>>>(x, bytes, chunk, S, 1, nload, out, dummy, 0, w); });
        printf("PROBE ldgsts chunk %6d S %2d nload %3d : %.4f ms %.0f GB/s\n", chunk, S, nload, ms, bytes / ms / 1e6);
      }
    }
  for (int rows : {64, 128, 256})
    for (int S : {2, 4, 6})
      for (int extra : {0, 2})
        for (int promo = 0; promo < 2; ++promo) {
          const int chunk = rows * 128;
          if ((size_t)S * chunk > 200 * 1024) continue;
          CUtensorMap m = make_map(rows, promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE);
          float ms = time_ms([&] { ring_kernel<<<nsm, 1024, 200 * 1024 + 8192 + 512>>>(x, bytes, chunk, S, 3, 0, out, m, extra, w); });
          printf("PROBE tensor rows %3d S %d extra %d promo %d : %.4f ms %.0f GB/s\n", rows, S, extra, promo, ms, bytes / ms / 1e6);
        }
  CK(cudaGetLastError());
  for (int cps : {1, 2, 4, 8})
    for (int threads : {256, 512, 1024}) {
      if (cps * threads > 2048) continue;
      float m1 = time_ms([&] { ldg_kernel<1><<<nsm * cps, threads>>>((const float4*)x, bytes / 16, out); });
      float m4 = time_ms([&] { ldg_kernel<4><<<nsm * cps, threads>>>((const float4*)x, bytes / 16, out); });
      float m8 = time_ms([&] { ldg_kernel<8><<<nsm * cps, threads>>>((const float4*)x, bytes / 16, out); });
      printf("PROBE ldg    ctas/sm %d threads %4d : U1 %.0f  U4 %.0f  U8 %.0f GB/s\n", cps, threads, bytes / m1 / 1e6, bytes / m4 / 1e6, bytes / m8 / 1e6);
    }
  CK(cudaDeviceSynchronize());
  return 0;
}
