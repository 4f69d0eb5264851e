// Micro-benchmarks that size the EdgeConv kernel's staging pipeline on B200 (diagnostic, not product code):
//   1. L2-resident read bandwidth with 128-bit LDG            (what direct neighbour gathers compete for)
//   2. L2 -> shared memory through cp.async.bulk (TMA engine) in pieces of 128 B / 256 B / 512 B / 16 KB
//      (row-granular staging of neighbour rows vs. tile-granular weight streaming)
//   3. shared-memory LDS.128 bandwidth                          (the max-aggregation's read path)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membw membw.cu ; run: ./membw
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(512) l2_read_kernel(const uint4* __restrict__ buf, size_t n_vec, int iters, uint4* sink) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; ++it) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
      uint4 v;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + i));
      acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
    }
  }
  if (acc.x == 0x12345678u) sink[0] = acc;
}

// one producer warp per CTA issues bulk copies of `piece` bytes into a ring of NBUF buffers of `bufbytes`
__global__ void __launch_bounds__(128) bulk_kernel(const uint8_t* __restrict__ buf, size_t nbytes, int piece, int bufbytes, int rounds, int* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int pieces = bufbytes / piece;
  // each CTA walks its own window of the buffer (L2-resident overall)
  const size_t window = (nbytes / gridDim.x) & ~(size_t)1023;
  const uint8_t* base = buf + (size_t)blockIdx.x * window;
  if (warp == 0) {
    for (int r = 0; r < rounds; ++r) {
      const int s = r & 1;
      if (r >= 2) {  // wait until round r-2 landed (single consumer = this warp, so this is also "buffer free")
        uint32_t ok = 0;
        const uint32_t par = ((r - 2) >> 1) & 1;
        while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(&full[s])), "r"(par) : "memory");
      }
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(pieces * piece) : "memory");
      __syncwarp();
      const size_t off0 = ((size_t)r * bufbytes) % (window - bufbytes);
      for (int p = lane; p < pieces; p += 32) {
        // scatter the source rows a little (stride 2*piece) like a row gather would
        const size_t src = (off0 + (size_t)p * piece) & ~(size_t)15;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + s * bufbytes + p * piece)), "l"(base + src), "r"(piece), "r"(smem_u32(&full[s])) : "memory");
      }
    }
    for (int r = rounds - 2; r < rounds; ++r) {
      if (r < 0) continue;
      const int s = r & 1;
      uint32_t ok = 0;
      const uint32_t par = (r >> 1) & 1;
      while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(&full[s])), "r"(par) : "memory");
    }
  }
  __syncthreads();
  if (smem[threadIdx.x] == 0xAB && threadIdx.x == 1000) sink[0] = 1;
}

__global__ void __launch_bounds__(512) lds_kernel(int iters, uint4* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  for (int i = threadIdx.x; i < 65536 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(i, i, i, i);
  __syncthreads();
  uint4 acc = make_uint4(0, 0, 0, 0);
  const int lane = threadIdx.x & 31;
  // quarter-warp reads one contiguous 128-byte row; 4 different rows per instruction (like the aggregation)
  uint32_t row = (threadIdx.x * 7) & 511;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const uint4 v = *reinterpret_cast<const uint4*>(smem + ((row + k * 13) & 511) * 128 + (lane & 7) * 16);
      acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
    }
    row = (row + acc.x) & 511;
  }
  if (acc.x == 0x12345678u) sink[0] = acc;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, L2 %d MB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20);
  const int sms = prop.multiProcessorCount;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  uint4* sink;
  CK(cudaMalloc(&sink, 64));
  float ms;

  for (size_t mb : {16, 32, 64, 96, 256, 2048}) {
    const size_t bytes = mb << 20;
    uint4* buf;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 1, bytes));
    const int iters = mb <= 256 ? 20 : 3;
    l2_read_kernel<<<sms * 4, 512>>>(buf, bytes / 16, 2, sink);  // warm
    CK(cudaEventRecord(e0));
    l2_read_kernel<<<sms * 4, 512>>>(buf, bytes / 16, iters, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("LDG.128 read, %4zu MB working set: %8.1f GB/s\n", mb, (double)bytes * iters / ms / 1e6);
    CK(cudaFree(buf));
  }

  {
    const size_t bytes = (size_t)48 << 20;
    uint8_t* buf;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 1, bytes));
    CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536));
    for (int piece : {128, 256, 512, 1024, 16384}) {
      const int bufbytes = 32768;
      const int rounds = 400;
      for (int ctas_per_sm : {1, 2}) {
        bulk_kernel<<<sms * ctas_per_sm, 128, 2 * bufbytes>>>(buf, bytes, piece, bufbytes, 20, (int*)sink);
        CK(cudaEventRecord(e0));
        bulk_kernel<<<sms * ctas_per_sm, 128, 2 * bufbytes>>>(buf, bytes, piece, bufbytes, rounds, (int*)sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double tot = (double)sms * ctas_per_sm * rounds * bufbytes;
        printf("cp.async.bulk L2->smem, piece %5d B, %d CTA/SM (1 producer warp, 2x32KB ring): %8.1f GB/s  (%.1f B/clk/SM @1.9GHz)\n", piece,
               ctas_per_sm, tot / ms / 1e6, tot / ms / 1e6 / sms / 1.9);
      }
    }
    CK(cudaFree(buf));
  }

  {
    CK(cudaFuncSetAttribute(lds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    const int iters = 2000;
    lds_kernel<<<sms, 512, 65536>>>(10, sink);
    CK(cudaEventRecord(e0));
    lds_kernel<<<sms, 512, 65536>>>(iters, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double tot = (double)sms * 512 * iters * 32 * 16;
    printf("LDS.128 (quarter-warp rows): %8.1f GB/s  (%.1f B/clk/SM @1.9GHz)\n", tot / ms / 1e6, tot / ms / 1e6 / sms / 1.9);
  }
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
