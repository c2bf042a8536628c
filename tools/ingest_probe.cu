// L2 -> SM operand ingest probe (sm_100a): how many bytes per clock can ONE SM pull from L2 by TMA — and does it depend on how many
// SMs pull, or on the pitch of the box rows?  The GEMM kernels' operand traffic (64 KB per
// k-block per SM against 1536 tensor-pipe cycles) needs 42 B/clk/SM at full rate; the fused decoder's 8-SM cluster measured ~24.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ingest_probe ingest_probe.cu -lcuda
//   ./ingest_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int BOX_ROWS = 128, BOX_BYTES = BOX_ROWS * 128;  // the GEMMs' operand box: 128 rows x 128 bytes, SWIZZLE_128B
constexpr int STAGES = 4;                                  // 2 boxes per stage: 32 KB, 128 KB in flight
constexpr int LSU_DEPTH = 3;                               // cp.async rounds of 32 KB in flight
constexpr int THREADS = 288;                               // warp 0: TMA, warps 1-8: cp.async

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const void* map, uint64_t* bar, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map),
               "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}

// mode bit 0: TMA stream (warp 0), bit 1: cp.async stream (warps 1-8), each `iters` rounds of 32 KB over its own slice of the buffer
__global__ void __launch_bounds__(THREADS, 1) probe(const __grid_constant__ CUtensorMap map, const uint8_t* __restrict__ src, long long rows_total,
                                                    int iters, int mode, int pitch, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* lsu_all = sm + STAGES * 2 * BOX_BYTES;  // LSU_DEPTH x 32 KB landing zones of the cp.async stream
  __shared__ uint64_t full[STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0)
    for (int i = 0; i < STAGES; ++i) mbar_init(&full[i], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const long long rows_per_cta = rows_total / gridDim.x;
  const long long row0 = (long long)blockIdx.x * rows_per_cta;
  long long t0, t1;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0)::"memory");
  if (warp == 0 && (mode & 1)) {
    if (lane == 0) {
      // keep STAGES stages in flight: wait for the oldest, reissue
      for (int i = 0; i < iters + STAGES; ++i) {
        const int s = i % STAGES;
        if (i >= STAGES) mbar_wait(&full[s], ((i / STAGES) - 1) & 1);
        if (i < iters) {
          const long long r = row0 + ((long long)i * 2 * BOX_ROWS) % (rows_per_cta - 2 * BOX_ROWS);
          mbar_expect(&full[s], 2 * BOX_BYTES);
          tma_2d(smem_u32(sm + s * 2 * BOX_BYTES), &map, &full[s], 0, (int)r);
          tma_2d(smem_u32(sm + s * 2 * BOX_BYTES + BOX_BYTES), &map, &full[s], 0, (int)r + BOX_ROWS);
        }
      }
    }
  } else if (warp >= 1 && (mode & 2)) {
    // 256 threads x 16 B = 4 KB per instruction; 8 instructions = one 32 KB round; LSU_DEPTH rounds in flight (commit groups)
    const int t = threadIdx.x - 32;
    for (int i = 0; i < iters; ++i) {
      const long long r = row0 + rows_per_cta / 2 + ((long long)i * 2 * BOX_ROWS) % (rows_per_cta / 2 - 2 * BOX_ROWS);
      const uint8_t* g = src + (r + (t >> 3)) * (long long)pitch + (t & 7) * 16;  // 32 rows x 128 B per instruction
      uint8_t* lsu_sm = lsu_all + (i % LSU_DEPTH) * 32768;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(lsu_sm + k * 4096 + t * 16)), "l"(g + (long long)k * 32 * pitch) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 2;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1)::"memory");
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const long long bytes_total = 64ll << 20;  // 64 MB: L2-resident on a 126 MB L2
  uint8_t* buf;
  CK(cudaMalloc(&buf, bytes_total));
  CK(cudaMemset(buf, 1, bytes_total));
  unsigned long long* out;
  CK(cudaMalloc(&out, 148 * sizeof(unsigned long long)));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  const int smem = (STAGES * 2 + LSU_DEPTH * 2) * BOX_BYTES + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int iters = 2000;  // 64 MB per stream per CTA
  const char* names[4] = {"", "TMA only", "cp.async only", "TMA + cp.async"};
  for (int pitch : {128, 1024, 4096}) {  // bytes between the 128-byte rows of a box: contiguous, a 256-channel P32 row, a 1024-channel one
    const long long rows_total = bytes_total / pitch;
    CUtensorMap map;
    cuuint64_t dims[2] = {64, (cuuint64_t)rows_total}, strides[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {64, BOX_ROWS}, estr[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    printf("L2 -> SM ingest per SM (B/clk): rows of 128 B at a pitch of %d B, %d rounds of 32 KB per stream, buffer 64 MB (L2-resident)\n", pitch, iters);
    for (int grid : {1, 8, 74, 148}) {
      if (rows_total / grid < 4 * 2 * BOX_ROWS + 8) continue;
      for (int mode = 1; mode <= 1; ++mode) {  // (a cp.async stream, mode bit 1, is built but its timing was never validated: not run)
        for (int rep = 0; rep < 2; ++rep) {  // first pass warms L2
          probe<<<grid, THREADS, smem>>>(map, buf, rows_total, iters, mode, pitch, out);
          CK(cudaDeviceSynchronize());
        }
        unsigned long long h[148];
        CK(cudaMemcpy(h, out, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        double worst = 0, sum = 0;
        for (int i = 0; i < grid; ++i) { sum += (double)h[i]; if ((double)h[i] > worst) worst = (double)h[i]; }
        const double bytes = (double)iters * 32768.0 * ((mode & 1) + ((mode >> 1) & 1));
        printf("  grid %3d  %-16s  %6.1f B/clk/SM (slowest CTA %6.1f)   chip %7.0f B/clk\n", grid, names[mode], bytes / (sum / grid), bytes / worst,
               bytes * grid / worst);
      }
    }
  }
  return 0;
}
