// Micro-benchmark (evidence for DESIGN.md §kernels): how should a sweep write the 64x64 valid tile of
// each 66x66-pitch plane of a level arena?  All variants are persistent (148 CTAs x 512 threads), each
// CTA owning a contiguous range of planes, like the streaming sweep.
//   0 dense      : write whole 66x66 planes contiguously (upper bound, not a legal sweep output)
//   1 row64      : warp stores 32 consecutive doubles of a valid row (8 B-misaligned 512 B rows)
//   2 row128     : lanes 0..30 store aligned 16 B pairs (STG.128), lane 31 the two odd ends
//   3 tma tile   : cp.async.bulk.tensor store of the 64x64x1 box from a dense smem tile
//   4 tma rows4  : per-warp cp.async.bulk.tensor stores of 64x4x1 boxes (no CTA-wide sync)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_patterns store_patterns.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int N0 = 66, PLANE = 66 * 66, V = 64;

__global__ void __launch_bounds__(512, 1) k_dense(double* out, int nplanes) {
  const int p0 = (long long)blockIdx.x * nplanes / gridDim.x, p1 = (long long)(blockIdx.x + 1) * nplanes / gridDim.x;
  for (int p = p0; p < p1; ++p) {
    double* o = out + (size_t)p * PLANE;
    for (int i = threadIdx.x; i < PLANE; i += 512) o[i] = (double)p;
  }
}
__global__ void __launch_bounds__(512, 1) k_row64(double* out, int nplanes) {
  const int p0 = (long long)blockIdx.x * nplanes / gridDim.x, p1 = (long long)(blockIdx.x + 1) * nplanes / gridDim.x;
  for (int p = p0; p < p1; ++p) {
    double* o = out + (size_t)p * PLANE;
    if (p % 66 == 0 || p % 66 == 65) continue;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int idx = threadIdx.x + m * 512, i = idx % V, j = idx / V;
      o[(j + 1) * N0 + i + 1] = (double)p;
    }
  }
}
__global__ void __launch_bounds__(512, 1) k_row128(double* out, int nplanes) {
  const int p0 = (long long)blockIdx.x * nplanes / gridDim.x, p1 = (long long)(blockIdx.x + 1) * nplanes / gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int p = p0; p < p1; ++p) {
    double* o = out + (size_t)p * PLANE;
    if (p % 66 == 0 || p % 66 == 65) continue;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int j = warp + 16 * m;
      double* row = o + (j + 1) * N0;          // row[0] ghost, row[1..64] valid, row start 16 B aligned
      if (lane < 31) *reinterpret_cast<double2*>(row + 2 + 2 * lane) = make_double2((double)p, (double)p);
      else { row[1] = (double)p; row[64] = (double)p; }
    }
  }
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(512, 1) k_tma_tile(const __grid_constant__ CUtensorMap tm, int nplanes) {
  extern __shared__ __align__(128) unsigned char dsm[];
  double (*tile)[V * V] = reinterpret_cast<double (*)[V * V]>(dsm);
  const int p0 = (long long)blockIdx.x * nplanes / gridDim.x, p1 = (long long)(blockIdx.x + 1) * nplanes / gridDim.x;
  int buf = 0;
  for (int p = p0; p < p1; ++p) {
    if (p % 66 == 0 || p % 66 == 65) continue;
    // the store issued two planes ago must have finished READING this buffer
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 8; ++m) tile[buf][threadIdx.x + m * 512] = (double)p;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&tm), "r"(1),
                   "r"(1), "r"(p), "r"(s32(tile[buf]))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    buf ^= 1;
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(512, 1) k_tma_rows4(const __grid_constant__ CUtensorMap tm, int nplanes) {
  extern __shared__ __align__(128) unsigned char dsm[];      // per warp, double buffered: 64 KB
  double (*stage)[2][V * 4] = reinterpret_cast<double (*)[2][V * 4]>(dsm);
  const int p0 = (long long)blockIdx.x * nplanes / gridDim.x, p1 = (long long)(blockIdx.x + 1) * nplanes / gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int buf = 0;
  for (int p = p0; p < p1; ++p) {
    if (p % 66 == 0 || p % 66 == 65) continue;
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 8; ++m) stage[warp][buf][lane + 32 * m] = (double)p;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&tm), "r"(1),
                   "r"(1 + 4 * warp), "r"(p), "r"(s32(stage[warp][buf]))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    buf ^= 1;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn enc, double* base, int nplanes, int bx, int by) {
  CUtensorMap tm;
  cuuint64_t dims[3] = {66, 66, (cuuint64_t)nplanes};
  cuuint64_t strides[2] = {66 * 8, 66 * 66 * 8};
  cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
  return tm;
}

int main() {
  const int nboxes = 512, nplanes = nboxes * 66;
  double* out;
  const size_t bytes = (size_t)nplanes * PLANE * 8;
  CK(cudaMalloc(&out, bytes + 4096));
  CK(cudaMemset(out, 0, bytes));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  EncodeFn enc = (EncodeFn)fn;
  CUtensorMap tm_tile = make_map(enc, out, nplanes, 64, 64), tm_rows = make_map(enc, out, nplanes, 64, 4);
  CK(cudaFuncSetAttribute(k_tma_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_tma_rows4, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  const double valid_gb = (double)nboxes * 64 * 64 * 64 * 8 / 1e9;
  for (int v = 0; v < 5; ++v) {
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
      CK(cudaEventRecord(a));
      switch (v) {
        case 0: k_dense<<<148, 512>>>(out, nplanes); break;
        case 1: k_row64<<<148, 512>>>(out, nplanes); break;
        case 2: k_row128<<<148, 512>>>(out, nplanes); break;
        case 3: k_tma_tile<<<148, 512, 65536>>>(tm_tile, nplanes); break;
        case 4: k_tma_rows4<<<148, 512, 65536>>>(tm_rows, nplanes); break;
      }
      CK(cudaEventRecord(b));
      CK(cudaEventSynchronize(b));
      CK(cudaGetLastError());
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      if (rep > 0 && ms < best) best = ms;
    }
    const char* names[] = {"dense", "row64", "row128", "tma_tile", "tma_rows4"};
    const double gb = v == 0 ? bytes / 1e9 : valid_gb;
    printf("%-10s %.3f ms  %.0f GB/s written\n", names[v], best, gb / (best * 1e-3));
  }
  // verify variant outputs agree on a sample (last run = tma_rows4): plane 70 row 5 must hold 70.0 in valid cells
  double h[66];
  CK(cudaMemcpy(h, out + (size_t)70 * PLANE + 6 * N0, sizeof h, cudaMemcpyDeviceToHost));
  printf("check: ghost %.1f valid %.1f %.1f ghost %.1f\n", h[0], h[1], h[64], h[65]);
  return 0;
}
