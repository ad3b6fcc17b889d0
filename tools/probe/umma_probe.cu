// Hardware probe (sm_100a): how does tcgen05.mma address a 128B-swizzled K-major A operand whose start
// address is shifted by whole 128-byte rows and whose 8-row-group stride (SBO) is not a multiple of 1024?
// A tile G[R][64] bf16 is TMA-loaded (SWIZZLE_128B) to a 1024-aligned smem buffer; B = identity (64x64), so
// D[m][n] = A_seen[m][n].  The host prints, for every variant, which row of G each D row equals.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../diffusion_model_project_b200/csrc/b2d_ptx.cuh"
using namespace b2d;

constexpr int R = 320;  // rows of G resident in smem (40 KB)

struct Params {
  CUtensorMap tmA, tmB;
  float* out;      // [128][64]
  int row_off;     // start row shift
  int sbo;         // bytes
  int base_off;    // descriptor base_offset field
  int box_rows;    // rows per TMA box of A
};

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // R x 128 B
  uint8_t* sB = smem + R * 128;       // 64 x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 64 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], R * 128 + 64 * 128);
    for (int r = 0; r < R; r += p.box_rows) tma_load_2d(sA + r * 128, &p.tmA, &bars[0], 0, r);
    tma_load_2d(sB, &p.tmB, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 64);
    uint64_t ad = umma_smem_desc(smem_u32(sA + p.row_off * 128), (uint32_t)p.sbo, 2) | (uint64_t(p.base_off & 7) << 49);
    uint64_t bd = umma_smem_desc(smem_u32(sB), 1024, 2);
    for (int k = 0; k < 4; ++k) umma_bf16(tb, ad + 2 * k, bd + 2 * k, idesc, k > 0);
    umma_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tb + (uint32_t(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.out[threadIdx.x * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 64); }
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                        CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  PFN enc = (PFN)fp;
  // G[r][c] = r + c/64 (exact in bf16 for r < 256? use r*0.5 style: value = r, c encoded separately)
  std::vector<__nv_bfloat16> hG(R * 64), hB(64 * 64);
  for (int r = 0; r < R; ++r) for (int c = 0; c < 64; ++c) hG[r * 64 + c] = __float2bfloat16((float)(r % 256) + (c == 1 ? 0.f : 0.f) + (c >= 2 ? 0.f : 0.f));
  // column c of row r: bf16 can hold integers up to 256 exactly; encode row in col 0, (row>=256) in col 1, col index in others
  for (int r = 0; r < R; ++r) for (int c = 0; c < 64; ++c) {
    float v = (c == 0) ? (float)(r % 256) : (c == 1) ? (float)(r / 256) : (float)c;
    hG[r * 64 + c] = __float2bfloat16(v);
  }
  for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *dG, *dB; float* dO;
  cudaMalloc(&dG, hG.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dG, hG.data(), hG.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int box_rows_list[2] = {64, 10};
  for (int bi = 0; bi < 2; ++bi) {
    Params p;
    p.box_rows = box_rows_list[bi];
    {
      cuuint64_t dims[2] = {64, (cuuint64_t)R}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, (cuuint32_t)p.box_rows}; cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dG, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode A failed %d\n", (int)r); return 1; }
      cuuint64_t dimsb[2] = {64, 64}; cuuint32_t boxb[2] = {64, 64};
      r = enc(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsb, str, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode B failed %d\n", (int)r); return 1; }
    }
    p.out = dO;
    const int sbos[3] = {1024, 1280, 2048};
    for (int si = 0; si < 3; ++si)
      for (int off = 0; off < 4; ++off)
        for (int bo = 0; bo < 2; ++bo) {
          if (bo == 1 && off == 0) continue;
          p.row_off = off; p.sbo = sbos[si]; p.base_off = bo ? (off & 7) : 0;
          cudaMemset(dO, 0, 128 * 64 * 4);
          probe_kernel<<<1, 128, R * 128 + 64 * 128 + 64 + 1024>>>(p);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("box %d sbo %d off %d bo %d: CUDA error %s\n", p.box_rows, p.sbo, off, p.base_off, cudaGetErrorString(e)); return 2; }
          std::vector<float> hO(128 * 64);
          cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
          // expected row map under the "absolute address" model
          int ok_rows = 0, clean_rows = 0;
          printf("box_rows %2d sbo %4d off %d base_off %d : ", p.box_rows, p.sbo, off, p.base_off);
          std::vector<int> got(128);
          for (int m = 0; m < 128; ++m) {
            const float* d = &hO[m * 64];
            int row = (int)d[0] + 256 * (int)d[1];
            bool clean = true;
            for (int c = 2; c < 64; ++c) if (d[c] != (float)c) clean = false;
            const int expect = off + (m / 8) * (p.sbo / 128) + (m % 8);
            got[m] = clean ? row : -1;
            clean_rows += clean;
            ok_rows += (clean && row == expect);
          }
          printf("rows matching absolute-address model %3d/128, column-clean rows %3d/128 | first 24 rows:", ok_rows, clean_rows);
          for (int m = 0; m < 24; ++m) printf(" %d", got[m]);
          printf("\n");
        }
  }
  return 0;
}
