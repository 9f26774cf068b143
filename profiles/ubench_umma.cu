// Micro-benchmarks behind the K2 tile-shape decisions (B200, sm_100a):
//   1. cycles per tcgen05.mma (cta_group::2, M=256, K=16, bf16) by N and by where A lives
//      (shared-memory descriptor vs TMEM), issued back to back by one thread, no loads in flight;
//   2. TMA (cp.async.bulk.tensor) throughput per SM by box shape, all SMs loading from L2-resident data.
// Build + run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include \
//                    -o gpurun_out/ubench_umma profiles/ubench_umma.cu -lcuda && gpurun_out/ubench_umma
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../quadruplet-sentence-transformer_b200/csrc/sm100_ptx.cuh"

using namespace qst;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// ------------------------------------------------------------------------------------------
// 1. MMA issue/throughput
// ------------------------------------------------------------------------------------------
template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) mma_bench(int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const uint32_t rank = ptx::cluster_ctarank();
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
  ptx::fence_proxy_async();
  if (threadIdx.x < 32) ptx::tmem_alloc_pair(ptx::smem_u32(&s_tmem), 512);
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 0 && rank == 0) {
    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(256, N);
    const uint64_t da = ptx::make_sw128_kmajor_desc(base);            // A: 128 rows x 64 k (16 KB)
    const uint64_t db = ptx::make_sw128_kmajor_desc(base + 16384);    // B: up to 128 rows x 64 k
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (TS) ptx::umma_bf16_pair_ts(tmem + 256, tmem + (uint32_t)(8 * k), db + 2 * k, idesc, 1u);
        else ptx::umma_bf16_pair(tmem + 256, da + 2 * k, db + 2 * k, idesc, 1u);
      }
    }
    ptx::umma_commit_pair(ptx::smem_u32(&bar), 1);
    const long long t1 = clock64();
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc_pair(tmem, 512); }
}

template <int N, bool TS>
static void run_mma(int grid, long long* d_out) {
  const int reps = 4096;
  auto kern = mma_bench<N, TS>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 64 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int it = 0; it < 2; ++it) CK(cudaLaunchKernelEx(&cfg, kern, reps, d_out));
  CK(cudaDeviceSynchronize());
  long long h[2];
  CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
  printf("mma pair M=256 N=%3d A=%s grid=%3d: issue %.1f cyc/mma, complete %.1f cyc/mma (ideal %d)\n", N, TS ? "tmem" : "smem",
         grid, (double)h[0] / (reps * 4), (double)h[1] / (reps * 4), N / 2);
}

// ------------------------------------------------------------------------------------------
// 2. TMA throughput by box shape.  Every CTA loads boxes from a 64 MB (L2-resident after the first
//    pass) bf16 matrix [rows, 768] into a ring of `depth` smem slots.
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled encode_fn() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  return reinterpret_cast<PFN_encodeTiled>(p);
}

template <int RANK>
__global__ void __launch_bounds__(128, 1) tma_bench(const __grid_constant__ CUtensorMap map, int box_bytes, int box_rows,
                                                    int kb_per_box, int n_boxes, int depth, int producers, int rows_total,
                                                    long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[4][16];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int p = 0; p < 4; ++p)
      for (int i = 0; i < 16; ++i) ptx::mbar_init(ptx::smem_u32(&full[p][i]), 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&map);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && warp < producers) {
    const uint32_t my_base = base + (uint32_t)(warp * depth * box_bytes);
    const long long t0 = clock64();
    int row = (int)(((long long)(blockIdx.x * 4 + warp) * 4099) % (rows_total - box_rows));
    int kb = 0, slot = 0;
    uint32_t ph = 0;
    long long t_issue = 0;
    for (int i = 0; i < n_boxes + depth; ++i) {
      if (i >= depth) ptx::mbar_wait(ptx::smem_u32(&full[warp][slot]), ph ^ 1u);
      if (i < n_boxes) {
        const long long a = clock64();
        ptx::mbar_arrive_expect_tx(ptx::smem_u32(&full[warp][slot]), (uint32_t)box_bytes);
        if (RANK == 2) {
          ptx::tma_load_2d(my_base + slot * box_bytes, &map, ptx::smem_u32(&full[warp][slot]), kb * 64, row, ptx::kEvictNormal);
        } else {
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
              " [%0], [%1, {%3, %4, %5}], [%2], %6;"
              ::"r"(my_base + slot * box_bytes), "l"(reinterpret_cast<uint64_t>(&map)), "r"(ptx::smem_u32(&full[warp][slot])),
                "r"(0), "r"(row), "r"(kb), "l"(ptx::kEvictNormal)
              : "memory");
        }
        t_issue += clock64() - a;
        kb += kb_per_box;
        if (kb >= 12) { kb = 0; row += box_rows; if (row > rows_total - box_rows) row = 0; }
      }
      if (++slot == depth) { slot = 0; ph ^= 1u; }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && warp == 0) { out[0] = t1 - t0; out[1] = t_issue; }
  }
}

static void run_tma(PFN_encodeTiled enc, void* data, int rows_total, int rank, int box_rows, int kb_per_box, int depth,
                    int producers, long long* d_out) {
  CUtensorMap map;
  const int D = 768;
  CUresult r;
  if (rank == 2) {
    cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)rows_total};
    cuuint64_t gstr[1] = {(cuuint64_t)D * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, data, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t gdim[3] = {64, (cuuint64_t)rows_total, (cuuint64_t)(D / 64)};
    cuuint64_t gstr[2] = {(cuuint64_t)D * 2, 128};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, (cuuint32_t)kb_per_box};
    cuuint32_t es[3] = {1, 1, 1};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, data, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) { printf("tma rank=%d box_rows=%d kb=%d: encode failed (%d)\n", rank, box_rows, kb_per_box, (int)r); return; }
  const int box_bytes = box_rows * 128 * (rank == 3 ? kb_per_box : 1);
  const int n_boxes = (8 << 20) / box_bytes / producers;   // 8 MB per CTA
  const size_t smem = (size_t)producers * depth * box_bytes + 1024;
  if (smem > 200 * 1024) return;
  auto k2 = tma_bench<2>;
  auto k3 = tma_bench<3>;
  CK(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int it = 0; it < 2; ++it) {
    if (rank == 2) k2<<<148, 128, smem>>>(map, box_bytes, box_rows, 1, n_boxes, depth, producers, rows_total, d_out);
    else k3<<<148, 128, smem>>>(map, box_bytes, box_rows, kb_per_box, n_boxes, depth, producers, rows_total, d_out);
  }
  CK(cudaDeviceSynchronize());
  long long h[2];
  CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
  printf("tma %dD box %3d rows x %d kb (%5d B) depth %2d x %d producers: %.1f cyc/box/SM, %.1f B/clk/SM, issue %.1f cyc/box\n",
         rank, box_rows, rank == 3 ? kb_per_box : 1, box_bytes, depth, producers, (double)h[0] / (n_boxes * producers),
         (double)box_bytes * n_boxes * producers / (double)h[0], (double)h[1] / n_boxes);
}

int main() {
  long long* d_out;
  CK(cudaMalloc(&d_out, 64));
  for (int grid : {2, 148}) {
    run_mma<64, false>(grid, d_out);
    run_mma<64, true>(grid, d_out);
    run_mma<128, false>(grid, d_out);
    run_mma<128, true>(grid, d_out);
    run_mma<256, false>(grid, d_out);
    run_mma<256, true>(grid, d_out);
  }
  const int rows_total = 40000;   // x 1536 B = 61 MB: L2-resident
  void* data;
  CK(cudaMalloc(&data, (size_t)rows_total * 768 * 2));
  CK(cudaMemset(data, 0, (size_t)rows_total * 768 * 2));
  PFN_encodeTiled enc = encode_fn();
  for (int producers : {1, 2, 4}) {
    run_tma(enc, data, rows_total, 2, 32, 1, 4, producers, d_out);
    run_tma(enc, data, rows_total, 2, 128, 1, 4, producers, d_out);
    run_tma(enc, data, rows_total, 2, 256, 1, 2, producers, d_out);
    run_tma(enc, data, rows_total, 3, 64, 4, 2, producers, d_out);
    run_tma(enc, data, rows_total, 3, 128, 2, 2, producers, d_out);
  }
  run_tma(enc, data, rows_total, 3, 128, 4, 2, 1, d_out);
  run_tma(enc, data, rows_total, 3, 128, 4, 3, 1, d_out);
  run_tma(enc, data, rows_total, 3, 64, 4, 4, 1, d_out);
  run_tma(enc, data, rows_total, 3, 64, 4, 6, 1, d_out);
  run_tma(enc, data, rows_total, 2, 128, 1, 8, 1, d_out);
  run_tma(enc, data, rows_total, 2, 128, 1, 12, 1, d_out);
  return 0;
}
