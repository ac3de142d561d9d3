// Dev microbenchmark (NOT part of libdtk_b200.so): how fast can the 8 consumer warps of one CTA turn shared-memory resident
// 8 KB weight tiles (16 rows x 256 k, bf16) into GEMV partial sums on the legacy tensor pipe? Variants isolate what bounds
// the decode kernel's tile loop: ldmatrix vs 128-bit LDS of fragment-ordered tiles, number of accumulator chains, and the
// per-tile bookkeeping (partial-sum store + shared-memory atomic). One CTA per SM, nothing touches HBM in the timed loop.
#include "../../detikzify_b200/csrc/common.cuh"

namespace dtk {
namespace {
constexpr int NW = 8;

template <int VAR>
__global__ void __launch_bounds__(NW * 32, 1) tile_bench_kernel(int iters, int nslots, float* out, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* tpart = reinterpret_cast<float*>(smem + (size_t)nslots * 8192);
  int* gcnt = reinterpret_cast<int*>(tpart + 64 * 16);
  uint4* xb = reinterpret_cast<uint4*>(gcnt + 64);
  for (int i = tid; i < nslots * 8192 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 7);
  for (int i = tid; i < 64 * 4; i += blockDim.x) xb[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3c003c00u, 0x3c003c00u);
  if (tid < 64) gcnt[tid] = 0;
  __syncthreads();
  const uint32_t ring = smem_u32(smem);
  float total = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int sl = (warp + it * NW) % nslots;
    const uint32_t ta = ring + sl * 8192 + lane * 16;
    const uint2* xp = reinterpret_cast<const uint2*>(xb + (size_t)(it & 7) * 0 + (lane & 3)) + ((lane >> 2) & 1);
    uint2 b[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) b[s] = xp[(s & 15) * 8];
    uint32_t a[16][4];
    if (VAR & 1) {   // fragment-ordered tile: one 128-bit LDS per lane and k-step
#pragma unroll
      for (int s = 0; s < 16; ++s)
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(a[s][0]), "=r"(a[s][1]), "=r"(a[s][2]), "=r"(a[s][3]) : "r"(ta + s * 512));
    } else {
#pragma unroll
      for (int s = 0; s < 16; ++s) ldmatrix_x4(a[s][0], a[s][1], a[s][2], a[s][3], ta + s * 512);
    }
    float acc[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[c][e] = 0.f;
    if (!(VAR & 8)) {
      if (VAR & 2) {   // four accumulator chains
#pragma unroll
        for (int s = 0; s < 16; ++s) mma_bf16_16816(acc[s & 3], a[s], b[s].x, b[s].y);
      } else {         // two chains
#pragma unroll
        for (int s = 0; s < 16; ++s) mma_bf16_16816(acc[s & 1], a[s], b[s].x, b[s].y);
      }
    } else {           // no mma: consume the fragments with integer adds
#pragma unroll
      for (int s = 0; s < 16; ++s) acc[0][0] += __uint_as_float((a[s][0] ^ a[s][1] ^ a[s][2] ^ a[s][3]) & 0x3fffffffu);
    }
    const float r0 = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]) + (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
    const float r2 = (acc[0][2] + acc[1][2]) + (acc[2][2] + acc[3][2]) + (acc[0][3] + acc[1][3]) + (acc[2][3] + acc[3][3]);
    if (!(VAR & 4)) {  // bookkeeping of the decode kernel: partial sums to shared memory + group counter
      if ((lane & 3) == 0) {
        float* tp = tpart + ((warp + it * NW) & 63) * 16;
        tp[lane >> 2] = r0;
        tp[(lane >> 2) + 8] = r2;
      }
      __syncwarp();
      int last = 0;
      if (lane == 0) {
        __threadfence_block();
        last = (atomicAdd(&gcnt[it & 63], 1) == 7);
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last) {
        float v = 0.f;
        if (lane < 16)
          for (int t = 0; t < 8; ++t) v += *reinterpret_cast<volatile float*>(tpart + ((t + it * NW) & 63) * 16 + lane);
        total += v;
        if (lane == 0) gcnt[it & 63] = 0;
      }
    } else {
      total += r0 + r2;
    }
  }
  const long long t1 = clock64();
  if (total == 123.456f) out[0] = total;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

// Two tiles per warp iteration, half-tiles (8 k-steps) interleaved through three fragment buffers: the shared-memory
// reads of one tile overlap the mma of the other inside ONE warp (two warps per scheduler cannot hide it by themselves).
// BOOK = per-tile bookkeeping of the decode kernel (partial sums + shared-memory atomic) for both tiles at the end.
template <bool BOOK>
__global__ void __launch_bounds__(NW * 32, 1) tile_bench2_kernel(int iters, int nslots, float* out, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* tpart = reinterpret_cast<float*>(smem + (size_t)nslots * 8192);
  int* gcnt = reinterpret_cast<int*>(tpart + 64 * 16);
  uint4* xb = reinterpret_cast<uint4*>(gcnt + 64);
  for (int i = tid; i < nslots * 8192 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 7);
  for (int i = tid; i < 64 * 4; i += blockDim.x) xb[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3c003c00u, 0x3c003c00u);
  if (tid < 64) gcnt[tid] = 0;
  __syncthreads();
  const uint32_t ring = smem_u32(smem);
  float total = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it += 2) {
    const uint32_t ta0 = ring + ((warp + it * NW) % nslots) * 8192 + lane * 16;
    const uint32_t ta1 = ring + ((warp + (it + 1) * NW) % nslots) * 8192 + lane * 16;
    const uint2* xp = reinterpret_cast<const uint2*>(xb + (lane & 3)) + ((lane >> 2) & 1);
    uint2 b[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) b[s] = xp[s * 8];
    uint32_t X[8][4], Y[8][4], Z[8][4];
    float acc0[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, acc1[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int s = 0; s < 8; ++s) ldmatrix_x4(X[s][0], X[s][1], X[s][2], X[s][3], ta0 + s * 512);
#pragma unroll
    for (int s = 0; s < 8; ++s) ldmatrix_x4(Y[s][0], Y[s][1], Y[s][2], Y[s][3], ta1 + s * 512);
#pragma unroll
    for (int s = 0; s < 8; ++s) mma_bf16_16816(acc0[s & 1], X[s], b[s].x, b[s].y);
#pragma unroll
    for (int s = 0; s < 8; ++s) ldmatrix_x4(Z[s][0], Z[s][1], Z[s][2], Z[s][3], ta0 + (s + 8) * 512);
#pragma unroll
    for (int s = 0; s < 8; ++s) mma_bf16_16816(acc1[s & 1], Y[s], b[s].x, b[s].y);
#pragma unroll
    for (int s = 0; s < 8; ++s) ldmatrix_x4(X[s][0], X[s][1], X[s][2], X[s][3], ta1 + (s + 8) * 512);
#pragma unroll
    for (int s = 0; s < 8; ++s) mma_bf16_16816(acc0[s & 1], Z[s], b[s + 8].x, b[s + 8].y);
#pragma unroll
    for (int s = 0; s < 8; ++s) mma_bf16_16816(acc1[s & 1], X[s], b[s + 8].x, b[s + 8].y);
    const float r00 = (acc0[0][0] + acc0[1][0]) + (acc0[0][1] + acc0[1][1]), r02 = (acc0[0][2] + acc0[1][2]) + (acc0[0][3] + acc0[1][3]);
    const float r10 = (acc1[0][0] + acc1[1][0]) + (acc1[0][1] + acc1[1][1]), r12 = (acc1[0][2] + acc1[1][2]) + (acc1[0][3] + acc1[1][3]);
    if (BOOK) {
      if ((lane & 3) == 0) {
        float* tp = tpart + ((warp + it * NW) & 63) * 16;
        tp[lane >> 2] = r00; tp[(lane >> 2) + 8] = r02;
        float* tq = tpart + ((warp + (it + 1) * NW) & 63) * 16;
        tq[lane >> 2] = r10; tq[(lane >> 2) + 8] = r12;
      }
      __syncwarp();
      int l0 = 0, l1 = 0;
      if (lane == 0) {
        __threadfence_block();
        l0 = atomicAdd(&gcnt[it & 63], 1);
        l1 = atomicAdd(&gcnt[(it + 1) & 63], 1);
      }
      const int last0 = __shfl_sync(0xffffffffu, l0 == 7, 0), last1 = __shfl_sync(0xffffffffu, l1 == 7, 0);
      if (last0 | last1) {
        float v = 0.f;
        if (lane < 16)
          for (int t = 0; t < 8; ++t) v += *reinterpret_cast<volatile float*>(tpart + ((t + (it + (last1 ? 1 : 0)) * NW) & 63) * 16 + lane);
        total += v;
        if (lane == 0) { if (last0) gcnt[it & 63] = 0; if (last1) gcnt[(it + 1) & 63] = 0; }
      }
    } else {
      total += r00 + r02 + r10 + r12;
    }
  }
  const long long t1 = clock64();
  if (total == 123.456f) out[0] = total;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}
}  // namespace
}  // namespace dtk

extern "C" __attribute__((visibility("default"))) int dtk_dbg_tile_bench(int variant, int iters, int nslots, int grid, float* out,
                                                                          long long* cycles, void* stream) {
  using namespace dtk;
  const int smem = nslots * 8192 + 64 * 16 * 4 + 64 * 4 + 64 * 4 * 16 + 256;
#define LAUNCH(V)                                                                                                        \
  case V:                                                                                                                \
    if (cudaFuncSetAttribute(tile_bench_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -2; \
    tile_bench_kernel<V><<<grid, NW * 32, smem, (cudaStream_t)stream>>>(iters, nslots, out, cycles);                     \
    break;
  if (variant == 16 || variant == 20) {
    if (variant == 16) {
      if (cudaFuncSetAttribute(tile_bench2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -2;
      tile_bench2_kernel<true><<<grid, NW * 32, smem, (cudaStream_t)stream>>>(iters, nslots, out, cycles);
    } else {
      if (cudaFuncSetAttribute(tile_bench2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -2;
      tile_bench2_kernel<false><<<grid, NW * 32, smem, (cudaStream_t)stream>>>(iters, nslots, out, cycles);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
  }
  switch (variant) {
    LAUNCH(0) LAUNCH(1) LAUNCH(2) LAUNCH(3) LAUNCH(4) LAUNCH(5) LAUNCH(6) LAUNCH(7) LAUNCH(8) LAUNCH(9) LAUNCH(12) LAUNCH(13)
    default: return -1;
  }
#undef LAUNCH
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
