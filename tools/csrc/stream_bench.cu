// Dev microbenchmark (NOT part of libdtk_b200.so: built by tools/stream_bench.py into tools/libdtk_dev.so): how fast can one persistent CTA per SM stream a large buffer from
// HBM, (mode 0) with 1-D TMA bulk copies into a shared-memory ring, or (mode 1) with 128-bit LDG by all
// warps? Used to size the ring / chunking of the persistent decode kernel. Not on the product path.
#include "../../detikzify_b200/csrc/common.cuh"

namespace dtk {
namespace {

DTK_DEV void sb_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count)); }
DTK_DEV void sb_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory"); }
DTK_DEV void sb_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory"); }
DTK_DEV void sb_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = clock64();
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && clock64() - t0 > 4000000000ll) __trap();
  }
}
DTK_DEV void sb_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, int hint) {
  if (hint) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
  } else {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
  }
}

// mode 0: warps [0, ncw) consume, warps [ncw, ncw+npw) produce (lane 0). chunk k -> CTA k % G.
// mode 1: every warp streams chunks with LDG (chunk k -> global warp k % (G * nwarps)).
__global__ void __launch_bounds__(512, 1) stream_bench_kernel(const uint8_t* __restrict__ buf, unsigned long long bytes, int mode,
                                                             int chunk, int nslots, int ncw, int npw, int read_smem, int hint,
                                                             float* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int c = blockIdx.x, G = gridDim.x;
  const unsigned long long nchunks = bytes / (unsigned long long)chunk;
  float acc = 0.f;
  if (mode == 1) {
    const unsigned long long gw = (unsigned long long)c * nwarps + warp, GW = (unsigned long long)G * nwarps;
    for (unsigned long long k = gw; k < nchunks; k += GW) {
      const uint4* p = reinterpret_cast<const uint4*>(buf + k * chunk);
      const int n16 = chunk >> 4;
      for (int i = lane; i < n16; i += 32 * 8) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (i + u * 32 < n16) ? ldg_stream(p + i + u * 32) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += __uint_as_float(v[u].x ^ v[u].y ^ v[u].z ^ v[u].w);
      }
    }
    if (acc == 123.456f) sink[0] = acc;
    return;
  }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nslots * chunk);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + nslots), ring = smem_u32(smem);
  if (tid == 0) {
    for (int s = 0; s < nslots; ++s) { sb_mbar_init(full0 + 8 * s, 1); sb_mbar_init(empty0 + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  // local chunk n of this CTA = global chunk c + n * G ; slot n % nslots ; producer n % npw ; consumer n % ncw
  const unsigned long long nlocal = (nchunks > (unsigned long long)c) ? (nchunks - c + G - 1) / G : 0;
  if (warp >= ncw) {
    const int pw = warp - ncw;
    if (pw < npw && lane == 0) {
      for (unsigned long long n = pw; n < nlocal; n += npw) {
        const uint32_t s = (uint32_t)(n % nslots), use = (uint32_t)(n / nslots);
        if (use > 0) sb_wait(empty0 + 8 * s, (use - 1) & 1);
        sb_expect_tx(full0 + 8 * s, chunk);
        sb_bulk(ring + s * chunk, buf + ((unsigned long long)c + n * G) * chunk, chunk, full0 + 8 * s, hint);
      }
    }
    return;
  }
  for (unsigned long long n = warp; n < nlocal; n += ncw) {
    const uint32_t s = (uint32_t)(n % nslots), use = (uint32_t)(n / nslots);
    sb_wait(full0 + 8 * s, use & 1);
    if (read_smem) {
      const uint4* p = reinterpret_cast<const uint4*>(smem + (size_t)s * chunk);
      for (int i = lane; i < (chunk >> 4); i += 32) { uint4 v = p[i]; acc += __uint_as_float(v.x ^ v.y ^ v.z ^ v.w); }
    }
    __syncwarp();
    if (lane == 0) sb_arrive(empty0 + 8 * s);
  }
  if (acc == 123.456f) sink[0] = acc;
}

}  // namespace
}  // namespace dtk

extern "C" __attribute__((visibility("default"))) int dtk_dbg_stream_bench(const void* buf, uint64_t bytes, int mode, int chunk, int nslots, int ncw, int npw,
                                     int read_smem, int hint, int grid, float* sink, void* stream) {
  using namespace dtk;
  if (chunk <= 0 || (chunk & 15) || nslots <= 0 || ncw <= 0 || npw < 0 || (ncw + npw) * 32 > 512) return -1;
  if (mode == 0 && ((nslots % ncw) || (nslots % npw))) return -1;  // fixed slot ownership
  const int smem = mode == 0 ? nslots * chunk + 2 * nslots * 8 + 64 : 0;
  if (cudaFuncSetAttribute(stream_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem > 1024 ? smem : 1024) != cudaSuccess) return -2;
  stream_bench_kernel<<<grid, (ncw + npw) * 32, smem, (cudaStream_t)stream>>>((const uint8_t*)buf, bytes, mode, chunk, nslots, ncw, npw, read_smem, hint, sink);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
