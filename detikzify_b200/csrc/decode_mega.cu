// Persistent weight-streaming decode kernel: ONE cooperative launch per generated token.
//
// Why: batch-1 decode streams every decoder weight once per token (2.56 GB for ds-1.3b) through ~120
// dependent GEMV-sized steps of a few microseconds each. Launched as separate kernels (even from a
// CUDA graph) the HBM pipe drains at every step boundary and the chain is launch/ramp bound
// (measured 0.37 of the HBM roofline). Here the weight stream is decoupled from the dependency chain:
//
//   * grid = one CTA per SM, resident for the whole token (cooperative launch);
//   * 4 PRODUCER warps per CTA (one issuing lane each, registers handed back with setmaxnreg.dec) walk the
//     CTA's statically known list of weight tiles for ALL layers and phases and stream them with 1-D TMA bulk
//     copies (cp.async.bulk + mbarrier complete_tx) into a ~190 KB shared-memory ring of 8 KB slots, never
//     waiting for activations — weights do not depend on them — so HBM stays busy across phase boundaries
//     (measured: the ring sustains 7.2 TB/s);
//   * 8 CONSUMER warps (setmaxnreg.inc 232: no spills, a whole tile in flight) take tiles in order. A tile is
//     16 output rows x 256 k, pre-arranged in HBM (launch_retile, once at load) so that it lands in shared
//     memory exactly in ldmatrix.x4 order; the dot products run on the tensor pipe (mma.sync m16n8k16, fp32
//     accumulate) with the activation vector split into bf16 hi + lo parts (x = hi + lo to 2^-17) that occupy
//     alternating columns of the B operand: one HMMA per k-step, fp32-grade GEMV; the ring slot is handed back
//     as soon as its shared-memory reads are issued;
//   * rows are grouped so that one thread's two accumulator rows (g, g+8) are a RoPE pair (i, i+64) or a
//     SwiGLU pair (gate_i, up_i): RMSNorm prologue, RoPE + KV-cache write, SiLU*mul and residual add are
//     all fused; partial sums of a group's k-tiles are combined in a fixed order (deterministic);
//   * phases are separated by a hand-rolled grid barrier (release-reduction + acquire poll) among the
//     consumer threads; the ring depth (~4 us of streaming per SM) covers part of the barrier + activation
//     re-staging bubble;
//   * a phase's 16-row groups are cut into equal blocks over as many CTAs as needed (the participating set
//     rotates from phase to phase), so participants finish together and idle CTAs' producers run ahead.
//
// Per layer: P1 qkv(+RMSNorm, RoPE, KV write) | P2 split-KV attention (the CTA's KV rows are L2-prefetched at
// the start of the layer and read with plain loads; the last CTA of a head to arrive merges the partials) |
// P3 o-proj + residual | P4 gate/up + SiLU*mul (+RMSNorm) | P5 down + residual; finally lm_head (+final
// RMSNorm). DESIGN.md section 4 lists the measured alternatives that were rejected.
//
// Replaces the per-token HF eager path (modeling_llama.py:303-333, ~900 launches per token).
#include "common.cuh"
#include "launch.h"

namespace dtk {
namespace {

constexpr int NCW = 8;                       // consumer warps
constexpr int NPW = 4;                       // producer warps (one issuing lane each): ~500 cycles per bulk copy
constexpr int MEGA_THREADS = (NCW + NPW) * 32;
constexpr int CONSUMER_THREADS = NCW * 32;
static_assert(NCW % NPW == 0, "slot ownership: NPW must divide NCW");
constexpr int TILE_BYTES = 8192;             // ring slot = one weight tile = one 16-key K+V attention item
constexpr int NT = 112;                      // per-tile partial-sum entries (>= max tiles/group + tiles in flight)
constexpr int NG = 48;                       // per-group arrival counters / prefetched residual rows (>= groups in flight)
constexpr long long SPIN_CYCLES = 4000000000ll;  // bounded waits (~2 s): trap instead of hanging the GPU

// ------------------------------------------------------------------ mbarrier / bulk-copy PTX
DTK_DEV void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
DTK_DEV void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
DTK_DEV void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
DTK_DEV void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  long long t0 = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && (++spins & 1023u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > SPIN_CYCLES) __trap();
    }
  }
}
DTK_DEV void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
DTK_DEV void consumer_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(CONSUMER_THREADS) : "memory"); }
DTK_DEV float ldcg_f(const float* p) { return __ldcg(p); }

// grid barrier over the consumer threads of all CTAs (producer warps never take part).
// bar.sync makes the CTA's writes visible to thread 0 (cta scope); its release-reduction publishes them
// cumulatively at gpu scope; the acquire poll + bar.sync orders every thread's later ld.cg reads.
DTK_DEV void grid_barrier(unsigned long long* counter, unsigned long long target, int flags) {
  consumer_sync();
  if (flags & 2) return;
  if (threadIdx.x == 0) {
    if (flags & 4) asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;\n" ::"l"(counter), "l"(1ull) : "memory");
    else asm volatile("red.release.gpu.global.add.u64 [%0], %1;\n" ::"l"(counter), "l"(1ull) : "memory");
    uint32_t spins = 0;
    long long t0 = 0;
    unsigned long long v;
    do {
      if (flags & 8) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(counter) : "memory");
      else asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(counter) : "memory");
      if (v < target && (++spins & 1023u) == 0) {
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > SPIN_CYCLES) __trap();
      }
    } while (v < target);
  }
  consumer_sync();
}

// ------------------------------------------------------------------ work description
enum { PH_QKV = 0, PH_O = 2, PH_GU = 3, PH_DOWN = 4, PH_LM = 5 };

// attention split: CTA c handles head c % heads, key range index c / heads (cph ranges per head)
struct AttnSplit {
  int active, head, j0, j1, last;  // keys [j0, j1) among the OLD keys [0, pos); `last` also takes key `pos`
  int cph;                         // CTAs per head
  int n_items;                     // 16-key items
};
DTK_DEV AttnSplit attn_split(const MegaArgs& p, int c, int G, int pos) {
  AttnSplit a;
  int cph = G / p.heads;
  if (cph < 1) cph = 1;            // (heads > G is rejected on the host)
  if (cph > 16) cph = 16;
  a.cph = cph;
  a.active = c < cph * p.heads;
  a.head = c % p.heads;
  const int r = c / p.heads;
  int per = (pos + cph - 1) / cph;
  per = (per + 15) & ~15;
  a.j0 = min(pos, r * per);
  a.j1 = min(pos, a.j0 + per);
  a.last = a.active && (r == cph - 1);
  a.n_items = a.active ? (a.j1 - a.j0 + 15) / 16 : 0;
  return a;
}

// sum over the 256 consumer threads
DTK_DEV float consumer_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  consumer_sync();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < NCW; ++i) t += red[i];
  consumer_sync();
  return t;
}

// Stage a K-vector (optionally RMS-normalised) into shared memory as the B operand of mma.m16n8k16:
// entry [kstep S][t] (uint4) = { hi(x[16S+2t], x[16S+2t+1]), hi(x[16S+2t+8], +9), lo(..2t..), lo(..2t+8..) }
// where hi = bf16(x), lo = bf16(x - hi). All 8 columns of B are the same vector, so every lane of a quad
// column reads entry t = lane & 3. Entries for k >= K (padding up to Kp) are zero.
DTK_DEV void stage_xb(const float* src_f32, const bf16* src_bf16, int K, int Kp, const bf16* norm_w, float eps,
                      uint4* xb, float* red) {
  const int tid = threadIdx.x, nsteps = Kp >> 4;
  constexpr int MAXS = 2;  // k-steps held in registers per thread when normalising (K <= 8192)
  float v[MAXS][16];
  float ss = 0.f;
  if (norm_w) {
#pragma unroll
    for (int u = 0; u < MAXS; ++u) {
      const int S = tid + u * CONSUMER_THREADS;
      if (S < nsteps) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int k = S * 16 + q4 * 4;
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k < K) {
            if (src_bf16) {
              const uint2 raw = *reinterpret_cast<const uint2*>(src_bf16 + k);
              const float2 lo2 = unpack_bf16x2(raw.x), hi2 = unpack_bf16x2(raw.y);
              a = make_float4(lo2.x, lo2.y, hi2.x, hi2.y);
            } else {
              a = __ldcg(reinterpret_cast<const float4*>(src_f32 + k));
            }
            const uint2 wr = *reinterpret_cast<const uint2*>(norm_w + k);
            const float2 w0 = unpack_bf16x2(wr.x), w1 = unpack_bf16x2(wr.y);
            ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
            a.x *= w0.x; a.y *= w0.y; a.z *= w1.x; a.w *= w1.y;   // weight now, 1/rms after the reduction
          }
          v[u][q4 * 4 + 0] = a.x; v[u][q4 * 4 + 1] = a.y; v[u][q4 * 4 + 2] = a.z; v[u][q4 * 4 + 3] = a.w;
        }
      }
    }
    const float r = rsqrtf(consumer_sum(ss, red) / K + eps);
#pragma unroll
    for (int u = 0; u < MAXS; ++u) {
      const int S = tid + u * CONSUMER_THREADS;
      if (S < nsteps) {
        // HF order is (x * rsqrt) * w; here (x * w) * rsqrt — same value up to one fp32 rounding
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = v[u][2 * j] * r, b = v[u][2 * j + 1] * r;
          const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
          hi[j] = pack_bf16x2(__bfloat162float(ah), __bfloat162float(bh));
          lo[j] = pack_bf16x2(a - __bfloat162float(ah), b - __bfloat162float(bh));
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) xb[S * 4 + t] = make_uint4(hi[t], hi[t + 4], lo[t], lo[t + 4]);
      }
    }
  } else {
    for (int S = tid; S < nsteps; S += CONSUMER_THREADS) {
      float w16[16];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int k = S * 16 + q4 * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K) a = __ldcg(reinterpret_cast<const float4*>(src_f32 + k));
        w16[q4 * 4 + 0] = a.x; w16[q4 * 4 + 1] = a.y; w16[q4 * 4 + 2] = a.z; w16[q4 * 4 + 3] = a.w;
      }
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = w16[2 * j], b = w16[2 * j + 1];
        const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
        hi[j] = pack_bf16x2(__bfloat162float(ah), __bfloat162float(bh));
        lo[j] = pack_bf16x2(a - __bfloat162float(ah), b - __bfloat162float(bh));
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) xb[S * 4 + t] = make_uint4(hi[t], hi[t + 4], lo[t], lo[t + 4]);
    }
  }
  consumer_sync();
}

__global__ void __launch_bounds__(MEGA_THREADS, 1) decode_mega_kernel(const MegaArgs p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x, G = gridDim.x;
  const int nslots = p.nslots;
  uint8_t* ring = smem;
  float* actf = reinterpret_cast<float*>(smem + (size_t)nslots * TILE_BYTES);
  uint4* xb = reinterpret_cast<uint4*>(actf);
  uint64_t* bars = reinterpret_cast<uint64_t*>(actf + p.act_floats);
  float* red = reinterpret_cast<float*>(bars + 2 * nslots);  // 16 floats
  float* rope_s = red + 16;                                   // [64][2] cos/sin of this position
  float* tpart = rope_s + 128;                                // [NT][16] per-tile partial sums
  int* gcnt = reinterpret_cast<int*>(tpart + NT * 16);        // [NG] tiles finished per group
  float* rbuf = reinterpret_cast<float*>(gcnt + NG);          // [NG][16] residuals prefetched at a group's first tile
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + nslots);
  const uint32_t ring_u32 = smem_u32(ring);

  if (tid == 0) {
    for (int s = 0; s < nslots; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  const int pos = p.pos[0], slot = p.slots[0];
  int tok = p.tok[0];
  if (tok < 0 || tok >= p.V) tok = 0;
  const int qd = p.heads * 128, kd = p.kv_heads * 128;
  if (tid < 128) rope_s[tid] = p.rope_cs[(int64_t)pos * 128 + tid];
  if (tid < NG) gcnt[tid] = 0;
  __syncthreads();
  const AttnSplit as = attn_split(p, c, G, pos);
  const int kvh = as.head / (p.heads / p.kv_heads);

  // ---- item ownership. The CTA's local TILE sequence (all phases, in order) is dealt to agents by index:
  // local tile n -> ring slot n % nslots, producer warp n % NPW, consumer warp n % NCW (nslots is a multiple of
  // both, so a slot always has the same producer and the same consumer -> mbarrier parity waits never alias).
  struct Walk {
    uint32_t nb = 0;     // local tiles before the current phase
    uint32_t gb = 0;     // local groups before the current phase
    uint32_t rot = 0;    // rotation of the participating CTA set
  };
  // A weight phase with `groups` 16-row groups is cut into equal blocks of per = ceil(groups / G) groups; only
  // ceil(groups / per) CTAs take part (same amount of work each, so they reach the barrier together), the others
  // idle for that phase while their producers prefetch ahead. The participating set rotates from phase to phase.
  auto phase_span = [&](const Walk& w, int groups, int& g0, int& cnt, int& nact) {
    const int per = (groups + G - 1) / G;
    nact = (groups + per - 1) / per;
    const int ci = (int)(((uint32_t)c + (uint32_t)G - w.rot) % (uint32_t)G);
    g0 = ci * per;
    cnt = (ci < nact) ? min(per, groups - g0) : 0;
  };

  if (warp >= NCW) {
    // =============================================================== PRODUCERS
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
    const uint32_t pw = (uint32_t)(warp - NCW);
    if (lane == 0) {
      Walk w;
      // visit own tiles j = j0, j0 + NPW, ... of a phase with `ntiles` local tiles (tpg tiles per group)
      auto for_own = [&](int ntiles, int tpg, auto&& issue) {
        uint32_t j = (pw + NPW - (w.nb & (NPW - 1))) & (NPW - 1);
        if ((int)j < ntiles) {
          const uint32_t n0 = w.nb + j;
          uint32_t sl = n0 % (uint32_t)nslots, use = n0 / (uint32_t)nslots;
          uint32_t k = j / (uint32_t)tpg, ks = j - k * (uint32_t)tpg;
          for (; (int)j < ntiles; j += NPW) {
            if (use > 0) mbar_wait(empty0 + 8 * sl, (use - 1) & 1);
            issue((int)k, (int)ks, ring_u32 + sl * TILE_BYTES, full0 + 8 * sl);
            sl += NPW;
            if (sl >= (uint32_t)nslots) { sl -= nslots; ++use; }
            ks += NPW;
            while (ks >= (uint32_t)tpg) { ks -= tpg; ++k; }
          }
        }
        w.nb += ntiles;
      };
      auto stream_phase = [&](const MegaMat& m, int layer) {
        int g0, cnt, nact;
        phase_span(w, m.groups, g0, cnt, nact);
        const bf16* base = m.base + (int64_t)layer * m.layer_stride;
        for_own(cnt * m.tpg, m.tpg, [&](int k, int ks, uint32_t dst, uint32_t fb) {
          mbar_expect_tx(fb, TILE_BYTES);
          bulk_g2s(dst, base + ((int64_t)(g0 + k) * m.tpg + ks) * MEGA_TILE_ELEMS, TILE_BYTES, fb);
        });
        w.gb += cnt;
        w.rot = (w.rot + (uint32_t)nact) % (uint32_t)G;
      };
      for (int l = 0; l < p.L; ++l) {
        stream_phase(p.qkv, l);
        stream_phase(p.o, l);
        stream_phase(p.gu, l);
        stream_phase(p.down, l);
      }
      stream_phase(p.lm, 0);
    }
    return;
  }

  // ================================================================= CONSUMERS
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
  unsigned long long bar_target = *p.bar_base;  // barriers completed before this launch (x G)
  Walk w;
  // visit own tiles of a phase; body(j, k, ks, smem address of the slot) runs after the bytes landed and must
  // finish READING the slot before calling release()
  uint32_t cur_slot = 0;
  auto release = [&]() {
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * cur_slot);
  };
  auto for_own = [&](int ntiles, int tpg, auto&& body) {
    uint32_t j = ((uint32_t)warp + NCW - (w.nb & (NCW - 1))) & (NCW - 1);
    if ((int)j < ntiles) {
      const uint32_t n0 = w.nb + j;
      uint32_t sl = n0 % (uint32_t)nslots, use = n0 / (uint32_t)nslots;
      uint32_t k = j / (uint32_t)tpg, ks = j - k * (uint32_t)tpg;
      for (; (int)j < ntiles; j += NCW) {
        mbar_wait(full0 + 8 * sl, use & 1);
        cur_slot = sl;
        body((int)j, (int)k, (int)ks, sl);
        sl += NCW;
        if (sl >= (uint32_t)nslots) { sl -= nslots; ++use; }
        ks += NCW;
        while (ks >= (uint32_t)tpg) { ks -= tpg; ++k; }
      }
    }
    w.nb += ntiles;
  };

  // optional phase timestamps (globaltimer ns, comparable across SMs): [CTA][phase][4] = {start, staged, items done, barrier done}
  long long* dbg = p.dbg ? p.dbg + (int64_t)c * (p.L * 5 + 1) * 4 : nullptr;
  int dbg_i = 0;
  auto stamp = [&](int k) {
    if (dbg && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
      dbg[dbg_i * 4 + k] = (long long)t;
    }
  };

  // one weight phase: every tile = 16 k-steps of (ldmatrix.x4, LDS.128, 2 x mma); the warp that finishes a
  // group's last tile sums the group's partials in k order and runs the epilogue for its 16 rows
  auto run_phase = [&](const MegaMat& m, int ph, int layer) {
    int g0, cnt, nact;
    phase_span(w, m.groups, g0, cnt, nact);
    const uint32_t nb0 = w.nb, gb0 = w.gb;
    const int tpg = m.tpg;
    for_own(cnt * tpg, tpg, [&](int j, int k, int ks, uint32_t sl) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      {
        const uint32_t ta = ring_u32 + sl * TILE_BYTES + lane * 16;
        // B operand: even columns of the 16 x 8 B tile carry the hi part of x, odd columns the lo part (column = lane >> 2),
        // so ONE mma per k-step yields W.hi in accumulator column 0 and W.lo in column 1 (legacy HMMA issues only once
        // per ~16 cycles per SM sub-partition on sm_100: the tensor pipe, not memory, paces the post-barrier burst in
        // which a full ring is drained from shared memory). Two independent chains (k-step parity) hide the HMMA latency.
        const uint2* xp = reinterpret_cast<const uint2*>(xb + (size_t)ks * 64 + (lane & 3)) + ((lane >> 2) & 1);
        float c1[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t a[16][4];
        uint2 b[16];
#pragma unroll
        for (int s = 0; s < 16; ++s) {
          ldmatrix_x4(a[s][0], a[s][1], a[s][2], a[s][3], ta + s * 512);
          b[s] = xp[s * 8];
        }
        release();   // every lane's shared-memory reads of the slot are issued; the arrive is ordered after them
#pragma unroll
        for (int s = 0; s < 16; s += 2) {
          mma_bf16_16816(acc, a[s], b[s].x, b[s].y);
          mma_bf16_16816(c1, a[s + 1], b[s + 1].x, b[s + 1].y);
        }
        // lanes with (lane & 3) == 0 hold columns 0 (hi) and 1 (lo) of rows g (c[0], c[1]) and g + 8 (c[2], c[3])
        acc[0] = (acc[0] + c1[0]) + (acc[1] + c1[1]);
        acc[2] = (acc[2] + c1[2]) + (acc[3] + c1[3]);
      }

      const uint32_t gslot = (gb0 + (uint32_t)k) % NG;
      if (ks == 0 && (ph == PH_O || ph == PH_DOWN) && lane < 16) {
        // residual of row (group, lane), fetched at the group's FIRST tile so that its L2 latency is off the
        // critical path of the group's epilogue (the value is final: it was produced before the last barrier)
        const int row = (g0 + k) * 16 + lane;
        float b = 0.f;
        if (row < p.H) b = (ph == PH_O && layer == 0) ? __bfloat162float(p.embed[(int64_t)tok * p.H + row]) : ldcg_f(p.x + row);
        rbuf[gslot * 16 + lane] = b;
      }
      const uint32_t n = nb0 + (uint32_t)j;
      if ((lane & 3) == 0) {
        float* tp = tpart + (n % NT) * 16;
        tp[lane >> 2] = acc[0];
        tp[(lane >> 2) + 8] = acc[2];
      }
      __syncwarp();
      int last = 0;
      if (lane == 0) {
        __threadfence_block();
        last = (atomicAdd(&gcnt[gslot], 1) == tpg - 1);
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (!last) return;
      __threadfence_block();
      // ---- group epilogue (this warp saw the last tile of group k)
      const uint32_t n0 = nb0 + (uint32_t)k * tpg;
      float v = 0.f;
      if (lane < 16)
        for (int t = 0; t < tpg; ++t) v += *reinterpret_cast<volatile float*>(tpart + ((n0 + t) % NT) * 16 + lane);
      const float v1 = __shfl_down_sync(0xffffffffu, v, 8);
      if (lane == 0) gcnt[gslot] = 0;
      if (lane >= 8) return;
      const int gi = g0 + k, r = lane;
      if (ph == PH_QKV) {
        const int hb = gi >> 3, i = ((gi & 7) << 3) + r;      // 128-row block, index inside the half
        const int row0 = hb * 128 + i;
        if (row0 < qd + kd) {
          const float2 csn = *reinterpret_cast<const float2*>(rope_s + i * 2);
          const float y0 = v * csn.x - v1 * csn.y, y1 = v1 * csn.x + v * csn.y;
          if (row0 < qd) { p.q[row0] = y0; p.q[row0 + 64] = y1; }
          else {
            const int kh = (row0 - qd) >> 7;
            bf16* dd = p.kv + (int64_t)slot * p.kv_slot_stride + (int64_t)layer * p.kv_layer_stride + ((int64_t)kh * p.max_len + pos) * 128;
            dd[i] = __float2bfloat16_rn(y0);
            dd[i + 64] = __float2bfloat16_rn(y1);
          }
        } else {
          const int kh = (row0 - qd - kd) >> 7;
          bf16* dd = p.kv + (int64_t)slot * p.kv_slot_stride + (int64_t)layer * p.kv_layer_stride + p.kv_v_offset + ((int64_t)kh * p.max_len + pos) * 128;
          dd[i] = __float2bfloat16_rn(v);
          dd[i + 64] = __float2bfloat16_rn(v1);
        }
      } else if (ph == PH_O) {
        const int r0 = gi * 16 + r, r1 = r0 + 8;
        const float b0 = *reinterpret_cast<volatile float*>(rbuf + gslot * 16 + r), b1 = *reinterpret_cast<volatile float*>(rbuf + gslot * 16 + r + 8);
        if (r0 < p.H) p.x[r0] = b0 + v;
        if (r1 < p.H) p.x[r1] = b1 + v1;
      } else if (ph == PH_GU) {
        const int i = gi * 8 + r;
        if (i < p.I) p.h[i] = silu(v) * v1;
      } else if (ph == PH_DOWN) {
        const int r0 = gi * 16 + r, r1 = r0 + 8;
        const float b0 = *reinterpret_cast<volatile float*>(rbuf + gslot * 16 + r), b1 = *reinterpret_cast<volatile float*>(rbuf + gslot * 16 + r + 8);
        if (r0 < p.H) p.x[r0] = b0 + v;
        if (r1 < p.H) p.x[r1] = b1 + v1;
      } else {
        const int r0 = gi * 16 + r, r1 = r0 + 8;
        if (r0 < p.V) p.logits[r0] = v;
        if (r1 < p.V) p.logits[r1] = v1;
      }
    });
    w.gb += cnt;
    w.rot = (w.rot + (uint32_t)nact) % (uint32_t)G;
  };

  const int Hp = p.qkv.tpg * 256, Qp = p.o.tpg * 256, Ip = p.down.tpg * 256;
  for (int l = 0; l < p.L; ++l) {
    const int64_t no = (int64_t)l * p.norm_stride;
    // Pull into L2 what the consumers will read with plain loads later in this layer: the CTA's KV rows (P2) and
    // the layer's norm weights (P1/P4 staging) — their DRAM latency then hides behind P1 instead of sitting on the
    // critical path after a barrier (126 MB L2 holds ~1.2 layers of weight stream, the lines stay resident).
    if (as.active) {
      const bf16* kb = p.kv + (int64_t)slot * p.kv_slot_stride + (int64_t)l * p.kv_layer_stride + ((int64_t)kvh * p.max_len + as.j0) * 128;
      const int lines = (as.j1 - as.j0) * 2;   // 128-byte lines per K (and per V) range
      for (int i = tid; i < lines; i += CONSUMER_THREADS) {
        asm volatile("prefetch.global.L2 [%0];\n" ::"l"(kb + (int64_t)i * 64));
        asm volatile("prefetch.global.L2 [%0];\n" ::"l"(kb + p.kv_v_offset + (int64_t)i * 64));
      }
    }
    for (int i = tid * 64; i < p.H; i += CONSUMER_THREADS * 64)
      asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p.norm2_0 + no + i));
    // ---------------- P1: RMSNorm + qkv + RoPE + KV write
    stamp(0);
    stage_xb(l == 0 ? nullptr : p.x, l == 0 ? p.embed + (int64_t)tok * p.H : nullptr, p.H, Hp, p.norm1_0 + no, p.eps, xb, red);
    stamp(1);
    run_phase(p.qkv, PH_QKV, l);
    stamp(2);
    bar_target += G;
    grid_barrier(p.bar_count, bar_target, p.dbg_flags);
    stamp(3); ++dbg_i;

    // ---------------- P2: attention over this CTA's key range of its head
    stamp(0); stamp(1);
    if (as.active) {
      const int hw = lane >> 4, l16 = lane & 15;
      const float sl2 = 0.08838834764831845f * 1.4426950408889634f;  // 128^-1/2 * log2(e)
      float q[8];
      {
        const float* qp = p.q + as.head * 128 + l16 * 8;
        const float4 a = __ldcg(reinterpret_cast<const float4*>(qp)), b = __ldcg(reinterpret_cast<const float4*>(qp + 4));
        q[0] = a.x * sl2; q[1] = a.y * sl2; q[2] = a.z * sl2; q[3] = a.w * sl2;
        q[4] = b.x * sl2; q[5] = b.y * sl2; q[6] = b.z * sl2; q[7] = b.w * sl2;
      }
      // the key/value of the token being decoded (written in P1 of this launch): fetch early
      uint4 knew = make_uint4(0, 0, 0, 0), vnew = make_uint4(0, 0, 0, 0);
      if (as.last && warp == 0) {
        const bf16* kb = p.kv + (int64_t)slot * p.kv_slot_stride + (int64_t)l * p.kv_layer_stride + ((int64_t)kvh * p.max_len + pos) * 128;
        knew = __ldcg(reinterpret_cast<const uint4*>(kb + l16 * 8));
        vnew = __ldcg(reinterpret_cast<const uint4*>(kb + p.kv_v_offset + l16 * 8));
      }
      float m = -INFINITY, lsum = 0.f, o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
      auto key_update = [&](const uint4& kraw, const uint4& vraw, bool valid) {
        float kf[8];
        unpack8(kraw, kf);
        float s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s2 += q[i] * kf[i];
        s2 += __shfl_xor_sync(0xffffffffu, s2, 8);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 4);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
        if (valid) {
          const float mn = fmaxf(m, s2), alpha = exp2f(m - mn), pj = exp2f(s2 - mn);
          float vf[8];
          unpack8(vraw, vf);
          lsum = lsum * alpha + pj;
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = o[i] * alpha + pj * vf[i];
          m = mn;
        }
      };
      {
        // keys j0 + hid, j0 + hid + 16, ... for half-warp hid (16 half-warps per CTA): the rows were pulled into L2
        // at the start of the layer; up to 8 keys (16 loads) are in flight per lane, i.e. one L2 round trip per
        // 128 keys of the range
        const bf16* kb = p.kv + (int64_t)slot * p.kv_slot_stride + (int64_t)l * p.kv_layer_stride + (int64_t)kvh * p.max_len * 128;
        const bf16* vb = kb + p.kv_v_offset;
        const int hid = warp * 2 + hw;
        for (int jb = as.j0; jb < as.j1; jb += 128) {   // warp-uniform trip count
          uint4 kr[8], vr[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int j = jb + u * 16 + hid;
            kr[u] = vr[u] = make_uint4(0, 0, 0, 0);
            if (j < as.j1) {
              kr[u] = __ldcg(reinterpret_cast<const uint4*>(kb + (int64_t)j * 128 + l16 * 8));
              vr[u] = __ldcg(reinterpret_cast<const uint4*>(vb + (int64_t)j * 128 + l16 * 8));
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) key_update(kr[u], vr[u], jb + u * 16 + hid < as.j1);
        }
      }
      if (as.last && warp == 0) key_update(knew, vnew, hw == 0);
      // merge the 16 half-warp states -> one partial per CTA
      float* sm_m = actf;            // [16]
      float* sm_l = actf + 16;       // [16]
      float* sm_o = actf + 32;       // [16][128]
      const int hidx = warp * 2 + hw;
      if (l16 == 0) { sm_m[hidx] = m; sm_l[hidx] = lsum; }
#pragma unroll
      for (int i = 0; i < 8; ++i) sm_o[hidx * 128 + l16 * 8 + i] = o[i];
      consumer_sync();
      if (tid < 128) {
        float M = -INFINITY;
#pragma unroll
        for (int h = 0; h < 16; ++h) M = fmaxf(M, sm_m[h]);
        float Lt = 0.f, O = 0.f;
#pragma unroll
        for (int h = 0; h < 16; ++h) {
          const float wgt = (sm_m[h] == -INFINITY) ? 0.f : exp2f(sm_m[h] - M);
          Lt += sm_l[h] * wgt;
          O += sm_o[h * 128 + tid] * wgt;
        }
        float* pp = p.part + (int64_t)c * 132;
        pp[tid] = O;
        if (tid == 0) { pp[128] = M; pp[129] = Lt; }
      }
      // the LAST CTA of this head to get here merges the head's partials into the normalised output
      consumer_sync();
      if (tid == 0) {
        unsigned prev;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;\n" : "=r"(prev) : "l"(p.head_cnt + as.head) : "memory");
        red[8] = (prev == (unsigned)as.cph - 1u) ? 1.f : 0.f;
      }
      consumer_sync();
      if (red[8] != 0.f) {
        if (tid < 128) {
          float ms[16], lv[16], ov[16];
#pragma unroll
          for (int r = 0; r < 16; ++r) {   // all loads are independent: one L2 round trip
            const float* pp = p.part + (int64_t)(r * p.heads + as.head) * 132;
            const bool ok = r < as.cph;
            ms[r] = ok ? ldcg_f(pp + 128) : -INFINITY;
            lv[r] = ok ? ldcg_f(pp + 129) : 0.f;
            ov[r] = ok ? ldcg_f(pp + tid) : 0.f;
          }
          float M = -INFINITY;
#pragma unroll
          for (int r = 0; r < 16; ++r) M = fmaxf(M, ms[r]);
          float Lt = 0.f, O = 0.f;
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const float wgt = (ms[r] == -INFINITY) ? 0.f : exp2f(ms[r] - M);
            Lt += lv[r] * wgt;
            O += ov[r] * wgt;
          }
          p.attn[as.head * 128 + tid] = O / Lt;
        }
        if (tid == 0) p.head_cnt[as.head] = 0u;  // self-resetting (next use is a grid barrier away)
      }
    }
    stamp(2);
    bar_target += G;
    grid_barrier(p.bar_count, bar_target, p.dbg_flags);
    stamp(3); ++dbg_i;

    // ---------------- P3: o-proj + residual on the merged attention output
    stamp(0);
    stage_xb(p.attn, nullptr, qd, Qp, nullptr, 0.f, xb, red);
    stamp(1);
    run_phase(p.o, PH_O, l);
    stamp(2);
    bar_target += G;
    grid_barrier(p.bar_count, bar_target, p.dbg_flags);
    stamp(3); ++dbg_i;

    // ---------------- P4: RMSNorm + gate/up + SiLU*mul
    stamp(0);
    stage_xb(p.x, nullptr, p.H, Hp, p.norm2_0 + no, p.eps, xb, red);
    stamp(1);
    run_phase(p.gu, PH_GU, l);
    stamp(2);
    bar_target += G;
    grid_barrier(p.bar_count, bar_target, p.dbg_flags);
    stamp(3); ++dbg_i;

    // ---------------- P5: down + residual
    stamp(0);
    stage_xb(p.h, nullptr, p.I, Ip, nullptr, 0.f, xb, red);
    stamp(1);
    for (int i = tid * 64; i < p.H; i += CONSUMER_THREADS * 64)   // next stage's norm weights -> L2 (the stream evicted them since)
      asm volatile("prefetch.global.L2 [%0];\n" ::"l"((l + 1 < p.L ? p.norm1_0 + no + p.norm_stride : p.final_norm) + i));
    run_phase(p.down, PH_DOWN, l);
    stamp(2);
    bar_target += G;
    grid_barrier(p.bar_count, bar_target, p.dbg_flags);
    stamp(3); ++dbg_i;
  }
  // ---------------- final RMSNorm + lm_head
  stamp(0);
  stage_xb(p.x, nullptr, p.H, Hp, p.final_norm, p.eps, xb, red);
  stamp(1);
  run_phase(p.lm, PH_LM, 0);
  stamp(2); stamp(3);
  // publish the barrier epoch for the next launch (stream-ordered): every CTA executed 5L barriers
  if (c == 0 && tid == 0 && !(p.dbg_flags & 2)) *p.bar_base = bar_target;
}

// ------------------------------------------------------------------ one-time weight re-tiling
// dst chunk q (16 B) = tile (group, ks) -> [kstep s][matrix m][row r]: rows-half = m & 1, k-half = m >> 1
__global__ void __launch_bounds__(256) retile_kernel(const bf16* __restrict__ src, int N, int K, int mode, int groups,
                                                     int tpg, bf16* __restrict__ dst) {
  const int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t total = (int64_t)groups * tpg * 512;
  if (q >= total) return;
  const int r = (int)(q & 7), m = (int)((q >> 3) & 3), s = (int)((q >> 5) & 15);
  const int64_t tile = q >> 9;
  const int ks = (int)(tile % tpg), gi = (int)(tile / tpg);
  const int ar = (m & 1) * 8 + r;                      // A-operand row 0..15
  const int col = ks * 256 + s * 16 + (m >> 1) * 8;
  int row;
  if (mode == TILE_SEQ) row = gi * 16 + ar;
  else if (mode == TILE_ROPE) row = (gi >> 3) * 128 + ((gi & 7) << 3) + (ar & 7) + (ar >> 3) * 64;
  else row = (ar < 8) ? 2 * (gi * 8 + ar) : 2 * (gi * 8 + ar - 8) + 1;  // source rows are interleaved (gate, up)
  uint4 v = make_uint4(0, 0, 0, 0);
  if (row < N && col < K) v = *reinterpret_cast<const uint4*>(src + (int64_t)row * K + col);
  *reinterpret_cast<uint4*>(dst + q * 8) = v;
}

}  // namespace

int64_t mega_tiled_elems(int N, int K, int mode, int* groups, int* tpg) {
  int g = (mode == TILE_GLU) ? (N / 2 + 7) / 8 : (N + 15) / 16;
  int t = (K + 255) / 256;
  if (groups) *groups = g;
  if (tpg) *tpg = t;
  return (int64_t)g * t * MEGA_TILE_ELEMS;
}

cudaError_t launch_retile(const bf16* src, int N, int K, int mode, bf16* dst, cudaStream_t s) {
  if ((K & 7) || (mode == TILE_ROPE && (N & 127))) return cudaErrorInvalidValue;
  int groups, tpg;
  mega_tiled_elems(N, K, mode, &groups, &tpg);
  const int64_t chunks = (int64_t)groups * tpg * 512;
  retile_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, s>>>(src, N, K, mode, groups, tpg, dst);
  return cudaGetLastError();
}

int mega_smem_bytes(const MegaArgs& a) {
  return a.nslots * TILE_BYTES + a.act_floats * 4 + 2 * a.nslots * 8 + (16 + 128 + NT * 16) * 4 + NG * 4 + NG * 16 * 4;
}

cudaError_t mega_configure(MegaArgs& a, int H, int I, int heads, int max_smem_optin, int num_sms, int* grid_out) {
  auto pad = [](int k) { return (k + 255) / 256 * 256; };
  int actf = pad(H) > pad(I) ? pad(H) : pad(I);
  if (pad(heads * 128) > actf) actf = pad(heads * 128);
  if (actf < 32 + 16 * 128) actf = 32 + 16 * 128;  // attention merge scratch
  actf = (actf + 31) & ~31;
  a.act_floats = actf;
  if ((I + 255) / 256 > NT - 44) return cudaErrorInvalidValue;  // partial-sum window must cover a group + tiles in flight
  const int fixed = actf * 4 + (16 + 128 + NT * 16) * 4 + NG * 4 + NG * 16 * 4 + 64;
  int nslots = (max_smem_optin - fixed) / (TILE_BYTES + 16);
  if (nslots > 32) nslots = 32;
  // every ring slot must always be filled by the same producer warp and drained by the same consumer warp
  // (slot s <-> producer s % NPW, consumer s % NCW): mbarrier parity waits are only alias-free when the
  // successive uses of one barrier are ordered inside one thread.
  nslots &= ~(NCW - 1);
  if (nslots < NCW) return cudaErrorInvalidValue;
  a.nslots = nslots;
  if (heads > num_sms || H > 8192) return cudaErrorInvalidValue;  // normed vector is register-staged (K <= 8192)
  *grid_out = num_sms;
  return cudaSuccess;
}

cudaError_t launch_decode_mega(const MegaArgs& a, int grid, cudaStream_t s, uint64_t* counter) {
  const int smem = mega_smem_bytes(a);
  cudaError_t e = cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  void* args[] = {(void*)&a};
  e = cudaLaunchCooperativeKernel((const void*)decode_mega_kernel, dim3(grid), dim3(MEGA_THREADS), args, (size_t)smem, s);
  if (counter) ++*counter;
  return e;
}

}  // namespace dtk
