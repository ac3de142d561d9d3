// Persistent weight-streaming decode kernel: ONE cooperative launch per generated token.
//
// Why: batch-1 decode streams every decoder weight once per token (2.56 GB for ds-1.3b) through ~120
// dependent GEMV-sized steps of a few microseconds each. Launched as separate kernels (even from a
// CUDA graph) the HBM pipe drains at every step boundary and the chain is launch/ramp bound
// (measured 0.37 of the HBM roofline). Here the weight stream is decoupled from the dependency chain:
//
//   * grid = one CTA per SM, resident for the whole token (cooperative launch);
//   * 4 PRODUCER warps per CTA walk the CTA's statically known list of weight tiles for ALL layers and
//     phases and stream them with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx) into a
//     ~190 KB shared-memory ring of 8 KB slots, never waiting for activations — weights do not depend on
//     them — so HBM stays busy across phase boundaries (measured: the ring sustains 7.2 TB/s);
//   * 8 CONSUMER warps take tiles in order. A tile is 16 output rows x 256 k, pre-arranged in HBM
//     (launch_retile, once at load) so that it lands in shared memory exactly in ldmatrix.x4 order; the
//     dot products run on the tensor pipe (mma.sync m16n8k16, fp32 accumulate) with the activation vector
//     split into bf16 hi + lo parts (x = hi + lo to 2^-17), i.e. fp32-grade GEMV at ~1/6 of the issue slots
//     of a CUDA-core unpack+FMA loop (which measured consumer-bound);
//   * rows are grouped so that one thread's two accumulator rows (g, g+8) are a RoPE pair (i, i+64) or a
//     SwiGLU pair (gate_i, up_i): RMSNorm prologue, RoPE + KV-cache write, SiLU*mul and residual add are
//     all fused; partial sums of a group's k-tiles are combined in a fixed order (deterministic);
//   * phases are separated by a hand-rolled grid barrier (release-reduction + acquire poll); the ring
//     depth (~4 us of streaming per SM) covers the barrier + activation re-staging bubble;
//   * 16-row groups are dealt round-robin over CTAs with a running offset across phases, so the cumulative
//     bytes per CTA never differ by more than one group.
//
// Per layer: P1 qkv(+RMSNorm, RoPE, KV write) | P2 split-KV attention (old keys streamed through the
// same ring; the new key read after the barrier; last CTA of a head merges) | P3 o-proj + residual |
// P4 gate/up + SiLU*mul (+RMSNorm) | P5 down + residual; finally lm_head (+final RMSNorm).
//
// Replaces the per-token HF eager path (modeling_llama.py:303-333, ~900 launches per token).
#include <cstdio>

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
constexpr int NG = 48;                       // per-group arrival counters / residual rows (>= groups in flight)
constexpr long long SPIN_CYCLES = 1000000000ll;  // bounded waits (~2 s): trap instead of hanging the GPU

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
      else if (now - t0 > SPIN_CYCLES) {
        if ((threadIdx.x & 31) == 0) printf("[dtk] mbarrier wait timed out: cta %d tid %d bar %u parity %u\n", (int)blockIdx.x, (int)threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
DTK_DEV void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
DTK_DEV void consumer_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(CONSUMER_THREADS) : "memory"); }

// ---- data-carried synchronisation. A cross-CTA value is an 8-byte {fp32 bits, tag} pair: aligned 8-byte
// accesses are single-copy atomic, so a reader sees the old or the new pair, never a mix. Writers use one
// plain store; readers poll the L2 copy (ld.cg) until the tag of the expected phase shows up. Write-after-read
// hazards are excluded by the data-flow itself (a buffer is only rewritten by work that transitively depends
// on every reader of the previous version; see DESIGN.md "decode synchronisation").
// Tagged values are exchanged with morally-strong, gpu-scope 8-byte accesses: weak ld/st (even .cg) carry no
// coherence guarantee between SMs — on the two-die B200 a weak polling load can keep hitting a stale
// die-local L2 copy forever (observed). A u64 access is single-copy atomic, so {value, tag} never tears.
DTK_DEV uint2 ld_poll2(const uint2* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return make_uint2((uint32_t)v, (uint32_t)(v >> 32));
}
// two consecutive pairs
DTK_DEV uint4 ld_poll4(const uint4* p) {
  const uint2 a = ld_poll2(reinterpret_cast<const uint2*>(p)), b = ld_poll2(reinterpret_cast<const uint2*>(p) + 1);
  return make_uint4(a.x, a.y, b.x, b.y);
}
DTK_DEV void st_tag(uint2* p, float v, uint32_t tag) {
  const unsigned long long w = (unsigned long long)__float_as_uint(v) | ((unsigned long long)tag << 32);
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(w) : "memory");
}
DTK_DEV void backoff(int cycles) {
  const long long t = clock64();
  while (clock64() - t < cycles) {}
}
struct Spin {
  uint32_t n = 0;
  long long t0 = 0;
  DTK_DEV void tick(const void* what = nullptr, uint32_t want = 0, uint32_t seen = 0) {
    if ((++n & 255u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > SPIN_CYCLES) {
        if ((threadIdx.x & 31) == 0)
          printf("[dtk] tag wait timed out: cta %d tid %d addr %p want %u seen %u\n", (int)blockIdx.x, (int)threadIdx.x, what, want, seen);
        __trap();
      }
    }
  }
};
DTK_DEV float ld_tag(const uint2* p, uint32_t tag, int nowait) {
  Spin sp;
  uint2 u = ld_poll2(p);
  while (u.y != tag && !(nowait & 2)) { sp.tick(p, tag, u.y); if (!(nowait & 4)) backoff(200); u = ld_poll2(p); }
  return __uint_as_float(u.x);
}
// 8 consecutive tagged elements (64 B): four 16-byte loads in flight per attempt (one round trip once the data
// is there), short back-off between failed attempts to keep polling traffic off the L2.
DTK_DEV void ld_tag8(const uint2* p, uint32_t tag, int nowait, float (&out)[8]) {
  Spin sp;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = ld_poll4(q), b = ld_poll4(q + 1), c = ld_poll4(q + 2), d = ld_poll4(q + 3);
  while (!(nowait & 2) && (a.y != tag || a.w != tag || b.y != tag || b.w != tag || c.y != tag || c.w != tag || d.y != tag || d.w != tag)) {
    sp.tick(p, tag, a.y);
    if (!(nowait & 4)) backoff(200);
    a = ld_poll4(q); b = ld_poll4(q + 1); c = ld_poll4(q + 2); d = ld_poll4(q + 3);
  }
  out[0] = __uint_as_float(a.x); out[1] = __uint_as_float(a.z); out[2] = __uint_as_float(b.x); out[3] = __uint_as_float(b.z);
  out[4] = __uint_as_float(c.x); out[5] = __uint_as_float(c.z); out[6] = __uint_as_float(d.x); out[7] = __uint_as_float(d.z);
}
// 16 consecutive tagged elements (one k-step, 128 B): eight loads in flight per attempt
DTK_DEV void ld_tag16(const uint2* p, uint32_t tag, int nowait, float (&out)[16]) {
  Spin sp;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 u[8];
  bool ok;
  do {
    ok = true;
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] = ld_poll4(q + i);
#pragma unroll
    for (int i = 0; i < 8; ++i) ok = ok && (u[i].y == tag) && (u[i].w == tag);
    if (nowait & 2) break;
    if (!ok) { sp.tick(p, tag, u[0].y); if (!(nowait & 4)) backoff(200); }
  } while (!ok);
#pragma unroll
  for (int i = 0; i < 8; ++i) { out[2 * i] = __uint_as_float(u[i].x); out[2 * i + 1] = __uint_as_float(u[i].z); }
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
DTK_DEV void stage_xb(const uint2* src_t, uint32_t tag, int nowait, const bf16* src_bf16, int K, int Kp, const bf16* norm_w,
                      float eps, uint4* xb, float* red) {
  const int tid = threadIdx.x, nsteps = Kp >> 4;
  consumer_sync();         // every warp of the CTA is done reading the previous phase's xb
  constexpr int MAXS = 2;  // k-steps held in registers per thread when normalising (K <= 8192)
  float v[MAXS][16];
  float ss = 0.f;
  if (norm_w) {
    // norm weights do not depend on other CTAs: fetch them first so their latency overlaps the tag polling
    uint4 nw[MAXS][2];
#pragma unroll
    for (int u = 0; u < MAXS; ++u) {
      const int S = tid + u * CONSUMER_THREADS;
      nw[u][0] = nw[u][1] = make_uint4(0, 0, 0, 0);
      if (S < nsteps && S * 16 < K) {
        nw[u][0] = *reinterpret_cast<const uint4*>(norm_w + S * 16);
        nw[u][1] = *reinterpret_cast<const uint4*>(norm_w + S * 16 + 8);
      }
    }
#pragma unroll
    for (int u = 0; u < MAXS; ++u) {
      const int S = tid + u * CONSUMER_THREADS;
      if (S < nsteps) {
        float x16[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x16[i] = 0.f;
        if (S * 16 < K) {   // K is a multiple of 16 for every normed vector (hidden size)
          if (src_bf16) {
            float f[8];
            unpack8(*reinterpret_cast<const uint4*>(src_bf16 + S * 16), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) x16[i] = f[i];
            unpack8(*reinterpret_cast<const uint4*>(src_bf16 + S * 16 + 8), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) x16[8 + i] = f[i];
          } else {
            ld_tag16(src_t + S * 16, tag, nowait, x16);
          }
        }
        float wv[16];
        {
          float f[8];
          unpack8(nw[u][0], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) wv[i] = f[i];
          unpack8(nw[u][1], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) wv[8 + i] = f[i];
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          ss += x16[i] * x16[i];
          v[u][i] = x16[i] * wv[i];   // weight now, 1/rms after the reduction
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
    for (int S0 = tid; S0 < nsteps; S0 += 2 * CONSUMER_THREADS) {
      // two k-steps per thread with all 16 loads in flight (one round trip)
      float w16[2][16];
      const int S1 = S0 + CONSUMER_THREADS;
      const bool full0 = S0 * 16 + 16 <= K, full1 = S1 < nsteps && S1 * 16 + 16 <= K;
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
        for (int i = 0; i < 16; ++i) w16[h2][i] = 0.f;
      if (full0 && full1 && !(nowait & 16)) {
        Spin sp;
        const uint4* q0 = reinterpret_cast<const uint4*>(src_t + S0 * 16);
        const uint4* q1 = reinterpret_cast<const uint4*>(src_t + S1 * 16);
        uint4 u0[8], u1[8];
        bool ok;
        do {
          ok = true;
#pragma unroll
          for (int i = 0; i < 8; ++i) { u0[i] = ld_poll4(q0 + i); u1[i] = ld_poll4(q1 + i); }
#pragma unroll
          for (int i = 0; i < 8; ++i) ok = ok && u0[i].y == tag && u0[i].w == tag && u1[i].y == tag && u1[i].w == tag;
          if (nowait & 2) break;
          if (!ok) { sp.tick(); if (!(nowait & 4)) backoff(200); }
        } while (!ok);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          w16[0][2 * i] = __uint_as_float(u0[i].x); w16[0][2 * i + 1] = __uint_as_float(u0[i].z);
          w16[1][2 * i] = __uint_as_float(u1[i].x); w16[1][2 * i + 1] = __uint_as_float(u1[i].z);
        }
      } else {
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int S = h2 ? S1 : S0;
          if (S < nsteps) {
            if (S * 16 + 16 <= K) ld_tag16(src_t + S * 16, tag, nowait, w16[h2]);
            else if (S * 16 < K) {   // K is a multiple of 8: ragged last k-step
              float f[8];
              ld_tag8(src_t + S * 16, tag, nowait, f);
#pragma unroll
              for (int i = 0; i < 8; ++i) w16[h2][i] = f[i];
            }
          }
        }
      }
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int S = h2 ? S1 : S0;
        if (S < nsteps) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float a = w16[h2][2 * j], b = w16[h2][2 * j + 1];
            const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
            hi[j] = pack_bf16x2(__bfloat162float(ah), __bfloat162float(bh));
            lo[j] = pack_bf16x2(a - __bfloat162float(ah), b - __bfloat162float(bh));
          }
#pragma unroll
          for (int t = 0; t < 4; ++t) xb[S * 4 + t] = make_uint4(hi[t], hi[t + 4], lo[t], lo[t + 4]);
        }
      }
    }
  }
  consumer_sync();
}

// epoch tags of one launch: tag(l, k) = base + 8 l + k + 1
enum { TG_QKV = 0, TG_PART = 1, TG_ATTN = 2, TG_XO = 3, TG_H = 4, TG_XD = 5 };

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
  if (tid == 0 && (p.dbg_flags & 64) && (c == 30 || c == 0 || c == 52)) printf("[dtk] cta %d G %d pos %d: active %d head %d j0 %d j1 %d last %d cph %d\n", c, G, pos, as.active, as.head, as.j0, as.j1, as.last, as.cph);
  const int kvh = as.head / (p.heads / p.kv_heads);
  const int nowait = (p.dbg_flags & 2) | ((p.dbg_flags & 4) ? 0 : 4) | (p.dbg_flags & 16);   // dev flags: 2 = never wait for tags, 4 = poll WITH back-off

  // ---- work assignment. A weight phase with `groups` 16-row groups is cut into equal blocks of
  // per = ceil(groups / G) groups; only ceil(groups / per) CTAs take part (all with the same amount of work, so
  // they finish together), the others idle for that phase while their producers prefetch ahead. The set of
  // participating CTAs rotates from phase to phase. The CTA's local TILE sequence (all phases, in order) is
  // dealt to agents by index: local tile n -> ring slot n % nslots, producer warp n % NPW, consumer warp n % NCW
  // (nslots is a multiple of both: a slot always has the same producer and consumer, so mbarrier parity waits
  // never alias).
  struct Walk {
    uint32_t nb = 0;     // local tiles before the current phase
    uint32_t gb = 0;     // local groups before the current phase
    uint32_t rot = 0;    // rotation of the participating CTA set
  };
  auto phase_span = [&](const Walk& w, int groups, int& g0, int& cnt, int& nact) {
    const int per = (groups + G - 1) / G;
    nact = (groups + per - 1) / per;
    const int ci = (int)(((uint32_t)c + (uint32_t)G - w.rot) % (uint32_t)G);
    g0 = ci * per;
    cnt = (ci < nact) ? min(per, groups - g0) : 0;
  };

  if (warp >= NCW) {
    // =============================================================== PRODUCERS
    const uint32_t pw = (uint32_t)(warp - NCW);
    if (lane == 0) {
      Walk w;
      // visit own tiles j = j0, j0 + NPW, ... of a phase with `ntiles` local tiles (tpg tiles per group)
      // Outstanding-copy throttle: any L2 round trip of this SM (tag polls, residual loads) queues behind the
      // SM's own outstanding bulk-copy responses, so the number of issued-but-not-landed copies is bounded
      // (depth per producer warp) while the ring may still hold many landed tiles.
      constexpr int MAXD = 8;
      uint32_t of_bar[MAXD], of_par[MAXD];
      int of_n = 0, of_head = 0;
      const int depth = p.prod_depth < 1 ? 1 : (p.prod_depth > MAXD ? MAXD : p.prod_depth);
      auto for_own = [&](int ntiles, int tpg, auto&& issue) {
        uint32_t j = (pw + NPW - (w.nb & (NPW - 1))) & (NPW - 1);
        if ((int)j < ntiles) {
          const uint32_t n0 = w.nb + j;
          uint32_t sl = n0 % (uint32_t)nslots, use = n0 / (uint32_t)nslots;
          uint32_t k = j / (uint32_t)tpg, ks = j - k * (uint32_t)tpg;
          for (; (int)j < ntiles; j += NPW) {
            if (of_n == depth) {   // wait until the oldest outstanding copy has landed
              mbar_wait(of_bar[of_head], of_par[of_head]);
              of_head = (of_head + 1 == depth) ? 0 : of_head + 1;
              --of_n;
            }
            if (use > 0) mbar_wait(empty0 + 8 * sl, (use - 1) & 1);
            issue((int)k, (int)ks, ring_u32 + sl * TILE_BYTES, full0 + 8 * sl);
            {
              int tail = of_head + of_n;
              if (tail >= depth) tail -= depth;
              of_bar[tail] = full0 + 8 * sl;
              of_par[tail] = use & 1;
              ++of_n;
            }
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
  const uint32_t tag0 = (uint32_t)(*p.epoch);   // tags of this launch: tag0 + 8 l + k + 1
  auto TAG = [&](int l, int k) -> uint32_t { return tag0 + (uint32_t)(8 * l + k + 1); };
  Walk w;
  uint32_t cur_slot = 0;
  auto release = [&]() {
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * cur_slot);
  };
  // visit own tiles of a phase; body(j, k, ks, slot) runs after the bytes landed and must finish READING the slot
  // before calling release()
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

  // optional phase timestamps (globaltimer ns, comparable across SMs): [CTA][phase][4] = {start, staged, items done, -}
  long long* dbg = p.dbg ? p.dbg + (int64_t)c * (p.L * 5 + 1) * 4 : nullptr;
  int dbg_i = 0;
  auto stamp = [&](int k) {
    if (dbg && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
      dbg[dbg_i * 4 + k] = (long long)t;
    }
  };

  // one weight phase. stage(): fills xb (only called when this CTA has work in the phase). Every tile = 16
  // k-steps of (ldmatrix.x4, LDS.128, 2 x mma); the warp that finishes a group's last tile sums the group's
  // partials in k order and runs the epilogue for its 16 rows.
  auto run_phase = [&](const MegaMat& m, int ph, int layer, auto&& stage) {
    int g0, cnt, nact;
    phase_span(w, m.groups, g0, cnt, nact);
    const uint32_t nb0 = w.nb, gb0 = w.gb;
    const int tpg = m.tpg;
    stamp(0);
    if (cnt > 0) stage();
    stamp(1);
    for_own(cnt * tpg, tpg, [&](int j, int k, int ks, uint32_t sl) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      if (!(p.dbg_flags & 1)) {
        const uint32_t ta = ring_u32 + sl * TILE_BYTES + lane * 16;
        const uint4* xp = xb + (size_t)ks * 64 + (lane & 3);
#pragma unroll
        for (int s = 0; s < 16; ++s) {
          uint32_t a[4];
          ldmatrix_x4(a[0], a[1], a[2], a[3], ta + s * 512);
          const uint4 b = xp[s * 4];
          mma_bf16_16816(acc, a, b.x, b.y);
          mma_bf16_16816(acc, a, b.z, b.w);
        }
      }
      release();
      const uint32_t gslot = (gb0 + (uint32_t)k) % NG;
      if (ks == 0 && (ph == PH_O || ph == PH_DOWN) && lane < 16 && !(p.dbg_flags & 8)) {
        // residual of row (group, lane): its producer finished a whole phase ago, so this never waits long;
        // fetched here (first tile of the group) its L2 latency is off the group's critical path
        const int row = (g0 + k) * 16 + lane;
        float b = 0.f;
        if (row < p.H) {
          if (ph == PH_O) b = (layer == 0) ? __bfloat162float(p.embed[(int64_t)tok * p.H + row]) : ld_tag(p.xt + row, TAG(layer - 1, TG_XD), nowait);
          else b = ld_tag(p.xt + row, TAG(layer, TG_XO), nowait);
        }
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
        const uint32_t tg = TAG(layer, TG_QKV);
        if (row0 < qd + kd) {
          const float2 csn = *reinterpret_cast<const float2*>(rope_s + i * 2);
          const float y0 = v * csn.x - v1 * csn.y, y1 = v1 * csn.x + v * csn.y;
          if (row0 < qd) { st_tag(p.qt + row0, y0, tg); st_tag(p.qt + row0 + 64, y1, tg); }
          else {
            const int kh = (row0 - qd) >> 7;
            bf16* dd = p.kv + (int64_t)slot * p.kv_slot_stride + (int64_t)layer * p.kv_layer_stride + ((int64_t)kh * p.max_len + pos) * 128;
            const __nv_bfloat16 h0 = __float2bfloat16_rn(y0), h1 = __float2bfloat16_rn(y1);
            dd[i] = h0;
            dd[i + 64] = h1;
            st_tag(p.kvt + kh * 128 + i, __bfloat162float(h0), tg);       // this token's key for the attention CTAs
            st_tag(p.kvt + kh * 128 + i + 64, __bfloat162float(h1), tg);
          }
        } else {
          const int kh = (row0 - qd - kd) >> 7;
          bf16* dd = p.kv + (int64_t)slot * p.kv_slot_stride + (int64_t)layer * p.kv_layer_stride + p.kv_v_offset + ((int64_t)kh * p.max_len + pos) * 128;
          const __nv_bfloat16 h0 = __float2bfloat16_rn(v), h1 = __float2bfloat16_rn(v1);
          dd[i] = h0;
          dd[i + 64] = h1;
          st_tag(p.kvt + kd + kh * 128 + i, __bfloat162float(h0), tg);
          st_tag(p.kvt + kd + kh * 128 + i + 64, __bfloat162float(h1), tg);
        }
      } else if (ph == PH_O) {
        const int r0 = gi * 16 + r, r1 = r0 + 8;
        float b0 = *reinterpret_cast<volatile float*>(rbuf + gslot * 16 + r), b1 = *reinterpret_cast<volatile float*>(rbuf + gslot * 16 + r + 8);
        if (p.dbg_flags & 8) {
          b0 = b1 = 0.f;
          if (layer == 0) {
            if (r0 < p.H) b0 = __bfloat162float(p.embed[(int64_t)tok * p.H + r0]);
            if (r1 < p.H) b1 = __bfloat162float(p.embed[(int64_t)tok * p.H + r1]);
          } else {
            if (r0 < p.H) b0 = ld_tag(p.xt + r0, TAG(layer - 1, TG_XD), nowait);
            if (r1 < p.H) b1 = ld_tag(p.xt + r1, TAG(layer - 1, TG_XD), nowait);
          }
        }
        if (r0 < p.H) st_tag(p.xt + r0, b0 + v, TAG(layer, TG_XO));
        if (r1 < p.H) st_tag(p.xt + r1, b1 + v1, TAG(layer, TG_XO));
      } else if (ph == PH_GU) {
        const int i = gi * 8 + r;
        if (i < p.I) st_tag(p.ht + i, silu(v) * v1, TAG(layer, TG_H));
      } else if (ph == PH_DOWN) {
        const int r0 = gi * 16 + r, r1 = r0 + 8;
        float b0 = *reinterpret_cast<volatile float*>(rbuf + gslot * 16 + r), b1 = *reinterpret_cast<volatile float*>(rbuf + gslot * 16 + r + 8);
        if (p.dbg_flags & 8) {
          b0 = b1 = 0.f;
          if (r0 < p.H) b0 = ld_tag(p.xt + r0, TAG(layer, TG_XO), nowait);
          if (r1 < p.H) b1 = ld_tag(p.xt + r1, TAG(layer, TG_XO), nowait);
        }
        if (r0 < p.H) st_tag(p.xt + r0, b0 + v, TAG(layer, TG_XD));
        if (r1 < p.H) st_tag(p.xt + r1, b1 + v1, TAG(layer, TG_XD));
      } else {
        const int r0 = gi * 16 + r, r1 = r0 + 8;
        if (r0 < p.V) p.logits[r0] = v;
        if (r1 < p.V) p.logits[r1] = v1;
      }
    });
    stamp(2); stamp(3); ++dbg_i;
    w.gb += cnt;
    w.rot = (w.rot + (uint32_t)nact) % (uint32_t)G;
  };

  const int Hp = p.qkv.tpg * 256, Qp = p.o.tpg * 256, Ip = p.down.tpg * 256;
  for (int l = 0; l < p.L; ++l) {
    const int64_t no = (int64_t)l * p.norm_stride;
    // KV rows of this CTA's (head, key range) are read with plain loads in P2: pull them into L2 now so that
    // their DRAM latency hides behind P1 (126 MB L2 holds ~1.2 layers of weight stream, they stay resident)
    if (as.active) {
      const bf16* kb = p.kv + (int64_t)slot * p.kv_slot_stride + (int64_t)l * p.kv_layer_stride + ((int64_t)kvh * p.max_len + as.j0) * 128;
      const int lines = (as.j1 - as.j0) * 2;   // 128-byte lines per K (and per V) range
      for (int i = tid; i < lines; i += CONSUMER_THREADS) {
        asm volatile("prefetch.global.L2 [%0];\n" ::"l"(kb + (int64_t)i * 64));
        asm volatile("prefetch.global.L2 [%0];\n" ::"l"(kb + p.kv_v_offset + (int64_t)i * 64));
      }
    }
    // ---------------- P1: RMSNorm + qkv + RoPE + KV write
    run_phase(p.qkv, PH_QKV, l, [&]() {
      stage_xb(p.xt, l == 0 ? 0u : TAG(l - 1, TG_XD), nowait, l == 0 ? p.embed + (int64_t)tok * p.H : nullptr, p.H, Hp,
               p.norm1_0 + no, p.eps, xb, red);
    });

    // ---------------- P2: attention over this CTA's key range of its head
    stamp(0);
    if (as.active) {
      const int hw = lane >> 4, l16 = lane & 15;
      const float sl2 = 0.08838834764831845f * 1.4426950408889634f;  // 128^-1/2 * log2(e)
      float q[8];
      ld_tag8(p.qt + as.head * 128 + l16 * 8, TAG(l, TG_QKV), nowait, q);
      stamp(1);
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] *= sl2;
      float m = -INFINITY, lsum = 0.f, o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
      auto key_update = [&](const float (&kf)[8], const float (&vf)[8], bool valid) {
        float s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s2 += q[i] * kf[i];
        s2 += __shfl_xor_sync(0xffffffffu, s2, 8);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 4);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
        if (valid) {
          const float mn = fmaxf(m, s2), alpha = exp2f(m - mn), pj = exp2f(s2 - mn);
          lsum = lsum * alpha + pj;
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = o[i] * alpha + pj * vf[i];
          m = mn;
        }
      };
      {
        // keys j0 + hid, j0 + hid + 16, ... for half-warp hid (16 half-warps per CTA); up to 8 keys (16 loads) in
        // flight per lane so that a 128-key range costs ONE L2 round trip; the newest key (this token's, from P1)
        // is handled last: its tag has long been set by then
        const bf16* kb = p.kv + (int64_t)slot * p.kv_slot_stride + (int64_t)l * p.kv_layer_stride + (int64_t)kvh * p.max_len * 128;
        const bf16* vb = kb + p.kv_v_offset;
        const int hid = warp * 2 + hw;
        const int nb8 = (p.dbg_flags & 32) ? 4 : 8;   // keys per batch (dev flag 32: 4)
        for (int jb = as.j0; jb < as.j1; jb += 16 * nb8) {   // warp-uniform trip count
          uint4 kr[8], vr[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int j = jb + u * 16 + hid;
            kr[u] = vr[u] = make_uint4(0, 0, 0, 0);
            if (u < nb8 && j < as.j1) {
              kr[u] = __ldcg(reinterpret_cast<const uint4*>(kb + (int64_t)j * 128 + l16 * 8));
              vr[u] = __ldcg(reinterpret_cast<const uint4*>(vb + (int64_t)j * 128 + l16 * 8));
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int j = jb + u * 16 + hid;
            float kf[8], vf[8];
            unpack8(kr[u], kf);
            unpack8(vr[u], vf);
            key_update(kf, vf, u < nb8 && j < as.j1);
          }
        }
        if (as.last && warp == 0) {   // the key/value of the token being decoded (produced in P1 of this launch)
          float nkf[8], nvf[8];
          ld_tag8(p.kvt + kvh * 128 + l16 * 8, TAG(l, TG_QKV), nowait, nkf);
          ld_tag8(p.kvt + kd + kvh * 128 + l16 * 8, TAG(l, TG_QKV), nowait, nvf);
          key_update(nkf, nvf, hw == 0);
        }
      }
      // merge the 16 half-warp states -> one partial per CTA
      float* sm_m = actf;            // [16]
      float* sm_l = actf + 16;       // [16]
      float* sm_o = actf + 32;       // [16][128]
      const int hidx = warp * 2 + hw;
      stamp(3);
      consumer_sync();               // the scratch aliases xb: all warps are done with the previous phase's tiles
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
        if (!as.last) {
          // publish this CTA's partial (m, l, o[128]) for the head's merger
          uint2* pp = p.partt + (int64_t)c * 132;
          const uint32_t tg = TAG(l, TG_PART);
          st_tag(pp + tid, O, tg);
          if (tid == 0) { st_tag(pp + 128, M, tg); st_tag(pp + 129, Lt, tg); }
        } else {
          // the CTA that also owns the newest key merges the head: poll the other ranges' partials
          const uint32_t tg = TAG(l, TG_PART);
          float ms[16], lv[16], ov[16];
#pragma unroll
          for (int r = 0; r < 16; ++r) { ms[r] = -INFINITY; lv[r] = 0.f; ov[r] = 0.f; }
          Spin sp;
          bool ok = false;
          while (!ok) {
            ok = true;
#pragma unroll
            for (int r = 0; r < 15; ++r) {
              if (r < as.cph - 1) {
                const uint2* pp = p.partt + (int64_t)(r * p.heads + as.head) * 132;
                const uint2 a = ld_poll2(pp + 128), b = ld_poll2(pp + 129), d = ld_poll2(pp + tid);
                ok = ok && a.y == tg && b.y == tg && d.y == tg;
                ms[r] = __uint_as_float(a.x); lv[r] = __uint_as_float(b.x); ov[r] = __uint_as_float(d.x);
              }
            }
            if (nowait & 2) break;
            if (!ok) { sp.tick(p.partt, tg, 12345u); if (!(nowait & 4)) backoff(200); }
          }
          float MM = M;
#pragma unroll
          for (int r = 0; r < 15; ++r) if (r < as.cph - 1) MM = fmaxf(MM, ms[r]);
          const float wown = (M == -INFINITY) ? 0.f : exp2f(M - MM);
          float LL = Lt * wown, OO = O * wown;
#pragma unroll
          for (int r = 0; r < 15; ++r) {
            if (r < as.cph - 1) {
              const float wgt = (ms[r] == -INFINITY) ? 0.f : exp2f(ms[r] - MM);
              LL += lv[r] * wgt;
              OO += ov[r] * wgt;
            }
          }
          st_tag(p.attnt + as.head * 128 + tid, OO / LL, TAG(l, TG_ATTN));
          if (tid == 0 && (p.dbg_flags & 64)) printf("[dtk] merger cta %d layer %d head %d wrote attn tag %u val %f (M %f L %f)\n", c, l, as.head, TAG(l, TG_ATTN), OO / LL, MM, LL);
        }
      }
      consumer_sync();   // sm_* scratch (aliases xb) is free again
    }
    stamp(2); ++dbg_i;

    // ---------------- P3: o-proj + residual on the merged attention output
    run_phase(p.o, PH_O, l, [&]() { stage_xb(p.attnt, TAG(l, TG_ATTN), nowait, nullptr, qd, Qp, nullptr, 0.f, xb, red); });
    // ---------------- P4: RMSNorm + gate/up + SiLU*mul
    run_phase(p.gu, PH_GU, l, [&]() { stage_xb(p.xt, TAG(l, TG_XO), nowait, nullptr, p.H, Hp, p.norm2_0 + no, p.eps, xb, red); });
    // ---------------- P5: down + residual
    run_phase(p.down, PH_DOWN, l, [&]() { stage_xb(p.ht, TAG(l, TG_H), nowait, nullptr, p.I, Ip, nullptr, 0.f, xb, red); });
  }
  // ---------------- final RMSNorm + lm_head
  run_phase(p.lm, PH_LM, 0, [&]() { stage_xb(p.xt, TAG(p.L - 1, TG_XD), nowait, nullptr, p.H, Hp, p.final_norm, p.eps, xb, red); });
  // next launch uses fresh tags (stream order makes this visible to it)
  if (c == 0 && tid == 0) *p.epoch = (unsigned long long)(tag0 + 8u * (uint32_t)p.L + 8u);
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
  if (heads > num_sms || H > 8192 || (H & 15) || (I & 7)) return cudaErrorInvalidValue;  // normed vector is register-staged
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
