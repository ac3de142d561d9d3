// Persistent weight-streaming decode kernel: ONE cooperative launch per generated token.
//
// Why: batch-1 decode streams every decoder weight once per token (2.56 GB for ds-1.3b) through ~120
// dependent GEMV-sized steps of a few microseconds each. Launched as separate kernels (even from a
// CUDA graph) the HBM pipe drains at every step boundary and the chain is launch/ramp bound
// (measured 0.37 of the HBM roofline). Here the weight stream is decoupled from the dependency chain:
//
//   * grid = one CTA per SM, resident for the whole token (cooperative launch);
//   * warp 8 of every CTA is a PRODUCER: it walks the CTA's statically known list of weight rows for
//     ALL layers and phases and streams them with 1-D TMA bulk copies (cp.async.bulk, mbarrier
//     complete_tx) into a ~176 KB shared-memory ring, never waiting for activations — weights do not
//     depend on them — so HBM stays busy across phase boundaries;
//   * warps 0-7 are CONSUMERS: they take ring slots in order, do the fp32-accumulated dot products
//     against the activation vector held in shared memory and run the fused epilogues
//     (RMSNorm prologue, RoPE + KV-cache write, SiLU*mul, residual add);
//   * phases are separated by a hand-rolled grid barrier (monotonic atomic counter); the ring depth
//     (~4 us of streaming per SM) covers the barrier + activation re-staging bubble;
//   * work items are dealt round-robin over CTAs with a running offset across phases, so the
//     cumulative bytes per CTA never differ by more than one item.
//
// Per layer: P1 qkv(+RMSNorm, RoPE, KV write) | P2 split-KV attention (old keys streamed through the
// same ring; the new key read after the barrier) | P3 o-proj + residual (prologue merges the attention
// partials) | P4 gate/up + SiLU*mul (+RMSNorm) | P5 down + residual; finally lm_head (+final RMSNorm).
//
// Replaces the per-token HF eager path (modeling_llama.py:303-333, ~900 launches per token).
#include <cooperative_groups.h>

#include "common.cuh"
#include "launch.h"

namespace dtk {
namespace {

constexpr int NCW = 8;                       // consumer warps
constexpr int NPW = 4;                       // producer warps (one issuing lane each): per-item issue cost ~0.3 us
constexpr int MEGA_THREADS = (NCW + NPW) * 32;
constexpr int CONSUMER_THREADS = NCW * 32;
static_assert(NCW % NPW == 0, "slot ownership: NPW must divide NCW");
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
DTK_DEV void grid_barrier(unsigned long long* counter, unsigned long long target, int skip = 0) {
  consumer_sync();
  if (skip) return;
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;\n" ::"l"(counter), "l"(1ull) : "memory");
    uint32_t spins = 0;
    long long t0 = 0;
    unsigned long long v;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(counter) : "memory");
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
enum { PH_QKV = 0, PH_ATTN = 1, PH_O = 2, PH_GU = 3, PH_DOWN = 4, PH_LM = 5 };

struct Phase {
  const bf16* W;   // weight matrix [N, K]
  int N, K;
  int mode;        // 0: contiguous row pairs (2i, 2i+1); 1: rope pairs (i, i+64) inside 128-row groups; 2: single rows
  int n_items;
};

DTK_DEV Phase make_phase(const MegaArgs& p, int layer, int ph) {
  Phase d;
  const int64_t lo = (int64_t)layer * p.layer_stride;
  const int qd = p.heads * 128, kd = p.kv_heads * 128;
  switch (ph) {
    case PH_QKV: d.W = p.wqkv0 + lo; d.N = qd + 2 * kd; d.K = p.H; d.mode = 1; d.n_items = d.N / 2; break;
    case PH_O: d.W = p.wo0 + lo; d.N = p.H; d.K = qd; d.mode = 0; d.n_items = d.N / 2; break;
    case PH_GU: d.W = p.wgu0 + lo; d.N = 2 * p.I; d.K = p.H; d.mode = 0; d.n_items = d.N / 2; break;
    case PH_DOWN: d.W = p.wd0 + lo; d.N = p.H; d.K = p.I; d.mode = 2; d.n_items = d.N; break;
    default: d.W = p.lm_head; d.N = p.V; d.K = p.H; d.mode = 0; d.n_items = d.N / 2; break;
  }
  return d;
}

// rows of item `it`: (r0, r1); r1 < 0 for single-row items
DTK_DEV void item_rows(const Phase& d, int it, int& r0, int& r1) {
  if (d.mode == 0) { r0 = 2 * it; r1 = r0 + 1; }
  else if (d.mode == 1) { r0 = (it >> 6) * 128 + (it & 63); r1 = r0 + 64; }
  else { r0 = it; r1 = -1; }
}

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

// ------------------------------------------------------------------ consumer helpers
// activation vector in shared memory as two float4 planes (conflict-free LDS.128): lo[c] = x[8c..8c+3], hi[c] = x[8c+4..8c+7]
struct ActView {
  float4* lo;
  float4* hi;
};

DTK_DEV void dot_rows(const uint8_t* row0, const uint8_t* row1, int KC, const ActView& x, int lane, float& a0, float& a1) {
  a0 = 0.f; a1 = 0.f;
  const uint4* w0 = reinterpret_cast<const uint4*>(row0);
  const uint4* w1 = reinterpret_cast<const uint4*>(row1);
#pragma unroll 4
  for (int c = lane; c < KC; c += 32) {
    const uint4 v0 = w0[c];
    const float4 xl = x.lo[c], xh = x.hi[c];
    float f[8];
    unpack8(v0, f);
    a0 += f[0] * xl.x + f[1] * xl.y + f[2] * xl.z + f[3] * xl.w + f[4] * xh.x + f[5] * xh.y + f[6] * xh.z + f[7] * xh.w;
    if (row1) {
      const uint4 v1 = w1[c];
      unpack8(v1, f);
      a1 += f[0] * xl.x + f[1] * xl.y + f[2] * xl.z + f[3] * xl.w + f[4] * xh.x + f[5] * xh.y + f[6] * xh.z + f[7] * xh.w;
    }
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
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

// stage a K-vector (optionally RMS-normalised) into the planes. src_f32 (ld.cg) or src_bf16 (embedding row).
// With a norm the vector (K <= 8192) is held in registers: the x and norm-weight loads are issued together
// (one L2/HBM round trip) and the scaled values are written to shared memory once.
DTK_DEV void stage_vector(const float* src_f32, const bf16* src_bf16, int K, const bf16* norm_w, float eps,
                          const ActView& x, float* red) {
  const int KC = K >> 3, tid = threadIdx.x;
  if (!norm_w) {
    for (int c = tid; c < KC; c += CONSUMER_THREADS) {
      x.lo[c] = __ldcg(reinterpret_cast<const float4*>(src_f32 + c * 8));
      x.hi[c] = __ldcg(reinterpret_cast<const float4*>(src_f32 + c * 8 + 4));
    }
    consumer_sync();
    return;
  }
  float4 a[4], b[4];
  uint4 nw[4];
  float ss = 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int c = tid + u * CONSUMER_THREADS;
    if (c < KC) {
      nw[u] = *reinterpret_cast<const uint4*>(norm_w + c * 8);
      if (src_bf16) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(src_bf16 + c * 8), f);
        a[u] = make_float4(f[0], f[1], f[2], f[3]);
        b[u] = make_float4(f[4], f[5], f[6], f[7]);
      } else {
        a[u] = __ldcg(reinterpret_cast<const float4*>(src_f32 + c * 8));
        b[u] = __ldcg(reinterpret_cast<const float4*>(src_f32 + c * 8 + 4));
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int c = tid + u * CONSUMER_THREADS;
    if (c < KC)
      ss += a[u].x * a[u].x + a[u].y * a[u].y + a[u].z * a[u].z + a[u].w * a[u].w + b[u].x * b[u].x + b[u].y * b[u].y +
            b[u].z * b[u].z + b[u].w * b[u].w;
  }
  const float r = rsqrtf(consumer_sum(ss, red) / K + eps);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int c = tid + u * CONSUMER_THREADS;
    if (c < KC) {
      float w[8];
      unpack8(nw[u], w);
      x.lo[c] = make_float4(a[u].x * r * w[0], a[u].y * r * w[1], a[u].z * r * w[2], a[u].w * r * w[3]);
      x.hi[c] = make_float4(b[u].x * r * w[4], b[u].y * r * w[5], b[u].z * r * w[6], b[u].w * r * w[7]);
    }
  }
  consumer_sync();
}

__global__ void __launch_bounds__(MEGA_THREADS, 1) decode_mega_kernel(const MegaArgs p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x, G = gridDim.x;
  const int nslots = p.nslots, slot_bytes = p.slot_bytes;
  uint8_t* ring = smem;
  float* actf = reinterpret_cast<float*>(smem + (size_t)nslots * slot_bytes);
  const int act_floats = p.act_floats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(actf + act_floats);
  float* red = reinterpret_cast<float*>(bars + 2 * nslots);  // 16 floats
  float* rope_s = red + 16;                                   // [64][2] cos/sin of this position
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
  __syncthreads();

  const int pos = p.pos[0], slot = p.slots[0];
  int tok = p.tok[0];
  if (tok < 0 || tok >= p.V) tok = 0;
  const int qd = p.heads * 128, kd = p.kv_heads * 128;
  if (tid < 128) rope_s[tid] = p.rope_cs[(int64_t)pos * 128 + tid];
  __syncthreads();
  const AttnSplit as = attn_split(p, c, G, pos);
  const int kvh = as.head / (p.heads / p.kv_heads);

  // ---- item ownership. The CTA's local item sequence (all phases, in order) is dealt to agents by index:
  // local item n -> ring slot n % nslots, producer warp n % NPW, consumer warp n % NCW (nslots is a multiple of
  // both, so a slot always has the same producer and the same consumer). An agent visits ONLY its own items.
  struct Walk {
    uint32_t nb = 0;     // local items before the current phase
    uint32_t gmod = 0;   // (global item counter) mod G -> round-robin offset of the current phase
  };
  // items of this CTA in a weight phase: it = first + k * G, k in [0, cnt)
  auto phase_span = [&](const Walk& w, int n_items, int& first, int& cnt) {
    first = (int)(((uint32_t)c + (uint32_t)G - w.gmod) % (uint32_t)G);
    cnt = first < n_items ? (n_items - 1 - first) / G + 1 : 0;
  };

  if (warp >= NCW) {
    // =============================================================== PRODUCERS
    const uint32_t pw = (uint32_t)(warp - NCW);
    if (lane == 0) {
      Walk w;
      // visit own items k = k0, k0 + NPW, ... of a phase with cnt local items
      auto for_own = [&](int cnt, auto&& issue) {
        uint32_t k = (pw + NPW - (w.nb & (NPW - 1))) & (NPW - 1);
        if ((int)k < cnt) {
          const uint32_t n0 = w.nb + k;
          uint32_t sl = n0 % (uint32_t)nslots, use = n0 / (uint32_t)nslots;
          for (; (int)k < cnt; k += NPW) {
            if (use > 0) mbar_wait(empty0 + 8 * sl, (use - 1) & 1);
            issue((int)k, ring_u32 + sl * slot_bytes, full0 + 8 * sl);
            sl += NPW;
            if (sl >= (uint32_t)nslots) { sl -= nslots; ++use; }
          }
        }
        w.nb += cnt;
      };
      auto stream_phase = [&](const Phase& d) {
        const uint32_t rb = (uint32_t)d.K * 2;
        int first, cnt;
        phase_span(w, d.n_items, first, cnt);
        for_own(cnt, [&](int k, uint32_t dst, uint32_t fb) {
          int r0, r1;
          item_rows(d, first + k * G, r0, r1);
          const bf16* src = d.W + (int64_t)r0 * d.K;
          if (d.mode == 0) {           // rows 2i, 2i+1 are contiguous in memory: one copy
            mbar_expect_tx(fb, 2 * rb);
            bulk_g2s(dst, src, 2 * rb, fb);
          } else if (d.mode == 1) {
            mbar_expect_tx(fb, 2 * rb);
            bulk_g2s(dst, src, rb, fb);
            bulk_g2s(dst + rb, src + (int64_t)64 * d.K, rb, fb);
          } else {
            mbar_expect_tx(fb, rb);
            bulk_g2s(dst, src, rb, fb);
          }
        });
        w.gmod = (w.gmod + (uint32_t)d.n_items) % (uint32_t)G;
      };
      for (int l = 0; l < p.L; ++l) {
        stream_phase(make_phase(p, l, PH_QKV));
        {  // old keys/values of this CTA's (head, range): 16-key items, K rows then V rows
          const bf16* kb = p.kv + (int64_t)slot * p.kv_slot_stride + (int64_t)l * p.kv_layer_stride + (int64_t)kvh * p.max_len * 128;
          const bf16* vb = kb + p.kv_v_offset;
          for_own(as.n_items, [&](int i, uint32_t dst, uint32_t fb) {
            const int k0 = as.j0 + i * 16, nk = min(16, as.j1 - k0);
            mbar_expect_tx(fb, (uint32_t)nk * 512);
            bulk_g2s(dst, kb + (int64_t)k0 * 128, (uint32_t)nk * 256, fb);
            bulk_g2s(dst + 16 * 256, vb + (int64_t)k0 * 128, (uint32_t)nk * 256, fb);
          });
        }
        stream_phase(make_phase(p, l, PH_O));
        stream_phase(make_phase(p, l, PH_GU));
        stream_phase(make_phase(p, l, PH_DOWN));
      }
      stream_phase(make_phase(p, 0, PH_LM));
    }
    return;
  }

  // ================================================================= CONSUMERS
  ActView X;
  X.lo = reinterpret_cast<float4*>(actf);
  unsigned long long bar_target = *p.bar_base;  // barriers completed before this launch (x G)
  Walk w;
  // visit own items of a phase with cnt local items; body(k, smem pointer) runs after the bytes landed and
  // must finish reading the slot before returning (the slot is released right after)
  auto for_own = [&](int cnt, auto&& pre, auto&& body) {
    uint32_t k = ((uint32_t)warp + NCW - (w.nb & (NCW - 1))) & (NCW - 1);
    if ((int)k < cnt) {
      const uint32_t n0 = w.nb + k;
      uint32_t sl = n0 % (uint32_t)nslots, use = n0 / (uint32_t)nslots;
      for (; (int)k < cnt; k += NCW) {
        pre((int)k);
        mbar_wait(full0 + 8 * sl, use & 1);
        body((int)k, ring + (size_t)sl * slot_bytes);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * sl);
        sl += NCW;
        if (sl >= (uint32_t)nslots) { sl -= nslots; ++use; }
      }
    }
    w.nb += cnt;
  };

  // optional phase timestamps (CTA-local clock64): [phase][4] = {start, staged, items done, barrier done}
  long long* dbg = (p.dbg && (c == 0 || c == G / 2 || c == G - 1)) ? p.dbg + (int64_t)(c == 0 ? 0 : (c == G - 1 ? 2 : 1)) * (p.L * 5 + 1) * 4 : nullptr;
  int dbg_i = 0;
  auto stamp = [&](int k) { if (dbg && tid == 0) dbg[dbg_i * 4 + k] = clock64(); };

  auto run_phase = [&](const Phase& d, int ph, int layer) {
    const int KC = d.K >> 3;
    X.hi = X.lo + KC;
    int first, cnt;
    phase_span(w, d.n_items, first, cnt);
    int r0 = 0, r1 = 0;
    float b0 = 0.f, b1 = 0.f;
    for_own(cnt,
      [&](int k) {  // before waiting on the weights: rows + residuals (their L2 latency hides behind the wait)
        item_rows(d, first + k * G, r0, r1);
        if (lane == 0) {
          if (ph == PH_O) {
            if (layer == 0) {  // residual stream starts as the token embedding
              b0 = __bfloat162float(p.embed[(int64_t)tok * p.H + r0]);
              b1 = __bfloat162float(p.embed[(int64_t)tok * p.H + r1]);
            } else { b0 = ldcg_f(p.x + r0); b1 = ldcg_f(p.x + r1); }
          } else if (ph == PH_DOWN) {
            b0 = ldcg_f(p.x + r0);
          }
        }
      },
      [&](int k, const uint8_t* base) {
        float a0 = 0.f, a1 = 0.f;
        if (!(p.dbg_flags & 1)) dot_rows(base, r1 >= 0 ? base + (size_t)d.K * 2 : nullptr, KC, X, lane, a0, a1);
        if (lane == 0) {
          if (ph == PH_QKV) {
            const int i = r0 & 127;
            if (r0 < qd + kd) {
              const float2 csn = *reinterpret_cast<const float2*>(rope_s + i * 2);
              const float y0 = a0 * csn.x - a1 * csn.y, y1 = a1 * csn.x + a0 * csn.y;
              if (r0 < qd) { p.q[r0] = y0; p.q[r1] = y1; }
              else {
                const int kh = (r0 - qd) >> 7;
                bf16* dd = p.kv + (int64_t)slot * p.kv_slot_stride + (int64_t)layer * p.kv_layer_stride + ((int64_t)kh * p.max_len + pos) * 128;
                dd[i] = __float2bfloat16_rn(y0);
                dd[i + 64] = __float2bfloat16_rn(y1);
              }
            } else {
              const int kh = (r0 - qd - kd) >> 7;
              bf16* dd = p.kv + (int64_t)slot * p.kv_slot_stride + (int64_t)layer * p.kv_layer_stride + p.kv_v_offset + ((int64_t)kh * p.max_len + pos) * 128;
              dd[i] = __float2bfloat16_rn(a0);
              dd[i + 64] = __float2bfloat16_rn(a1);
            }
          } else if (ph == PH_O) {
            p.x[r0] = b0 + a0;
            p.x[r1] = b1 + a1;
          } else if (ph == PH_GU) {
            p.h[first + k * G] = silu(a0) * a1;
          } else if (ph == PH_DOWN) {
            p.x[r0] = b0 + a0;
          } else {
            p.logits[r0] = a0;
            p.logits[r1] = a1;
          }
        }
      });
    w.gmod = (w.gmod + (uint32_t)d.n_items) % (uint32_t)G;
  };

  for (int l = 0; l < p.L; ++l) {
    const int64_t lo = (int64_t)l * p.layer_stride;
    // ---------------- P1: RMSNorm + qkv + RoPE + KV write
    stamp(0);
    X.hi = X.lo + (p.H >> 3);
    stage_vector(l == 0 ? nullptr : p.x, l == 0 ? p.embed + (int64_t)tok * p.H : nullptr, p.H, p.norm1_0 + lo, p.eps, X, red);
    stamp(1);
    run_phase(make_phase(p, l, PH_QKV), PH_QKV, l);
    stamp(2);
    bar_target += G;
    grid_barrier(p.bar_count, bar_target, p.dbg_flags & 2);
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
      for_own(as.n_items, [&](int) {}, [&](int i, const uint8_t* base) {
        const int k0 = as.j0 + i * 16, nk = min(16, as.j1 - k0);
#pragma unroll 4
        for (int kk = 0; kk < 8; ++kk) {
          const int key = kk * 2 + hw;
          const bool valid = key < nk;
          uint4 kraw = make_uint4(0, 0, 0, 0), vraw = make_uint4(0, 0, 0, 0);
          if (valid) {
            kraw = *reinterpret_cast<const uint4*>(base + key * 256 + l16 * 16);
            vraw = *reinterpret_cast<const uint4*>(base + 16 * 256 + key * 256 + l16 * 16);
          }
          key_update(kraw, vraw, valid);
        }
      });
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
          const float w = (sm_m[h] == -INFINITY) ? 0.f : exp2f(sm_m[h] - M);
          Lt += sm_l[h] * w;
          O += sm_o[h * 128 + tid] * w;
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
          float ms[16], M = -INFINITY;
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            ms[r] = (r < as.cph) ? ldcg_f(p.part + (int64_t)(r * p.heads + as.head) * 132 + 128) : -INFINITY;
            M = fmaxf(M, ms[r]);
          }
          float ov[16], lv[16];
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const float* pp = p.part + (int64_t)(r * p.heads + as.head) * 132;
            ov[r] = (r < as.cph) ? ldcg_f(pp + tid) : 0.f;
            lv[r] = (r < as.cph) ? ldcg_f(pp + 129) : 0.f;
          }
          float Lt = 0.f, O = 0.f;
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const float w = (ms[r] == -INFINITY) ? 0.f : exp2f(ms[r] - M);
            Lt += lv[r] * w;
            O += ov[r] * w;
          }
          p.attn[as.head * 128 + tid] = O / Lt;
        }
        if (tid == 0) p.head_cnt[as.head] = 0u;  // self-resetting (next use is a grid barrier away)
      }
    }
    stamp(2);
    bar_target += G;
    grid_barrier(p.bar_count, bar_target, p.dbg_flags & 2);
    stamp(3); ++dbg_i;

    // ---------------- P3: o-proj + residual on the merged attention output
    stamp(0);
    X.hi = X.lo + (qd >> 3);
    stage_vector(p.attn, nullptr, qd, nullptr, 0.f, X, red);
    stamp(1);
    run_phase(make_phase(p, l, PH_O), PH_O, l);
    stamp(2);
    bar_target += G;
    grid_barrier(p.bar_count, bar_target, p.dbg_flags & 2);
    stamp(3); ++dbg_i;

    // ---------------- P4: RMSNorm + gate/up + SiLU*mul
    stamp(0);
    X.hi = X.lo + (p.H >> 3);
    stage_vector(p.x, nullptr, p.H, p.norm2_0 + lo, p.eps, X, red);
    stamp(1);
    run_phase(make_phase(p, l, PH_GU), PH_GU, l);
    stamp(2);
    bar_target += G;
    grid_barrier(p.bar_count, bar_target, p.dbg_flags & 2);
    stamp(3); ++dbg_i;

    // ---------------- P5: down + residual
    stamp(0);
    X.hi = X.lo + (p.I >> 3);
    stage_vector(p.h, nullptr, p.I, nullptr, 0.f, X, red);
    stamp(1);
    run_phase(make_phase(p, l, PH_DOWN), PH_DOWN, l);
    stamp(2);
    bar_target += G;
    grid_barrier(p.bar_count, bar_target, p.dbg_flags & 2);
    stamp(3); ++dbg_i;
  }
  // ---------------- final RMSNorm + lm_head
  stamp(0);
  X.hi = X.lo + (p.H >> 3);
  stage_vector(p.x, nullptr, p.H, p.final_norm, p.eps, X, red);
  stamp(1);
  run_phase(make_phase(p, 0, PH_LM), PH_LM, 0);
  stamp(2); stamp(3);
  // publish the barrier epoch for the next launch (stream-ordered): every CTA executed 5L barriers
  if (c == 0 && tid == 0 && !(p.dbg_flags & 2)) *p.bar_base = bar_target;
}

}  // namespace

int mega_smem_bytes(const MegaArgs& a) { return a.nslots * a.slot_bytes + a.act_floats * 4 + 2 * a.nslots * 8 + (16 + 128 + a.heads * 16) * 4; }

cudaError_t mega_configure(MegaArgs& a, int H, int I, int heads, int max_smem_optin, int num_sms, int* grid_out) {
  // slot = the largest work item: a pair of K=H rows or one K=I row; 16-key attention item = 8 KB
  int slot = 2 * H * 2;
  if (I * 2 > slot) slot = I * 2;
  if (heads * 128 * 2 * 2 > slot) slot = heads * 128 * 2 * 2;  // o-proj pair (K = heads*128)
  if (slot < 16 * 512) slot = 16 * 512;
  slot = (slot + 127) & ~127;
  int actf = H > I ? H : I;
  if (heads * 128 > actf) actf = heads * 128;
  if (actf < 32 + 16 * 128) actf = 32 + 16 * 128;  // attention merge scratch
  actf = (actf + 31) & ~31;
  a.slot_bytes = slot;
  a.act_floats = actf;
  const int fixed = actf * 4 + (16 + 128 + heads * 16) * 4 + 64;
  int nslots = (max_smem_optin - fixed) / (slot + 16);
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
