// Persistent weight-streaming decode kernel: ONE cooperative launch per generated token.
//
// Why: batch-1 decode streams every decoder weight once per token (2.56 GB for ds-1.3b) through ~120
// dependent GEMV-sized steps of a few microseconds each. Launched as separate kernels (even from a
// CUDA graph) the HBM pipe drains at every step boundary and the chain is launch/ramp bound
// (measured 0.37 of the HBM roofline). Here the weight stream is decoupled from the dependency chain:
//
//   * grid = one CTA per SM, resident for the whole token (cooperative launch);
//   * 4 PRODUCER warps per CTA (one issuing lane each, registers handed back with setmaxnreg.dec) walk the
//     CTA's statically known list of weight tiles AND cached key/value items for ALL layers and phases and stream
//     them with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx) into a ~190 KB shared-memory ring of
//     8 KB slots, never waiting for activations — neither weights nor the cache depend on this token — so HBM
//     stays busy across phase boundaries (the ring alone sustains 7.2 TB/s, profiles/r2_stream_per_sm.txt);
//   * 8 CONSUMER warps (setmaxnreg.inc: the kernel calls no function, or ptxas would ignore setmaxnreg) take
//     tiles in order. A tile is 16 output rows x 256 k, pre-arranged in HBM (launch_retile, once at load) so that
//     it lands in shared memory exactly in ldmatrix.x4 order; the dot products run on the tensor pipe (mma.sync
//     m16n8k16, fp32 accumulate) with the activation vector split into bf16 hi + lo parts (x = hi + lo to 2^-17)
//     that occupy alternating columns of the B operand: one HMMA per k-step, fp32-grade GEMV; the ring slot is
//     handed back as soon as its shared-memory reads are issued;
//   * rows are grouped so that one thread's two accumulator rows (g, g+8) are a RoPE pair (i, i+64) or a
//     SwiGLU pair (gate_i, up_i): RMSNorm scale, RoPE + KV-cache write, SiLU*mul and residual add are all fused
//     into the group epilogue; partial sums of a group's k-tiles are combined in a fixed order (deterministic);
//     the epilogues of a phase run in PARALLEL behind a CTA barrier (warp w takes groups w, w + 8, ...), not on whichever
//     warp finishes a group last;
//   * values that cross CTAs travel as 8-byte TAGGED words {fp32, phase tag} (see below): no grid-wide barrier or
//     counter anywhere, readers poll the words they need;
//   * the input vector of a phase is staged per 256-element SLICE by the warp whose tile needs it (stage_slice):
//     one coalesced L2 round trip per warp, no CTA-wide two-pass staging; phases end with a CTA barrier only;
//   * a phase's 16-row groups are cut into equal blocks over as many CTAs as needed (the participating set
//     rotates from phase to phase), so participants finish together and idle CTAs' producers run ahead.
//
// Per layer: P1 qkv(+RMSNorm, RoPE, KV write) | P2 split-KV attention (the CTA's share of the cached keys/values
// arrives through the ring; a fixed owner CTA per head merges the partials) | P3 o-proj + residual | P4 gate/up +
// SiLU*mul (+RMSNorm) | P5 down + residual; finally lm_head (+final RMSNorm, + greedy argmax and token publication
// in the kernel tail). DESIGN.md section 4 lists the measured alternatives that were rejected.
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
constexpr int NT = 104;                      // per-tile partial-sum entries (>= max tiles/group + tiles in flight)
constexpr int NS = 72;                       // 256-column slices of a staged input vector (>= max tiles per group)
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
DTK_DEV uint32_t mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}
// slow path of a wait: bounded spin, trap instead of hanging the GPU
DTK_DEV void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try(bar, parity)) {
    if ((++spins & 1023u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > SPIN_CYCLES) __trap();
    }
  }
}
DTK_DEV void mbar_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try(bar, parity)) mbar_wait_slow(bar, parity);
}
DTK_DEV void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
DTK_DEV void consumer_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(CONSUMER_THREADS) : "memory"); }

// ------------------------------------------------------------------ tagged activation words
// Every activation value that crosses CTAs (residual stream, q, the new key/value row, attention partials and output,
// the SwiGLU vector) travels as ONE 8-byte word {fp32 value, 32-bit phase tag}: aligned 8-byte accesses are single-copy
// atomic, so a reader that sees the expected tag also sees the value written with it — no release fence on the
// producer side (it cost ~0.8 us of every phase) and no acquire on the consumer side. Writers use st.relaxed.gpu.
// Readers first try a WEAK coalesced load (ld.cg: may be served by the SM's own L2 partition) and only re-read with
// ld.relaxed.gpu the words whose tag is still old — a stale copy is harmless because it carries a stale tag.
// The grid-wide counter below is therefore only a HINT that says when reading is worthwhile; correctness rests on the
// tags. A buffer is overwritten one layer later, after a chain of data dependencies that runs through every CTA which
// read it (write-after-read safe).
typedef unsigned long long u64;
DTK_DEV void st_tag(u64* p, float v, uint32_t tag) {
  const u64 w = ((u64)tag << 32) | (u64)__float_as_uint(v);
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(w) : "memory");
}
DTK_DEV ulonglong2 ld_strong2(const u64* p) {
  ulonglong2 r;
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];\n" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
  return r;
}
DTK_DEV ulonglong2 ld_weak2(const u64* p) {
  ulonglong2 r;
  asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];\n" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
  return r;
}
DTK_DEV u64 ld_strong1(const u64* p) {
  u64 r;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(r) : "l"(p) : "memory");
  return r;
}
DTK_DEV u64 ld_weak1(const u64* p) {
  u64 r;
  asm volatile("ld.global.cg.u64 %0, [%1];\n" : "=l"(r) : "l"(p) : "memory");
  return r;
}
DTK_DEV bool tag_ok(u64 w, uint32_t tag) { return (uint32_t)(w >> 32) == tag; }
DTK_DEV float tag_val(u64 w) { return __uint_as_float((uint32_t)w); }
struct Spin {   // bounded polling with a short back-off: trap instead of hanging the GPU
  uint32_t n = 0;
  long long t0 = 0;
  DTK_DEV void tick() {
    __nanosleep(32);   // (0 .. 96 ns measured equal, 256 ns +1 %, 512 ns +3 %; pipelined re-polls slower: profiles/r2_decode_poll_sweep.txt)
    if ((++n & 255u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > SPIN_CYCLES) __trap();
    }
  }
};
// make a (weakly loaded) pair valid: re-read with coherent loads until both tags match (slow path out of line)
DTK_DEV u64 poll1(const u64* p, uint32_t tag) {
  Spin sp;
  for (;;) {
    const u64 w = ld_strong1(p);
    if (tag_ok(w, tag)) return w;
    sp.tick();
  }
}
DTK_DEV float settle1(u64 w, const u64* p, uint32_t tag, bool nowait) {
  if (!tag_ok(w, tag) && !nowait) w = poll1(p, tag);
  return tag_val(w);
}

// N tagged pairs per thread: weak loads first, then every pair whose tag is still old is re-read coherently, all of them per
// round (one L2 round trip per round however many are late)
template <int N>
DTK_DEV void ld_pairs(const u64* const (&ptr)[N], const bool (&on)[N], uint32_t tag, bool nowait, float2 (&out)[N]) {
  ulonglong2 w[N];
#pragma unroll
  for (int u = 0; u < N; ++u)
    if (on[u]) w[u] = ld_weak2(ptr[u]);
  Spin sp;
  for (;;) {
    bool bad[N], any = false;
#pragma unroll
    for (int u = 0; u < N; ++u) {
      bad[u] = on[u] && !(tag_ok(w[u].x, tag) && tag_ok(w[u].y, tag));
      any = any || bad[u];
    }
    if (!any || nowait) break;
    sp.tick();
#pragma unroll
    for (int u = 0; u < N; ++u)
      if (bad[u]) w[u] = ld_strong2(ptr[u]);
  }
#pragma unroll
  for (int u = 0; u < N; ++u) out[u] = on[u] ? make_float2(tag_val(w[u].x), tag_val(w[u].y)) : make_float2(0.f, 0.f);
}

// ------------------------------------------------------------------ work description
enum { PH_QKV = 0, PH_ATTN = 1, PH_O = 2, PH_GU = 3, PH_DOWN = 4, PH_LM = 5 };

// attention split: CTA c handles head c % heads, key range index c / heads (cph ranges per head)
struct AttnSplit {
  int active, head, j0, j1, last;  // keys [j0, j1) among the OLD keys [0, pos); `last` also takes key `pos`
  int cph;                         // CTAs per head
  int n_items;                     // 16-key ring items (K rows + V rows of 16 consecutive positions)
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

// The B operand of mma.m16n8k16 for a GEMV: the input vector lives in shared memory as
// entry [kstep S][t] (uint4) = { hi(x[16S+2t], x[16S+2t+1]), hi(x[16S+2t+8], +9), lo(..2t..), lo(..2t+8..) }
// where hi = bf16(x), lo = bf16(x - hi). All 8 columns of B are the same vector, so every lane of a quad column reads
// entry t = lane & 3. Entries for k >= K (padding up to the 256-column tile) are zero.
//
// DATAFLOW STAGING (round 2): a tile (group, ks) needs only the 256-element SLICE ks of the vector, so the warp that owns
// the tile stages that slice itself, when it gets there — ONE warp: 4 tagged pairs per lane in one coalesced L2 round trip,
// re-read coherently until their tags match, converted in registers (the two halves of a B entry meet through one shuffle)
// and written straight to the entries. No CTA-wide barrier in front of a phase: a warp starts multiplying as soon as ITS
// slice has arrived, and when the slowest producer CTA of the previous phase finally publishes its rows, one warp per CTA
// has a few tiles left instead of every warp a whole phase. With norm_w the staged slice is x * w (RMSNorm gain) WITHOUT the
// 1/rms factor (the GEMV is linear: the epilogue scales by r); the slice's sum of squares goes to slice_ss[ks], and
// r = rsqrt(sum over slices / K + eps) is formed by the epilogue warp (same order in every CTA: identical r everywhere).
// The vector buffer is single: a slice may only be overwritten when every warp of the CTA is past the previous weight
// phase (phase_done counts warps x phases); the L2 round trip comes first, so that wait is normally free.
// Inlined (one call site): ptxas ignores setmaxnreg in a kernel that calls functions (C7507); fully inlined measured 0.8 % faster.
DTK_DEV void stage_slice(const u64* src, uint32_t in_tag, bool nowait, const bf16* src_bf16, int K, int ks,
                                         const bf16* norm_w, uint4* xb, float* slice_ss, volatile uint32_t* slice_tag,
                                         uint32_t my_tag) {
  const int lane = threadIdx.x & 31;
  const int npair = K >> 1;
  float2 v[4];
  uint32_t gw[4];
  bool on[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int pi = ks * 128 + u * 32 + lane;
    on[u] = pi < npair;
    gw[u] = (norm_w && on[u]) ? *reinterpret_cast<const uint32_t*>(norm_w + 2 * pi) : 0u;
  }
  if (src) {
    ulonglong2 w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (on[u]) w[u] = ld_weak2(src + 2 * (ks * 128 + u * 32 + lane));
    Spin sp;
    for (;;) {
      bool bad[4], any = false;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        bad[u] = on[u] && !(tag_ok(w[u].x, in_tag) && tag_ok(w[u].y, in_tag));
        any = any || bad[u];
      }
      if (!any || nowait) break;
      sp.tick();
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (bad[u]) w[u] = ld_strong2(src + 2 * (ks * 128 + u * 32 + lane));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = on[u] ? make_float2(tag_val(w[u].x), tag_val(w[u].y)) : make_float2(0.f, 0.f);
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u] = make_float2(0.f, 0.f);
      if (on[u]) {
        const uint32_t e = *reinterpret_cast<const uint32_t*>(src_bf16 + 2 * (ks * 128 + u * 32 + lane));
        v[u] = make_float2(__uint_as_float(e << 16), __uint_as_float(e & 0xffff0000u));
      }
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    float a = v[u].x, b = v[u].y;
    if (norm_w) {
      ss += a * a + b * b;
      a *= __uint_as_float(gw[u] << 16);
      b *= __uint_as_float(gw[u] & 0xffff0000u);
    }
    const float ah = __bfloat162float(__float2bfloat16_rn(a)), bh = __bfloat162float(__float2bfloat16_rn(b));
    const uint32_t hi = pack_bf16x2(ah, bh), lo = pack_bf16x2(a - ah, b - bh);
    const uint32_t hi2 = __shfl_xor_sync(0xffffffffu, hi, 4), lo2 = __shfl_xor_sync(0xffffffffu, lo, 4);
    if (!(lane & 4)) xb[(size_t)(ks * 16 + u * 4 + (lane >> 3)) * 4 + (lane & 3)] = make_uint4(hi, hi2, lo, lo2);
  }
  if (norm_w) {
    ss = warp_sum(ss);
    if (lane == 0) slice_ss[ks] = ss;
  }
  __syncwarp();
  if (lane == 0) {
    __threadfence_block();
    slice_tag[ks] = my_tag;
  }
  __syncwarp();
}
// r = rsqrt(mean(x^2) + eps) from the slices' sums of squares (fixed order: identical in every warp and CTA)
DTK_DEV float slices_rn(const float* slice_ss, int nslice, int K, float eps) {
  const int lane = threadIdx.x & 31;
  float t = 0.f;
  for (int i = lane; i < nslice; i += 32) t += *reinterpret_cast<const volatile float*>(slice_ss + i);
  t = warp_sum(t);
  return rsqrtf(t / K + eps);
}

// DBG = true: dev instrumentation (phase stamps, per-tile trace, timing-experiment flags) compiled in.
// Both roles run ONE rolled loop over the 5 L + 1 phases of the token (qkv | attention | o | gate/up | down per layer,
// then lm_head) with a single copy of the tile code and a run-time phase switch in the epilogue: the per-token
// instruction footprint of a warp stays inside the SM's 32 KB instruction cache (the fully specialised version was
// ~160 KB, re-fetched from L2 every layer: the first tiles of every phase ran 3-6x slower than the steady state).
template <bool DBG>
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
  int* gcnt = reinterpret_cast<int*>(tpart + NT * 16);        // [NG] (unused since the epilogues run behind a barrier; keeps the layout)
  float* rbuf = reinterpret_cast<float*>(gcnt + NG);          // [NG][16] residuals prefetched at a group's first tile
  float* qkn = rbuf + NG * 16;                                // [3][128] q | new key | new value of the CTA's head
  float* slice_ss = qkn + 384;                                // [NS] sum of squares of each staged slice (normed phases)
  uint32_t* slice_tag = reinterpret_cast<uint32_t*>(slice_ss + NS);   // [NS] phase tag of the data a slice holds
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
  if (tid < NS) slice_tag[tid] = (uint32_t)p.bar_base[1];   // the epoch: never a phase tag of this launch
  if (tid < 128) rope_s[tid] = p.rope_cs[(int64_t)pos * 128 + tid];
  __syncthreads();
  const AttnSplit as = attn_split(p, c, G, pos);
  const int kvh = as.head / (p.heads / p.kv_heads);
  const bf16* kv_slot = p.kv + (int64_t)slot * p.kv_slot_stride;
  // shared KV prefix: positions [0, shlen) (a multiple of 16, so a 16-position ring item never straddles) live in another slot
  const int shlen = p.share_len[0];
  const bf16* kv_share = p.kv + (int64_t)p.share_slot[0] * p.kv_slot_stride;
  const int nphase = 5 * p.L + 1;

  // ---- item ownership. The CTA's local TILE sequence (all phases, in order) is dealt to agents by index:
  // local tile n -> ring slot n % nslots, producer warp n % NPW, consumer warp n % NCW (nslots is a multiple of
  // both, so a slot always has the same producer and the same consumer -> mbarrier parity waits never alias).
  struct Walk {
    uint32_t nb = 0;     // local tiles before the current phase
    uint32_t gb = 0;     // local groups before the current phase
    uint32_t rot = 0;    // rotation of the participating CTA set
  };
  // A weight phase with `groups` 16-row groups is cut into equal blocks of per = ceil(groups / G) groups; only
  // ceil(groups / per) CTAs take part (same amount of work each, so they finish together), the others idle for that
  // phase. The participating set rotates from phase to phase.
  // (per / nact come from the host and every modulo below is a conditional subtraction: no division on the per-phase path)
  auto phase_span = [&](const Walk& w, const MegaMat& m, int& g0, int& cnt, int& nact) {
    nact = m.nact;
    int ci = c + G - (int)w.rot;   // rot < G
    if (ci >= G) ci -= G;
    g0 = ci * m.per;
    cnt = (ci < nact) ? min(m.per, m.groups - g0) : 0;
  };
  // phase it -> (layer, kind, matrix index into p.mat)
  auto mat_of = [](int ph) { return ph == PH_QKV ? 0 : ph == PH_O ? 1 : ph == PH_GU ? 2 : ph == PH_DOWN ? 3 : 4; };

  if (warp >= NCW) {
    // =============================================================== PRODUCERS
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;\n");
    const uint32_t pw = (uint32_t)(warp - NCW);
    if (lane == 0) {
      Walk w;
      uint32_t sl = pw, use = 0;   // ring slot / use count of this lane's next tile: its tiles are n = pw, pw + NPW, ... across ALL phases
      long long* tr = nullptr;   // dev trace: row (local tile index within the traced layer), column 0 = issue clock
      uint32_t tr_nb0 = 0;
      int l = 0, ph = 0;
      for (int it = 0; it < nphase; ++it) {
        if (it == nphase - 1) { ph = PH_LM; l = 0; }
        if (DBG && ph == 0) {
          tr = (p.dbg2 && l == p.dbg_layer && it != nphase - 1) ? p.dbg2 + (int64_t)c * MEGA_DBG2_ROWS * 4 : nullptr;
          if (tr) tr_nb0 = w.nb;
        }
        // this phase's local tile list: weight tiles of the CTA's row groups, or (attention) the CTA's share of the
        // cached keys / values of the layer: item i = positions [j0 + 16 i, +16) of the CTA's kv head, K rows at slot
        // offset 0 and V rows at 4096 (rows of one head are contiguous in the cache). Neither depends on this token, so
        // both stream ahead of the dependency chain and the phases read shared memory only.
        const bool attn = ph == PH_ATTN;
        const MegaMat& m = p.mat[mat_of(ph)];
        int g0 = 0, cnt = 0, nact = 0;
        if (!attn) phase_span(w, m, g0, cnt, nact);
        const int tpg = attn ? 1 : m.tpg;
        const int ntiles = attn ? as.n_items : cnt * tpg;
        const int64_t kvo = (int64_t)l * p.kv_layer_stride + (int64_t)kvh * p.max_len * 128;
        const bf16* base = attn ? kv_slot + kvo
                                : m.base + (int64_t)l * m.layer_stride + (int64_t)g0 * tpg * MEGA_TILE_ELEMS;
        // own tiles j = j0, j0 + NPW, ...
        uint32_t j = (pw + NPW - (w.nb & (NPW - 1))) & (NPW - 1);
        if ((int)j < ntiles) {
          for (; (int)j < ntiles; j += NPW) {
            if (use > 0) mbar_wait(empty0 + 8 * sl, (use - 1) & 1);
            const uint32_t dst = ring_u32 + sl * TILE_BYTES, fb = full0 + 8 * sl;
            if (attn) {
              const int key0 = as.j0 + (int)j * 16;
              const uint32_t bytes = (uint32_t)min(16, p.max_len - key0) * 256u;
              const bf16* kb = (key0 < shlen ? kv_share + kvo : base) + (int64_t)key0 * 128;
              mbar_expect_tx(fb, 2 * bytes);
              bulk_g2s(dst, kb, bytes, fb);
              bulk_g2s(dst + 4096, kb + p.kv_v_offset, bytes, fb);
            } else {
              mbar_expect_tx(fb, TILE_BYTES);
              bulk_g2s(dst, base + (int64_t)j * MEGA_TILE_ELEMS, TILE_BYTES, fb);
            }
            if (DBG && tr) {
              const uint32_t row = w.nb + j - tr_nb0;
              if (row < 160u) tr[row * 4] = clock64();
            }
            sl += NPW;
            if (sl >= (uint32_t)nslots) { sl -= nslots; ++use; }
          }
        }
        w.nb += ntiles;
        if (!attn) {
          w.gb += cnt;
          w.rot += (uint32_t)nact;
          if (w.rot >= (uint32_t)G) w.rot -= G;
        }
        if (++ph == 5) { ph = 0; ++l; }
      }
    }
    return;
  }

  // ================================================================= CONSUMERS
  asm volatile("setmaxnreg.inc.sync.aligned.u32 224;\n");
  const int dflags = DBG ? p.dbg_flags : 0;
  const bool nowait = (dflags & 2) != 0;
  unsigned long long bar_target = p.bar_base[0];   // arrivals counted before this launch
  const uint32_t epoch = (uint32_t)p.bar_base[1];  // tag of phase it = epoch + it + 1
  Walk w;
  uint32_t sl = (uint32_t)warp, use = 0;   // ring slot / use count of this warp's next tile or item: n = warp, warp + NCW, ... across ALL phases
  uint32_t cur_slot = 0;
  auto release = [&]() {   // hand the ring slot back: every lane's shared-memory reads of it are issued, the arrive is ordered after them
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * cur_slot);
  };
  // optional phase timestamps (globaltimer ns, comparable across SMs): [CTA][phase][4] = {start, staged, items done, barrier done}
  long long* dbg = (DBG && p.dbg) ? p.dbg + (int64_t)c * nphase * 4 : nullptr;
  long long* ctr = nullptr;   // dev trace of one layer (see MegaArgs::dbg2)
  uint32_t ctr_nb0 = 0;
  int it = 0, l = 0, ph = 0;
  auto stamp = [&](int k) {
    if (!DBG) return;
    if (dbg && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
      dbg[it * 4 + k] = (long long)t;
    }
    if (ctr && tid == 0 && ph < 5) ctr[(160 + ph) * 4 + k] = clock64();
  };

  // tagged cross-CTA vectors (MegaArgs::tg): residual stream after attention (xa) / after the MLP (xb), q, the new
  // key and value rows, merged attention output, SwiGLU vector, per-CTA attention partials
  u64* const t_xa = p.tg;
  u64* const t_xb = t_xa + p.tg_H;
  u64* const t_q = t_xb + p.tg_H;
  u64* const t_kn = t_q + qd;
  u64* const t_vn = t_kn + kd;
  u64* const t_att = t_vn + kd;
  u64* const t_h = t_att + qd;
  u64* const t_part = t_h + p.tg_I;

  float best_v = -INFINITY;   // greedy tail: this thread's best (logit, index) over the lm_head rows it finished
  int best_i = 0x7fffffff;
  for (it = 0; it < nphase; ++it) {
    if (it == nphase - 1) { ph = PH_LM; l = 0; }
    const uint32_t tag = epoch + (uint32_t)it + 1u;   // tags of this phase's outputs; its inputs carry tag - 1
    if (DBG && ph == 0) {
      ctr = (p.dbg2 && l == p.dbg_layer && it != nphase - 1) ? p.dbg2 + (int64_t)c * MEGA_DBG2_ROWS * 4 : nullptr;
      ctr_nb0 = w.nb;
    }
    stamp(0);
    if (ph == PH_ATTN) {
      // ---------------- attention over this CTA's key range of its head. No grid-wide wait in front of it: the cached
      // keys/values arrive through the ring, and q / the new key and value are polled (by one warp) as tagged words of
      // this head only.
      if (as.active) {
        const int hw = lane >> 4, l16 = lane & 15;
        if (warp == 0) {   // one warp fetches q (and, in the head's last key range, the new key / value row): 64 pairs each
          const float sl2 = 0.08838834764831845f * 1.4426950408889634f;  // 128^-1/2 * log2(e)
          float2* q2 = reinterpret_cast<float2*>(qkn);
          const u64* qsrc = t_q + as.head * 128;
          const u64* const ptr[6] = {qsrc + 2 * lane, qsrc + 2 * (lane + 32), t_kn + kvh * 128 + 2 * lane, t_kn + kvh * 128 + 2 * (lane + 32),
                                     t_vn + kvh * 128 + 2 * lane, t_vn + kvh * 128 + 2 * (lane + 32)};
          const bool lastr = as.last != 0;
          const bool on[6] = {true, true, lastr, lastr, lastr, lastr};
          float2 v[6];
          ld_pairs<6>(ptr, on, tag - 1, nowait, v);
          q2[lane] = make_float2(v[0].x * sl2, v[0].y * sl2);
          q2[lane + 32] = make_float2(v[1].x * sl2, v[1].y * sl2);
          if (lastr) { q2[64 + lane] = v[2]; q2[96 + lane] = v[3]; q2[128 + lane] = v[4]; q2[160 + lane] = v[5]; }
        }
        consumer_sync();   // q is staged; every warp is past its last qkv tile (the merge scratch below aliases the staged vector)
        stamp(1);
        float q[8];
        {
          const float4 a = *reinterpret_cast<const float4*>(qkn + l16 * 8), b = *reinterpret_cast<const float4*>(qkn + l16 * 8 + 4);
          q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
        }
        float m = -INFINITY, lsum = 0.f, o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = 0.f;
        // online softmax over FOUR keys at a time (per half-warp): the four dot products and their shuffle reductions are
        // independent chains, one rescale per batch instead of one per key
        auto keys4 = [&](const uint4 (&kr)[4], const uint4 (&vr)[4], const bool (&valid)[4]) {
          float s2[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float kf[8];
            unpack8(kr[u], kf);
            float d = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) d += q[i] * kf[i];
            s2[u] = d;
          }
#pragma unroll
          for (int st = 8; st > 0; st >>= 1)
#pragma unroll
            for (int u = 0; u < 4; ++u) s2[u] += __shfl_xor_sync(0xffffffffu, s2[u], st);
          float mx = m;
#pragma unroll
          for (int u = 0; u < 4; ++u) { s2[u] = valid[u] ? s2[u] : -INFINITY; mx = fmaxf(mx, s2[u]); }
          if (mx == -INFINITY) return;           // nothing valid so far
          const float alpha = exp2f(m - mx);      // m = -inf -> 0
          float pj[4], ps = 0.f;
#pragma unroll
          for (int u = 0; u < 4; ++u) { pj[u] = exp2f(s2[u] - mx); ps += pj[u]; }
          lsum = lsum * alpha + ps;
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] *= alpha;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (valid[u]) {   // (positions past the range may hold uninitialised cache memory: 0 * NaN must not reach o)
              float vf[8];
              unpack8(vr[u], vf);
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] += pj[u] * vf[i];
            }
          }
          m = mx;
        };
        // ring items: 16 positions each; half-warp hw takes positions hw, hw + 2, ... of the item
        {
          uint32_t j = ((uint32_t)warp + NCW - (w.nb & (NCW - 1))) & (NCW - 1);
          if ((int)j < as.n_items) {
            for (; (int)j < as.n_items; j += NCW) {
              long long* trow = nullptr;
              if (DBG && ctr) {
                const uint32_t row = w.nb + j - ctr_nb0;
                if (row < 160u) trow = ctr + row * 4;
              }
              if (DBG && trow && lane == 0) trow[3] = clock64();
              mbar_wait(full0 + 8 * sl, use & 1);
              if (DBG && trow && lane == 0) trow[1] = clock64();
              cur_slot = sl;
              const uint32_t base = ring_u32 + sl * TILE_BYTES + l16 * 16;
              const int key0 = as.j0 + (int)j * 16;
              uint4 kr[2][4], vr[2][4];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const uint32_t a = base + (uint32_t)(u * 2 + hw) * 256u;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(kr[u >> 2][u & 3].x), "=r"(kr[u >> 2][u & 3].y), "=r"(kr[u >> 2][u & 3].z), "=r"(kr[u >> 2][u & 3].w) : "r"(a));
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(vr[u >> 2][u & 3].x), "=r"(vr[u >> 2][u & 3].y), "=r"(vr[u >> 2][u & 3].z), "=r"(vr[u >> 2][u & 3].w) : "r"(a + 4096u));
              }
              release();
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const bool valid[4] = {key0 + (h * 4 + 0) * 2 + hw < as.j1, key0 + (h * 4 + 1) * 2 + hw < as.j1,
                                       key0 + (h * 4 + 2) * 2 + hw < as.j1, key0 + (h * 4 + 3) * 2 + hw < as.j1};
                keys4(kr[h], vr[h], valid);
              }
              if (DBG && trow && lane == 0) trow[2] = clock64();
              sl += NCW;
              if (sl >= (uint32_t)nslots) { sl -= nslots; ++use; }
            }
          }
          w.nb += as.n_items;
        }
        if (as.last && warp == 0 && hw == 0) {   // the key / value of the token being decoded (published by the qkv phase of this launch)
          float d = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) d += q[i] * qkn[128 + l16 * 8 + i];
#pragma unroll
          for (int st = 8; st > 0; st >>= 1) d += __shfl_xor_sync(0x0000ffffu, d, st);
          const float mx = fmaxf(m, d), alpha = exp2f(m - mx), pj = exp2f(d - mx);
          lsum = lsum * alpha + pj;
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = o[i] * alpha + pj * qkn[256 + l16 * 8 + i];
          m = mx;
        }
        // merge the 16 half-warp states -> one partial per CTA
        float* sm_m = actf;            // [16]
        float* sm_l = actf + 16;       // [16]
        float* sm_o = actf + 32;       // [16][128]
        const int hidx = warp * 2 + hw;
        if (l16 == 0) { sm_m[hidx] = m; sm_l[hidx] = lsum; }
#pragma unroll
        for (int i = 0; i < 8; ++i) sm_o[hidx * 128 + l16 * 8 + i] = o[i];
        consumer_sync();
        const bool owner = c < p.heads;   // the CTA of the head's first key range folds the head's partials
        float* mg = actf + 32 + 16 * 128;   // [cph][132] (owner only)
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
          if (owner) {
            mg[tid] = O;
            if (tid == 0) { mg[128] = M; mg[129] = Lt; }
          } else {
            u64* pp = t_part + (int64_t)c * 132;
            st_tag(pp + tid, O, tag);
            if (tid == 0) { st_tag(pp + 128, M, tag); st_tag(pp + 129, Lt, tag); }
          }
        }
        if (owner) {
          // fixed owner instead of "last CTA to arrive": no atomic round trip on the critical path; the partials are tagged,
          // the owner polls them (coalesced, with back-off) as soon as its own share is done
          const int nw = (as.cph - 1) * 130;
          for (int i0 = 0; i0 < nw; i0 += 2 * CONSUMER_THREADS) {
            u64 wv2[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int i = i0 + u * CONSUMER_THREADS + tid;
              if (i < nw) wv2[u] = ld_weak1(t_part + (int64_t)((i / 130 + 1) * p.heads + as.head) * 132 + (i % 130));
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int i = i0 + u * CONSUMER_THREADS + tid;
              if (i < nw) mg[(i / 130 + 1) * 132 + (i % 130)] = settle1(wv2[u], t_part + (int64_t)((i / 130 + 1) * p.heads + as.head) * 132 + (i % 130), tag, nowait);
            }
          }
          consumer_sync();
          if (tid < 128) {
            float M = -INFINITY;
            for (int r = 0; r < as.cph; ++r) M = fmaxf(M, mg[r * 132 + 128]);
            float Lt = 0.f, O = 0.f;
#pragma unroll 1
            for (int r = 0; r < as.cph; ++r) {
              const float mr = mg[r * 132 + 128];
              const float wgt = (mr == -INFINITY) ? 0.f : exp2f(mr - M);
              Lt += mg[r * 132 + 129] * wgt;
              O += mg[r * 132 + tid] * wgt;
            }
            st_tag(t_att + as.head * 128 + tid, O / Lt, tag);
          }
        }
      } else {
        stamp(1);
      }
      stamp(2);
    } else {
      // ---------------- weight phase: multiply the CTA's row groups (each warp stages the input slices of its own tiles), fused epilogue
      const MegaMat& m = p.mat[mat_of(ph)];
      int g0, cnt, nact;
      phase_span(w, m, g0, cnt, nact);
      const int tpg = m.tpg;
      const int64_t no = (int64_t)l * p.norm_stride;
      const u64* src = ph == PH_QKV ? (l == 0 ? nullptr : t_xb) : ph == PH_O ? t_att : ph == PH_GU ? t_xa : ph == PH_DOWN ? t_h : t_xb;
      const int K = ph == PH_O ? qd : ph == PH_DOWN ? p.I : p.H;
      const bf16* nw = ph == PH_QKV ? p.norm1_0 + no : ph == PH_GU ? p.norm2_0 + no : ph == PH_LM ? p.final_norm : nullptr;
      float rn = 0.f;   // RMSNorm scale of this phase's input: formed by the first epilogue this warp runs (0 = not yet)
      stamp(1);
      const uint32_t nb0 = w.nb, gb0 = w.gb;
      // residual source of the O / DOWN epilogues: the row's previous value in the other residual buffer
      const u64* res_src = (ph == PH_O) ? t_xb : t_xa;
      const uint32_t res_tag = (ph == PH_O) ? tag - 3 : tag - 2;   // O: x after the previous layer's MLP; DOWN: this layer's xa
      const int ntiles = cnt * tpg;
      uint32_t j = ((uint32_t)warp + NCW - (nb0 & (NCW - 1))) & (NCW - 1);
      // The warps walk their tiles in ROUNDS (tile j + 8 r in round r) and meet every sync_every rounds: a ring slot belongs
      // to one consumer warp, so nothing else stops a warp from running many groups ahead of another one, and a warp that
      // gets NT tiles ahead would overwrite its own partial sums of a group whose epilogue has not run yet (seen at the
      // v2-8b shape: 880 lm_head tiles per CTA, scattered wrong logits). The interval is as long as the windows allow:
      // drift (interval rounds) + one group + the round in flight <= NT partial sums, groups in flight <= NG / 2.
      // (QKV -> attention has no barrier in between, but the attention phase writes no partial sums.)
      const int rounds = (ntiles + NCW - 1) / NCW;
      const int sync_every = max(1, min((NT - tpg - 2 * NCW) / NCW, NG * tpg / (2 * NCW)));
      int since_sync = 0;
      // ---- group epilogues. They do NOT run inside the tile loop: with "the warp that finishes a group's last tile runs its
      // epilogue" the warp that is last in one round starts its next tile late, is last again, and ends up with ALL epilogues of
      // the phase in series (gate/up: it finished 2 us after the other seven warps, profiles/r2_decode_trace_own_cost.txt). The
      // tiles only leave their partial sums in shared memory; after a CTA barrier the complete groups are dealt to the warps
      // (warp w: groups k_ep + w, + 8, ...) and run in parallel — no per-tile counter, atomic or fence either. Partials are
      // summed in k order (deterministic).
      uint32_t k_ep = 0;   // groups of this phase whose epilogue has run
      auto run_epilogues = [&](uint32_t k_hi) {
        for (uint32_t ek = k_ep + (uint32_t)warp; ek < k_hi; ek += NCW) {
          const uint32_t egs = (gb0 + ek) % NG;
              const uint32_t n0 = nb0 + ek * tpg;
              float v = 0.f;
              if (lane < 16)
                for (int t = 0; t < tpg; ++t) v += tpart[((n0 + t) % NT) * 16 + lane];
              const float v1 = __shfl_down_sync(0xffffffffu, v, 8);
              if (nw && rn == 0.f) rn = slices_rn(slice_ss, tpg, K, p.eps);   // (every slice of the vector is staged: the group is complete)
              if (lane < 8) {
                const int gi = g0 + (int)ek, r = lane;
                if (ph == PH_QKV) {
                  const int hb = gi >> 3, i = ((gi & 7) << 3) + r;      // 128-row block, index inside the half
                  const int row0 = hb * 128 + i;
                  const float a0 = v * rn, a1 = v1 * rn;
                  if (row0 < qd + kd) {
                    const float2 csn = *reinterpret_cast<const float2*>(rope_s + i * 2);
                    const float y0 = a0 * csn.x - a1 * csn.y, y1 = a1 * csn.x + a0 * csn.y;
                    if (row0 < qd) { st_tag(t_q + row0, y0, tag); st_tag(t_q + row0 + 64, y1, tag); }
                    else {
                      const int kh = (row0 - qd) >> 7;
                      bf16* dd = p.kv + (int64_t)slot * p.kv_slot_stride + (int64_t)l * p.kv_layer_stride + ((int64_t)kh * p.max_len + pos) * 128;
                      const bf16 z0 = __float2bfloat16_rn(y0), z1 = __float2bfloat16_rn(y1);
                      dd[i] = z0;
                      dd[i + 64] = z1;
                      st_tag(t_kn + kh * 128 + i, __bfloat162float(z0), tag);      // the cache row as this launch's attention reads it
                      st_tag(t_kn + kh * 128 + i + 64, __bfloat162float(z1), tag);
                    }
                  } else {
                    const int kh = (row0 - qd - kd) >> 7;
                    bf16* dd = p.kv + (int64_t)slot * p.kv_slot_stride + (int64_t)l * p.kv_layer_stride + p.kv_v_offset + ((int64_t)kh * p.max_len + pos) * 128;
                    const bf16 z0 = __float2bfloat16_rn(a0), z1 = __float2bfloat16_rn(a1);
                    dd[i] = z0;
                    dd[i + 64] = z1;
                    st_tag(t_vn + kh * 128 + i, __bfloat162float(z0), tag);
                    st_tag(t_vn + kh * 128 + i + 64, __bfloat162float(z1), tag);
                  }
                } else if (ph == PH_O || ph == PH_DOWN) {
                  u64* dst = (ph == PH_O) ? t_xa : t_xb;
                  const int r0 = gi * 16 + r, r1 = r0 + 8;
                  const float b0 = rbuf[egs * 16 + r], b1 = rbuf[egs * 16 + r + 8];
                  if (r0 < p.H) st_tag(dst + r0, b0 + v, tag);
                  if (r1 < p.H) st_tag(dst + r1, b1 + v1, tag);
                } else if (ph == PH_GU) {
                  const int i = gi * 8 + r;
                  if (i < p.I) st_tag(t_h + i, silu(v * rn) * (v1 * rn), tag);
                } else {
                  const int r0 = gi * 16 + r, r1 = r0 + 8;
                  const float l0 = v * rn, l1 = v1 * rn;
                  if (r0 < p.V) p.logits[r0] = l0;
                  if (r1 < p.V) p.logits[r1] = l1;
                  if (p.fuse_greedy) {   // rows are visited in increasing order per thread: '>' keeps the lowest index on ties
                    if (r0 < p.V && r0 != p.bad_token && l0 > best_v) { best_v = l0; best_i = r0; }
                    if (r1 < p.V && r1 != p.bad_token && l1 > best_v) { best_v = l1; best_i = r1; }
                  }
                }
              }
        }
        k_ep = k_hi;
      };
      {
        uint32_t k = 0, ks = j;
        while (ks >= (uint32_t)tpg) { ks -= tpg; ++k; }
        for (int rd = 0; rd <= rounds; ++rd) {
          // a meeting point: every sync_every rounds (partial-sum window) and once after the last round (rd == rounds). All tiles
          // j < 8 rd are done: the groups they complete get their epilogues (ONE call site: the epilogue code exists once).
          const bool at_end = rd == rounds;
          if (at_end || ++since_sync > sync_every) {
            consumer_sync();
            run_epilogues(at_end ? (uint32_t)cnt : (uint32_t)(NCW * rd) / (uint32_t)tpg);
            since_sync = 1;
          }
          if (at_end) break;
          if ((int)j >= ntiles) { j += NCW; continue; }
          long long* trow = nullptr;
          if (DBG && ctr) {
            const uint32_t row = nb0 + j - ctr_nb0;
            if (row < 160u) trow = ctr + row * 4;
          }
          if (DBG && trow && lane == 0) trow[3] = clock64();
          // the tile's slice of the input vector (normally staged by this very warp at its previous tile of the same ks)
          if (*reinterpret_cast<volatile uint32_t*>(slice_tag + ks) != tag)
            stage_slice(src, tag - 1, nowait, p.embed + (int64_t)tok * p.H, K, (int)ks, nw, xb, slice_ss, slice_tag, tag);
          mbar_wait(full0 + 8 * sl, use & 1);
          if (DBG && trow && lane == 0) trow[1] = clock64();
          // ---- one tile = 16 k-steps of (ldmatrix.x4, mma). B operand: even columns of the 16 x 8 B tile carry the hi
          // part of x, odd columns the lo part (column = lane >> 2), so ONE mma per k-step yields W.hi in accumulator
          // column 0 and W.lo in column 1. The A fragments are loaded in batches interleaved with the mma of earlier
          // batches, so that the tensor pipe starts while the rest of the tile is still being read (shared-memory returns
          // are in order; all 16 ldmatrix in front of the first mma made the two pipes take turns: 0.53 us per tile).
          float rA0, rA2;
          {
            cur_slot = sl;
            float acc[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
            const uint32_t ta = ring_u32 + sl * TILE_BYTES + lane * 16;
            const uint2* xp = reinterpret_cast<const uint2*>(xb + (size_t)ks * 64 + (lane & 3)) + ((lane >> 2) & 1);
            uint2 b[16];
#pragma unroll
            for (int s = 0; s < 16; ++s) b[s] = xp[s * 8];
            uint32_t a[16][4];
#pragma unroll
            for (int s = 0; s < 8; ++s) ldmatrix_x4(a[s][0], a[s][1], a[s][2], a[s][3], ta + s * 512);
#pragma unroll
            for (int bq = 0; bq < 4; ++bq) {
              if (bq < 2) {
#pragma unroll
                for (int s = 8 + 4 * bq; s < 12 + 4 * bq; ++s) ldmatrix_x4(a[s][0], a[s][1], a[s][2], a[s][3], ta + s * 512);
              }
              if (bq == 1) release();
              if (!(dflags & 1)) {
#pragma unroll
                for (int s = 4 * bq; s < 4 * bq + 4; s += 2) {
                  mma_bf16_16816(acc, a[s], b[s].x, b[s].y);
                  mma_bf16_16816(c1, a[s + 1], b[s + 1].x, b[s + 1].y);
                }
              }
            }
            // lanes with (lane & 3) == 0 hold columns 0 (hi) and 1 (lo) of rows g (c[0], c[1]) and g + 8 (c[2], c[3])
            rA0 = (acc[0] + c1[0]) + (acc[1] + c1[1]);
            rA2 = (acc[2] + c1[2]) + (acc[3] + c1[3]);
          }
          const uint32_t gsA = (gb0 + k) % NG;
          if ((ph == PH_O || ph == PH_DOWN) && lane < 16 && ks == 0) {
            // residual of row (group, lane), fetched at the group's FIRST tile so that its L2 latency is off the
            // critical path of the group's epilogue (the value was published two or more phases ago)
            const int row = (g0 + (int)k) * 16 + lane;
            float bres = 0.f;
            if (row < p.H)
              bres = (ph == PH_O && l == 0) ? __bfloat162float(p.embed[(int64_t)tok * p.H + row])
                                            : settle1(ld_weak1(res_src + row), res_src + row, res_tag, nowait);
            rbuf[gsA * 16 + lane] = bres;
          }
          if ((lane & 3) == 0) {
            float* tp = tpart + ((nb0 + j) % NT) * 16;
            tp[lane >> 2] = rA0;
            tp[(lane >> 2) + 8] = rA2;
          }
          if (DBG && trow && lane == 0) trow[2] = clock64();
          j += NCW;
          sl += NCW;
          if (sl >= (uint32_t)nslots) { sl -= nslots; ++use; }
          ks += NCW;
          while (ks >= (uint32_t)tpg) { ks -= tpg; ++k; }
        }
      }
      w.nb += ntiles;
      w.gb += cnt;
      w.rot += (uint32_t)nact;
      if (w.rot >= (uint32_t)G) w.rot -= G;
      stamp(2);
    }
    // Phase boundary: a CTA-wide barrier only (none after qkv: the attention phase meets after polling its head's q; none
    // after lm_head). No grid-wide arrival counter: the next phase's warps poll the tagged words of their own slices
    // (0.80 vs 0.91 ms per token with a counter, profiles/r2_decode_variants.txt). The CTA barrier is what allows the single
    // vector buffer (slices are overwritten by the next phase; the attention merge scratch aliases it), and it is also FASTER
    // than letting the warps drift (0.788 vs 0.809 ms per token, profiles/r2_decode_ab.txt).
    if (ph != PH_QKV && ph != PH_LM) consumer_sync();
    stamp(3);
    if (++ph == 5) { ph = 0; ++l; }
  }
  if (p.fuse_greedy) {
    // ---- greedy tail (HF argmax after NoBadWords; lowest index wins ties): CTA-local argmax, one 64-bit atomicMax per CTA
    // on {order-preserving logit bits, ~index}; the last CTA to arrive publishes the token exactly as the sampler does
    // (device state of the loop + ONE 8-byte store to the mapped host ring) — no second launch per token.
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
    }
    int* red_i = reinterpret_cast<int*>(red + 8);
    consumer_sync();   // every warp has read the final norm's partial sums out of `red`
    if (lane == 0) { red[warp] = best_v; red_i[warp] = best_i; }
    consumer_sync();
    if (tid == 0) {
      for (int wv = 1; wv < NCW; ++wv)
        if (red[wv] > best_v || (red[wv] == best_v && red_i[wv] < best_i)) { best_v = red[wv]; best_i = red_i[wv]; }
      uint32_t u = __float_as_uint(best_v);
      u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
      const unsigned long long key = ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (uint32_t)best_i);
      atomicMax(p.amax, key);
      unsigned long long prev;
      asm volatile("atom.acq_rel.gpu.global.add.u64 %0, [%1], 1;\n" : "=l"(prev) : "l"(p.amax + 1) : "memory");
      if (prev == (unsigned long long)G - 1ull) {
        const unsigned long long best = atomicExch(p.amax, 0ull);
        const int token = (int)(0xffffffffu - (uint32_t)(best & 0xffffffffull));
        const unsigned long long gstep = *p.gen_step;
        p.gen_tok[0] = token;
        p.gen_pos[0] = min(pos + 1, p.max_pos);
        const unsigned long long entry = ((gstep + 1ull) << 32) | (unsigned long long)(unsigned)token;
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;\n" ::"l"(p.host_ring + (gstep % (unsigned long long)p.ring)), "l"(entry) : "memory");
        *p.gen_step = gstep + 1ull;
        p.amax[1] = 0ull;
      }
    }
  }
  // publish the arrival count and the tag epoch for the next launch (stream-ordered): every CTA made 4L arrivals and
  // the launch used tags epoch + 1 .. epoch + 5L + 1
  if (c == 0 && tid == 0 && !(dflags & 2)) {
    p.bar_base[0] = bar_target;
    p.bar_base[1] = (unsigned long long)(uint32_t)(epoch + (uint32_t)nphase);
  }
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
  return a.nslots * TILE_BYTES + a.act_floats * 4 + 2 * a.nslots * 8 + (16 + 128 + NT * 16) * 4 + NG * 4 + NG * 16 * 4 + 384 * 4 + NS * 8;
}

cudaError_t mega_configure(MegaArgs& a, int H, int I, int heads, int max_smem_optin, int num_sms, int* grid_out) {
  auto pad = [](int k) { return (k + 255) / 256 * 256; };
  int actf = pad(H) > pad(I) ? pad(H) : pad(I);
  if (pad(heads * 128) > actf) actf = pad(heads * 128);
  if (actf < 32 + 16 * 128 + 16 * 132) actf = 32 + 16 * 128 + 16 * 132;  // attention merge scratch (half-warp states + the head's partials)
  actf = (actf + 31) & ~31;
  a.act_floats = actf;
  a.tg_H = pad(H);
  a.tg_I = pad(I);
  if ((I + 255) / 256 > NT - 44 || (I + 255) / 256 > NS || (H + 255) / 256 > NS) return cudaErrorInvalidValue;  // partial-sum window must cover a group + tiles in flight; slice table
  const int fixed = actf * 4 + (16 + 128 + NT * 16) * 4 + NG * 4 + NG * 16 * 4 + 384 * 4 + NS * 8 + 64;
  int nslots = (max_smem_optin - fixed) / (TILE_BYTES + 16);
  if (nslots > 32) nslots = 32;
  // every ring slot must always be filled by the same producer warp and drained by the same consumer warp
  // (slot s <-> producer s % NPW, consumer s % NCW): mbarrier parity waits are only alias-free when the
  // successive uses of one barrier are ordered inside one thread.
  nslots &= ~(NCW - 1);
  if (nslots < NCW) return cudaErrorInvalidValue;
  a.nslots = nslots;
  if (heads > num_sms) return cudaErrorInvalidValue;
  *grid_out = num_sms;
  return cudaSuccess;
}

cudaError_t launch_decode_mega(const MegaArgs& a, int grid, cudaStream_t s, uint64_t* counter) {
  const int smem = mega_smem_bytes(a);
  const bool dbgk = a.dbg != nullptr || a.dbg2 != nullptr || a.dbg_flags != 0;
  const void* fn = dbgk ? (const void*)decode_mega_kernel<true> : (const void*)decode_mega_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  void* args[] = {(void*)&a};
  e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(MEGA_THREADS), args, (size_t)smem, s);
  if (counter) ++*counter;
  return e;
}

}  // namespace dtk
