// Host-side launcher declarations (one per kernel family). Each launcher returns the CUDA error
// of the launch and bumps *counter (kernel launches issued; bench.py's gpu_launches).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef __nv_bfloat16 bf16;

namespace dtk {

// ---------------------------------------------------------------- dense GEMM  C = A * W^T
enum { ACT_NONE = 0, ACT_GELU_TANH = 1, ACT_GELU_ERF = 2 };

struct GemmArgs {
  const bf16* A;              // [M, K] bf16, row stride lda (elements)
  int64_t lda;
  int a_rows_per_batch;       // 0 = plain; else row m lives at A + (m / rpb) * a_batch_stride + (m % rpb) * lda
  int64_t a_batch_stride;
  const bf16* W;              // [N, K] bf16 row-major (K contiguous), row stride ldw
  int64_t ldw;
  int M, N, K;                // K % 8 == 0, N % 2 == 0
  // epilogue: v = acc (+bias[n]) ; v = act(v) ; (+ rowbias[m % rowbias_mod, n]) ; (+ resid[m, n])
  const bf16* bias;           // [N] or null
  const bf16* rowbias;        // [rowbias_mod, N] or null (ViT position embedding)
  int rowbias_mod;
  const float* resid;         // fp32 [M, ldr] or null (may alias out_f32)
  int64_t ldr;
  int act;
  int glu;                    // 1: columns come in (gate, up) pairs -> out[m, n/2] = silu(gate) * up
  float* out_f32;             // exactly one of out_f32 / out_bf16 is non-null
  bf16* out_bf16;
  int64_t ldo;
};
cudaError_t launch_gemm(const GemmArgs& a, cudaStream_t s, uint64_t* counter);   // dispatches tcgen05 / mma.sync
cudaError_t launch_gemm_mma(const GemmArgs& a, cudaStream_t s, uint64_t* counter);
cudaError_t launch_gemm_tc(const GemmArgs& a, cudaStream_t s, uint64_t* counter);
bool gemm_tc_supported(const GemmArgs& a);
void set_gemm_swap_split(int v);    // dev: 0 (default) = heuristic split-K factor of the swapped tile, 1..8 = forced
void set_gemm_skinny_swap(int v);   // dev: 1 (default) = swapped-operand tcgen05 tile for M < 64, 0 = 128 x 32 tile
void set_gemm_impl(int impl);   // 0 = mma.sync everywhere, 1 = tcgen05 where supported (process-wide dev switch)
int get_gemm_impl();

// ---------------------------------------------------------------- flash attention (mma.sync)
struct AttnArgs {
  const bf16 *q, *k, *v;
  bf16* o;
  int64_t q_bs, q_hs, q_rs;   // batch / head / row strides in elements
  int64_t k_bs, k_hs, k_rs;
  int64_t v_bs, v_hs, v_rs;
  int64_t o_bs, o_hs, o_rs;
  int B, heads, kv_group;     // kv head = head / kv_group
  int Tq, Tk;
  int q_pos0;                 // causal: query i sits at position q_pos0 + i; key j visible iff j <= pos
  int causal;
  int head_dim;               // 72 or 128
  float scale;
  // shared KV prefix (prefill on a sequence that borrows positions [0, split_row) from another slot): key rows below
  // split_row are read from k2 / v2 (same strides), the rest from k / v. split_row = 0: everything from k / v.
  const bf16 *k2, *v2;
  int split_row;
  // partial mode (shared-prefix attention of a batched decode step; part_o != null): the Tq <= 64 query rows are the
  // rollouts of one figure, blockIdx.x selects a range of part_tiles key tiles, and instead of the normalised output the
  // kernel exports the flash state of that range in the convention of decode_attn_kernel's partials:
  //   part_ml[(row * heads + head) * part_np + part_idx0 + blockIdx.x] = { max score * scale * log2(e), sum exp2 },
  //   part_o[... * 128 + d] = unnormalised output.
  float *part_o, *part_ml;
  int part_np, part_idx0, part_tiles;
};
cudaError_t launch_flash_attn(const AttnArgs& a, cudaStream_t s, uint64_t* counter);

// ---------------------------------------------------------------- ViT attention on tcgen05 (attn_tc.cu): head_dim 72, non-causal
// qkv bf16 [B*N, 3*heads*72] (q | k | v column blocks); vT: scratch bf16 [B*heads*80, attn_tc_vt_cols(N)]; o bf16 [B*N, heads*72]
bool attn_tc_supported();
int attn_tc_vt_cols(int N);
cudaError_t launch_attn_tc(const bf16* qkv, bf16* vT, bf16* o, int B, int heads, int N, float scale, cudaStream_t s, uint64_t* counter);

// ---------------------------------------------------------------- row-wise / elementwise
// y = LN(x) * w + b  (fp32 stats); x fp32 [M, D]; writes bf16 and/or fp32 outputs
cudaError_t launch_layernorm(const float* x, const bf16* w, const bf16* b, float eps, int M, int D,
                             bf16* out_bf16, float* out_f32, cudaStream_t s, uint64_t* counter);
// y = x * rsqrt(mean(x^2)+eps) * w ; x fp32 [M, D] row stride ldx; out bf16 [M, D]
cudaError_t launch_rmsnorm(const float* x, int64_t ldx, const bf16* w, float eps, int M, int D, bf16* out,
                           cudaStream_t s, uint64_t* counter);
// pixels fp32 [B,3,S,S] -> patches bf16 [B*N, KP] (channel-major (c, py, px) like conv weight; zero pad)
cudaError_t launch_im2col(const float* pixels, int B, int S, int P, int KP, bf16* out, cudaStream_t s,
                          uint64_t* counter);
cudaError_t launch_cast_f32_bf16(const float* in, bf16* out, int64_t n, cudaStream_t s, uint64_t* counter);
// x[t, :] = ids[t] == image_token ? img[(start_pos + t) - img_start, :] : embed[ids[t], :]
cudaError_t launch_embed_splice(const int64_t* ids, int T, int start_pos, const bf16* embed, int H, int vocab,
                                int image_token, const float* img, int img_start, int n_img, float* x,
                                cudaStream_t s, uint64_t* counter);
// prefill: qkv fp32 [T, qd+2kd] -> roped q bf16 [T, qd]; K/V bf16 into the cache at positions start_pos+t
cudaError_t launch_rope_kv_decode(const float* qkv, int B, const int* slots, const int* pos, int heads, int kv_heads,
                                  const float* rope_cs, float* q_out, bf16* kv_base, int64_t kv_slot_stride,
                                  int64_t kv_v_offset, int max_len, cudaStream_t s, uint64_t* counter, bf16* q_bf16 = nullptr);
cudaError_t launch_rope_kv_prefill(const float* qkv, int T, int start_pos, int heads, int kv_heads,
                                   const float* rope_cs, bf16* q_out, bf16* kcache, bf16* vcache,
                                   int max_len, cudaStream_t s, uint64_t* counter);

// ---------------------------------------------------------------- decode (batch of single tokens)
enum { GEMV_STORE = 0, GEMV_ADD = 1, GEMV_GLU = 2, GEMV_QKV = 3 };
struct GemvArgs {
  int mode;
  const bf16* W;              // [N, K]
  int N, K;
  const float* x;             // [B, x_stride] fp32
  int64_t x_stride;
  const bf16* norm_w;         // fused RMSNorm on x (null = none)
  float eps;
  float* out;                 // STORE/ADD: [B, out_stride]; GLU: [B, out_stride] (N/2 valid); QKV: q [B, out_stride]
  int64_t out_stride;
  int B;
  // QKV mode
  const int* slots;           // device int[B]
  const int* pos;             // device int[B] : position of the token being processed
  const float* rope_cs;       // fp32 [max_len, 64, 2] (cos, sin)
  bf16* kv_base;              // cache base of this layer for slot 0: K then V
  int64_t kv_slot_stride;     // elements between slots
  int64_t kv_v_offset;        // elements from K to V of the same layer
  int q_dim, kv_dim, max_len;
};
cudaError_t launch_gemv(const GemvArgs& a, cudaStream_t s, uint64_t* counter);

// x[b, :] = embed[tok[b], :]  (fp32 out)
cudaError_t launch_embed_tokens(const int* tok32, const int64_t* tok64, int B, const bf16* embed, int H,
                                int vocab, float* x, cudaStream_t s, uint64_t* counter);

struct DecodeAttnArgs {
  const float* q;             // [B, q_dim] fp32 (roped)
  int64_t q_stride;
  const bf16* kv_base;        // layer base for slot 0
  int64_t kv_slot_stride, kv_v_offset;
  const int* slots;           // device int[B]
  const int* pos;             // device int[B]; keys [0, pos] are attended
  const int* share_slot;      // device int[B]: slot that holds positions [0, share_len[b]) of sequence b (shared prefix)
  const int* share_len;       // device int[B]: 0 = nothing shared
  int B, heads, kv_group, max_len, nsplit;
  float scale;
  float* part_o;              // [B, heads, np, 128]
  float* part_ml;             // [B, heads, np, 2]
  unsigned int* counters;     // [B * heads], zero-initialised, self-resetting
  float* out;                 // [B, q_dim] fp32
  int64_t out_stride;
  // shared-prefix ("cascade") mode: keys [0, key_begin) were already reduced by launch_flash_attn in partial mode into the
  // partial slots [nsplit, np); this kernel covers [key_begin, pos] and its merge adds all np partials. 0 / nsplit = off.
  int key_begin, np;
  bf16* out_bf16;             // optional bf16 copy of the output (the o-proj GEMM operand), same layout
};
cudaError_t launch_decode_attn(const DecodeAttnArgs& a, cudaStream_t s, uint64_t* counter);

struct SampleSeq {
  int suppress;
  uint32_t step;
  uint32_t seq_id;
};
struct SampleArgs {
  const float* logits;        // [B, V]
  int B, V;
  float temperature, top_p;
  float top_p_limit;              // (float)(1.0 - top_p): ascending cumulative mass <= limit is removed
  int top_k, do_sample, bad_token, bs_token;
  uint64_t seed;
  const unsigned long long* seed_dev;   // generation loop: the seed lives in a device word (written by gen_begin) so that a
                                        // captured decode graph does not depend on it; null = use `seed`
  float* scratch;             // [B, V] fp32 work buffer (receives the final probability vector)
  int want_probs;             // greedy only: also write softmax(masked logits) to scratch (sampling always writes it)
  int64_t* out_ids;           // device int64[B] or null
  // generation-loop state (all optional, device): when gen_tok != null the sampler also advances
  // the loop: gen_tok[b] = token, gen_pos[b] += 1, and publishes the token to the pinned host ring.
  int* gen_tok;
  int* gen_pos;
  unsigned long long* gen_step;   // single counter (device); RNG counter + ring row
  unsigned long long* host_ring;  // mapped pinned u64 [ring, B]: ((step + 1) << 32) | token, one store per token
  int ring;
  int max_pos;                    // gen_pos is clamped to this (max_len - 1)
  unsigned int* done_counter;     // device, zero-initialised, self-resetting
  SampleSeq seq[64];
};
cudaError_t launch_sample(const SampleArgs& a, cudaStream_t s, uint64_t* counter);
void set_sample_impl(int impl);   // 0 = register-resident kernel when V <= 32768 (default), 1 = generic kernel
int get_sample_impl();

// ---------------------------------------------------------------- persistent decode kernel (B = 1)
// Decode-side weight copy: every matrix is re-tiled once at load into 8 KB tiles of 16 rows x 256 k that a
// single 1-D bulk copy lands in shared memory exactly as ldmatrix.x4 wants them ([kstep 16][matrix 4][row 8][8 bf16]).
// Row groups are permuted so that the two accumulator rows (g, g+8) of one thread are a RoPE pair (i, i+64)
// or a SwiGLU pair (gate_i, up_i).
enum { TILE_SEQ = 0, TILE_ROPE = 1, TILE_GLU = 2 };
constexpr int MEGA_TILE_ELEMS = 4096;  // 16 x 256 bf16 = 8 KB
constexpr int MEGA_DBG2_ROWS = 168;     // dev trace: rows 0..159 = local tiles of the traced layer, 160..164 = phase stamps
struct MegaMat {
  const bf16* base;       // tiled copy, layer 0
  int64_t layer_stride;   // elements between layers
  int N, K;               // logical rows / cols of the matrix (GLU: N = 2I interleaved source rows)
  int groups, tpg;        // 16-row groups, 256-column tiles per group
  int per, nact;          // work split over the grid (host-computed: no division in the kernel): per = ceil(groups / grid) groups per participating CTA, nact = ceil(groups / per) participants
  int mode;
};
struct MegaArgs {
  int H, I, L, heads, kv_heads, V, max_len;
  float eps;
  const bf16 *embed, *final_norm;
  const bf16 *norm1_0, *norm2_0;                               // row-major arena; layer l = ptr + l * norm_stride
  int64_t norm_stride;
  MegaMat mat[5];                                             // qkv, o, gate/up, down (per layer) and lm_head
  const int *tok, *pos, *slots;                               // device state of the sequence being decoded
  const int *share_slot, *share_len;                          // shared KV prefix: positions [0, share_len[0]) live in share_slot[0] (multiple of 16)
  bf16* kv;
  int64_t kv_slot_stride, kv_layer_stride, kv_v_offset;
  const float* rope_cs;
  float* logits;
  // tagged cross-CTA activation words {fp32 value, phase tag} (zero-initialised): xa[tg_H] xb[tg_H] q[heads*128]
  // knew[kv_heads*128] vnew[kv_heads*128] attn[heads*128] h[tg_I] part[grid][132]
  unsigned long long* tg;
  int tg_H, tg_I;                                             // H and I padded to the 256-column tile (mega_configure)
  unsigned int* head_cnt;                                     // [heads] arrival counters (zero-initialised, self-resetting)
  unsigned long long *bar_count, *bar_base;                   // arrival counter; bar_base[0] = arrivals, [1] = tag epoch of past launches
  int nslots, act_floats;                                     // shared-memory ring geometry (mega_configure)
  // greedy generation loop: argmax of the masked logits in the kernel tail + token publication (replaces the sampler launch)
  int fuse_greedy, bad_token, ring, max_pos;
  unsigned long long* amax;                                   // [0] packed (ordered logit, ~index) max cell, [1] arrival counter; zero-initialised, self-resetting
  int *gen_tok, *gen_pos;
  unsigned long long *gen_step, *host_ring;
  int variant;                                                // dev A/B switches (option "mega_variant"): no switches at present: the round-2 A/B variants were decided, see profiles/r2_decode_ab.txt
  int dbg_flags;                                              // dev only: 1 = skip tile math, 2 = skip grid barriers, 4/8 = relaxed arrive/poll
  long long* dbg;                                             // optional: [grid][5L+1][4] globaltimer stamps (null = off)
  long long* dbg2;                                            // optional: [grid][MEGA_DBG2_ROWS][4] clock64 per-tile trace of layer dbg_layer
  int dbg_layer;
};
int mega_smem_bytes(const MegaArgs& a);
cudaError_t mega_configure(MegaArgs& a, int H, int I, int heads, int max_smem_optin, int num_sms, int* grid_out);
cudaError_t launch_decode_mega(const MegaArgs& a, int grid, cudaStream_t s, uint64_t* counter);
// one-time re-tiling of a row-major [N, K] (ld = K) matrix into the decode layout
int64_t mega_tiled_elems(int N, int K, int mode, int* groups, int* tpg);
cudaError_t launch_retile(const bf16* src, int N, int K, int mode, bf16* dst, cudaStream_t s);

// device image preprocessing (Pillow-exact 8-bit bicubic resize + rescale + normalise): rgb uint8 [h, w, 3] -> fp32 [3, S, S]
cudaError_t launch_image_preprocess(const uint8_t* rgb, int h, int w, int S, const int* bounds_h, const int* coef_h, int ksize_h,
                                    const int* bounds_v, const int* coef_v, int ksize_v, float rescale, const float* mean,
                                    const float* std, uint8_t* tmp, float* out, uint8_t* out_u8, cudaStream_t s, uint64_t* counter);

// single-query attention of the SigLIP attention-pool head: q fp32 [heads*72] (shared by all images),
// kv bf16 [B*N, 2*D] -> out bf16 [B, D]
cudaError_t launch_pool_attn(const float* q, const bf16* kv, int B, int N, int D, int heads, float scale,
                             bf16* out, cudaStream_t s, uint64_t* counter);

}  // namespace dtk
