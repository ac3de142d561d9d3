/* Plain-C client of the C ABI (C99, no CUDA/torch headers): what a cgo / JNI / ctypes binding sees. Exercises the
 * host-only entry points; every other declared symbol must resolve at link time (the table below takes their addresses).
 * Built and run by tests/test_cpu_abi.py. */
#include <stdio.h>
#include <string.h>

#include "detikzify_b200.h"

int main(void) {
  dtk_config c;
  memset(&c, 0, sizeof c);
  /* detikzify-ds-1.3b + SigLIP so400m/14@384 */
  c.hidden = 2048; c.inter = 5504; c.layers = 24; c.heads = 16; c.kv_heads = 16; c.head_dim = 128; c.vocab = 32256; c.max_len = 2048;
  c.rms_eps = 1e-6f; c.rope_theta = 100000.f; c.rope_factor = 4.f; c.rope_type = 0;
  c.v_hidden = 1152; c.v_inter = 4304; c.v_layers = 27; c.v_heads = 16; c.v_image = 384; c.v_patch = 14; c.v_act = 0; c.v_eps = 1e-6f;
  c.concat = 3; c.image_token_id = 32013; c.eos_token_id = 32021; c.max_seqs = 2; c.max_batch = 1;

  if (dtk_abi_version() != DTK_ABI_VERSION) { fprintf(stderr, "abi version\n"); return 1; }
  int n = dtk_weight_count(&c);
  if (n <= 0) { fprintf(stderr, "weight count %d\n", n); return 2; }
  unsigned long long total = 0, last_end = 0;
  for (int i = 0; i < n; ++i) {
    dtk_weight_info w;
    if (dtk_weight_get(&c, i, &w) != DTK_OK) { fprintf(stderr, "weight_get %d\n", i); return 3; }
    if (w.offset % 256 != 0 || w.offset < last_end) { fprintf(stderr, "layout %s\n", w.name); return 4; }
    if (w.nbytes != (unsigned long long)w.rows * (unsigned long long)w.cols * 2ull) { fprintf(stderr, "nbytes %s\n", w.name); return 5; }
    last_end = w.offset + w.nbytes;
    total += w.nbytes;
  }
  unsigned long long arena = dtk_arena_bytes(&c);
  if (arena < last_end) { fprintf(stderr, "arena\n"); return 6; }
  /* algorithmic decode bytes: 2.561 GB of weights + 196608 B per cached position */
  unsigned long long b0 = dtk_decode_bytes(&c, 0), b1 = dtk_decode_bytes(&c, 1000);
  if (b1 - b0 != 1000ull * 196608ull) { fprintf(stderr, "kv bytes %llu\n", b1 - b0); return 7; }
  if (b0 < 2560000000ull || b0 > 2563000000ull) { fprintf(stderr, "weight bytes %llu\n", b0); return 8; }
  c.head_dim = 64;                                   /* unsupported shape must be refused, not crash */
  if (dtk_arena_bytes(&c) != 0) { fprintf(stderr, "bad config accepted\n"); return 9; }
  /* link-time presence of the device entry points (not called: no GPU in this test) */
  void* syms[] = {(void*)dtk_create, (void*)dtk_destroy, (void*)dtk_last_error, (void*)dtk_vit_encode, (void*)dtk_project, (void*)dtk_image_preprocess,
                  (void*)dtk_seq_alloc, (void*)dtk_seq_free, (void*)dtk_seq_fork, (void*)dtk_seq_share, (void*)dtk_prefill, (void*)dtk_decode,
                  (void*)dtk_sample, (void*)dtk_gen_begin, (void*)dtk_gen_step, (void*)dtk_gen_wait, (void*)dtk_gen_end,
                  (void*)dtk_set_option, (void*)dtk_get_option, (void*)dtk_launch_count};
  for (unsigned i = 0; i < sizeof syms / sizeof syms[0]; ++i) if (!syms[i]) return 10;
  printf("ok weights=%d arena=%llu decode_bytes0=%llu\n", n, arena, b0);
  return 0;
}
