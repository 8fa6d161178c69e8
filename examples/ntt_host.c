/* Minimal C client of libboojum_b200 (no Python, no torch): a forward coset NTT and its inverse on host buffers, and a
 * Poseidon2 Merkle cap.  Shows that the drop-in boundary is a plain C ABI.
 *   gcc -std=c99 -I include examples/ntt_host.c -L era_boojum_b200 -lboojum_b200 -Wl,-rpath,$PWD/era_boojum_b200 -o ntt_host */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "boojum_b200.h"

#define CHECK(call)                                                                   \
  do {                                                                                \
    int32_t st_ = (call);                                                             \
    if (st_ != BJ_OK) {                                                               \
      fprintf(stderr, "%s -> %d (%s): %s\n", #call, st_, bj_status_string(st_), ctx ? bj_last_error(ctx) : ""); \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

int main(void) {
  bj_ctx* ctx = NULL;
  const uint32_t log_n = 16, n_cols = 4;
  const size_t n = (size_t)1 << log_n;
  CHECK(bj_ctx_create(0, NULL, &ctx)); /* BJ_ERR_NO_DEVICE without a GPU: there is no CPU fallback */
  uint64_t* h = NULL;
  CHECK(bj_alloc_host_pinned(sizeof(uint64_t) * n * n_cols, (void**)&h));
  uint64_t* ref = (uint64_t*)malloc(sizeof(uint64_t) * n * n_cols);
  for (size_t i = 0; i < n * n_cols; i++) ref[i] = h[i] = (i * 0x9E3779B97F4A7C15ull) % BJ_GOLDILOCKS_P;
  CHECK(bj_ntt_natural_to_bitreversed_host(ctx, h, log_n, n_cols, 7));
  /* back: bit-reversed values -> natural order on the device, then the inverse transform */
  uint64_t* d = NULL;
  CHECK(bj_alloc(ctx, sizeof(uint64_t) * n * n_cols, (void**)&d));
  CHECK(bj_upload(ctx, d, h, sizeof(uint64_t) * n * n_cols));
  CHECK(bj_bitreverse(ctx, d, log_n, n_cols, n));
  CHECK(bj_intt_natural_to_natural(ctx, d, log_n, n_cols, n, 7));
  CHECK(bj_download(ctx, h, d, sizeof(uint64_t) * n * n_cols));
  CHECK(bj_ctx_synchronize(ctx));
  printf("round trip %s\n", memcmp(h, ref, sizeof(uint64_t) * n * n_cols) == 0 ? "ok" : "MISMATCH");
  /* Merkle cap over the four columns */
  const uint64_t* srcs[4] = {d, d + n, d + 2 * n, d + 3 * n};
  uint64_t *leaf_hashes = NULL, *nodes = NULL, cap[16 * 4];
  CHECK(bj_alloc(ctx, sizeof(uint64_t) * 4 * n, (void**)&leaf_hashes));
  CHECK(bj_alloc(ctx, sizeof(uint64_t) * 4 * (n - 16), (void**)&nodes));
  CHECK(bj_merkle_build_poseidon2(ctx, srcs, 4, n, 1, 16, leaf_hashes, nodes));
  CHECK(bj_download(ctx, cap, nodes + 4 * (n - 32), sizeof(cap)));
  CHECK(bj_ctx_synchronize(ctx));
  printf("cap[0] = %016llx %016llx %016llx %016llx, %llu kernel launches\n", (unsigned long long)cap[0], (unsigned long long)cap[1],
         (unsigned long long)cap[2], (unsigned long long)cap[3], (unsigned long long)bj_launch_count(ctx));
  bj_free(ctx, d);
  bj_free(ctx, leaf_hashes);
  bj_free(ctx, nodes);
  bj_free_host_pinned(h);
  free(ref);
  bj_ctx_destroy(ctx);
  return 0;
}
