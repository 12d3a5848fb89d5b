/* fake_executor.c -- TEST INFRASTRUCTURE: a recording stand-in for the spp_executor_* entry points of
 * include/salient_b200.h, so the native host path (csrc/host_session.cpp) can be driven on a machine
 * without a GPU.  "Device" pointers are host memory here (the test hands CPU tensors to the
 * HostSession); submit() plays the GPU's part with a deterministic pattern a test can predict:
 *
 *   T_0 = batch_size, E_h = min(2 * T_h + h, out_col_cap[h]), T_{h+1} = T_h + E_h / 2
 *   out_rowptr[h][i] = 1000 * (h + 1) + i          (i <= T_h)
 *   out_col[h][e]    = seed0 + 10 * (h + 1) + e    (e <  E_h),  seed0 = first seed of the batch
 *   n_id_out[i]      = seed0 + i                   (i <  N_b = T_L)
 *   x_out[i, :]      = bytes (seed0 + i) & 0xff
 *   y_out[i]         = 3 * seeds[i]                (y_row_bytes == 8)
 *   buckets (do_split, P = fmap.num_parts): counts[p] = N_b / (P + 1) for p < P, counts[P] = the rest,
 *   bucket_ids[pos] = 7 * pos + seed0, perm[i] = N_b - 1 - i, counts[P + 1] = N_b
 *
 * A ticket completes after `fx_polls_needed` polls (or a wait), which lets a test exercise the
 * not-ready / blocking branches.  Cites: spp_executor_submit / poll / wait, include/salient_b200.h. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "salient_b200.h"

#define FX_RING 64

typedef struct fx_ticket {
  uint64_t id;
  int polls;
  int waited;
} fx_ticket;

typedef struct fx_executor {
  uint64_t next_ticket;
  fx_ticket ring[FX_RING];
  int polls_needed;
  int fail_submit_at; /* ticket number whose submit fails (0: never) */
  int overflow_at;    /* ticket number whose size block reports SPP_META_OVERFLOW (0: never) */
  int64_t submitted, waits;
  spp_batch_job last;
} fx_executor;

static const char* g_err = "";

void* fx_create(int polls_needed) {
  fx_executor* ex = (fx_executor*)calloc(1, sizeof(fx_executor));
  ex->next_ticket = 1;
  ex->polls_needed = polls_needed;
  return ex;
}
void fx_destroy(void* e) { free(e); }
void fx_fail_submit_at(void* e, int t) { ((fx_executor*)e)->fail_submit_at = t; }
void fx_overflow_at(void* e, int t) { ((fx_executor*)e)->overflow_at = t; }
int64_t fx_submitted(void* e) { return ((fx_executor*)e)->submitted; }
int64_t fx_waits(void* e) { return ((fx_executor*)e)->waits; }
const spp_batch_job* fx_last_job(void* e) { return &((fx_executor*)e)->last; }
const char* fx_last_error(void) { return g_err; }

uint64_t fx_submit(void* e, const spp_batch_job* j) {
  fx_executor* ex = (fx_executor*)e;
  if (ex->fail_submit_at && (int)ex->next_ticket == ex->fail_submit_at) {
    g_err = "fx_submit: injected failure";
    return 0;
  }
  const uint64_t t = ex->next_ticket++;
  ex->submitted++;
  ex->last = *j;
  const int L = j->n_hops;
  const int64_t bs = j->batch_size;
  const int64_t* seeds = j->seeds_host ? j->seeds_host : j->seeds_dev;
  if (j->seeds_host && bs) memcpy(j->seeds_dev, j->seeds_host, (size_t)bs * 8); /* the H2D copy */
  const int64_t seed0 = bs ? seeds[0] : 0;
  int64_t* m = j->meta_host;
  memset(m, 0, sizeof(int64_t) * (SPP_META_WORDS + SPP_MAX_PARTS + 2));
  int64_t T = bs;
  for (int h = 0; h < L; ++h) {
    int64_t E = 2 * T + h;
    if (E > j->out_col_cap[h]) E = j->out_col_cap[h];
    m[SPP_META_NODES(h)] = T;
    m[SPP_META_EDGES(h)] = E;
    for (int64_t i = 0; i <= T; ++i) j->out_rowptr[h][i] = 1000 * (h + 1) + i;
    for (int64_t q = 0; q < E; ++q) j->out_col[h][q] = seed0 + 10 * (h + 1) + q;
    T += E / 2;
  }
  const int64_t nb = T;
  m[SPP_META_NODES(L)] = nb;
  if (ex->overflow_at && (int)t == ex->overflow_at) m[SPP_META_OVERFLOW] = 1;
  if (j->n_id_out)
    for (int64_t i = 0; i < nb; ++i) j->n_id_out[i] = seed0 + i;
  if (j->x_out && j->feature_mode)
    for (int64_t i = 0; i < nb; ++i) memset((char*)j->x_out + i * j->row_bytes, (int)((seed0 + i) & 0xff), (size_t)j->row_bytes);
  if (j->y_out && j->y_row_bytes == 8)
    for (int64_t i = 0; i < bs; ++i) ((int64_t*)j->y_out)[i] = 3 * seeds[i];
  if (j->do_split) {
    const int P = j->fmap.num_parts;
    int64_t* c = m + SPP_META_WORDS;
    int64_t used = 0;
    for (int p = 0; p < P; ++p) {
      c[p] = nb / (P + 1);
      used += c[p];
    }
    c[P] = nb - used;
    c[P + 1] = nb;
    for (int64_t pos = 0; pos < nb; ++pos) j->bucket_ids[pos] = 7 * pos + seed0;
    for (int64_t i = 0; i < nb; ++i) j->perm[i] = nb - 1 - i;
  }
  fx_ticket* k = &ex->ring[t % FX_RING];
  k->id = t;
  k->polls = 0;
  k->waited = 0;
  return t;
}

int fx_poll(void* e, uint64_t t) {
  fx_executor* ex = (fx_executor*)e;
  fx_ticket* k = &ex->ring[t % FX_RING];
  if (k->id != t) {
    g_err = "fx_poll: unknown ticket";
    return -1;
  }
  if (k->waited) return 1;
  return ++k->polls > ex->polls_needed ? 1 : 0;
}

int fx_wait(void* e, uint64_t t) {
  fx_executor* ex = (fx_executor*)e;
  fx_ticket* k = &ex->ring[t % FX_RING];
  if (k->id != t) {
    g_err = "fx_wait: unknown ticket";
    return -1;
  }
  k->waited = 1;
  ex->waits++;
  return 0;
}
