/* vcalloc_driver.c — a C consumer of include/vcalloc.h (no Python, no ctypes): loads a session dumped by
 * volcano_b200.snapshot.Snapshot.dump(), drives libvcalloc.so end to end — vc_init, vc_snapshot_create,
 * vc_snapshot_upload, vc_allocate_run, result accessors — and prints the number of decisions / visits / fit errors and
 * an FNV-1a hash over (task, node, kind, visit) of every decision and (job, outcome, first_op, n_ops) of every visit, so
 * that a test can compare it with the same session run through the Python binding.
 *
 *   gcc -O2 -I include tools/cdriver/vcalloc_driver.c -o tools/cdriver/vcalloc_driver -L volcano_b200 -lvcalloc \
 *       -Wl,-rpath,'$ORIGIN/../../volcano_b200'
 *   tools/cdriver/vcalloc_driver session.bin [device]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vcalloc.h"

static void *read_array(FILE *f) {
  uint64_t n = 0;
  if (fread(&n, 8, 1, f) != 1) { fprintf(stderr, "truncated dump\n"); exit(2); }
  void *p = malloc(n ? n : 1);
  if (n && fread(p, 1, n, f) != n) { fprintf(stderr, "truncated dump\n"); exit(2); }
  return p;
}

static uint64_t fnv(uint64_t h, const void *data, size_t n) {
  const unsigned char *b = (const unsigned char *)data;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

#define CHECK(x)                                                                    \
  do {                                                                              \
    int rc_ = (x);                                                                  \
    if (rc_ != VC_OK) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, vc_last_error()); return 1; } \
  } while (0)

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s session.bin [device]\n", argv[0]); return 2; }
  FILE *f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  char magic[8];
  vc_dims dims;
  vc_conf conf;
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "VCSNAP01", 8) != 0) { fprintf(stderr, "not a VCSNAP01 dump\n"); return 2; }
  if (fread(&dims, sizeof dims, 1, f) != 1 || fread(&conf, sizeof conf, 1, f) != 1) { fprintf(stderr, "truncated dump\n"); return 2; }
  if (vc_abi_version() != VC_ABI_VERSION) { fprintf(stderr, "header / library ABI mismatch\n"); return 2; }

  vc_nodes nd;
  nd.allocatable = read_array(f); nd.idle = read_array(f); nd.used = read_array(f); nd.releasing = read_array(f);
  nd.pipelined = read_array(f); nd.k8s_allocatable = read_array(f); nd.k8s_requested = read_array(f);
  nd.k8s_nonzero_requested = read_array(f); nd.max_tasks = read_array(f); nd.pod_count = read_array(f);
  nd.label_bits = read_array(f); nd.taint_hard = read_array(f); nd.taint_soft = read_array(f); nd.flags = read_array(f);
  nd.revocable_zone = read_array(f); nd.zone_active = read_array(f);
  vc_tasks tk;
  tk.resreq = read_array(f); tk.req_has = read_array(f); tk.k8s_req = read_array(f); tk.k8s_nonzero_req = read_array(f);
  tk.job = read_array(f); tk.klass = read_array(f); tk.role = read_array(f); tk.priority = read_array(f);
  tk.pod_index = read_array(f); tk.creation_ts = read_array(f); tk.uid_rank = read_array(f);
  vc_classes cl;
  cl.selector = read_array(f); cl.n_affinity = read_array(f); cl.affinity = read_array(f); cl.tolerated_hard = read_array(f);
  cl.tolerated_soft = read_array(f); cl.n_preferred = read_array(f); cl.preferred = read_array(f);
  cl.preferred_weight = read_array(f); cl.flags = read_array(f);
  vc_jobs jb;
  jb.queue = read_array(f); jb.min_available = read_array(f); jb.priority = read_array(f); jb.creation_ts = read_array(f);
  jb.uid_rank = read_array(f); jb.flags = read_array(f); jb.n_tasks_total = read_array(f); jb.ready_num = read_array(f);
  jb.waiting_num = read_array(f); jb.pending_besteffort = read_array(f); jb.valid_num = read_array(f);
  jb.task_min_total = read_array(f); jb.role_off = read_array(f); jb.allocated = read_array(f); jb.role_min = read_array(f);
  jb.role_occupied = read_array(f); jb.role_pipelined = read_array(f); jb.role_pending_other = read_array(f);
  jb.role_valid = read_array(f); jb.role_flags = read_array(f);
  vc_queues qu;
  qu.weight = read_array(f); qu.priority = read_array(f); qu.creation_ts = read_array(f); qu.uid_rank = read_array(f);
  qu.flags = read_array(f); qu.capability = read_array(f); qu.capability_has = read_array(f); qu.guarantee = read_array(f);
  qu.guarantee_has = read_array(f); qu.allocated = read_array(f); qu.request = read_array(f); qu.request_has = read_array(f);
  qu.allocated_has = read_array(f);
  fclose(f);

  CHECK(vc_init(argc > 2 ? atoi(argv[2]) : 0));
  vc_snapshot *s = NULL;
  CHECK(vc_snapshot_create(&dims, &s));
  CHECK(vc_snapshot_upload(s, &nd, &tk, &cl, &jb, &qu, &conf));
  vc_result *r = NULL;
  CHECK(vc_allocate_run(s, &r));
  const size_t nd_ = vc_result_num_decisions(r), nv = vc_result_num_visits(r), nf = vc_result_num_fit_errors(r);
  const vc_decision *dec = vc_result_decisions(r);
  const vc_visit *vis = vc_result_visits(r);
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < nd_; ++i) {
    const int32_t rec[4] = {dec[i].task, dec[i].node, dec[i].kind, dec[i].visit};
    h = fnv(h, rec, sizeof rec);
  }
  for (size_t i = 0; i < nv; ++i) h = fnv(h, &vis[i], sizeof vis[i]);
  const vc_stats *st = vc_result_stats(r);
  printf("{\"decisions\": %zu, \"visits\": %zu, \"fit_errors\": %zu, \"hash\": \"%016llx\", \"commit_ms\": %.3f, \"upload_ms\": %.3f}\n",
         nd_, nv, nf, (unsigned long long)h, st->commit_ms, st->upload_ms);
  vc_result_free(r);
  vc_snapshot_destroy(s);
  return 0;
}
