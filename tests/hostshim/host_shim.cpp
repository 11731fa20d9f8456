// Test shim: compiles the PRODUCT's host-side session-open logic (volcano_b200/csrc/vc_host.hpp, plain C++17, no CUDA)
// into a small shared object so that `pytest -m "not gpu"` exercises it on the CPU (the GPU box runs the same code
// inside libvcalloc.so). Not part of the product; nothing here is an alternative compute path.
#include "../../volcano_b200/csrc/vc_host.hpp"

extern "C" {

double vh_go_pow_uint(double x, unsigned n) { return vch::go_pow_uint(x, n); }

int32_t vh_num_feasible_nodes_to_find(int32_t n, int32_t pct, int32_t min_nodes, int32_t min_pct) {
  return vch::num_feasible_nodes_to_find(n, pct, min_nodes, min_pct);
}

int vh_job_valid(const vc_conf *c, const vc_jobs *jb, int j) { return vch::job_valid(*c, *jb, j) ? 1 : 0; }

// proportion OnSessionOpen on ssn.TotalResource = sum of node allocatable: deserved [R][Q] and share [Q]
void vh_proportion_open(const vc_dims *d, const vc_nodes *nd, const vc_jobs *jb, const vc_queues *qu, double *deserved_out,
                        double *share_out) {
  const int R = d->n_dims, N = d->n_nodes, Q = d->n_queues;
  vch::HRes total;
  for (int k = 0; k < R; ++k) {
    double acc = 0;
    for (int n = 0; n < N; ++n) acc += nd->allocatable[(size_t)k * N + n];
    total.v[k] = acc;
    if (k >= 2 && N > 0) { total.has |= 1u << k; total.nil = false; }
  }
  std::vector<vch::QAttr> qa;
  vch::proportion_open(*d, *jb, *qu, total, qa);
  for (int q = 0; q < Q; ++q) {
    for (int k = 0; k < R; ++k)
      deserved_out[(size_t)k * Q + q] = (k < 2 || qa[q].deserved.k(k)) ? qa[q].deserved.v[k] : 0.0;
    share_out[q] = qa[q].share;
  }
}

// pop order of a util.PriorityQueue over `items` (pushed in the given order) under ssn.TaskOrderFn
void vh_task_heap_order(const vc_tasks *tk, int by_priority, int32_t *items, int n) {
  std::vector<int> v(items, items + n);
  vch::go_heap_order(v, vch::TaskLess{tk, by_priority != 0});
  for (int i = 0; i < n; ++i) items[i] = v[i];
}

int vh_less_equal_zero(const double *l, uint32_t l_has, const double *r, uint32_t r_has, int R) {
  vch::HRes a = vch::HRes::load(l, 1, 0, R, l_has), b = vch::HRes::load(r, 1, 0, R, r_has);
  return a.less_equal_zero(b, R) ? 1 : 0;
}

void vh_diff_zero(const double *l, uint32_t l_has, const double *r, uint32_t r_has, int R, double *inc_out, uint32_t *inc_has,
                  double *dec_out, uint32_t *dec_has) {
  vch::HRes a = vch::HRes::load(l, 1, 0, R, l_has), b = vch::HRes::load(r, 1, 0, R, r_has), inc, dec;
  vch::hdiff(a, b, inc, dec, R);
  for (int d = 0; d < R; ++d) { inc_out[d] = inc.v[d]; dec_out[d] = dec.v[d]; }
  *inc_has = inc.has; *dec_has = dec.has;
}
void vh_min_dimension(const double *l, uint32_t l_has, const double *r, uint32_t r_has, int R, int infinity, double *out,
                      uint32_t *out_has) {
  vch::HRes a = vch::HRes::load(l, 1, 0, R, l_has), b = vch::HRes::load(r, 1, 0, R, r_has);
  a.min_dim(b, infinity != 0, R);
  for (int d = 0; d < R; ++d) out[d] = a.v[d];
  *out_has = a.has;
}

// pickUpPendingTasks of the backfill action after `ops` (allocate's kept operations): the Keep record is filled exactly
// as vc_snapshot_upload fills it
int vh_backfill_pick(const vc_dims *d, const vc_conf *conf, const vc_nodes *nd, const vc_tasks *tk, const vc_jobs *jb,
                     const vc_queues *qu, const vc_tasks *bt, int n_bf, const vc_decision *ops, int n_ops, int alloc_ran, int32_t *order_out) {
  const size_t R = d->n_dims, K = d->n_kdims, N = d->n_nodes, T = d->n_tasks, J = d->n_jobs, Q = d->n_queues, NR = d->n_roles,
               B = (size_t)n_bf;
  vch::HRes total;
  for (size_t k = 0; k < R; ++k) {
    double acc = 0;
    for (size_t n = 0; n < N; ++n) acc += nd->allocatable[k * N + n];
    total.v[k] = acc;
    if (k >= 2 && N > 0) { total.has |= 1u << k; total.nil = false; }
  }
  const bool has_drf = vch::has_plugin(*conf, VC_PLUGIN_DRF), has_prop = vch::has_plugin(*conf, VC_PLUGIN_PROPORTION);
  std::vector<vch::QAttr> qattr;
  if (has_prop) vch::proportion_open(*d, *jb, *qu, total, qattr);
  else qattr.assign(Q, vch::QAttr());
  vch::BackfillTasks bf;
  bf.n = n_bf;
  bf.req.assign(bt->resreq, bt->resreq + R * B); bf.kreq.assign(bt->k8s_req, bt->k8s_req + K * B);
  bf.knz.assign(bt->k8s_nonzero_req, bt->k8s_nonzero_req + 2 * B);
  bf.has.assign(bt->req_has, bt->req_has + B); bf.uid.assign(bt->uid_rank, bt->uid_rank + B);
  bf.job.assign(bt->job, bt->job + B); bf.klass.assign(bt->klass, bt->klass + B); bf.role.assign(bt->role, bt->role + B);
  bf.prio.assign(bt->priority, bt->priority + B); bf.podidx.assign(bt->pod_index, bt->pod_index + B);
  bf.ts.assign(bt->creation_ts, bt->creation_ts + B);
  vch::BackfillKeep k;
  k.j_queue.assign(jb->queue, jb->queue + J); k.j_min.assign(jb->min_available, jb->min_available + J);
  k.j_prio.assign(jb->priority, jb->priority + J); k.j_ready0.assign(jb->ready_num, jb->ready_num + J);
  k.j_pbe.assign(jb->pending_besteffort, jb->pending_besteffort + J);
  k.j_taskmintotal.assign(jb->task_min_total, jb->task_min_total + J); k.j_roleoff.assign(jb->role_off, jb->role_off + J + 1);
  k.r_min.assign(jb->role_min, jb->role_min + NR); k.r_occ0.assign(jb->role_occupied, jb->role_occupied + NR);
  k.r_flags.assign(jb->role_flags, jb->role_flags + NR);
  k.q_prio.assign(qu->priority, qu->priority + Q); k.t_role.assign(tk->role, tk->role + T);
  k.j_flags.assign(jb->flags, jb->flags + J);
  vch::rank_by(jb->creation_ts, jb->uid_rank, J, k.j_rank);
  vch::rank_by(qu->creation_ts, qu->uid_rank, Q, k.q_rank);
  k.j_valid.resize(J);
  for (size_t j = 0; j < J; ++j) k.j_valid[j] = vch::job_valid(*conf, *jb, (int)j) ? 1 : 0;
  k.j_alloc0.assign(jb->allocated, jb->allocated + R * J);
  vch::BackfillPick p = vch::backfill_pick(*conf, alloc_ran != 0, R, T, J, Q, B, has_drf, has_prop, total.v, total.has, bf, k, ops, (size_t)n_ops,
                                           tk->job, tk->resreq, tk->req_has, qattr);
  for (size_t i = 0; i < p.order.size(); ++i) order_out[i] = p.order[i];
  return (int)p.order.size();
}

// preconditions of the run-length batches (vch::runs_exact, vch::future_rows_exact): 1 = m placements leave exactly row -/+ m * request
int vh_runs_exact(int D, int N, int T, const double *row_a, const double *row_b, const double *req) {
  return vch::runs_exact((size_t)D, (size_t)N, (size_t)T, {row_a, row_b}, {req}) ? 1 : 0;
}
int vh_future_rows_exact(int D, int N, int T, const double *idle, const double *rel, const double *pip, const double *req) {
  return vch::future_rows_exact((size_t)D, (size_t)N, (size_t)T, idle, rel, pip, req) ? 1 : 0;
}

}  // extern "C"
