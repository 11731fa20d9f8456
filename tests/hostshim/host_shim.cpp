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

}  // extern "C"
