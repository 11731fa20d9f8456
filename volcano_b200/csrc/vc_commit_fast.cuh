// vc_commit_fast.cuh — K2f, the incremental variant of the persistent commit kernel.
//
// Same exact semantics as k_commit (vc_commit.cuh), for the common session shape: no Releasing / Pipelined
// resources at open (FutureIdle == Idle, so only the idle gradient exists), no normalising batch scorer,
// R <= 8. It exploits the one structural fact of the greedy loop: a placement changes ONE node, so between
// two consecutive tasks with the same (class, request) record every other (task, node) verdict and score
// is unchanged. Per CTA it keeps, for the group being placed,
//     c_cat[i], c_score[i]      verdict + total score of each of its nodes          (shared memory)
//     sl_score/node/cnt[cta]    every CTA's current best (score, node) + #candidates (shared memory)
// and then a step is:
//     group changed  -> full sweep: every thread re-evaluates its node, all-gather of the CTA bests
//     same group     -> the CTA that owns the node changed by the previous placement re-evaluates that one
//                       node, rescans its cache and PUBLISHES one 16-byte record into a ring in L2; every
//                       other CTA reads that single record. The owner never waits for anybody, so a run of
//                       placements inside one CTA proceeds at shared-memory speed and the L2 round trip is
//                       only paid when the winner moves to another CTA.
// Control state is replicated exactly as in k_commit; only warp 0 of each CTA runs it, the other warps
// sleep on the block barrier and serve full sweeps / rollbacks on command.
#pragma once
#include "vc_commit.cuh"

#define RING_DEPTH 1024
#define RING_STRIDE 4  // uint4 per ring entry (64 bytes: one entry per cache line)
#define CMD_SWEEP 1
#define CMD_DISCARD 2
#define CMD_EXIT 3
#define VC_JOBX_PURE 0x100u  // host-computed: every named role of the job maps to a single group

struct FastSmem {
  double *alloc, *idle, *used, *kalloc, *kreq, *knz;
  int32_t *max_tasks, *pod_count, *nerr_stamp, *c_cat;
  unsigned long long *nerr;
  double *c_score;
  double *sl_score;
  int32_t *sl_node, *sl_cnt;
  int cap;
};
struct FastNodeView {
  const FastSmem &s;
  int i;
  __device__ __forceinline__ double alloc(int d) const { return s.alloc[d * s.cap + i]; }
  __device__ __forceinline__ double idle(int d) const { return s.idle[d * s.cap + i]; }
  __device__ __forceinline__ double used(int d) const { return s.used[d * s.cap + i]; }
  __device__ __forceinline__ double kalloc(int k) const { return s.kalloc[k * s.cap + i]; }
  __device__ __forceinline__ double kreq(int k) const { return s.kreq[k * s.cap + i]; }
  __device__ __forceinline__ double knz(int k) const { return s.knz[k * s.cap + i]; }
};

struct Best {
  double score;
  int node;
  int cnt;
};
__device__ __forceinline__ void best_fold(Best &a, double s, int n, int cnt) {
  if (n >= 0 && (a.node < 0 || better(s, n, a.score, a.node))) { a.score = s; a.node = n; }
  a.cnt += cnt;
}
__device__ __forceinline__ void best_warp_reduce(Best &b) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    double os = __shfl_xor_sync(0xffffffffu, b.score, o);
    int on = __shfl_xor_sync(0xffffffffu, b.node, o);
    int oc = __shfl_xor_sync(0xffffffffu, b.cnt, o);
    best_fold(b, os, on, oc);
  }
}
__device__ __forceinline__ uint4 pack_best(const Best &b, unsigned tag) {
  unsigned long long sb = (unsigned long long)__double_as_longlong(b.score);
  return make_uint4((unsigned)sb, (unsigned)(sb >> 32), (unsigned)b.node, (tag << 2) | (unsigned)min(b.cnt, 2));
}
__device__ __forceinline__ Best unpack_best(const uint4 &v) {
  Best b;
  b.score = __longlong_as_double((long long)((unsigned long long)v.x | ((unsigned long long)v.y << 32)));
  b.node = (int)v.z;
  b.cnt = (int)(v.w & 3u);
  return b;
}

// all-gather of the CTA bests (warp 0 of every CTA); fills the slot table
__device__ __forceinline__ void exchange_all_fast(const K2Params &p, const Best &mine, unsigned ag, FastSmem &fs) {
  const int lane = threadIdx.x & 31;
  const int G = p.n_cta;
  const unsigned tag = (ag + 1u) & 0x3fffffffu;
  uint4 *base = p.mbox + (size_t)(ag & 1u) * G * MBOX_STRIDE;
  if (lane == 0) mbox_store(base + (size_t)blockIdx.x * MBOX_STRIDE, pack_best(mine, tag));
  for (int s0 = 0; s0 < G; s0 += 32 * 4) {
    uint4 a[4];
    bool need[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) need[k] = (s0 + k * 32 + lane) < G;
    bool pending;
    do {
      pending = false;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (need[k]) a[k] = mbox_load(base + (size_t)(s0 + k * 32 + lane) * MBOX_STRIDE);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (!need[k]) continue;
        if ((a[k].w >> 2) != tag) { pending = true; continue; }
        const int s = s0 + k * 32 + lane;
        Best b = unpack_best(a[k]);
        fs.sl_score[s] = b.score; fs.sl_node[s] = b.node; fs.sl_cnt[s] = b.cnt;
        need[k] = false;
      }
    } while (pending);
  }
  __syncwarp();
}

// best of this CTA from its verdict/score cache (warp 0)
__device__ __forceinline__ Best scan_cache(const FastSmem &fs, int nmine, int nbase) {
  const int lane = threadIdx.x & 31;
  Best b{0.0, -1, 0};
  for (int i = lane; i < nmine; i += 32)
    if (fs.c_cat[i] == 0) best_fold(b, fs.c_score[i], nbase + i, 1);
  best_warp_reduce(b);
  return b;
}

__global__ void __launch_bounds__(256, 1) k_commit_fast(K2Params p) {
  const DevConf &c = p.c;
  const int R = p.d.R, K = p.d.K, N = p.d.N, J = p.d.J, Q = p.d.Q, NR = p.d.NR, T = p.d.T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int cta = blockIdx.x;
  const int nbase = p.d.node_begin + cta * p.npc;
  const int nmine = max(0, min(p.npc, p.d.node_end - nbase));
  const int cap = p.npc;
  const int G = p.n_cta;

  // ---- carve shared memory ----
  unsigned char *sp = k2_smem;
  Ctl &S = *reinterpret_cast<Ctl *>(sp);
  sp += (sizeof(Ctl) + 15) & ~(size_t)15;
  FastSmem fs;
  fs.cap = cap;
  auto take = [&](int rows) { double *q = reinterpret_cast<double *>(sp); sp += (size_t)rows * cap * sizeof(double); return q; };
  fs.alloc = take(R); fs.idle = take(R); fs.used = take(R);
  fs.kalloc = take(K); fs.kreq = take(K); fs.knz = take(2);
  fs.c_score = take(1);
  fs.nerr = reinterpret_cast<unsigned long long *>(sp); sp += (size_t)cap * 8;
  fs.sl_score = reinterpret_cast<double *>(sp); sp += (size_t)G * 8;
  fs.max_tasks = reinterpret_cast<int32_t *>(sp); sp += (size_t)cap * 4;
  fs.pod_count = reinterpret_cast<int32_t *>(sp); sp += (size_t)cap * 4;
  fs.nerr_stamp = reinterpret_cast<int32_t *>(sp); sp += (size_t)cap * 4;
  fs.c_cat = reinterpret_cast<int32_t *>(sp); sp += (size_t)cap * 4;
  fs.sl_node = reinterpret_cast<int32_t *>(sp); sp += (size_t)G * 4;
  fs.sl_cnt = reinterpret_cast<int32_t *>(sp); sp += (size_t)G * 4;

  for (int i = tid; i < nmine; i += blockDim.x) {
    const int n = nbase + i;
    for (int d = 0; d < R; ++d) {
      fs.alloc[d * cap + i] = p.alloc[(size_t)d * N + n];
      fs.idle[d * cap + i] = p.idle[(size_t)d * N + n];
      fs.used[d * cap + i] = p.used[(size_t)d * N + n];
    }
    for (int k = 0; k < K; ++k) {
      fs.kalloc[k * cap + i] = p.kalloc[(size_t)k * N + n];
      fs.kreq[k * cap + i] = p.kreq[(size_t)k * N + n];
    }
    for (int k = 0; k < 2; ++k) fs.knz[k * cap + i] = p.knz[(size_t)k * N + n];
    fs.max_tasks[i] = p.max_tasks[n];
    fs.pod_count[i] = p.pod_count[n];
    fs.nerr[i] = 0ull;
    fs.nerr_stamp[i] = -1;
    fs.c_cat[i] = 2;
    fs.c_score[i] = 0.0;
  }
  for (int s = tid; s < G; s += blockDim.x) { fs.sl_score[s] = 0.0; fs.sl_node[s] = -1; fs.sl_cnt[s] = 0; }

  // ---- per-CTA replica of the control state (same layout as k_commit) ----
  int32_t *ri = p.rep_i32 + (size_t)cta * p.rep_i32_stride;
  double *rf = p.rep_f64 + (size_t)cta * p.rep_f64_stride;
  HeapEnt *heap = p.rep_heap + (size_t)cta * p.rep_heap_stride;
  int32_t *j_ready = ri; ri += J;
  int32_t *j_waiting = ri; ri += J;
  int32_t *j_cursor = ri; ri += J;
  int32_t *r_occ = ri; ri += NR;
  int32_t *r_pip = ri; ri += NR;
  int32_t *r_pending = ri; ri += NR;
  int32_t *r_failed = ri; ri += NR;
  int32_t *q_active = ri; ri += Q;
  int32_t *q_scursor = ri; ri += Q;
  int32_t *q_hsize = ri; ri += Q;
  uint32_t *q_alloc_has = reinterpret_cast<uint32_t *>(ri); ri += Q;
  uint32_t *q_flags2 = reinterpret_cast<uint32_t *>(ri); ri += Q;
  int32_t *ops = ri; ri += (size_t)p.max_job_tasks * 3;
  double *j_share = rf; rf += J;
  double *j_alloc = rf; rf += (size_t)R * J;
  double *q_alloc = rf; rf += (size_t)R * Q;
  double *q_share = rf; rf += Q;
  double *ops_score = rf; rf += p.max_job_tasks;

  for (int j = tid; j < J; j += blockDim.x) {
    j_ready[j] = p.j_ready0[j];
    j_waiting[j] = p.j_waiting0[j];
    j_cursor[j] = 0;
    j_share[j] = p.j_share0[j];
    for (int d = 0; d < R; ++d) j_alloc[(size_t)d * J + j] = p.j_alloc0[(size_t)d * J + j];
  }
  for (int r = tid; r < NR; r += blockDim.x) {
    r_occ[r] = p.r_occ0[r]; r_pip[r] = p.r_pip0[r]; r_pending[r] = p.r_pending0[r]; r_failed[r] = 0;
  }
  for (int q = tid; q < Q; q += blockDim.x) {
    q_active[q] = (p.qjobs_off[q + 1] > p.qjobs_off[q]) ? 1 : 0;
    q_scursor[q] = 0;
    q_hsize[q] = 0;
    q_alloc_has[q] = p.q_alloc_has0[q];
    q_flags2[q] = p.q_flags2[q];
    q_share[q] = p.q_share0[q];
    for (int d = 0; d < R; ++d) q_alloc[(size_t)d * Q + q] = p.q_alloc0[(size_t)d * Q + q];
  }
  if (tid == 0) {
    S.seq = 0;
    S.n_dec = S.n_vis = S.n_fit = S.n_steps = 0;
    for (int k = 0; k < 8; ++k) S.prof[k] = 0;
    S.prof_last = clock64();
    S.cmd = 0; S.visit_id = 0; S.cur_group = -1; S.cache_group = -1; S.dirty_node = -1;
    S.ag = 0; S.pc = 0; S.since_sync = 0; S.n_full = 0; S.n_incr = 0;
  }
  __syncthreads();

  // full evaluation of this thread's nodes for the staged group record (command CMD_SWEEP)
  auto sweep_part = [&]() {
    const TaskRec &trec = S.trec;
    const uint32_t *cs_row = p.cstat + (size_t)trec.klass * N + nbase;
    const int rl = S.sweep_rl;
    const bool use_cache = S.sweep_use_cache != 0;
    const int vid = S.visit_id;
    Best b{0.0, -1, 0};
    for (int i = tid; i < nmine; i += blockDim.x) {
      FastNodeView nv{fs, i};
      const uint32_t cs = cs_row[i];
      if (use_cache && fs.nerr_stamp[i] != vid) { fs.nerr[i] = 0ull; fs.nerr_stamp[i] = vid; }
      int cat = 2;
      double sc = 0.0;
      if (!(use_cache && ((fs.nerr[i] >> rl) & 1ull))) {
        cat = eval_pair_fast(c, R, K, trec, nv, cs, c.pred_predicates && fs.max_tasks[i] <= fs.pod_count[i], &sc);
        if (cat == 2 && use_cache) fs.nerr[i] |= (1ull << rl);
      }
      fs.c_cat[i] = cat;
      fs.c_score[i] = sc;
      if (cat == 0) best_fold(b, sc, nbase + i, 1);
    }
    best_warp_reduce(b);
    if (lane == 0) { S.w_score[0][warp] = b.score; S.w_node[0][warp] = b.node; S.w_cnt[0][warp] = b.cnt; }
  };
  // stmt.Discard() for this thread's nodes (command CMD_DISCARD), statement.go:357-381
  auto discard_part = [&]() {
    const int n_ops = S.n_ops;
    for (int k = n_ops - 1; k >= 0; --k) {
      const int ot = ops[k * 3 + 0], on = ops[k * 3 + 1];
      if (on >= nbase && on < nbase + nmine && ((on - nbase) % blockDim.x) == tid) {
        const int i = on - nbase;
        for (int d = 0; d < R; ++d) {
          double rq = p.req[(size_t)d * T + ot];
          fs.idle[d * cap + i] += rq;
          fs.used[d * cap + i] -= rq;
        }
        if (c.has_predicates) {
          fs.pod_count[i] -= 1;
          for (int kk = 0; kk < K; ++kk) fs.kreq[kk * cap + i] -= p.tkreq[(size_t)kk * T + ot];
          for (int kk = 0; kk < 2; ++kk) fs.knz[kk * cap + i] -= p.tknz[(size_t)kk * T + ot];
        }
      }
    }
  };

  if (warp != 0) {
    // ---- worker warps: serve block-wide commands ----
    for (;;) {
      __syncthreads();  // B1: command posted
      const int cmd = S.cmd;
      if (cmd == CMD_EXIT) break;
      if (cmd == CMD_SWEEP) sweep_part();
      else if (cmd == CMD_DISCARD) discard_part();
      __syncthreads();  // B2: command done
    }
  } else {
    // ===================================================================================
    // warp 0: the replicated control program (allocate.go:283-348, :558-694)
    // ===================================================================================
    const bool out_cta = (cta == 0);
    bool qorder_prop = false, overused_prop = false;
    for (int i = 0; i < c.n_plugins; ++i) {
      if ((c.enabled[i] & VC_EN_QUEUE_ORDER) && c.plugin[i] == VC_PLUGIN_PROPORTION) qorder_prop = true;
      if ((c.enabled[i] & VC_EN_OVERUSED) && c.plugin[i] == VC_PLUGIN_PROPORTION) overused_prop = true;
    }
    for (;;) {
      // ---- queues.Pop() ----
      int bq = -1, bprio = 0;
      double bshare = 0.0;
      uint32_t brank = 0;
      for (int q = lane; q < Q; q += 32) {
        if (!q_active[q]) continue;
        int pr = qorder_prop ? p.q_prio[q] : 0;
        double sh = qorder_prop ? q_share[q] : 0.0;
        uint32_t rk = p.q_rank[q];
        bool lt = bq < 0 || pr > bprio || (pr == bprio && (sh < bshare || (sh == bshare && rk < brank)));
        if (lt) { bq = q; bprio = pr; bshare = sh; brank = rk; }
      }
      for (int o = 16; o; o >>= 1) {
        int oq = __shfl_xor_sync(0xffffffffu, bq, o);
        int opr = __shfl_xor_sync(0xffffffffu, bprio, o);
        double osh = __shfl_xor_sync(0xffffffffu, bshare, o);
        uint32_t ork = __shfl_xor_sync(0xffffffffu, brank, o);
        bool lt = oq >= 0 && (bq < 0 || opr > bprio || (opr == bprio && (osh < bshare || (osh == bshare && ork < brank))));
        if (lt) { bq = oq; bprio = opr; bshare = osh; brank = ork; }
      }
      const int q = bq;
      if (q < 0) break;
      __syncwarp();
      // ---- queue attr, ssn.Overused, jobs.Pop(), job state ----
      if (lane == 0) {
        q_active[q] = 0;
        S.queue = q;
        S.qflags = p.q_flags[q];
        S.qflags2 = q_flags2[q];
        S.qalloc_has = q_alloc_has[q];
        S.qdes_has = p.q_des_has[q];
        S.qshare = q_share[q];
        for (int d = 0; d < R; ++d) { S.qalloc[d] = q_alloc[(size_t)d * Q + q]; S.qdes[d] = p.q_des[(size_t)d * Q + q]; }
        bool over = false;
        if (overused_prop && (S.qflags2 & 1u)) {
          over = le_eps(S.qdes[0], S.qalloc[0]) && le_eps(S.qdes[1], S.qalloc[1]);
          for (int d = 2; d < R && over; ++d) {
            if (!(S.qdes_has & (1u << d))) continue;
            double rv = (S.qalloc_has & (1u << d)) ? S.qalloc[d] : 0.0;
            if (!le_eps(S.qdes[d], rv)) over = false;
          }
        }
        int j = -1;
        if (!over) {
          const int sbeg = p.qjobs_off[q], send = p.qjobs_off[q + 1];
          const int sc = sbeg + q_scursor[q];
          HeapEnt *h = heap + sbeg;
          int hs = q_hsize[q];
          bool have_s = sc < send, have_h = hs > 0, take_heap = false;
          if (have_s && have_h) {
            int js = p.qjobs[sc];
            JobKey ks;
            ks.share = j_share[js]; ks.prio = p.j_prio[js]; ks.rank = p.j_rank[js];
            ks.ready = j_ready[js] + p.j_pbe[js] >= p.j_min[js];
            ks.preempt = (p.j_flags[js] & VC_JOB_PREEMPTABLE) != 0;
            take_heap = job_less(c, key_of(h[0]), ks);
          } else if (have_h) {
            take_heap = true;
          }
          if (take_heap) {
            j = h[0].job;
            HeapEnt last = h[--hs];
            q_hsize[q] = hs;
            int i = 0;
            for (;;) {
              int l = 2 * i + 1;
              if (l >= hs) break;
              int m = l;
              if (l + 1 < hs && job_less(c, key_of(h[l + 1]), key_of(h[l]))) m = l + 1;
              if (!job_less(c, key_of(h[m]), key_of(last))) break;
              h[i] = h[m];
              i = m;
            }
            if (hs > 0) h[i] = last;
          } else if (have_s) {
            j = p.qjobs[sc];
            q_scursor[q] += 1;
          }
        }
        S.job = j;
        if (j >= 0) {
          S.cursor = p.job_task_off[j] + j_cursor[j];
          S.task_end = p.job_task_off[j + 1];
          S.ready = j_ready[j]; S.waiting = j_waiting[j]; S.pbe = p.j_pbe[j]; S.minav = p.j_min[j];
          S.ntasks_total = p.j_ntasks[j]; S.taskmintotal = p.j_taskmintotal[j]; S.jflags = p.j_flags[j];
          S.role_base = p.j_roleoff[j];
          S.nroles = p.j_roleoff[j + 1] - p.j_roleoff[j];
          S.jshare = j_share[j];
          for (int d = 0; d < R; ++d) S.jalloc[d] = j_alloc[(size_t)d * J + j];
          for (int r = 0; r < S.nroles; ++r) {
            int gr = S.role_base + r;
            S.r_occ[r] = r_occ[gr]; S.r_pip[r] = r_pip[gr]; S.r_pending[r] = r_pending[gr];
            S.r_min[r] = p.r_min[gr]; S.r_flags[r] = p.r_flags[gr]; S.r_failed[r] = (uint8_t)r_failed[gr];
          }
          S.n_ops = 0;
          S.visit_id += 1;  // util.NewPredicateHelper(): a fresh error cache per visit
        }
      }
      __syncwarp();
      if (S.job < 0) continue;
      const int j = S.job;
      const bool pure = (S.jflags & VC_JOBX_PURE) != 0;
      PROF_MARK(0);

      // ---- allocateResourcesForTasks, allocate.go:558-694 ----
      for (;;) {
        if (S.cursor >= S.task_end) break;
        PROF_MARK(4);
        const int4 meta = p.tmeta[S.cursor];
        const int t = meta.x, grp = meta.y, rl = meta.z - S.role_base;
        __syncwarp();
        if (grp != S.cur_group) {  // stage the group's request record
          if (lane < R) S.trec.req[lane] = p.g_req[(size_t)lane * p.n_groups + grp];
          if (lane >= 16 && lane < 16 + K) S.trec.kreq[lane - 16] = p.g_kreq[(size_t)(lane - 16) * p.n_groups + grp];
          if (lane >= 24 && lane < 26) S.trec.knz[lane - 24] = p.g_knz[(size_t)(lane - 24) * p.n_groups + grp];
          if (lane == 31) { S.trec.has = p.g_has[grp]; S.trec.klass = p.g_class[grp]; }
        }
        if (lane == 0) { S.task = t; S.role_local = rl; S.cursor += 1; S.cur_group = grp; }
        __syncwarp();
        if (!ctl_allocatable(p, S)) continue;
        const bool named_role = !(S.r_flags[rl] & VC_ROLE_EMPTY_NAME);
        if (named_role && S.r_failed[rl]) {
          if (lane == 0) { if (out_cta) p.fit_errors[S.n_fit] = t; S.n_fit += 1; }
          __syncwarp();
          continue;
        }
        // For a 'pure' job the role-level error cache can never change a verdict (same record, node
        // resources only shrink inside a visit), so it is skipped; other jobs take exact full sweeps.
        const bool use_cache = c.enable_ecache && named_role && !pure;
        PROF_MARK(1);

        if (pure && grp == S.cache_group) {
          // -------- incremental step --------
          if (S.dirty_node >= 0) {
            const int dn = S.dirty_node;
            const int o = (dn - p.d.node_begin) / p.npc;
            const unsigned tag = (S.pc + 1u) & 0x3fffffffu;
            uint4 *ent = p.ring + (size_t)(S.pc % RING_DEPTH) * RING_STRIDE;
            Best nb;
            if (o == cta) {
              const int i = dn - nbase;
              FastNodeView nv{fs, i};
              const uint32_t cs = p.cstat[(size_t)S.trec.klass * N + dn];
              double sc = 0.0;
              int cat = eval_pair_fast(c, R, K, S.trec, nv, cs, c.pred_predicates && fs.max_tasks[i] <= fs.pod_count[i], &sc);
              if (lane == 0) { fs.c_cat[i] = cat; fs.c_score[i] = sc; }
              __syncwarp();
              nb = scan_cache(fs, nmine, nbase);
              if (lane == 0) mbox_store(ent, pack_best(nb, tag));
            } else {
              uint4 v;
              do { v = mbox_load(ent); } while ((v.w >> 2) != tag);
              nb = unpack_best(v);
            }
            if (lane == 0) {
              fs.sl_score[o] = nb.score; fs.sl_node[o] = nb.node; fs.sl_cnt[o] = min(nb.cnt, 2);
              S.pc += 1; S.since_sync += 1; S.dirty_node = -1;
            }
            __syncwarp();
          }
          if (S.since_sync >= RING_DEPTH / 2) {  // keep the publication ring from being overrun
            Best mine = scan_cache(fs, nmine, nbase);
            exchange_all_fast(p, mine, S.ag, fs);
            if (lane == 0) { S.ag += 1; S.since_sync = 0; }
            __syncwarp();
          }
          if (lane == 0) S.n_incr += 1;
          PROF_MARK(2);
        } else {
          // -------- full sweep --------
          if (lane == 0) { S.cmd = CMD_SWEEP; S.sweep_rl = rl; S.sweep_use_cache = use_cache ? 1 : 0; }
          __syncthreads();  // B1
          sweep_part();
          __syncthreads();  // B2
          PROF_MARK(2);
          Best mine{0.0, -1, 0};
          if (lane < nwarps) best_fold(mine, S.w_score[0][lane], S.w_node[0][lane], S.w_cnt[0][lane]);
          best_warp_reduce(mine);
          exchange_all_fast(p, mine, S.ag, fs);
          if (lane == 0) {
            S.ag += 1; S.since_sync = 0; S.dirty_node = -1; S.n_full += 1;
            S.cache_group = pure ? grp : -1;  // verdicts taken under an error cache are not reusable
          }
          __syncwarp();
        }
        // ---- global arg-max over the slot table ----
        Best g{0.0, -1, 0};
        for (int s = lane; s < G; s += 32) best_fold(g, fs.sl_score[s], fs.sl_node[s], fs.sl_cnt[s]);
        best_warp_reduce(g);
        if (lane == 0) S.n_steps += 1;
        PROF_MARK(3);

        if (g.cnt == 0) {  // no feasible node, allocate.go:639-659
          if (lane == 0) { if (out_cta) p.fit_errors[S.n_fit] = t; S.n_fit += 1; S.r_failed[rl] = 1; }
          __syncwarp();
          if (ctl_need_continue(S)) continue;
          break;
        }
        const int best = g.node;
        const double score = g.cnt == 1 ? 0.0 : g.score;
        // ---- Statement.Allocate: node.AddTask on the owner CTA (api/node_info.go:435-484) ----
        if (best >= nbase && best < nbase + nmine) {
          const int i = best - nbase;
          if (lane < R) {
            fs.idle[lane * cap + i] -= S.trec.req[lane];
            fs.used[lane * cap + i] += S.trec.req[lane];
          }
          if (c.has_predicates) {  // predicates AllocateFunc, predicates.go:212-256
            if (lane == 16) fs.pod_count[i] += 1;
            if (lane >= 17 && lane < 17 + K) fs.kreq[(lane - 17) * cap + i] += S.trec.kreq[lane - 17];
            if (lane >= 24 && lane < 26) fs.knz[(lane - 24) * cap + i] += S.trec.knz[lane - 24];
          }
        }
        if (lane == 0) {
          const TaskRec &trec = S.trec;
          S.dirty_node = best;
          S.r_pending[rl] -= 1;
          S.r_occ[rl] += 1;
          S.ready += 1;
          if (c.has_drf) {
            for (int d = 0; d < R; ++d) S.jalloc[d] += trec.req[d];
            S.jshare = drf_share(p, S.jalloc);
          }
          if (c.has_proportion && (S.qflags2 & 1u)) {
            S.qalloc[0] += trec.req[0];
            S.qalloc[1] += trec.req[1];
            for (int d = 2; d < R; ++d)
              if (trec.has & (1u << d)) { S.qalloc[d] += trec.req[d]; S.qalloc_has |= 1u << d; S.qflags2 &= ~2u; }
            S.qshare = queue_share(R, S.qalloc, S.qalloc_has, S.qdes, S.qdes_has);
          }
          const int k = S.n_ops;
          ops[k * 3 + 0] = t; ops[k * 3 + 1] = best; ops[k * 3 + 2] = VC_OP_ALLOCATE;
          ops_score[k] = score;
          S.n_ops = k + 1;
        }
        __syncwarp();
        if (ctl_job_ready(c, S)) break;
      }
      PROF_MARK(4);

      // ---- statement outcome, allocate.go:681-693 and :330-337 ----
      const bool ready = ctl_job_ready(c, S);
      const bool stmt = ready || ctl_job_pipelined(c, S);
      const int n_ops = S.n_ops;
      if (!stmt && n_ops > 0) {
        if (lane == 0) S.cmd = CMD_DISCARD;
        __syncthreads();  // B1
        discard_part();
        __syncthreads();  // B2
        if (lane == 0) {
          for (int k = n_ops - 1; k >= 0; --k) {
            const int ot = ops[k * 3 + 0];
            const int orl = p.t_role[ot] - S.role_base;
            S.r_pending[orl] += 1; S.r_occ[orl] -= 1; S.ready -= 1;
            if (c.has_drf)
              for (int d = 0; d < R; ++d) S.jalloc[d] -= p.req[(size_t)d * T + ot];
            if (c.has_proportion && (S.qflags2 & 1u)) {
              const uint32_t oh = p.req_has[ot];
              S.qalloc[0] -= p.req[(size_t)0 * T + ot];
              S.qalloc[1] -= p.req[(size_t)1 * T + ot];
              if (!(S.qflags2 & 2u))
                for (int d = 2; d < R; ++d)
                  if (oh & (1u << d)) { S.qalloc[d] -= p.req[(size_t)d * T + ot]; S.qalloc_has |= 1u << d; }
            }
          }
          if (c.has_drf) S.jshare = drf_share(p, S.jalloc);
          if (c.has_proportion && (S.qflags2 & 1u)) S.qshare = queue_share(R, S.qalloc, S.qalloc_has, S.qdes, S.qdes_has);
          S.cache_group = -1;  // several nodes changed at once: drop the verdict cache
          S.dirty_node = -1;
        }
        __syncwarp();
      }
      // results (CTA 0): decisions copied lane-parallel
      if (out_cta && stmt) {
        for (int k = lane; k < n_ops; k += 32) {
          vc_decision dcs;
          dcs.task = ops[k * 3 + 0]; dcs.node = ops[k * 3 + 1]; dcs.kind = ops[k * 3 + 2];
          dcs.visit = S.n_vis; dcs.score = ops_score[k];
          p.decisions[S.n_dec + k] = dcs;
        }
      }
      __syncwarp();
      if (lane == 0) {
        j_ready[j] = S.ready;
        j_waiting[j] = S.waiting;
        j_cursor[j] = S.cursor - p.job_task_off[j];
        j_share[j] = S.jshare;
        for (int d = 0; d < R; ++d) j_alloc[(size_t)d * J + j] = S.jalloc[d];
        for (int r = 0; r < S.nroles; ++r) {
          int gr = S.role_base + r;
          r_occ[gr] = S.r_occ[r]; r_pip[gr] = S.r_pip[r]; r_pending[gr] = S.r_pending[r]; r_failed[gr] = S.r_failed[r];
        }
        for (int d = 0; d < R; ++d) q_alloc[(size_t)d * Q + q] = S.qalloc[d];
        q_alloc_has[q] = S.qalloc_has;
        q_flags2[q] = S.qflags2;
        q_share[q] = S.qshare;
        if (out_cta) {
          vc_visit v;
          v.job = j;
          v.outcome = stmt ? (ready ? VC_VISIT_COMMIT : VC_VISIT_KEEP) : VC_VISIT_DISCARD;
          v.first_op = S.n_dec;
          v.n_ops = stmt ? n_ops : 0;
          p.visits[S.n_vis] = v;
        }
        if (stmt) S.n_dec += n_ops;
        S.n_vis += 1;
        if (stmt && ready && S.cursor < S.task_end) {  // jobs.Push(job), allocate.go:334-336
          HeapEnt e;
          e.share = S.jshare; e.job = j; e.prio = p.j_prio[j]; e.rank = p.j_rank[j];
          e.bits = (ctl_is_ready(S) ? 1u : 0u) | ((S.jflags & VC_JOB_PREEMPTABLE) ? 2u : 0u);
          HeapEnt *h = heap + p.qjobs_off[q];
          int i = q_hsize[q]++;
          while (i > 0) {
            int par = (i - 1) / 2;
            if (!job_less(c, key_of(e), key_of(h[par]))) break;
            h[i] = h[par];
            i = par;
          }
          h[i] = e;
        }
        q_active[q] = 1;  // queues.Push(queue), allocate.go:346
      }
      __syncwarp();
      PROF_MARK(0);
    }
    if (lane == 0) S.cmd = CMD_EXIT;
    __syncthreads();  // B1 of the exit command
  }

  // ---- epilogue: node state back to HBM, counters ----
  __syncthreads();
  for (int i = tid; i < nmine; i += blockDim.x) {
    const int n = nbase + i;
    for (int d = 0; d < R; ++d) {
      p.idle[(size_t)d * N + n] = fs.idle[d * cap + i];
      p.used[(size_t)d * N + n] = fs.used[d * cap + i];
    }
    for (int k = 0; k < K; ++k) p.kreq[(size_t)k * N + n] = fs.kreq[k * cap + i];
    for (int k = 0; k < 2; ++k) p.knz[(size_t)k * N + n] = fs.knz[k * cap + i];
    p.pod_count[n] = fs.pod_count[i];
  }
  if (cta == 0 && tid == 0) {
    p.counters[0] = S.n_dec;
    p.counters[1] = S.n_vis;
    p.counters[2] = S.n_fit;
    p.counters[3] = S.n_steps;
    p.counters[5] = S.n_full;
    p.counters[6] = S.n_incr;
    for (int k = 0; k < 8; ++k) p.prof[k] = S.prof[k];
  }
}
