// vc_commit_fast.cuh — K2f, the incremental variant of the persistent commit kernel.
//
// Same exact semantics as k_commit (vc_commit.cuh), for the common session shape: no Releasing / Pipelined
// resources at open (FutureIdle == Idle, so only the idle gradient exists), no normalising batch scorer,
// R <= 8. It exploits the one structural fact of the greedy loop: a placement changes ONE node, so between
// two consecutive tasks with the same (class, request) record every other (task, node) verdict and score
// is unchanged. Per CTA it keeps, for the group being placed,
//     c_cat[i], c_score[i], c_cs[i]   verdict, total score, static word of each of its nodes (shared memory)
//     sl_score/node/cnt[cta]          every CTA's current best (score, node) + #candidates   (shared memory)
//     cta best / global best          maintained incrementally, rescanned only when the holder got worse
// and then a step is:
//     group changed  -> full sweep: every thread re-evaluates its node, all-gather of the CTA bests
//     same group     -> the CTA that owns the node changed by the previous placement re-evaluates that one
//                       node and PUBLISHES one 16-byte record into a ring in L2; every other CTA reads that
//                       single record. The owner never waits for anybody, so a run of placements inside one
//                       CTA proceeds at shared-memory speed and the L2 round trip is only paid when the
//                       winner moves to another CTA.
// Control state is replicated exactly as in k_commit, packed in per-job / per-queue records so that one
// visit costs a handful of L2 round trips (lanes of warp 0 load a record in one coalesced access). Only
// warp 0 of each CTA runs the control program; the other warps sleep on the block barrier and serve full
// sweeps / rollbacks on command.
#pragma once
#include "vc_commit.cuh"

#define RING_DEPTH 1024
#define RING_STRIDE 4  // uint4 per ring entry (64 bytes: one entry per cache line)
#define CMD_SWEEP 1
#define CMD_DISCARD 2
#define CMD_EXIT 3
#define VC_JOBX_PURE 0x100u  // host-computed: every named role of the job maps to a single group
#define FAST_R 8

struct JobStatic {  // 64 B, read-only
  int32_t min_available, n_tasks_total, pending_besteffort, task_min_total;
  int32_t role_off, n_roles, priority;
  uint32_t flags;
  uint32_t rank;
  int32_t task_off, task_end, queue;
  int32_t pad[4];
};
struct JobDyn {  // 96 B, per-CTA replica
  int32_t ready, waiting, cursor, pad;
  double share;
  double alloc[FAST_R];
  double pad2;
};
struct RoleStatic { int32_t min; uint32_t flags; };
struct RoleDyn { int32_t occ, pip, pending, failed; };
struct QueueStatic {  // 80 B
  int32_t prio;
  uint32_t rank, flags, des_has;
  double des[FAST_R];
};
struct QueueDyn {  // 96 B
  uint32_t alloc_has, flags2;
  int32_t active, scursor, hsize, pad;
  double share;
  double alloc[FAST_R];
};
static_assert(sizeof(JobStatic) == 64 && sizeof(JobDyn) == 96 && sizeof(QueueStatic) == 80 && sizeof(QueueDyn) == 96, "record layout");

struct FastParams {  // extra kernel arguments of the fast kernel
  const JobStatic *jstat;
  const RoleStatic *rstat;
  const QueueStatic *qstat;
  const double *q_share0;  // [Q]
};

struct FastSmem {
  double *alloc, *idle, *used, *kalloc, *kreq, *knz;
  int32_t *max_tasks, *pod_count, *nerr_stamp, *c_cat;
  uint32_t *c_cs;
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

// best of this CTA from its verdict/score cache (warp 0); cnt = exact number of candidates
__device__ __forceinline__ Best scan_cache(const FastSmem &fs, int nmine, int nbase) {
  const int lane = threadIdx.x & 31;
  Best b{0.0, -1, 0};
  for (int i = lane; i < nmine; i += 32)
    if (fs.c_cat[i] == 0) best_fold(b, fs.c_score[i], nbase + i, 1);
  best_warp_reduce(b);
  return b;
}
// arg-max over the slot table (warp 0)
__device__ __forceinline__ Best fold_slots(const FastSmem &fs, int G) {
  const int lane = threadIdx.x & 31;
  Best g{0.0, -1, 0};
  for (int s = lane; s < G; s += 32) best_fold(g, fs.sl_score[s], fs.sl_node[s], fs.sl_cnt[s]);
  best_warp_reduce(g);
  return g;
}

// lanes of warp 0 copy a record of `words` 32-bit words from global to shared memory in one access
__device__ __forceinline__ void load_record(void *dst_smem, const void *src_global, int words, bool read_only) {
  const int lane = threadIdx.x & 31;
  for (int w = lane; w < words; w += 32) {
    const int *src = reinterpret_cast<const int *>(src_global) + w;
    reinterpret_cast<int *>(dst_smem)[w] = read_only ? __ldg(src) : *src;
  }
}

struct CtlFast {  // shared-memory state of the fast kernel next to Ctl
  JobStatic js;
  JobDyn jd;
  QueueStatic qs;
  QueueDyn qd;
  RoleStatic rs[VC_MAX_JOB_ROLES];
  RoleDyn rd[VC_MAX_JOB_ROLES];
  double cta_best_score, g_best_score;
  int cta_best_node, cta_cnt, g_best_node, g_cnt;
};

__global__ void __launch_bounds__(256, 1) k_commit_fast(K2Params p, FastParams fp) {
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
  CtlFast &F = *reinterpret_cast<CtlFast *>(sp);
  sp += (sizeof(CtlFast) + 15) & ~(size_t)15;
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
  fs.c_cs = reinterpret_cast<uint32_t *>(sp); sp += (size_t)cap * 4;
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
    fs.c_cs[i] = 0u;
  }
  for (int s = tid; s < G; s += blockDim.x) { fs.sl_score[s] = 0.0; fs.sl_node[s] = -1; fs.sl_cnt[s] = 0; }

  // ---- per-CTA replica of the mutable control state, packed records ----
  unsigned char *rb = reinterpret_cast<unsigned char *>(p.rep_f64 + (size_t)cta * p.rep_f64_stride);
  JobDyn *jdyn = reinterpret_cast<JobDyn *>(rb); rb += (size_t)J * sizeof(JobDyn);
  QueueDyn *qdyn = reinterpret_cast<QueueDyn *>(rb); rb += (size_t)Q * sizeof(QueueDyn);
  RoleDyn *rdyn = reinterpret_cast<RoleDyn *>(rb); rb += (size_t)NR * sizeof(RoleDyn);
  double *ops_score = reinterpret_cast<double *>(rb);
  int32_t *ops = p.rep_i32 + (size_t)cta * p.rep_i32_stride;  // task, node, kind
  HeapEnt *heap = p.rep_heap + (size_t)cta * p.rep_heap_stride;

  for (int j = tid; j < J; j += blockDim.x) {
    JobDyn jd;
    jd.ready = p.j_ready0[j]; jd.waiting = p.j_waiting0[j]; jd.cursor = 0; jd.pad = 0;
    jd.share = p.j_share0[j];
    for (int d = 0; d < FAST_R; ++d) jd.alloc[d] = d < R ? p.j_alloc0[(size_t)d * J + j] : 0.0;
    jd.pad2 = 0.0;
    jdyn[j] = jd;
  }
  for (int r = tid; r < NR; r += blockDim.x) {
    RoleDyn rd;
    rd.occ = p.r_occ0[r]; rd.pip = p.r_pip0[r]; rd.pending = p.r_pending0[r]; rd.failed = 0;
    rdyn[r] = rd;
  }
  for (int q = tid; q < Q; q += blockDim.x) {
    QueueDyn qd;
    qd.alloc_has = p.q_alloc_has0[q]; qd.flags2 = p.q_flags2[q];
    qd.active = (p.qjobs_off[q + 1] > p.qjobs_off[q]) ? 1 : 0;  // buildAllocateContext: queues with a job
    qd.scursor = 0; qd.hsize = 0; qd.pad = 0;
    qd.share = fp.q_share0[q];
    for (int d = 0; d < FAST_R; ++d) qd.alloc[d] = d < R ? p.q_alloc0[(size_t)d * Q + q] : 0.0;
    qdyn[q] = qd;
  }
  if (tid == 0) {
    S.seq = 0;
    S.n_dec = S.n_vis = S.n_fit = S.n_steps = 0;
    for (int k = 0; k < 8; ++k) S.prof[k] = 0;
    S.prof_last = clock64();
    S.cmd = 0; S.visit_id = 0; S.cur_group = -1; S.cache_group = -1; S.dirty_node = -1;
    S.ag = 0; S.pc = 0; S.since_sync = 0; S.n_full = 0; S.n_incr = 0;
    F.cta_best_node = -1; F.cta_cnt = 0; F.g_best_node = -1; F.g_cnt = 0; F.cta_best_score = F.g_best_score = 0.0;
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
      const uint32_t cs = __ldg(cs_row + i);
      if (use_cache && fs.nerr_stamp[i] != vid) { fs.nerr[i] = 0ull; fs.nerr_stamp[i] = vid; }
      int cat = 2;
      double sc = 0.0;
      if (!(use_cache && ((fs.nerr[i] >> rl) & 1ull))) {
        cat = eval_pair_fast(c, R, K, trec, nv, cs, c.pred_predicates && fs.max_tasks[i] <= fs.pod_count[i], &sc);
        if (cat == 2 && use_cache) fs.nerr[i] |= (1ull << rl);
      }
      fs.c_cat[i] = cat;
      fs.c_score[i] = sc;
      fs.c_cs[i] = cs;
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
      // ---- queues.Pop(): arg-min by ssn.QueueOrderFn over the active queues ----
      int bq = -1, bprio = 0;
      double bshare = 0.0;
      uint32_t brank = 0;
      for (int q = lane; q < Q; q += 32) {
        if (!qdyn[q].active) continue;
        int pr = qorder_prop ? __ldg(&fp.qstat[q].prio) : 0;
        double sh = qorder_prop ? qdyn[q].share : 0.0;
        uint32_t rk = __ldg(&fp.qstat[q].rank);
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
      // ---- queue records (one coalesced access each), head of the static job list, heap top ----
      load_record(&F.qs, &fp.qstat[q], sizeof(QueueStatic) / 4, true);
      load_record(&F.qd, &qdyn[q], sizeof(QueueDyn) / 4, false);
      const int sbeg = __ldg(&p.qjobs_off[q]), send = __ldg(&p.qjobs_off[q + 1]);
      __syncwarp();
      // ssn.Overused: attr.deserved.LessEqual(attr.allocated, Zero), proportion.go:319-331
      bool over = false;
      if (overused_prop && (F.qd.flags2 & 1u)) {
        bool ok = true;
        if (lane < R) {
          const int d = lane;
          if (d < 2 || (F.qs.des_has & (1u << d))) {
            double rv = (d < 2 || (F.qd.alloc_has & (1u << d))) ? F.qd.alloc[d] : 0.0;
            ok = le_eps(F.qs.des[d], rv);
          }
        }
        over = __all_sync(0xffffffffu, ok);
      }
      int j = -1;
      bool from_heap = false;
      if (!over) {
        HeapEnt *h = heap + sbeg;
        const int sc = sbeg + F.qd.scursor;
        const int hs = F.qd.hsize;
        const bool have_s = sc < send, have_h = hs > 0;
        int js = have_s ? __ldg(&p.qjobs[sc]) : -1;
        if (have_s) {  // records of the static candidate: needed for the comparison and, if chosen, as the job state
          load_record(&F.js, &fp.jstat[js], sizeof(JobStatic) / 4, true);
          load_record(&F.jd, &jdyn[js], sizeof(JobDyn) / 4, false);
        }
        HeapEnt top;
        if (have_h) top = h[0];
        __syncwarp();
        if (have_s && have_h) {
          JobKey ks;
          ks.share = F.jd.share; ks.prio = F.js.priority; ks.rank = F.js.rank;
          ks.ready = F.jd.ready + F.js.pending_besteffort >= F.js.min_available;
          ks.preempt = (F.js.flags & VC_JOB_PREEMPTABLE) != 0;
          from_heap = job_less(c, key_of(top), ks);
        } else if (have_h) {
          from_heap = true;
        }
        if (from_heap) {
          j = top.job;
          if (lane == 0) {  // heap pop: sift-down (only jobs that were re-pushed live here)
            int n = hs - 1;
            HeapEnt last = h[n];
            int i = 0;
            for (;;) {
              int l = 2 * i + 1;
              if (l >= n) break;
              int m = l;
              if (l + 1 < n && job_less(c, key_of(h[l + 1]), key_of(h[l]))) m = l + 1;
              if (!job_less(c, key_of(h[m]), key_of(last))) break;
              h[i] = h[m];
              i = m;
            }
            if (n > 0) h[i] = last;
            F.qd.hsize = n;
          }
          __syncwarp();
          load_record(&F.js, &fp.jstat[j], sizeof(JobStatic) / 4, true);
          load_record(&F.jd, &jdyn[j], sizeof(JobDyn) / 4, false);
        } else if (have_s) {
          j = js;
          if (lane == 0) F.qd.scursor += 1;
        }
      }
      __syncwarp();
      if (lane == 0) F.qd.active = 0;
      if (j < 0) {  // queue dropped: overused, or no jobs left (allocate.go:295-305)
        __syncwarp();
        if (lane == 0) { qdyn[q].active = 0; qdyn[q].scursor = F.qd.scursor; qdyn[q].hsize = F.qd.hsize; }
        __syncwarp();
        continue;
      }
      // roles of the job
      {
        const int rb0 = F.js.role_off, nr = F.js.n_roles;
        for (int r = lane; r < nr; r += 32) { F.rs[r] = fp.rstat[rb0 + r]; F.rd[r] = rdyn[rb0 + r]; }
      }
      __syncwarp();
      if (lane == 0) {  // unpack into the scalar control block the ctl_* helpers read
        S.queue = q; S.job = j;
        S.qflags = F.qs.flags; S.qflags2 = F.qd.flags2; S.qalloc_has = F.qd.alloc_has; S.qdes_has = F.qs.des_has;
        S.qshare = F.qd.share;
        for (int d = 0; d < R; ++d) { S.qalloc[d] = F.qd.alloc[d]; S.qdes[d] = F.qs.des[d]; }
        S.cursor = F.js.task_off + F.jd.cursor;
        S.task_end = F.js.task_end;
        S.ready = F.jd.ready; S.waiting = F.jd.waiting; S.pbe = F.js.pending_besteffort; S.minav = F.js.min_available;
        S.ntasks_total = F.js.n_tasks_total; S.taskmintotal = F.js.task_min_total; S.jflags = F.js.flags;
        S.role_base = F.js.role_off; S.nroles = F.js.n_roles;
        S.jshare = F.jd.share;
        for (int d = 0; d < R; ++d) S.jalloc[d] = F.jd.alloc[d];
        for (int r = 0; r < S.nroles; ++r) {
          S.r_occ[r] = F.rd[r].occ; S.r_pip[r] = F.rd[r].pip; S.r_pending[r] = F.rd[r].pending;
          S.r_failed[r] = (uint8_t)F.rd[r].failed; S.r_min[r] = F.rs[r].min; S.r_flags[r] = F.rs[r].flags;
        }
        S.n_ops = 0;
        S.visit_id += 1;  // util.NewPredicateHelper(): a fresh error cache per visit
      }
      __syncwarp();
      const bool pure = (S.jflags & VC_JOBX_PURE) != 0;
      int4 meta = __ldg(&p.tmeta[S.cursor]);
      PROF_MARK(0);

      // ---- allocateResourcesForTasks, allocate.go:558-694 ----
      for (;;) {
        if (S.cursor >= S.task_end) break;
        PROF_MARK(4);
        const int t = meta.x, grp = meta.y, rl = meta.z - S.role_base;
        const int next_pos = S.cursor + 1;
        if (next_pos < S.task_end) meta = __ldg(&p.tmeta[next_pos]);  // prefetch the next task's record
        __syncwarp();
        if (grp != S.cur_group) {  // stage the group's request record
          if (lane < R) S.trec.req[lane] = __ldg(&p.g_req[(size_t)lane * p.n_groups + grp]);
          if (lane >= 16 && lane < 16 + K) S.trec.kreq[lane - 16] = __ldg(&p.g_kreq[(size_t)(lane - 16) * p.n_groups + grp]);
          if (lane >= 24 && lane < 26) S.trec.knz[lane - 24] = __ldg(&p.g_knz[(size_t)(lane - 24) * p.n_groups + grp]);
          if (lane == 31) { S.trec.has = __ldg(&p.g_has[grp]); S.trec.klass = __ldg(&p.g_class[grp]); }
        }
        if (lane == 0) { S.task = t; S.role_local = rl; S.cursor = next_pos; S.cur_group = grp; }
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
              double sc = 0.0;
              const int old_cat = fs.c_cat[i];
              int cat = eval_pair_fast(c, R, K, S.trec, nv, fs.c_cs[i], c.pred_predicates && fs.max_tasks[i] <= fs.pod_count[i], &sc);
              __syncwarp();
              if (lane == 0) { fs.c_cat[i] = cat; fs.c_score[i] = sc; }
              // CTA best, incrementally: rescan only when the holder got worse
              int cnt = F.cta_cnt + (cat == 0 ? 1 : 0) - (old_cat == 0 ? 1 : 0);
              bool rescan = false;
              double bs = F.cta_best_score;
              int bn = F.cta_best_node;
              if (bn == dn) {
                if (cat == 0 && sc >= bs) bs = sc;
                else rescan = true;
              } else if (cat == 0 && (bn < 0 || better(sc, dn, bs, bn))) {
                bs = sc; bn = dn;
              }
              __syncwarp();
              if (rescan) { Best r = scan_cache(fs, nmine, nbase); bs = r.score; bn = r.node; cnt = r.cnt; }
              nb.score = bs; nb.node = bn; nb.cnt = cnt;
              if (lane == 0) {
                F.cta_best_score = bs; F.cta_best_node = bn; F.cta_cnt = cnt;
                mbox_store(ent, pack_best(nb, tag));
              }
              nb.cnt = min(cnt, 2);
            } else {
              uint4 v;
              do { v = mbox_load(ent); } while ((v.w >> 2) != tag);
              nb = unpack_best(v);
            }
            // global best, incrementally
            const int old_cnt = fs.sl_cnt[o];
            int gcnt = F.g_cnt + nb.cnt - old_cnt;
            double gs = F.g_best_score;
            int gn = F.g_best_node;
            bool refold = false;
            const int g_owner = gn >= 0 ? (gn - p.d.node_begin) / p.npc : -1;
            if (g_owner == o) {
              if (nb.node >= 0 && (better(nb.score, nb.node, gs, gn) || (nb.node == gn && nb.score == gs))) { gs = nb.score; gn = nb.node; }
              else refold = true;
            } else if (nb.node >= 0 && (gn < 0 || better(nb.score, nb.node, gs, gn))) {
              gs = nb.score; gn = nb.node;
            }
            __syncwarp();
            if (lane == 0) {
              fs.sl_score[o] = nb.score; fs.sl_node[o] = nb.node; fs.sl_cnt[o] = nb.cnt;
              S.pc += 1; S.since_sync += 1; S.dirty_node = -1;
            }
            __syncwarp();
            if (refold) { Best g = fold_slots(fs, G); gs = g.score; gn = g.node; gcnt = g.cnt; }
            if (lane == 0) { F.g_best_score = gs; F.g_best_node = gn; F.g_cnt = gcnt; }
            __syncwarp();
          }
          if (S.since_sync >= RING_DEPTH / 2) {  // keep the publication ring from being overrun
            Best mine{F.cta_best_score, F.cta_best_node, F.cta_cnt};
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
          Best g = fold_slots(fs, G);
          if (lane == 0) {
            S.ag += 1; S.since_sync = 0; S.dirty_node = -1; S.n_full += 1;
            S.cache_group = pure ? grp : -1;  // verdicts taken under an error cache are not reusable
            F.cta_best_score = mine.score; F.cta_best_node = mine.node; F.cta_cnt = mine.cnt;
            F.g_best_score = g.score; F.g_best_node = g.node; F.g_cnt = g.cnt;
          }
          __syncwarp();
        }
        const double g_score = F.g_best_score;
        const int g_node = F.g_best_node, g_cnt = F.g_cnt;
        if (lane == 0) S.n_steps += 1;
        PROF_MARK(3);

        if (g_cnt == 0) {  // no feasible node, allocate.go:639-659
          if (lane == 0) { if (out_cta) p.fit_errors[S.n_fit] = t; S.n_fit += 1; S.r_failed[rl] = 1; }
          __syncwarp();
          if (ctl_need_continue(S)) continue;
          break;
        }
        const int best = g_node;
        const double score = g_cnt == 1 ? 0.0 : g_score;
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
        // event handlers: drf (drf.go:391-418) and proportion (proportion.go:475-497), one lane per dimension
        double new_jshare = 0.0, new_qshare = 0.0;
        if (c.has_drf) {
          double sh = 0.0;
          if (lane < R) {
            const int d = lane;
            const double al = S.jalloc[d] + S.trec.req[d];
            if ((d < 2 || (p.total_has & (1u << d))) && p.total[d] >= VC_MIN_RESOURCE) sh = share_of(al, p.total[d]);
          }
          for (int o = 16; o; o >>= 1) sh = fmax(sh, __shfl_xor_sync(0xffffffffu, sh, o));
          new_jshare = sh;
        }
        const bool prop = c.has_proportion && (S.qflags2 & 1u);
        if (prop) {
          double sh = 0.0;
          if (lane < R) {
            const int d = lane;
            const bool touched = d < 2 || (S.trec.has & (1u << d));
            const bool has = d < 2 || (S.qalloc_has & (1u << d)) || touched;
            const double al = has ? S.qalloc[d] + (touched ? S.trec.req[d] : 0.0) : 0.0;
            if ((d < 2 || (S.qdes_has & (1u << d))) && S.qdes[d] >= VC_MIN_RESOURCE) sh = share_of(al, S.qdes[d]);
          }
          for (int o = 16; o; o >>= 1) sh = fmax(sh, __shfl_xor_sync(0xffffffffu, sh, o));
          new_qshare = sh;
        }
        __syncwarp();
        if (lane == 0) {
          const TaskRec &trec = S.trec;
          S.dirty_node = best;
          S.r_pending[rl] -= 1;
          S.r_occ[rl] += 1;
          S.ready += 1;
          if (c.has_drf) {
            for (int d = 0; d < R; ++d) S.jalloc[d] += trec.req[d];
            S.jshare = new_jshare;
          }
          if (prop) {
            S.qalloc[0] += trec.req[0];
            S.qalloc[1] += trec.req[1];
            for (int d = 2; d < R; ++d)
              if (trec.has & (1u << d)) { S.qalloc[d] += trec.req[d]; S.qalloc_has |= 1u << d; S.qflags2 &= ~2u; }
            S.qshare = new_qshare;
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
      // ---- write the job / queue / role records back into the replica (one coalesced access each) ----
      if (lane == 0) {
        F.jd.ready = S.ready; F.jd.waiting = S.waiting; F.jd.cursor = S.cursor - F.js.task_off; F.jd.share = S.jshare;
        for (int d = 0; d < R; ++d) { F.jd.alloc[d] = S.jalloc[d]; F.qd.alloc[d] = S.qalloc[d]; }
        F.qd.alloc_has = S.qalloc_has; F.qd.flags2 = S.qflags2; F.qd.share = S.qshare;
        F.qd.active = 1;  // queues.Push(queue), allocate.go:346
        for (int r = 0; r < S.nroles; ++r) {
          F.rd[r].occ = S.r_occ[r]; F.rd[r].pip = S.r_pip[r]; F.rd[r].pending = S.r_pending[r]; F.rd[r].failed = S.r_failed[r];
        }
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
          e.share = S.jshare; e.job = j; e.prio = F.js.priority; e.rank = F.js.rank;
          e.bits = (ctl_is_ready(S) ? 1u : 0u) | ((S.jflags & VC_JOB_PREEMPTABLE) ? 2u : 0u);
          HeapEnt *h = heap + sbeg;
          int i = F.qd.hsize++;
          while (i > 0) {
            int par = (i - 1) / 2;
            if (!job_less(c, key_of(e), key_of(h[par]))) break;
            h[i] = h[par];
            i = par;
          }
          h[i] = e;
        }
      }
      __syncwarp();
      for (int w = lane; w < (int)(sizeof(JobDyn) / 4); w += 32) reinterpret_cast<int *>(&jdyn[j])[w] = reinterpret_cast<int *>(&F.jd)[w];
      for (int w = lane; w < (int)(sizeof(QueueDyn) / 4); w += 32) reinterpret_cast<int *>(&qdyn[q])[w] = reinterpret_cast<int *>(&F.qd)[w];
      for (int r = lane; r < S.nroles; r += 32) rdyn[S.role_base + r] = F.rd[r];
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
