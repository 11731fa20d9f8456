// vc_evict.cuh — device side of the preempt and reclaim actions (actions/preempt/preempt.go, actions/reclaim/reclaim.go).
//
// Per preemptor the node-wide part of the action runs here, one thread per node:
//   k_evict_rank   plugin predicates in their preempt reading (ssn.PredicateForPreemptAction, framework/session.go:
//                  679-697: only the static, unresolvable failures reject a node; the pod-count cap does not), the
//                  util.PrioritizeNodes total of the node, and an exact necessary condition of util.ValidateVictims
//                  (util/scheduler_helper.go:313-329): FutureIdle plus the requests of EVERY task the action's filter
//                  lets through must cover the preemptor — any victim set the plugins return is a subset of those.
//   k_evict_pick   preempt: util.SortNodes' first not-yet-tried candidate = arg-max (score, lowest index); reclaim: the
//                  lowest-index candidate (NodeList order, reclaim.go:180).
//   k_evict_apply  Statement.Evict / Statement.Pipeline on the node rows the ranking reads (Releasing, Pipelined, the
//                  upstream NodeInfo sums of the predicates plugin) once the host committed to a node.
// The victim selection on the one node under trial (tier votes, victim queue order, evict-until-fit) is O(tasks on the
// node) and runs in the host control loop (vc_evict.hpp), like the reference's action goroutine.
#pragma once
#include "vc_device.cuh"

#define EV_MODE_PREEMPT_INTER 0  // victims: other jobs of the preemptor's queue (preempt.go:198-214)
#define EV_MODE_PREEMPT_INTRA 1  // victims: the preemptor's own job (preempt.go:252-268)
#define EV_MODE_RECLAIM 2        // victims: Running tasks of other, reclaimable queues (reclaim.go:184-199)
#define EV_MAX_VICTIMS 256

struct EvictTask {  // the preemptor, staged per launch
  TaskRec rec;
  int klass, job, queue, mode;
  int job_prio, task_prio;
  // ssn.Allocatable of the preemptor's queue (proportion queueAllocatable, proportion.go:333-348) as a bound: with
  // quota_on the queue must admit the preemptor once EVERY candidate of the node is gone (they all sit in that queue)
  int quota_on, quota_open;
  uint32_t qalloc_has, qdes_has;
  double qalloc[VC_MAX_DIMS], qdes[VC_MAX_DIMS];
};
#define EV_PICK_K 32  // candidates one k_evict_pick launch hands to the host, in the action's node order
#define EV_CMD_OFF 80  // the apply command starts at this int of the mapped buffer (after the two sets of pick slots)
struct EvictParams {
  DevDims d;
  DevConf c;
  // node rows: Allocatable / upstream allocatable are inputs; the rest are the working copies the actions share
  const double *alloc, *kalloc;
  double *idle, *used, *rel, *pip, *kreq, *knz;
  int32_t *pod_count;
  const uint32_t *cstat;  // [C][N]
  // node.Tasks as CSR over nodes
  const int32_t *rt_off, *rt_idx;  // [N+1], [RT] running-task ids grouped by node, ascending
  const double *rt_req;            // [R][RT]
  const double *rt_kreq, *rt_knz;  // [K][RT]
  const uint32_t *rt_flags;        // [RT] VC_RT_*
  const int32_t *rt_job;           // [RT]
  uint8_t *rt_evicted;             // [RT] 1 = Releasing (evicted in this session)
  const int32_t *j_queue;          // [J]
  const uint32_t *q_flags;         // [Q]
  const int32_t *rt_prio;          // [RT] TaskInfo.Priority
  const int32_t *j_prio, *j_min;   // [J]
  const int32_t *j_ready;          // [J] ReadyTaskNum as the session stands (the host refreshes it after every eviction)
  const uint8_t *q_over;           // [Q] proportion: !allocated.LessEqual(deserved) as the session stands
  int exact_sums;                  // every request is an integer-valued double: sums of them are exact
  int RT;
  // per-preemptor scratch
  unsigned long long *key;  // [N] order-preserving score key of a candidate
  uint8_t *cand;            // [N] 1 = candidate, 2 = tried
  // pick result (mapped pinned memory): EV_PICK_K nodes in the action's order, -1 = no further candidate
  int32_t *pick_node;
  // apply command (mapped pinned memory): [0] node, [1] n_victims, [2..] victim running-task ids
  const int32_t *cmd;
};

struct EvNodeView {
  const EvictParams &p;
  int n;
  __device__ __forceinline__ double alloc(int d) const { return p.alloc[(size_t)d * p.d.N + n]; }
  __device__ __forceinline__ double idle(int d) const { return p.idle[(size_t)d * p.d.N + n]; }
  __device__ __forceinline__ double used(int d) const { return p.used[(size_t)d * p.d.N + n]; }
  __device__ __forceinline__ double rel(int d) const { return p.rel[(size_t)d * p.d.N + n]; }
  __device__ __forceinline__ double pip(int d) const { return p.pip[(size_t)d * p.d.N + n]; }
  __device__ __forceinline__ double kalloc(int k) const { return p.kalloc[(size_t)k * p.d.N + n]; }
  __device__ __forceinline__ double kreq(int k) const { return p.kreq[(size_t)k * p.d.N + n]; }
  __device__ __forceinline__ double knz(int k) const { return p.knz[(size_t)k * p.d.N + n]; }
};

__device__ __forceinline__ unsigned long long ev_score_key(double x) {  // monotone double -> u64
  unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

// the action's candidate filter on one task of node.Tasks
__device__ __forceinline__ bool ev_filter(const EvictParams &p, const EvictTask &t, int r) {
  if (p.rt_evicted[r]) return false;
  const uint32_t f = p.rt_flags[r];
  if (!(f & VC_RT_PREEMPTABLE)) return false;
  const int j = p.rt_job[r];
  if (t.mode == EV_MODE_RECLAIM) {
    if (!(f & VC_RT_RUNNING) || j < 0) return false;
    const int q = p.j_queue[j];
    return q >= 0 && q != t.queue && !(p.q_flags[q] & VC_QUEUE_NOT_RECLAIMABLE);
  }
  if (!(f & (VC_RT_RUNNING | VC_RT_BOUND))) return false;
  if (t.mode == EV_MODE_PREEMPT_INTRA) return j == t.job;
  return j >= 0 && j != t.job && p.j_queue[j] == t.queue;
}

// Could task r be among the victims ssn.Preemptable / ssn.Reclaimable returns (session_plugins.go:211-307)? Inside a
// tier the plugins' answers are intersected, but an empty intersection resets the list to nil and the next plugin's answer
// replaces it — so all that is certain about a victim is that it passed the LAST voting plugin of the tier that decided.
// Used here, per voter, is the part of its test that does not depend on the other candidates: conformance (not critical),
// gang (job above minAvailable, gang.go:97-129), priority (lower job / task priority, priority.go:110-148), proportion
// (queue above deserved, proportion.go:286-317); drf's test depends on the walk and counts as a yes.
__device__ __forceinline__ bool ev_may_be_victim(const EvictParams &p, const EvictTask &t, int r) {
  const bool reclaim = t.mode == EV_MODE_RECLAIM;
  const uint32_t flag = reclaim ? VC_EN_RECLAIMABLE : VC_EN_PREEMPTABLE;
  const int j = p.rt_job[r];
  int i = 0;
  while (i < p.c.n_plugins) {
    const int tier = p.c.tier[i];
    bool any = false, last = false;
    for (; i < p.c.n_plugins && p.c.tier[i] == tier; ++i) {
      if (!(p.c.enabled[i] & flag)) continue;
      const int pl = p.c.plugin[i];
      bool vote;
      if (pl == VC_PLUGIN_CONFORMANCE) vote = !(p.rt_flags[r] & VC_RT_CRITICAL);
      else if (pl == VC_PLUGIN_GANG) vote = j >= 0 && p.j_ready[j] > p.j_min[j];
      else if (pl == VC_PLUGIN_PRIORITY && !reclaim) vote = j >= 0 && (j != t.job ? p.j_prio[j] < t.job_prio : p.rt_prio[r] < t.task_prio);
      else if (pl == VC_PLUGIN_DRF && !reclaim) vote = j >= 0;
      else if (pl == VC_PLUGIN_PROPORTION && reclaim) vote = j >= 0 && p.j_queue[j] >= 0 && p.q_over[p.j_queue[j]];
      else continue;
      any = true;
      last = vote;
    }
    if (any && last) return true;
  }
  return false;
}

// LPN lanes per node (16: two nodes per warp), one lane per entry of node.Tasks (strided when a node runs more pods than
// that): the filters are chains of dependent loads (task -> job -> queue ...), so the walk is latency-bound and the lanes
// overlap it; the requests of the possible victims are then summed per dimension with a shuffle tree inside the group.
// The group's first lane scores the node when it is a candidate. RMAX bounds the resource dimensions (register budget).
template <int LPN, int RMAX>
__global__ void __launch_bounds__(256) k_evict_rank(EvictParams p, EvictTask t) {
  const size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const int n = (int)(gid / LPN);
  const int gl = (int)(threadIdx.x % LPN);  // lane inside the group
  const int N = p.d.N, R = p.d.R, K = p.d.K;
  const bool live = n < N;
  const uint32_t cs = live ? p.cstat[(size_t)t.klass * N + n] : 0u;
  const bool stat_ok = live && (cs & CS_STATIC_OK);
  double freed[RMAX];
#pragma unroll
  for (int d = 0; d < RMAX; ++d) freed[d] = 0.0;
  int n_pass = 0;
  if (stat_ok) {
    for (int k = p.rt_off[n] + gl; k < p.rt_off[n + 1]; k += LPN) {
      const int r = p.rt_idx[k];
      if (!ev_filter(p, t, r)) continue;
      n_pass += 1;
      if (!ev_may_be_victim(p, t, r)) continue;
#pragma unroll
      for (int d = 0; d < RMAX; ++d)
        if (d < R) freed[d] += p.rt_req[(size_t)d * p.RT + r];
    }
  }
  // sums over the group (every lane of the warp takes part in the shuffles; groups are aligned, xor < LPN stays inside)
#pragma unroll
  for (int o = LPN / 2; o; o >>= 1) {
    n_pass += __shfl_xor_sync(0xffffffffu, n_pass, o);
#pragma unroll
    for (int d = 0; d < RMAX; ++d)
      if (d < R) freed[d] += __shfl_xor_sync(0xffffffffu, freed[d], o);
  }
  if (!live || gl != 0) return;
  uint8_t cand = 0;
  unsigned long long key = 0ull;
  if (stat_ok) {
    // ValidateVictims can only pass when FutureIdle + every possible victim covers the request. The sums above are
    // exact for integer-valued requests (exact_sums); otherwise a relative slack keeps the test a necessary condition
    // whatever the order of the additions.
    const double slack = p.exact_sums ? 0.0 : 1e-9;
    bool fits = true;
#pragma unroll
    for (int d = 0; d < RMAX; ++d) {
      if (d >= R) continue;
      if (d >= 2 && !(t.rec.has & (1u << d))) continue;
      const double pot = (p.idle[(size_t)d * N + n] + p.rel[(size_t)d * N + n]) - p.pip[(size_t)d * N + n];
      const double have = pot + freed[d];
      if (!le_eps(t.rec.req[d], have + fabs(have) * slack)) fits = false;
    }
    if (t.quota_on && p.exact_sums && t.mode != EV_MODE_RECLAIM) {
      // the evict loop ends with ssn.Allocatable(queue, preemptor) (preempt.go:380, :405): even with every possible
      // victim of this node evicted the queue must stay within deserved on the requested dimensions
      if (!t.quota_open) fits = false;
      const uint32_t rq_has = t.rec.has & ~3u;
#pragma unroll
      for (int d = 0; d < RMAX; ++d) {
        if (d >= R) continue;
        const double rq = t.rec.req[d];
        if (!(rq > 0.0)) continue;
        if (d >= 2 && (!(rq_has & (1u << d)) || d == p.d.pods_dim)) continue;
        const double al = (d < 2 || (t.qalloc_has & (1u << d))) ? t.qalloc[d] : 0.0;
        const double de = (d < 2 || (t.qdes_has & (1u << d))) ? t.qdes[d] : 0.0;
        if ((al - freed[d]) + rq > de) fits = false;
      }
    }
    if (fits && (t.mode != EV_MODE_RECLAIM || n_pass > 0)) {
      cand = 1;
      if (t.mode != EV_MODE_RECLAIM) {  // reclaim walks NodeList order, no scores (reclaim.go:172-180)
        const EvNodeView nv{p, n};
        double order = 0.0;
        const bool has_order = node_order(p.c, R, K, t.rec, nv, cs, &order);
        key = ev_score_key(total_score(p.c, has_order, has_order ? order : 0.0, 0, 0));
      }
    }
  }
  p.cand[n] = cand;
  p.key[n] = key;
}

// one block: the next EV_PICK_K candidates in the action's node order (preempt: score descending, lowest index first
// among equals; reclaim: index ascending); the nodes handed out are marked tried. Tournament with incremental repair:
// every thread keeps the best of its own nodes in registers, every warp its best in shared memory; handing out a node
// makes ONE thread rescan its nodes and ONE warp refold, the other 31 entries stay.
// `cached` != 0: the launch carries N * 8 bytes of dynamic shared memory and the keys of the live candidates are staged there
// once (0 = not a candidate), so that repairs read shared memory instead of global memory.
extern __shared__ unsigned long long ev_pick_cache[];
// The last store of the launch is `seq` into pick_node[EV_PICK_K] (system scope): the host polls that word of the mapped
// buffer instead of paying a stream synchronisation per hand-out.
__global__ void k_evict_pick_big(EvictParams p, int mode, int cached, int seq) {
  __shared__ unsigned long long s_key[32];
  __shared__ int s_node[32];
  __shared__ int s_best;
  __shared__ int s_pick[EV_PICK_K];  // hand-outs, written to the mapped host buffer once at the end (a store to system memory
                                     // per iteration sits on the critical path of the block-wide barriers)
  const int N = p.d.N;
  if (threadIdx.x < EV_PICK_K) s_pick[threadIdx.x] = -1;
  if (cached) {  // key + 1 (the top bit of a score key is never clear with all others set), 0 = not a candidate
    for (int n = threadIdx.x; n < N; n += blockDim.x)
      ev_pick_cache[n] = p.cand[n] == 1 ? (mode == EV_MODE_RECLAIM ? 1ull : p.key[n] + 1ull) : 0ull;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  auto better_kn = [](unsigned long long ka, int na, unsigned long long kb, int nb) {
    return na >= 0 && (nb < 0 || ka > kb || (ka == kb && na < nb));
  };
  unsigned long long bk = 0ull;  // this thread's best remaining candidate
  int bn = -1;
  auto rescan = [&](int skip) {
    bk = 0ull; bn = -1;
    if (cached) {
      for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const unsigned long long k = ev_pick_cache[n];
        if (n == skip || k == 0ull) continue;
        if (bn < 0 || k > bk) { bk = k; bn = n; }
      }
      return;
    }
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      if (n == skip || p.cand[n] != 1) continue;
      const unsigned long long k = mode == EV_MODE_RECLAIM ? 0ull : p.key[n];
      if (bn < 0 || k > bk) { bk = k; bn = n; }  // ascending n per thread: the first of equal keys stays
    }
  };
  auto warp_fold = [&]() {  // this warp's best into its shared slot
    unsigned long long wk = bk;
    int wn = bn;
    for (int o = 16; o; o >>= 1) {
      const unsigned long long ok = __shfl_xor_sync(0xffffffffu, wk, o);
      const int on = __shfl_xor_sync(0xffffffffu, wn, o);
      if (better_kn(ok, on, wk, wn)) { wk = ok; wn = on; }
    }
    if (lane == 0) { s_key[warp] = wk; s_node[warp] = wn; }
  };
  rescan(-1);
  warp_fold();
  __syncthreads();
  for (int it = 0; it < EV_PICK_K; ++it) {
    if (warp == 0) {
      unsigned long long gk = lane < nw ? s_key[lane] : 0ull;
      int gn = lane < nw ? s_node[lane] : -1;
      for (int o = 16; o; o >>= 1) {
        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, gk, o);
        const int on = __shfl_xor_sync(0xffffffffu, gn, o);
        if (better_kn(ok, on, gk, gn)) { gk = ok; gn = on; }
      }
      if (lane == 0) {
        s_pick[it] = gn;
        if (gn >= 0) { p.cand[gn] = 2; if (cached) ev_pick_cache[gn] = 0ull; }
        s_best = gn;
      }
    }
    __syncthreads();
    const int g = s_best;
    if (g < 0) break;  // exhausted: the remaining slots keep their -1
    if ((g % (int)blockDim.x) >> 5 == warp) {  // the warp of the thread that owned the node
      if (g % (int)blockDim.x == (int)threadIdx.x) rescan(g);  // (cand[g] = 2 was written by another thread: skip g by name)
      warp_fold();
    }
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x < EV_PICK_K) {
    *reinterpret_cast<volatile int32_t *>(p.pick_node + threadIdx.x) = s_pick[threadIdx.x];
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    *reinterpret_cast<volatile int32_t *>(p.pick_node + EV_PICK_K) = seq;
  }
}

// k_evict_pick, the common case (the candidates' keys fit in shared memory, N * 8 bytes): two-level selection with two
// block barriers in all. Phase 1, every warp on its own 1/32 of the nodes: EV_PICK_K rounds of a warp arg-max (each lane
// keeps the best of its ~N/1024 nodes in registers, the winner rescans its shared-memory entries) give the warp's first
// EV_PICK_K candidates in order. Phase 2, one warp: a 32-way merge of those sorted lists. Same order as the sequential
// tournament: key descending, node ascending.
__global__ void __launch_bounds__(1024) k_evict_pick(EvictParams p, int mode, int seq) {
  __shared__ unsigned long long l_key[32][EV_PICK_K];
  __shared__ int l_node[32][EV_PICK_K];
  __shared__ int s_pick[EV_PICK_K];
  const int N = p.d.N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int n = threadIdx.x; n < N; n += blockDim.x)  // key + 1 (never wraps: see ev_score_key), 0 = not a candidate
    ev_pick_cache[n] = p.cand[n] == 1 ? (mode == EV_MODE_RECLAIM ? 1ull : p.key[n] + 1ull) : 0ull;
  if (threadIdx.x < EV_PICK_K) s_pick[threadIdx.x] = -1;
  __syncwarp();  // a lane only ever reads the entries it wrote
  auto better_kn = [](unsigned long long ka, int na, unsigned long long kb, int nb) {
    return na >= 0 && (nb < 0 || ka > kb || (ka == kb && na < nb));
  };
  unsigned long long bk = 0ull;
  int bn = -1;
  auto rescan = [&]() {
    bk = 0ull; bn = -1;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      const unsigned long long k = ev_pick_cache[n];
      if (k == 0ull) continue;
      if (bn < 0 || k > bk) { bk = k; bn = n; }  // ascending n per thread: the first of equal keys stays
    }
  };
  rescan();
  for (int it = 0; it < EV_PICK_K; ++it) {
    unsigned long long wk = bk;
    int wn = bn;
    for (int o = 16; o; o >>= 1) {
      const unsigned long long ok = __shfl_xor_sync(0xffffffffu, wk, o);
      const int on = __shfl_xor_sync(0xffffffffu, wn, o);
      if (better_kn(ok, on, wk, wn)) { wk = ok; wn = on; }
    }
    if (lane == it) { l_key[warp][it] = wk; l_node[warp][it] = wn; }
    if (wn < 0) {  // this warp is out of candidates: the rest of its list says so
      if (lane > it && lane < EV_PICK_K) { l_key[warp][lane] = 0ull; l_node[warp][lane] = -1; }
      break;
    }
    if (bn == wn) { ev_pick_cache[wn] = 0ull; rescan(); }
  }
  __syncthreads();
  if (warp == 0) {
    int hp = 0;  // lane l walks warp l's list
    unsigned long long ck = l_key[lane][0];
    int cn = l_node[lane][0];
    for (int it = 0; it < EV_PICK_K; ++it) {
      unsigned long long gk = ck;
      int gn = cn;
      for (int o = 16; o; o >>= 1) {
        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, gk, o);
        const int on = __shfl_xor_sync(0xffffffffu, gn, o);
        if (better_kn(ok, on, gk, gn)) { gk = ok; gn = on; }
      }
      if (gn < 0) break;
      if (lane == 0) { s_pick[it] = gn; p.cand[gn] = 2; }
      if (cn == gn) {
        hp += 1;
        ck = hp < EV_PICK_K ? l_key[lane][hp] : 0ull;
        cn = hp < EV_PICK_K ? l_node[lane][hp] : -1;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < EV_PICK_K) {
    *reinterpret_cast<volatile int32_t *>(p.pick_node + threadIdx.x) = s_pick[threadIdx.x];
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    *reinterpret_cast<volatile int32_t *>(p.pick_node + EV_PICK_K) = seq;
  }
}

// Statement.Evict for every victim (node.UpdateTask: Releasing += Resreq; predicates DeallocateFunc: RemovePod), then
// Statement.Pipeline of the preemptor (Pipelined += Resreq; predicates AllocateFunc: AddPod) on node cmd[0]
// `undo` != 0: the inverse (Statement.Discard: UnPipeline, then unevict in reverse order; the sums are integer-valued
// doubles, so the order of the subtractions does not show)
__global__ void k_evict_apply(EvictParams p, EvictTask t, int undo) {
  const int n = p.cmd[0], nv = p.cmd[1];
  const int N = p.d.N, R = p.d.R, K = p.d.K;
  const int d = threadIdx.x;
  const double sg = undo ? -1.0 : 1.0;
  for (int k = 0; k < nv; ++k) {
    const int r = p.cmd[2 + k];
    if (d < R) p.rel[(size_t)d * N + n] += sg * p.rt_req[(size_t)d * p.RT + r];
    if (p.c.has_predicates) {
      if (d >= 32 && d < 32 + K) p.kreq[(size_t)(d - 32) * N + n] -= sg * p.rt_kreq[(size_t)(d - 32) * p.RT + r];
      if (d >= 48 && d < 50) p.knz[(size_t)(d - 48) * N + n] -= sg * p.rt_knz[(size_t)(d - 48) * p.RT + r];
    }
    if (d == 63) p.rt_evicted[r] = undo ? 0 : 1;
  }
  if (d < R) p.pip[(size_t)d * N + n] += sg * t.rec.req[d];
  if (p.c.has_predicates) {
    if (d >= 32 && d < 32 + K) p.kreq[(size_t)(d - 32) * N + n] += sg * t.rec.kreq[d - 32];
    if (d >= 48 && d < 50) p.knz[(size_t)(d - 48) * N + n] += sg * t.rec.knz[d - 48];
    if (d == 62) p.pod_count[n] += undo ? nv - 1 : 1 - nv;
  }
}
