// 1F1B pipeline instruction planner (host code, no CUDA): the C++ restatement of the schedule the reference
// installs over DeepSpeed's TrainSchedule (reference: utils/patches.py:113-160, `train_schedule_steps`) together with
// the TrainSchedule / InferenceSchedule helper arithmetic of deepspeed==0.18.4 runtime/pipe/schedule.py
// (third-party, not vendored; restated from the published source: _step_to_micro_batch, _valid_micro_batch,
// _valid_stage, _buffer_idx, num_pipe_buffers).  Integer-exact; tests/test_schedule.py pins it against the golden
// traces produced by running the reference generator itself (tests/golden/make_golden_schedule.py) and against
// SURVEY.md Appendix C.
//
// Output: a flat instruction array; DPIPE_OP_TICK_END closes every schedule tick so the caller sees the same
// per-tick grouping (`yield cmds`) as the reference.
#include "host_util.h"

namespace {

struct Planner {
  int micro_batches, stages, stage_id;
  int prev_stage() const { return stage_id - 1; }
  int next_stage() const { return stage_id + 1; }
  bool valid_micro_batch(int m) const { return m >= 0 && m < micro_batches; }
  bool valid_stage(int s) const { return s >= 0 && s < stages; }
  int num_pipe_buffers() const {
    int b = stages - stage_id;
    if (micro_batches < b) b = micro_batches;
    return b < 2 ? 2 : b;
  }
  int buffer_idx(int m) const { return m % num_pipe_buffers(); }
  // (micro_batch_id, is_forward) for a schedule step
  void step_to_micro_batch(int step, int* mb, bool* fwd) const {
    const bool even_step = (step % 2) == 0, even_stage = (stage_id % 2) == 0;
    if (even_step && even_stage) { *mb = step / 2 - stage_id / 2; *fwd = true; }
    else if (!even_step && !even_stage) { *mb = (step - 1) / 2 - stage_id / 2; *fwd = true; }
    else if (even_step && !even_stage) { *mb = step / 2 - stages + (stage_id + 1) / 2; *fwd = false; }
    else { *mb = ((step - 1) / 2) - stages + 1 + stage_id / 2; *fwd = false; }
  }
};

struct Emit {
  dpipe_instr* out;
  int cap, n;
  bool overflow;
  void push(int op, int buf, int mb) {
    if (n < cap) { out[n].op = op; out[n].buffer = buf; out[n].micro_batch = mb; }
    else overflow = true;
    ++n;
  }
};

}  // namespace

extern "C" int dpipe_sched_num_pipe_buffers(int micro_batches, int stages, int stage_id) {
  if (micro_batches < 1 || stages < 1 || stage_id < 0 || stage_id >= stages)
    return dpipe::fail(DPIPE_EINVAL, "dpipe_sched_num_pipe_buffers: bad geometry");
  Planner p{micro_batches, stages, stage_id};
  return p.num_pipe_buffers();
}

extern "C" int dpipe_sched_train(int micro_batches, int stages, int stage_id, dpipe_instr* out, int capacity) {
  if (micro_batches < 1 || stages < 1 || stage_id < 0 || stage_id >= stages || (!out && capacity > 0))
    return dpipe::fail(DPIPE_EINVAL, "dpipe_sched_train: bad geometry M=%d S=%d stage=%d", micro_batches, stages, stage_id);
  Planner p{micro_batches, stages, stage_id};
  Emit e{out, capacity, 0, false};
  int prev_mb = -1, prev_buffer = -1, curr_buffer = -1;
  const int total_steps = 2 * (micro_batches + stages - 1);
  for (int step = 0; step < total_steps; ++step) {
    int mb;
    bool fwd;
    p.step_to_micro_batch(step, &mb, &fwd);
    if (p.valid_micro_batch(prev_mb)) prev_buffer = p.buffer_idx(prev_mb);
    if (p.valid_micro_batch(mb)) curr_buffer = p.buffer_idx(mb);
    // first / last stage loads come before any communication (the reference's patch)
    if (stage_id == 0 || stage_id == stages - 1) {
      if (fwd && p.valid_micro_batch(mb)) e.push(DPIPE_OP_LOAD_MICRO_BATCH, curr_buffer, mb);
    }
    if (fwd) {
      if (p.valid_micro_batch(prev_mb) && p.valid_stage(p.prev_stage())) e.push(DPIPE_OP_SEND_GRAD, prev_buffer, prev_mb);
      if (p.valid_micro_batch(mb) && p.valid_stage(p.prev_stage())) e.push(DPIPE_OP_RECV_ACTIVATION, curr_buffer, mb);
    } else {
      if (p.valid_micro_batch(mb) && p.valid_stage(p.next_stage())) e.push(DPIPE_OP_RECV_GRAD, curr_buffer, mb);
      if (p.valid_micro_batch(prev_mb) && p.valid_stage(p.next_stage())) e.push(DPIPE_OP_SEND_ACTIVATION, prev_buffer, prev_mb);
    }
    if (p.valid_micro_batch(mb)) e.push(fwd ? DPIPE_OP_FORWARD_PASS : DPIPE_OP_BACKWARD_PASS, curr_buffer, mb);
    if (step == total_steps - 1) {
      e.push(DPIPE_OP_REDUCE_TIED_GRADS, -1, -1);
      e.push(DPIPE_OP_REDUCE_GRADS, -1, -1);
      e.push(DPIPE_OP_OPTIMIZER_STEP, -1, -1);
    }
    prev_mb = mb;
    e.push(DPIPE_OP_TICK_END, -1, -1);
  }
  if (e.overflow && capacity > 0) return dpipe::fail(DPIPE_EINVAL, "dpipe_sched_train: capacity %d < %d", capacity, e.n);
  return e.n;
}

// forward-only schedule used by eval_batch (deepspeed InferenceSchedule: two alternating buffers)
extern "C" int dpipe_sched_infer(int micro_batches, int stages, int stage_id, dpipe_instr* out, int capacity) {
  if (micro_batches < 1 || stages < 1 || stage_id < 0 || stage_id >= stages || (!out && capacity > 0))
    return dpipe::fail(DPIPE_EINVAL, "dpipe_sched_infer: bad geometry");
  Planner p{micro_batches, stages, stage_id};
  Emit e{out, capacity, 0, false};
  const int total_steps = micro_batches + stages - 1;
  const bool even_stage = (stage_id % 2) == 0;
  for (int step = 0; step < total_steps; ++step) {
    const int mb = step - stage_id;
    const int recv_buf = even_stage ? step % 2 : (step + 1) % 2;
    const int send_buf = even_stage ? (step + 1) % 2 : step % 2;
    if (stage_id == 0 || stage_id == stages - 1) {
      if (p.valid_micro_batch(mb)) e.push(DPIPE_OP_LOAD_MICRO_BATCH, recv_buf, mb);
    }
    const bool do_send = p.valid_stage(p.next_stage()) && p.valid_micro_batch(mb - 1);
    const bool do_recv = p.valid_stage(p.prev_stage()) && p.valid_micro_batch(mb);
    if (even_stage) {
      if (do_send) e.push(DPIPE_OP_SEND_ACTIVATION, send_buf, mb - 1);
      if (do_recv) e.push(DPIPE_OP_RECV_ACTIVATION, recv_buf, mb);
    } else {
      if (do_recv) e.push(DPIPE_OP_RECV_ACTIVATION, recv_buf, mb);
      if (do_send) e.push(DPIPE_OP_SEND_ACTIVATION, send_buf, mb - 1);
    }
    if (p.valid_micro_batch(mb)) e.push(DPIPE_OP_FORWARD_PASS, recv_buf, mb);
    e.push(DPIPE_OP_TICK_END, -1, -1);
  }
  if (e.overflow && capacity > 0) return dpipe::fail(DPIPE_EINVAL, "dpipe_sched_infer: capacity %d < %d", capacity, e.n);
  return e.n;
}

// contiguous min-max partition of `weights` into `parts` pieces: bounds[0..parts] (reference: DeepSpeed's
// partition_balanced used by partition_method='parameters', train.py:81-90,606).  Among all partitions that
// minimise the heaviest part the earliest boundaries are taken (greedy left fill at the optimal bottleneck).
extern "C" int dpipe_partition_balanced(const int64_t* weights, int n, int parts, int* bounds) {
  if (!weights || !bounds || n < 0 || parts < 1) return dpipe::fail(DPIPE_EINVAL, "dpipe_partition_balanced: bad arguments");
  int64_t lo = 0, hi = 0;
  for (int i = 0; i < n; ++i) {
    if (weights[i] < 0) return dpipe::fail(DPIPE_EINVAL, "dpipe_partition_balanced: negative weight");
    if (weights[i] > lo) lo = weights[i];
    hi += weights[i];
  }
  auto parts_needed = [&](int64_t cap) {
    int used = 1;
    int64_t cur = 0;
    for (int i = 0; i < n; ++i) {
      if (cur + weights[i] > cap) { ++used; cur = 0; }
      cur += weights[i];
    }
    return used;
  };
  while (lo < hi) {
    const int64_t mid = lo + (hi - lo) / 2;
    if (parts_needed(mid) <= parts) hi = mid; else lo = mid + 1;
  }
  // fill greedily at bottleneck `lo`, but leave at least one layer for each remaining part when possible
  int idx = 0;
  bounds[0] = 0;
  for (int part = 0; part < parts; ++part) {
    int64_t cur = 0;
    const int remaining_parts = parts - part - 1;
    while (idx < n && cur + weights[idx] <= lo && (n - idx) > remaining_parts) { cur += weights[idx]; ++idx; }
    if (part == parts - 1) idx = n;
    bounds[part + 1] = idx;
  }
  return 0;
}

// -------------------------------------------------------------------------------------------------------------------
// Split-backward ("zero-bubble") schedule.  Not in the reference: it is loss-equivalent to 1F1B (same per-micro-batch
// arithmetic, gradients accumulated in micro-batch order) but separates the backward pass into
//   B = input-gradient pass (on the critical path: its result is what the previous stage waits for) and
//   W = weight-gradient pass (no consumer on another stage),
// so W work can fill what would be pipeline bubbles (Qi et al., "Zero Bubble Pipeline Parallelism", ZB-H1/ZB-2p).
//
// The order comes from a deterministic list-scheduling simulation of ALL stages: whenever a stage becomes free it
// runs, in priority order, the oldest ready B, else the next F if it holds fewer than `max_inflight` micro-batches
// (forward done, weight-gradient pass not yet done: that is what pins activation and output-gradient memory), else a
// pending W; it idles only when nothing is ready.  Stage s is charged (tf, tb, tw) * stage_weight[s] per operation, so
// uneven partitions (57 blocks over 8 stages) are planned for what they are.  List scheduling is a heuristic whose
// result depends on the cost model it is driven with, so the planner tries a few cost models (the true one, the true
// one without the per-stage weights, all-ones), replays each resulting order under the TRUE costs and keeps the order
// with the smallest simulated makespan.  Every rank runs the same search, so all stages agree without communication.
// Needs a one-sided stage link (IpcLink) or asynchronous sends: a stage pushes activations right after F and gradients
// right after B, whenever the neighbour will get to them.
// -------------------------------------------------------------------------------------------------------------------
namespace {

struct ZbCost { long long f, b, w; };
enum ZbOp : unsigned char { ZB_F = 1, ZB_B = 2, ZB_W = 3 };

// list-scheduling simulation under `cost`; fills order[s * 3M + i] (kinds; the micro-batch of the i-th F/B/W of a stage
// is its ordinal among operations of that kind)
void zb_list_schedule(int M, int S, const ZbCost* cost, int max_inflight, unsigned char* order) {
  struct St { long long free_at; int nf, nb, nw, n; bool done; };
  St* st = new St[S];
  long long* fin_f = new long long[(size_t)S * M];
  long long* fin_b = new long long[(size_t)S * M];
  const long long INF = (1LL << 60);
  for (int s = 0; s < S; ++s) st[s] = {0, 0, 0, 0, 0, false};
  for (size_t i = 0; i < (size_t)S * M; ++i) { fin_f[i] = INF; fin_b[i] = INF; }
  int remaining = S;
  while (remaining > 0) {
    int s = -1;   // the stage with the smallest clock acts next (ties: lowest index)
    for (int i = 0; i < S; ++i)
      if (!st[i].done && (s < 0 || st[i].free_at < st[s].free_at)) s = i;
    St& a = st[s];
    const long long now = a.free_at;
    auto f_ready = [&](int m) { return s == 0 ? 0 : fin_f[(size_t)(s - 1) * M + m]; };
    auto b_ready = [&](int m) { return s == S - 1 ? fin_f[(size_t)s * M + m] : fin_b[(size_t)(s + 1) * M + m]; };
    int op = 0;
    if (a.nb < M && a.nb < a.nf && b_ready(a.nb) <= now) op = ZB_B;
    else if (a.nf < M && (a.nf - a.nw) < max_inflight && f_ready(a.nf) <= now) op = ZB_F;
    else if (a.nw < a.nb) op = ZB_W;
    if (op == 0) {   // nothing ready: sleep until the next point in time at which something can have changed
      long long next = INF;
      if (a.nb < M && a.nb < a.nf && b_ready(a.nb) < INF) next = b_ready(a.nb);
      if (a.nf < M && (a.nf - a.nw) < max_inflight && f_ready(a.nf) < INF && f_ready(a.nf) < next) next = f_ready(a.nf);
      if (next == INF || next <= now) {
        for (int i = 0; i < S; ++i)
          if (i != s && !st[i].done && st[i].free_at > now && st[i].free_at < next) next = st[i].free_at;
      }
      if (next == INF || next <= now) next = now + 1;
      a.free_at = next;
      continue;
    }
    if (op == ZB_F) { a.free_at = now + cost[s].f; fin_f[(size_t)s * M + a.nf] = a.free_at; a.nf++; }
    else if (op == ZB_B) { a.free_at = now + cost[s].b; fin_b[(size_t)s * M + a.nb] = a.free_at; a.nb++; }
    else { a.free_at = now + cost[s].w; a.nw++; }
    order[(size_t)s * 3 * M + a.n++] = (unsigned char)op;
    if (a.nf == M && a.nb == M && a.nw == M) { a.done = true; --remaining; }
  }
  delete[] st; delete[] fin_f; delete[] fin_b;
}

// makespan of the per-stage orders under `cost` (dependencies: F(s,m) after F(s-1,m); B(s,m) after B(s+1,m), or after
// F(s,m) on the last stage; W(s,m) after B(s,m)).  Returns -2 if the orders deadlock (must never happen).
long long zb_replay(int M, int S, const ZbCost* cost, const unsigned char* order) {
  const long long INF = (1LL << 60);
  long long* ff = new long long[(size_t)S * M];
  long long* fb = new long long[(size_t)S * M];
  for (size_t i = 0; i < (size_t)S * M; ++i) { ff[i] = INF; fb[i] = INF; }
  int* pos = new int[S];
  int* cnt = new int[(size_t)S * 3];
  long long* t = new long long[S];
  for (int s = 0; s < S; ++s) { pos[s] = 0; t[s] = 0; cnt[s * 3] = cnt[s * 3 + 1] = cnt[s * 3 + 2] = 0; }
  bool progress = true;
  while (progress) {
    progress = false;
    for (int s = 0; s < S; ++s) {
      while (pos[s] < 3 * M) {
        const int kind = order[(size_t)s * 3 * M + pos[s]];
        const int m = cnt[s * 3 + kind - 1];
        long long ready;
        if (kind == ZB_F) ready = s == 0 ? 0 : ff[(size_t)(s - 1) * M + m];
        else if (kind == ZB_B) ready = s == S - 1 ? ff[(size_t)s * M + m] : fb[(size_t)(s + 1) * M + m];
        else ready = fb[(size_t)s * M + m];
        if (ready == INF) break;
        const long long start = ready > t[s] ? ready : t[s];
        t[s] = start + (kind == ZB_F ? cost[s].f : kind == ZB_B ? cost[s].b : cost[s].w);
        if (kind == ZB_F) ff[(size_t)s * M + m] = t[s];
        if (kind == ZB_B) fb[(size_t)s * M + m] = t[s];
        cnt[s * 3 + kind - 1]++;
        ++pos[s];
        progress = true;
      }
    }
  }
  long long best = 0;
  bool complete = true;
  for (int s = 0; s < S; ++s) { if (pos[s] != 3 * M) complete = false; if (t[s] > best) best = t[s]; }
  delete[] ff; delete[] fb; delete[] pos; delete[] cnt; delete[] t;
  return complete ? best : -2;
}

// the candidate search described above; returns the chosen order (new[]-allocated, S * 3M kinds) and its makespan
unsigned char* zb_best_order(int M, int S, int tf, int tb, int tw, int max_inflight, const int* stage_weight, long long* makespan) {
  ZbCost* truth = new ZbCost[S];
  ZbCost* model = new ZbCost[S];
  for (int s = 0; s < S; ++s) {
    const long long w = stage_weight ? stage_weight[s] : 1;
    truth[s] = {tf * w, tb * w, tw * w};
  }
  unsigned char* best = nullptr;
  long long best_ms = -1;
  unsigned char* cand = new unsigned char[(size_t)S * 3 * M];
  for (int c = 0; c < 3; ++c) {
    for (int s = 0; s < S; ++s) {
      if (c == 0) model[s] = truth[s];
      else if (c == 1) model[s] = {tf, tb, tw};
      else model[s] = {1, 1, 1};
    }
    zb_list_schedule(M, S, model, max_inflight, cand);
    const long long ms = zb_replay(M, S, truth, cand);
    if (ms >= 0 && (best_ms < 0 || ms < best_ms)) {
      best_ms = ms;
      if (!best) best = new unsigned char[(size_t)S * 3 * M];
      for (size_t i = 0; i < (size_t)S * 3 * M; ++i) best[i] = cand[i];
    }
  }
  delete[] cand; delete[] truth; delete[] model;
  *makespan = best_ms;
  return best;
}

bool zb_args_ok(int M, int S, int tf, int tb, int tw, int max_inflight, const int* stage_weight) {
  if (M < 1 || S < 1 || tf < 1 || tb < 1 || tw < 1 || max_inflight < 1) return false;
  if (stage_weight)
    for (int s = 0; s < S; ++s)
      if (stage_weight[s] < 1 || stage_weight[s] > (1 << 20)) return false;
  return true;
}

}  // namespace

extern "C" int dpipe_sched_zb_ex(int micro_batches, int stages, int stage_id, int tf, int tb, int tw, int max_inflight,
                                 const int* stage_weight, dpipe_instr* out, int capacity) {
  const int M = micro_batches, S = stages;
  if (!zb_args_ok(M, S, tf, tb, tw, max_inflight, stage_weight) || stage_id < 0 || stage_id >= S || (!out && capacity > 0))
    return dpipe::fail(DPIPE_EINVAL, "dpipe_sched_zb: bad arguments");
  long long ms = 0;
  unsigned char* order = zb_best_order(M, S, tf, tb, tw, max_inflight, stage_weight, &ms);
  if (!order) return dpipe::fail(DPIPE_EINVAL, "dpipe_sched_zb: no deadlock-free order found (internal error)");
  Emit e{out, capacity, 0, false};
  const int s = stage_id;
  int cnt[3] = {0, 0, 0};
  for (int i = 0; i < 3 * M; ++i) {
    const int kind = order[(size_t)s * 3 * M + i];
    const int m = cnt[kind - 1]++;
    if (kind == ZB_F) {
      if (s == 0 || s == S - 1) e.push(DPIPE_OP_LOAD_MICRO_BATCH, m, m);
      if (s > 0) e.push(DPIPE_OP_RECV_ACTIVATION, m, m);
      e.push(DPIPE_OP_FORWARD_PASS, m, m);
      if (s < S - 1) e.push(DPIPE_OP_SEND_ACTIVATION, m, m);
    } else if (kind == ZB_B) {
      if (s < S - 1) e.push(DPIPE_OP_RECV_GRAD, m, m);
      e.push(DPIPE_OP_BACKWARD_INPUT, m, m);
      if (s > 0) e.push(DPIPE_OP_SEND_GRAD, m, m);
    } else {
      e.push(DPIPE_OP_BACKWARD_WEIGHT, m, m);
    }
    e.push(DPIPE_OP_TICK_END, -1, -1);
  }
  delete[] order;
  e.push(DPIPE_OP_REDUCE_TIED_GRADS, -1, -1);
  e.push(DPIPE_OP_REDUCE_GRADS, -1, -1);
  e.push(DPIPE_OP_OPTIMIZER_STEP, -1, -1);
  e.push(DPIPE_OP_TICK_END, -1, -1);
  if (e.overflow && capacity > 0) return dpipe::fail(DPIPE_EINVAL, "dpipe_sched_zb: capacity %d < %d", capacity, e.n);
  return e.n;
}

extern "C" int dpipe_sched_zb(int micro_batches, int stages, int stage_id, int tf, int tb, int tw, int max_inflight,
                              dpipe_instr* out, int capacity) {
  return dpipe_sched_zb_ex(micro_batches, stages, stage_id, tf, tb, tw, max_inflight, nullptr, out, capacity);
}

// simulated makespan of the chosen order under the true costs, for tests / reporting
extern "C" long long dpipe_sched_zb_makespan_ex(int micro_batches, int stages, int tf, int tb, int tw, int max_inflight,
                                                const int* stage_weight) {
  if (!zb_args_ok(micro_batches, stages, tf, tb, tw, max_inflight, stage_weight)) return -1;
  long long ms = 0;
  unsigned char* order = zb_best_order(micro_batches, stages, tf, tb, tw, max_inflight, stage_weight, &ms);
  if (!order) return -2;
  delete[] order;
  return ms;
}

extern "C" long long dpipe_sched_zb_makespan(int micro_batches, int stages, int tf, int tb, int tw, int max_inflight) {
  return dpipe_sched_zb_makespan_ex(micro_batches, stages, tf, tb, tw, max_inflight, nullptr);
}
