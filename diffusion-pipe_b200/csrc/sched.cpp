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
