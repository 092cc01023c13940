// Pipeline-schedule executor: C++ owns the ORDER of a stage's instruction stream, the copy stream, the CUDA events and the
// stage-boundary copies; the host language keeps only what needs autograd (LoadMicroBatch / Forward / Backward / optimizer).
//
//   dpipe_exec_load_plan   takes the instruction array of the planner (dpipe_sched_train / _infer / _zb_ex)
//   dpipe_exec_next        walks it: SendActivation / RecvActivation / SendGrad / RecvGrad are EXECUTED here —
//                            send:  event(compute stream) -> copy stream waits -> wait free[k] >= w-1 (device-side) ->
//                                   cudaMemcpyPeerAsync of every tensor into the peer's slot k -> ready[k] = w (st.release.sys)
//                            recv:  compute stream waits ready[k] >= w (ld.acquire.sys); the tuple IS the slot (no second copy)
//                            release (after the consumer of a slot has been enqueued): free[k] = w on the compute stream
//                          and returns to the caller only at an instruction the host must run, or when a channel still needs
//                          its once-per-step host handshake (shapes / IPC handles over the gloo side group).
// Same protocol as round 1's Python link (pipe/ipc_link.py), now one C call per schedule instruction instead of 4-10.
//
// Replaces DeepSpeed's PipelineEngine._exec_schedule + _exec_send/recv_* (instruction stream: reference
// utils/patches.py:113-160, driven from train.py:918).
#include <string.h>

#include <vector>

#include "host_util.h"

extern "C" int dpipe_flag_write(void* flag, uint64_t value, void* stream);
extern "C" int dpipe_flag_wait_geq(const void* flag, uint64_t value, double timeout_s, void* stream);

namespace dpipe {

struct Channel {
  bool sending = false, bound = false, laid_out = false;
  // sender: own `free` flags + the peer's mailbox; receiver: own mailbox + the peer's `free` flags
  unsigned char* flags_local = nullptr;     // sender: free[]   receiver: mailbox base (ready[] first)
  unsigned char* remote = nullptr;          // sender: peer mailbox base   receiver: peer free[]
  int peer_device = -1;
  int64_t flag_bytes = 0, slot_bytes = 0;
  std::vector<int64_t> offsets, nbytes;
  std::vector<uint64_t> count;              // per slot: writes (sender) / reads (receiver)
  std::vector<std::vector<const void*>> staged;   // per pipe buffer: the tensors a coming Send will copy
};

struct Exec {
  int device = 0, nslots = 0;
  double timeout_s = 300.0;
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> events;
  size_t next_event = 0;
  Channel ch[4];
  std::vector<dpipe_instr> plan;
  size_t pc = 0;
  bool train = true;
  int is_first = 0, is_last = 0;
  // the instruction last handed to the host (its slot releases happen when the host comes back: its work is enqueued by then)
  dpipe_instr pending = {0, -1, -1};
};

static cudaEvent_t next_event(Exec* e) {
  if (e->events.size() < 64) {
    cudaEvent_t ev;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    e->events.push_back(ev);
    return ev;
  }
  cudaEvent_t ev = e->events[e->next_event % e->events.size()];
  ++e->next_event;
  return ev;
}

static int push(Exec* e, int c, int buffer, int mb, cudaStream_t compute) {
  Channel& ch = e->ch[c];
  if (!ch.bound || !ch.laid_out) return fail(DPIPE_EINVAL, "stage executor: channel %d pushed before its handshake", c);
  if (buffer < 0 || buffer >= (int)ch.staged.size() || ch.staged[buffer].size() != ch.offsets.size())
    return fail(DPIPE_EINVAL, "stage executor: no tensors staged for the send of buffer %d on channel %d", buffer, c);
  const int k = mb % e->nslots;
  const uint64_t w = ch.count[k] + 1;
  cudaEvent_t ev = next_event(e);
  if (!ev) return fail(DPIPE_ECUDA, "stage executor: cudaEventCreate failed");
  DPIPE_CUDA_CHECK(cudaEventRecord(ev, compute));
  DPIPE_CUDA_CHECK(cudaStreamWaitEvent(e->copy_stream, ev, 0));
  int rc = dpipe_flag_wait_geq(ch.flags_local + 8 * k, w - 1, e->timeout_s, e->copy_stream);
  if (rc) return rc;
  unsigned char* base = ch.remote + ch.flag_bytes + (int64_t)k * ch.slot_bytes;
  for (size_t i = 0; i < ch.offsets.size(); ++i) {
    if (ch.nbytes[i] == 0) continue;
    DPIPE_CUDA_CHECK(cudaMemcpyPeerAsync(base + ch.offsets[i], ch.peer_device, ch.staged[buffer][i], e->device,
                                         (size_t)ch.nbytes[i], e->copy_stream));
  }
  rc = dpipe_flag_write(ch.remote + 8 * k, w, e->copy_stream);
  if (rc) return rc;
  ch.count[k] = w;
  ch.staged[buffer].clear();
  return 0;
}

static int pull(Exec* e, int c, int mb, cudaStream_t compute) {
  Channel& ch = e->ch[c];
  if (!ch.bound || !ch.laid_out) return fail(DPIPE_EINVAL, "stage executor: channel %d pulled before its handshake", c);
  const int k = mb % e->nslots;
  const uint64_t w = ch.count[k] + 1;
  int rc = dpipe_flag_wait_geq(ch.flags_local + 8 * k, w, e->timeout_s, compute);
  if (rc) return rc;
  ch.count[k] = w;
  return 0;
}

static int release(Exec* e, int c, int mb, cudaStream_t compute) {
  Channel& ch = e->ch[c];
  if (!ch.bound) return 0;
  const int k = mb % e->nslots;
  return dpipe_flag_write(ch.remote + 8 * k, ch.count[k], compute);
}

}  // namespace dpipe

using namespace dpipe;

struct dpipe_exec { Exec e; };

extern "C" int dpipe_exec_create(int device, int nslots, double timeout_s, dpipe_exec** out) {
  if (!out || nslots < 1 || nslots > 4096) return fail(DPIPE_EINVAL, "dpipe_exec_create: bad arguments");
  dpipe_exec* x = new dpipe_exec();
  x->e.device = device;
  x->e.nslots = nslots;
  x->e.timeout_s = timeout_s;
  cudaError_t err = cudaStreamCreateWithFlags(&x->e.copy_stream, cudaStreamNonBlocking);
  if (err != cudaSuccess) { delete x; return fail(DPIPE_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(err)); }
  for (int c = 0; c < 4; ++c) { x->e.ch[c].sending = (c == DPIPE_CH_ACT_OUT || c == DPIPE_CH_GRAD_OUT); x->e.ch[c].count.assign(nslots, 0); }
  *out = x;
  return 0;
}

extern "C" int dpipe_exec_destroy(dpipe_exec* x) {
  if (!x) return 0;
  for (cudaEvent_t ev : x->e.events) cudaEventDestroy(ev);
  if (x->e.copy_stream) cudaStreamDestroy(x->e.copy_stream);
  delete x;
  return 0;
}

extern "C" void* dpipe_exec_copy_stream(dpipe_exec* x) { return x ? (void*)x->e.copy_stream : nullptr; }

extern "C" int dpipe_exec_bind(dpipe_exec* x, int channel, void* flags_local, void* remote, int peer_device,
                               int64_t flag_bytes, int64_t slot_bytes, int reset_counts) {
  if (!x || channel < 0 || channel > 3 || !flags_local || !remote || flag_bytes < 8 * x->e.nslots || slot_bytes < 0)
    return fail(DPIPE_EINVAL, "dpipe_exec_bind: bad arguments");
  Channel& ch = x->e.ch[channel];
  ch.flags_local = (unsigned char*)flags_local;
  ch.remote = (unsigned char*)remote;
  ch.peer_device = peer_device;
  ch.flag_bytes = flag_bytes;
  ch.slot_bytes = slot_bytes;
  ch.bound = true;
  if (reset_counts) ch.count.assign(x->e.nslots, 0);
  return 0;
}

extern "C" int dpipe_exec_set_layout(dpipe_exec* x, int channel, int n, const int64_t* offsets, const int64_t* nbytes) {
  if (!x || channel < 0 || channel > 3 || n < 0 || (n > 0 && (!offsets || !nbytes))) return fail(DPIPE_EINVAL, "dpipe_exec_set_layout: bad arguments");
  Channel& ch = x->e.ch[channel];
  ch.offsets.assign(offsets, offsets + n);
  ch.nbytes.assign(nbytes, nbytes + n);
  for (int i = 0; i < n; ++i)
    if (offsets[i] < 0 || nbytes[i] < 0 || offsets[i] + nbytes[i] > ch.slot_bytes)
      return fail(DPIPE_EINVAL, "dpipe_exec_set_layout: tensor %d does not fit the slot", i);
  ch.laid_out = true;
  return 0;
}

extern "C" int dpipe_exec_forget_layouts(dpipe_exec* x) {   /* engine.reset_activation_shape(): shapes may change (train.py:916) */
  if (!x) return fail(DPIPE_EINVAL, "dpipe_exec_forget_layouts: null");
  for (int c = 0; c < 4; ++c) x->e.ch[c].laid_out = false;
  return 0;
}

extern "C" int dpipe_exec_load_plan(dpipe_exec* x, const dpipe_instr* instrs, int n, int train, int is_first_stage,
                                    int is_last_stage, int num_buffers) {
  if (!x || n < 0 || (n > 0 && !instrs) || num_buffers < 1) return fail(DPIPE_EINVAL, "dpipe_exec_load_plan: bad arguments");
  x->e.plan.assign(instrs, instrs + n);
  x->e.pc = 0;
  x->e.train = train != 0;
  x->e.is_first = is_first_stage;
  x->e.is_last = is_last_stage;
  x->e.pending = {0, -1, -1};
  for (int c = 0; c < 4; ++c) {
    x->e.ch[c].staged.clear();
    x->e.ch[c].staged.resize(num_buffers);
  }
  return 0;
}

extern "C" int dpipe_exec_stage_send(dpipe_exec* x, int channel, int buffer, int n, const void* const* ptrs) {
  if (!x || (channel != DPIPE_CH_ACT_OUT && channel != DPIPE_CH_GRAD_OUT) || n < 0 || (n > 0 && !ptrs))
    return fail(DPIPE_EINVAL, "dpipe_exec_stage_send: bad arguments");
  Channel& ch = x->e.ch[channel];
  if (buffer < 0 || buffer >= (int)ch.staged.size()) return fail(DPIPE_EINVAL, "dpipe_exec_stage_send: buffer %d out of range", buffer);
  ch.staged[buffer].assign(ptrs, ptrs + n);
  return 0;
}

extern "C" int dpipe_exec_recv_base(dpipe_exec* x, int channel, int micro_batch, void** base) {
  if (!x || (channel != DPIPE_CH_ACT_IN && channel != DPIPE_CH_GRAD_IN) || !base || micro_batch < 0)
    return fail(DPIPE_EINVAL, "dpipe_exec_recv_base: bad arguments");
  Channel& ch = x->e.ch[channel];
  if (!ch.bound) return fail(DPIPE_EINVAL, "dpipe_exec_recv_base: channel %d is not bound", channel);
  *base = ch.flags_local + ch.flag_bytes + (int64_t)(micro_batch % x->e.nslots) * ch.slot_bytes;
  return 0;
}

// Returns 1 with *out = the instruction the HOST must execute now (LoadMicroBatch, ForwardPass, Backward*, Reduce*,
// OptimizerStep) or a pseudo-instruction op = DPIPE_OP_NEED_HANDSHAKE + channel (buffer / micro_batch of the instruction that
// needs it: do the host handshake, bind, set the layout, call again: the instruction is retried); 0 when the plan is finished;
// < 0 on error.
extern "C" int dpipe_exec_next(dpipe_exec* x, void* compute_stream, dpipe_instr* out) {
  if (!x || !out) return fail(DPIPE_EINVAL, "dpipe_exec_next: null argument");
  Exec& e = x->e;
  cudaStream_t cs = (cudaStream_t)compute_stream;
  int rc;
  // slot releases owed for the instruction the host has just run (everything it launched is on `cs` by now)
  if (e.pending.op == DPIPE_OP_BACKWARD_PASS || e.pending.op == DPIPE_OP_BACKWARD_INPUT) {
    if (!e.is_last && (rc = release(&e, DPIPE_CH_GRAD_IN, e.pending.micro_batch, cs))) return rc;
  } else if (e.pending.op == DPIPE_OP_FORWARD_PASS && !e.train) {
    if (e.is_last && !e.is_first && (rc = release(&e, DPIPE_CH_ACT_IN, e.pending.micro_batch, cs))) return rc;
  }
  e.pending = {0, -1, -1};
  while (e.pc < e.plan.size()) {
    const dpipe_instr in = e.plan[e.pc];
    switch (in.op) {
      case DPIPE_OP_TICK_END:
      case DPIPE_OP_REDUCE_TIED_GRADS:
        ++e.pc;
        break;
      case DPIPE_OP_SEND_ACTIVATION: {
        Channel& ch = e.ch[DPIPE_CH_ACT_OUT];
        if (!ch.bound || !ch.laid_out) { *out = {DPIPE_OP_NEED_HANDSHAKE + DPIPE_CH_ACT_OUT, in.buffer, in.micro_batch}; return 1; }
        if ((rc = push(&e, DPIPE_CH_ACT_OUT, in.buffer, in.micro_batch, cs))) return rc;
        if (!e.train && !e.is_first) {
          // forward-only: outputs may alias pass-through tensors living in our input slot — hand the slot back only after
          // the copy engine has read them
          cudaEvent_t ev = next_event(&e);
          if (!ev) return fail(DPIPE_ECUDA, "stage executor: cudaEventCreate failed");
          DPIPE_CUDA_CHECK(cudaEventRecord(ev, e.copy_stream));
          DPIPE_CUDA_CHECK(cudaStreamWaitEvent(cs, ev, 0));
          if ((rc = release(&e, DPIPE_CH_ACT_IN, in.micro_batch, cs))) return rc;
        }
        ++e.pc;
        break;
      }
      case DPIPE_OP_SEND_GRAD: {
        Channel& ch = e.ch[DPIPE_CH_GRAD_OUT];
        if (!ch.bound || !ch.laid_out) { *out = {DPIPE_OP_NEED_HANDSHAKE + DPIPE_CH_GRAD_OUT, in.buffer, in.micro_batch}; return 1; }
        if ((rc = push(&e, DPIPE_CH_GRAD_OUT, in.buffer, in.micro_batch, cs))) return rc;
        // the activation slot of this micro-batch has been fully consumed (its backward was enqueued before this send)
        if ((rc = release(&e, DPIPE_CH_ACT_IN, in.micro_batch, cs))) return rc;
        ++e.pc;
        break;
      }
      case DPIPE_OP_RECV_ACTIVATION:
      case DPIPE_OP_RECV_GRAD: {
        const int c = in.op == DPIPE_OP_RECV_ACTIVATION ? DPIPE_CH_ACT_IN : DPIPE_CH_GRAD_IN;
        Channel& ch = e.ch[c];
        if (!ch.bound || !ch.laid_out) { *out = {DPIPE_OP_NEED_HANDSHAKE + c, in.buffer, in.micro_batch}; return 1; }
        if ((rc = pull(&e, c, in.micro_batch, cs))) return rc;
        ++e.pc;
        *out = in;            // the host wraps the slot (dpipe_exec_recv_base) into its tensors: no device work, no copy
        e.pending = in;
        return 1;
      }
      default:
        ++e.pc;
        *out = in;
        e.pending = in;
        return 1;
    }
  }
  return 0;
}
