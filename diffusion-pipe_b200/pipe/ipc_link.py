"""IpcLink — the B200 stage-boundary transport: CUDA-IPC mailboxes + cudaMemcpyPeerAsync over NVLink + device-side
sequence flags (C ABI: dpipe_ipc_* / dpipe_peer_copy / dpipe_flag_*; csrc/p2p_ipc.cu).  No NCCL on the boundary.

Per direction (activations stage s -> s+1, gradients s+1 -> s) there is one channel:
  * the RECEIVER owns a mailbox  [ready flags | NSLOTS x slot]  in its own HBM and exports it once (re-exported only
    if a larger boundary tuple shows up: resolution buckets change shapes between steps, train.py:916);
  * the SENDER owns the matching `free` flag array.
  send  (sender's copy stream):   wait free[k] >= w-1 ; peer-copy every tensor into slot k ; ready[k] = w
  recv  (receiver's compute stream): wait ready[k] >= w ; the tensors ARE views of slot k (no second copy)
  release (receiver's compute stream, once the slot's consumer has run): free[k] = w
with k = micro_batch % NSLOTS and w the per-slot use count.  The host never blocks on the data path; it only exchanges
shapes / IPC handles over a gloo side group the first time after `reset_activation_shape()` (once per step, like the
reference's metadata handshake) — host metadata, no device synchronisation.

Replaces DeepSpeed's p2p send/recv of the boundary tuple (schedule at utils/patches.py:134-143).
"""
import ctypes
import os

import torch
import torch.distributed as tdist

from .. import _lib

_DTYPES = [torch.float32, torch.float16, torch.bfloat16, torch.int64, torch.int32, torch.int16, torch.int8, torch.uint8,
           torch.bool, torch.float64]
_DTYPE_ID = {d: i for i, d in enumerate(_DTYPES)}
_META_LEN = 192
_ALIGN = 256
_WAIT_TIMEOUT_S = float(os.environ.get('DPIPE_LINK_TIMEOUT_S', 300.0))   # device-side flag wait; 0 = wait forever


def _align(n):
    return (n + _ALIGN - 1) // _ALIGN * _ALIGN


class _DevMem:
    """exposes a raw device allocation to torch through __cuda_array_interface__"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {'shape': (nbytes,), 'typestr': '|u1', 'data': (ptr, False), 'version': 2}


def _lib_call(name, *args):
    _lib.check(getattr(_lib.lib(), name)(*args), name)


def _ipc_alloc(nbytes):
    ptr = ctypes.c_void_p()
    handle = (ctypes.c_ubyte * 64)()
    _lib_call('dpipe_ipc_alloc', nbytes, ctypes.byref(ptr), handle)
    return ptr.value, bytes(handle)


def _ipc_open(handle):
    ptr = ctypes.c_void_p()
    buf = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
    _lib_call('dpipe_ipc_open', buf, ctypes.byref(ptr))
    return ptr.value


def _layout(specs):
    """byte offset of every tensor inside a slot and the slot size"""
    offs, off = [], 0
    for dt, shape in specs:
        n = 1
        for s in shape:
            n *= s
        offs.append(off)
        off += _align(max(1, n * torch.empty((), dtype=dt).element_size()))
    return offs, off


class _Channel:
    def __init__(self, link, peer_rank, peer_device, sending, tag):
        self.link = link
        self.peer_rank, self.peer_device = peer_rank, peer_device
        self.sending = sending
        self.tag = tag
        self.nslots = link.nslots
        self.flag_bytes = _align(8 * self.nslots)
        self.slot_bytes = 0
        self.specs = None          # [(dtype, shape)] valid until the next reset
        self.offs = None
        self.count = [0] * self.nslots   # writes (sender) / reads (receiver) per slot
        self.handshaken = False
        if sending:
            self.free_ptr, self.free_handle = _ipc_alloc(self.flag_bytes)    # flow-control flags live with the sender
            self.remote_ptr = None                                            # receiver's mailbox, mapped
        else:
            self.mail_ptr = None                                              # own mailbox
            self.mail_handle = None
            self.mail_u8 = None
            self.remote_free = None                                           # sender's free flags, mapped

    # ---- host-side control plane (gloo) ----
    def _send_cpu(self, t, tag):
        tdist.send(t, self.peer_rank, group=self.link.ctrl_group, tag=self.tag + tag)

    def _recv_cpu(self, t, tag):
        tdist.recv(t, self.peer_rank, group=self.link.ctrl_group, tag=self.tag + tag)

    def _ensure_mailbox(self, slot_bytes):
        """receiver: (re)allocate the mailbox if the boundary tuple outgrew it; returns True if it changed"""
        if self.mail_ptr is not None and slot_bytes <= self.slot_bytes:
            return False
        if self.mail_ptr is not None:
            torch.cuda.synchronize()
            _lib_call('dpipe_ipc_free', ctypes.c_void_p(self.mail_ptr))
            self.count = [0] * self.nslots
        self.slot_bytes = _align(int(slot_bytes * 1.0))
        total = self.flag_bytes + self.nslots * self.slot_bytes
        self.mail_ptr, self.mail_handle = _ipc_alloc(total)
        self.mail_u8 = torch.as_tensor(_DevMem(self.mail_ptr, total), device=self.link.device)
        return True

    def _map_mailbox(self, handle, slot_bytes):
        if self.remote_ptr is not None:
            torch.cuda.synchronize()
            _lib_call('dpipe_ipc_close', ctypes.c_void_p(self.remote_ptr))
            self.count = [0] * self.nslots
            # the receiver starts the new mailbox with zeroed flags: restart our own flow-control flags as well
            self.link._zero(self.free_ptr, self.flag_bytes)
        self.remote_ptr = _ipc_open(handle)
        self.slot_bytes = slot_bytes

    # ---- activations: the sender describes the tuple ----
    def handshake_send_described(self, tensors):
        meta = torch.zeros(_META_LEN, dtype=torch.int64)
        vals = [len(tensors), 0 if self.handshaken else 1]
        for t in tensors:
            vals += [_DTYPE_ID[t.dtype], t.dim()] + list(t.shape)
        assert len(vals) <= _META_LEN, 'boundary tuple too complex for the metadata message'
        meta[:len(vals)] = torch.tensor(vals, dtype=torch.int64)
        self._send_cpu(meta, 0)
        if not self.handshaken:
            self._send_cpu(torch.frombuffer(bytearray(self.free_handle), dtype=torch.uint8), 1)
        reply = torch.zeros(4, dtype=torch.int64)
        self._recv_cpu(reply, 2)
        if int(reply[0]):
            h = torch.zeros(64, dtype=torch.uint8)
            self._recv_cpu(h, 3)
            self._map_mailbox(bytes(h.numpy().tobytes()), int(reply[1]))
        self.handshaken = True
        self.specs = [(t.dtype, tuple(t.shape)) for t in tensors]
        self.offs, _ = _layout(self.specs)

    def handshake_recv_described(self):
        meta = torch.zeros(_META_LEN, dtype=torch.int64)
        self._recv_cpu(meta, 0)
        m = meta.tolist()
        n, first = m[0], m[1]
        if first:
            h = torch.zeros(64, dtype=torch.uint8)
            self._recv_cpu(h, 1)
            self.remote_free = _ipc_open(bytes(h.numpy().tobytes()))
        specs, pos = [], 2
        for _ in range(n):
            dt, nd = _DTYPES[m[pos]], m[pos + 1]
            specs.append((dt, tuple(m[pos + 2:pos + 2 + nd])))
            pos += 2 + nd
        self.specs = specs
        self.offs, need = _layout(specs)
        changed = self._ensure_mailbox(need)
        self._send_cpu(torch.tensor([1 if changed else 0, self.slot_bytes, self.nslots, 0], dtype=torch.int64), 2)
        if changed:
            self._send_cpu(torch.frombuffer(bytearray(self.mail_handle), dtype=torch.uint8), 3)
        self.handshaken = True

    # ---- gradients: the receiver already knows the shapes (its own outputs) ----
    def handshake_recv_known(self, like):
        self.specs = [(t.dtype, tuple(t.shape)) for t in like]
        self.offs, need = _layout(self.specs)
        changed = self._ensure_mailbox(need)
        self._send_cpu(torch.tensor([1 if changed else 0, self.slot_bytes, self.nslots, 0 if self.handshaken else 1],
                                    dtype=torch.int64), 0)
        if changed:
            self._send_cpu(torch.frombuffer(bytearray(self.mail_handle), dtype=torch.uint8), 1)
        if not self.handshaken:
            h = torch.zeros(64, dtype=torch.uint8)
            self._recv_cpu(h, 2)
            self.remote_free = _ipc_open(bytes(h.numpy().tobytes()))
        self.handshaken = True

    def handshake_send_known(self, tensors):
        info = torch.zeros(4, dtype=torch.int64)
        self._recv_cpu(info, 0)
        if int(info[0]):
            h = torch.zeros(64, dtype=torch.uint8)
            self._recv_cpu(h, 1)
            self._map_mailbox(bytes(h.numpy().tobytes()), int(info[1]))
        if int(info[3]):
            self._send_cpu(torch.frombuffer(bytearray(self.free_handle), dtype=torch.uint8), 2)
        self.handshaken = True
        self.specs = [(t.dtype, tuple(t.shape)) for t in tensors]
        self.offs, _ = _layout(self.specs)

    # ---- data path ----
    def push(self, tensors, mb):
        link = self.link
        k = mb % self.nslots
        w = self.count[k] + 1
        s = link.copy_stream
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        s.wait_event(ev)
        sp = ctypes.c_void_p(s.cuda_stream)
        _lib_call('dpipe_flag_wait_geq', ctypes.c_void_p(self.free_ptr + 8 * k), w - 1, _WAIT_TIMEOUT_S, sp)
        base = self.remote_ptr + self.flag_bytes + k * self.slot_bytes
        for t, off in zip(tensors, self.offs):
            nbytes = t.numel() * t.element_size()
            if nbytes:
                _lib_call('dpipe_peer_copy', ctypes.c_void_p(base + off), self.peer_device, ctypes.c_void_p(t.data_ptr()),
                          link.device.index, nbytes, sp)
                t.record_stream(s)
        _lib_call('dpipe_flag_write', ctypes.c_void_p(self.remote_ptr + 8 * k), w, sp)
        self.count[k] = w

    def pull(self, mb):
        k = mb % self.nslots
        w = self.count[k] + 1
        sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib_call('dpipe_flag_wait_geq', ctypes.c_void_p(self.mail_ptr + 8 * k), w, _WAIT_TIMEOUT_S, sp)
        self.count[k] = w
        base = self.flag_bytes + k * self.slot_bytes
        out = []
        for (dt, shape), off in zip(self.specs, self.offs):
            n = 1
            for x in shape:
                n *= x
            nbytes = n * torch.empty((), dtype=dt).element_size()
            out.append(self.mail_u8[base + off: base + off + nbytes].view(dt).view(shape))
        return out

    def release(self, mb):
        k = mb % self.nslots
        sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib_call('dpipe_flag_write', ctypes.c_void_p(self.remote_free + 8 * k), self.count[k], sp)


class IpcLink:
    def __init__(self, engine, rank_devices=None):
        """rank_devices[r] = CUDA ordinal of global rank r (engine._make_link gathers them); None = ordinal follows the rank"""
        self.engine = engine
        self.device = engine.device
        assert self.device.type == 'cuda'
        grid = engine.grid
        # 1F1B keeps at most `stages` micro-batches in flight on a stage; the zero-bubble order holds up to
        # `zb_max_inflight` (default 2 x stages).  A receiver that may hold n un-released activations needs n slots:
        # with fewer, the sender would wait for a slot whose release is ordered AFTER the forward pass it is feeding.
        self.nslots = max(2, engine.num_stages)
        if engine.pipeline_schedule == 'zb':
            self.nslots = max(2 * self.nslots, int(engine.zb_max_inflight or 2 * engine.num_stages))
        self.copy_stream = torch.cuda.Stream(device=self.device)
        # host control plane: one gloo group per pipeline (every rank creates all of them, in the same order)
        self.ctrl_group = None
        for d in range(grid.data_parallel_size):
            ranks = [grid._topo.get_rank(s, d) for s in range(grid.pipe_parallel_size)]
            g = tdist.new_group(ranks=ranks, backend='gloo')
            if d == grid.data_parallel_id:
                self.ctrl_group = g
        # peers: same node, all GPUs visible in every process under the same ordinals (two stages may share one device:
        # the peer copy degenerates to a device-local copy — tests/test_stage_link_one_gpu.py)
        local_rank = self.device.index
        def peer(stage):
            r = grid.stage_to_global(stage)
            return r, (rank_devices[r] if rank_devices is not None else local_rank + (r - engine.global_rank))
        s = engine.stage_id
        self.act_out = _Channel(self, *peer(s + 1), sending=True, tag=100) if s + 1 < engine.num_stages else None
        self.act_in = _Channel(self, *peer(s - 1), sending=False, tag=100) if s > 0 else None
        self.grad_out = _Channel(self, *peer(s - 1), sending=True, tag=200) if s > 0 else None
        self.grad_in = _Channel(self, *peer(s + 1), sending=False, tag=200) if s + 1 < engine.num_stages else None
        self._fresh = {}
        self.reset()

    def _zero(self, ptr, nbytes):
        torch.as_tensor(_DevMem(ptr, nbytes), device=self.device).zero_()
        torch.cuda.synchronize()

    def reset(self):
        self._fresh = {'act_out': True, 'act_in': True, 'grad_out': True, 'grad_in': True}

    def send_activations(self, outputs, buf, mb, keep=True):
        ts = [t.contiguous() for t in outputs]
        if self._fresh['act_out']:
            self.act_out.handshake_send_described(ts)
            self._fresh['act_out'] = False
        elif self.act_out.specs != [(t.dtype, tuple(t.shape)) for t in ts]:
            # the slot layout on both sides comes from the handshake: a different tuple would be copied to wrong offsets
            raise RuntimeError('the boundary tuple changed shape or dtype since it was announced to the next stage: call '
                               'engine.reset_activation_shape() before a step whose micro-batches have new shapes '
                               '(train.py:916, train.py:181)')
        self.act_out.push(ts, mb)
        if not keep and self.act_in is not None:
            # forward-only schedule: outputs may alias pass-through tensors that live in our input slot, so the slot is
            # released only after the copy engine has read them
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
            torch.cuda.current_stream().wait_event(ev)
            self.act_in.release(mb)

    def recv_activations(self, buf, mb):
        if self._fresh['act_in']:
            self.act_in.handshake_recv_described()
            self._fresh['act_in'] = False
        return tuple(self.act_in.pull(mb))

    def send_grads(self, grads, buf, mb):
        ts = [g.contiguous() for g in grads]
        if self._fresh['grad_out']:
            self.grad_out.handshake_send_known(ts)
            self._fresh['grad_out'] = False
        self.grad_out.push(ts, mb)
        # the activation slot of this micro-batch has been fully consumed (its backward is enqueued before us)
        if self.act_in is not None:
            self.act_in.release(mb)

    def recv_grads(self, like, buf, mb):
        if self._fresh['grad_in']:
            self.grad_in.handshake_recv_known(like)
            self._fresh['grad_in'] = False
        return self.grad_in.pull(mb)

    def release_grads(self, buf, mb):
        if self.grad_in is not None:
            self.grad_in.release(mb)

    def release_activations(self, buf, mb):
        """forward-only schedules: the input slot is free as soon as the forward has been enqueued"""
        if self.act_in is not None:
            self.act_in.release(mb)

    def flush(self):
        pass
