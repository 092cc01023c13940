"""Pipeline schedules.  The instruction stream comes from the C++ planner in libdpipe_b200.so
(csrc/sched.cpp, C ABI dpipe_sched_train / dpipe_sched_infer); this module only wraps it in the instruction
classes whose names the reference imports from deepspeed.runtime.pipe.schedule (utils/patches.py:14-17)."""
import ctypes

from .. import _lib


class _Instr(ctypes.Structure):
    _fields_ = [('op', ctypes.c_int32), ('buffer', ctypes.c_int32), ('micro_batch', ctypes.c_int32)]


class PipeInstruction:
    def __init__(self, **kwargs):
        self.name = self.__class__.__name__
        self.kwargs = kwargs
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __repr__(self):
        args = ', '.join(f'{k}={v}' for k, v in self.kwargs.items() if k != 'micro_batch_id')
        return f'{self.name}({args})'

    def __eq__(self, other):
        return type(self) is type(other) and self.kwargs.get('buffer_id') == other.kwargs.get('buffer_id')

    def __hash__(self):
        return hash((self.name, self.kwargs.get('buffer_id')))


class BufferOpInstruction(PipeInstruction):
    def __init__(self, buffer_id, **kwargs):
        super().__init__(buffer_id=buffer_id, **kwargs)


class OptimizerStep(PipeInstruction):
    pass


class ReduceGrads(PipeInstruction):
    pass


class ReduceTiedGrads(PipeInstruction):
    pass


class LoadMicroBatch(BufferOpInstruction):
    pass


class ForwardPass(BufferOpInstruction):
    pass


class BackwardPass(BufferOpInstruction):
    pass


class SendActivation(BufferOpInstruction):
    pass


class RecvActivation(BufferOpInstruction):
    pass


class SendGrad(BufferOpInstruction):
    pass


class RecvGrad(BufferOpInstruction):
    pass


class BackwardInput(BufferOpInstruction):
    """split backward: input-gradient pass of one micro-batch (weight gradients are queued)"""


class BackwardWeight(BufferOpInstruction):
    """split backward: run the queued weight-gradient work of one micro-batch"""


_OP_CLASSES = {1: LoadMicroBatch, 2: SendActivation, 3: RecvActivation, 4: SendGrad, 5: RecvGrad, 6: ForwardPass,
               7: BackwardPass, 8: ReduceTiedGrads, 9: ReduceGrads, 10: OptimizerStep, 11: BackwardInput, 12: BackwardWeight}


def _raw_plan(fn_name, micro_batches, stages, stage_id, extra=()):
    """the planner's instruction array as it comes out of the C ABI (ctypes array of dpipe_instr)"""
    lib = _lib.lib()
    fn = getattr(lib, fn_name)
    n = fn(micro_batches, stages, stage_id, *extra, None, 0)
    _lib.check(0 if n >= 0 else n, fn_name)
    buf = (_Instr * n)()
    n2 = fn(micro_batches, stages, stage_id, *extra, buf, n)
    _lib.check(0 if n2 == n else -1, fn_name)
    return buf


def make_instruction(op, buffer, micro_batch):
    """the instruction object for one dpipe_instr (names as deepspeed.runtime.pipe.schedule, utils/patches.py:14-17)"""
    if 8 <= op <= 10:
        return _OP_CLASSES[op]()
    return _OP_CLASSES[op](buffer, micro_batch_id=micro_batch)


def _plan(fn_name, micro_batches, stages, stage_id, extra=()):
    buf = _raw_plan(fn_name, micro_batches, stages, stage_id, extra)
    ticks, cur = [], []
    for ins in buf:
        if ins.op == 0:
            ticks.append(cur)
            cur = []
        elif 8 <= ins.op <= 10:
            cur.append(_OP_CLASSES[ins.op]())
        else:
            cur.append(_OP_CLASSES[ins.op](ins.buffer, micro_batch_id=ins.micro_batch))
    return ticks


class PipeSchedule:
    def __init__(self, micro_batches, stages, stage_id):
        self.micro_batches = micro_batches
        self.stages = stages
        self.stage_id = stage_id
        self.prev_stage = stage_id - 1
        self.next_stage = stage_id + 1

    @property
    def is_first_stage(self):
        return self.stage_id == 0

    @property
    def is_last_stage(self):
        return self.stage_id == self.stages - 1

    def __iter__(self):
        return iter(self.steps())


class TrainSchedule(PipeSchedule):
    """1F1B with the reference's LoadMicroBatch-before-communication patch (utils/patches.py:113-160)."""

    def steps(self):
        return _plan('dpipe_sched_train', self.micro_batches, self.stages, self.stage_id)

    def raw(self):
        return _raw_plan('dpipe_sched_train', self.micro_batches, self.stages, self.stage_id)

    def num_pipe_buffers(self):
        n = _lib.lib().dpipe_sched_num_pipe_buffers(self.micro_batches, self.stages, self.stage_id)
        _lib.check(0 if n > 0 else n, 'dpipe_sched_num_pipe_buffers')
        return n


class InferenceSchedule(PipeSchedule):
    def steps(self):
        return _plan('dpipe_sched_infer', self.micro_batches, self.stages, self.stage_id)

    def raw(self):
        return _raw_plan('dpipe_sched_infer', self.micro_batches, self.stages, self.stage_id)

    def num_pipe_buffers(self):
        return 2


class ZeroBubbleSchedule(PipeSchedule):
    """Split-backward schedule from the C++ list-scheduling planner (csrc/sched.cpp: dpipe_sched_zb).  Not in the
    reference; loss-equivalent to TrainSchedule.  `costs` = relative (forward, input-grad, weight-grad) durations;
    `max_inflight` = micro-batches a stage may hold between forward and weight-grad pass, i.e. the bound on
    activation memory (default 2 * stages: ZB-2p-like); `stage_weights` = relative amount of work per stage (e.g. its
    number of transformer blocks; default: equal) — the planner charges stage s `costs * stage_weights[s]`."""

    def __init__(self, micro_batches, stages, stage_id, costs=(13, 17, 10), max_inflight=None, stage_weights=None):
        super().__init__(micro_batches, stages, stage_id)
        self.costs = tuple(int(c) for c in costs)
        self.max_inflight = int(max_inflight or 2 * stages)
        if stage_weights is not None:
            stage_weights = [int(w) for w in stage_weights]
            if len(stage_weights) != stages or min(stage_weights) < 1:
                raise ValueError(f'stage_weights must be {stages} positive integers, got {stage_weights}')
        self.stage_weights = stage_weights

    def _weights_arg(self):
        import ctypes
        if self.stage_weights is None:
            return None
        return (ctypes.c_int * self.stages)(*self.stage_weights)

    def steps(self):
        return _plan('dpipe_sched_zb_ex', self.micro_batches, self.stages, self.stage_id,
                     (*self.costs, self.max_inflight, self._weights_arg()))

    def raw(self):
        return _raw_plan('dpipe_sched_zb_ex', self.micro_batches, self.stages, self.stage_id,
                         (*self.costs, self.max_inflight, self._weights_arg()))

    def num_pipe_buffers(self):
        return self.micro_batches     # buffers are indexed by micro-batch id

    def simulated_makespan(self):
        v = _lib.lib().dpipe_sched_zb_makespan_ex(self.micro_batches, self.stages, *self.costs, self.max_inflight,
                                                  self._weights_arg())
        if v < 0:
            raise _lib.DpipeError(f'dpipe_sched_zb_makespan failed ({v})')
        return v
