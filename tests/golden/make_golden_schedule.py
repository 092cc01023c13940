"""Generates tests/golden/schedule_traces.json by running the REFERENCE's own 1F1B generator.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_schedule.py

The function `train_schedule_steps` is taken verbatim (source text, via ast) from /root/reference/utils/patches.py
(lines 113-160) and executed against a stand-in `self` that provides the DeepSpeed TrainSchedule helper methods
(deepspeed==0.18.4 runtime/pipe/schedule.py, not installed here; restated).  The instruction classes are
lightweight stand-ins that only record (name, buffer_id).
"""
import ast
import json
import os

REF = '/root/reference/utils/patches.py'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'schedule_traces.json')


def load_reference_generator():
    src = open(REF).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'train_schedule_steps')
    code = compile(ast.Module(body=[fn], type_ignores=[]), REF, 'exec')
    ns = {}

    def mk(name, has_buf=True):
        class I:
            def __init__(self, *a):
                self.rec = [name] + ([a[0]] if has_buf else [])
        I.__name__ = name
        return I
    for n in ('LoadMicroBatch', 'SendGrad', 'RecvActivation', 'RecvGrad', 'SendActivation', 'ForwardPass', 'BackwardPass'):
        ns[n] = mk(n)
    for n in ('ReduceTiedGrads', 'ReduceGrads', 'OptimizerStep'):
        ns[n] = mk(n, False)
    exec(code, ns)
    return ns['train_schedule_steps']


class DeepSpeedTrainScheduleHelpers:
    """deepspeed.runtime.pipe.schedule.TrainSchedule helper methods (restated)."""

    def __init__(self, micro_batches, stages, stage_id):
        self.micro_batches, self.stages, self.stage_id = micro_batches, stages, stage_id
        self.prev_stage, self.next_stage = stage_id - 1, stage_id + 1

    def _valid_micro_batch(self, m):
        return 0 <= m < self.micro_batches

    def _valid_stage(self, s):
        return 0 <= s < self.stages

    def num_pipe_buffers(self):
        return max(2, min(self.stages - self.stage_id, self.micro_batches))

    def _buffer_idx(self, m):
        assert self._valid_micro_batch(m)
        return m % self.num_pipe_buffers()

    def _step_to_micro_batch(self, step_id):
        even_step, even_stage = step_id % 2 == 0, self.stage_id % 2 == 0
        if even_step and even_stage:
            return step_id // 2 - self.stage_id // 2, True
        if not even_step and not even_stage:
            return (step_id - 1) // 2 - self.stage_id // 2, True
        if even_step and not even_stage:
            return step_id // 2 - self.stages + (self.stage_id + 1) // 2, False
        return ((step_id - 1) // 2) - self.stages + 1 + self.stage_id // 2, False


def main():
    gen = load_reference_generator()
    cases = [(1, 1), (4, 1), (1, 2), (2, 2), (4, 2), (4, 3), (16, 4), (16, 8), (3, 8), (21, 8), (5, 3), (7, 5), (2, 4)]
    out = {}
    for m, s in cases:
        for st in range(s):
            helper = DeepSpeedTrainScheduleHelpers(m, s, st)
            ticks = [[c.rec for c in cmds] for cmds in gen(helper)]
            out[f'{m},{s},{st}'] = ticks
    with open(OUT, 'w') as f:
        json.dump(out, f, separators=(',', ':'))
    print('wrote', OUT, len(out), 'traces')


if __name__ == '__main__':
    main()
