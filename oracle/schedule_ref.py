"""ORACLE — test infrastructure only.  Pure-Python restatement of the reference's pipeline schedules.

  train_schedule      follows utils/patches.py:113-160 (`train_schedule_steps`) line by line, with the helper
                      methods of deepspeed==0.18.4 runtime/pipe/schedule.py TrainSchedule (third-party, not vendored:
                      _step_to_micro_batch / _even_step_forward_id / _odd_step_forward_id / _even_step_backward_id /
                      _odd_step_backward_id / _valid_micro_batch / _valid_stage / _buffer_idx / num_pipe_buffers).
  inference_schedule  deepspeed InferenceSchedule.steps (same file), used by eval_batch (train.py:181-183).

Pinned by tests/golden/schedule_traces.json, which is produced by executing the reference's own generator text.
Instructions are returned as [name, buffer_id] / [name] lists, one inner list per schedule tick.
"""


def _step_to_micro_batch(step_id, stages, stage_id):
    even_step, even_stage = step_id % 2 == 0, stage_id % 2 == 0
    if even_step and even_stage:
        return step_id // 2 - stage_id // 2, True
    if not even_step and not even_stage:
        return (step_id - 1) // 2 - stage_id // 2, True
    if even_step and not even_stage:
        return step_id // 2 - stages + (stage_id + 1) // 2, False
    return ((step_id - 1) // 2) - stages + 1 + stage_id // 2, False


def num_pipe_buffers(micro_batches, stages, stage_id):
    return max(2, min(stages - stage_id, micro_batches))


def train_schedule(micro_batches, stages, stage_id):
    def valid_mb(m):
        return 0 <= m < micro_batches

    def valid_stage(s):
        return 0 <= s < stages

    nbuf = num_pipe_buffers(micro_batches, stages, stage_id)
    prev_stage, next_stage = stage_id - 1, stage_id + 1
    out = []
    prev_micro_batch_id = -1
    prev_buffer = curr_buffer = None
    total_steps = 2 * (micro_batches + stages - 1)
    for step_id in range(total_steps):
        micro_batch_id, is_forward = _step_to_micro_batch(step_id, stages, stage_id)
        if valid_mb(prev_micro_batch_id):
            prev_buffer = prev_micro_batch_id % nbuf
        if valid_mb(micro_batch_id):
            curr_buffer = micro_batch_id % nbuf
        cmds = []
        if stage_id == 0 or stage_id == stages - 1:
            if is_forward and valid_mb(micro_batch_id):
                cmds.append(['LoadMicroBatch', curr_buffer])
        if is_forward:
            if valid_mb(prev_micro_batch_id) and valid_stage(prev_stage):
                cmds.append(['SendGrad', prev_buffer])
            if valid_mb(micro_batch_id) and valid_stage(prev_stage):
                cmds.append(['RecvActivation', curr_buffer])
        else:
            if valid_mb(micro_batch_id) and valid_stage(next_stage):
                cmds.append(['RecvGrad', curr_buffer])
            if valid_mb(prev_micro_batch_id) and valid_stage(next_stage):
                cmds.append(['SendActivation', prev_buffer])
        if valid_mb(micro_batch_id):
            cmds.append(['ForwardPass' if is_forward else 'BackwardPass', curr_buffer])
        if step_id == total_steps - 1:
            cmds += [['ReduceTiedGrads'], ['ReduceGrads'], ['OptimizerStep']]
        prev_micro_batch_id = micro_batch_id
        out.append(cmds)
    return out


def inference_schedule(micro_batches, stages, stage_id):
    def valid_mb(m):
        return 0 <= m < micro_batches

    out = []
    total_steps = micro_batches + stages - 1
    even = stage_id % 2 == 0
    for step_id in range(total_steps):
        cmds = []
        micro_batch_id = step_id - stage_id
        if even:
            recv_buf, send_buf = step_id % 2, (step_id + 1) % 2
        else:
            recv_buf, send_buf = (step_id + 1) % 2, step_id % 2
        if stage_id == 0 or stage_id == stages - 1:
            if valid_mb(micro_batch_id):
                cmds.append(['LoadMicroBatch', recv_buf])
        send = 0 <= stage_id + 1 < stages and valid_mb(micro_batch_id - 1)
        recv = 0 <= stage_id - 1 < stages and valid_mb(micro_batch_id)
        if even:
            if send:
                cmds.append(['SendActivation', send_buf])
            if recv:
                cmds.append(['RecvActivation', recv_buf])
        else:
            if recv:
                cmds.append(['RecvActivation', recv_buf])
            if send:
                cmds.append(['SendActivation', send_buf])
        if valid_mb(micro_batch_id):
            cmds.append(['ForwardPass', recv_buf])
        out.append(cmds)
    return out
