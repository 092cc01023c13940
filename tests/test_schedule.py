"""CPU: the C++ 1F1B planner is bit-exact with (a) traces produced by the reference's own generator
(tests/golden/schedule_traces.json, made by tests/golden/make_golden_schedule.py from utils/patches.py:113-160),
(b) the pure-Python oracle restatement, and (c) structural invariants at BASELINE sizes."""
import json
import os

import pytest

from diffusion_pipe_b200.pipe.schedule import InferenceSchedule, TrainSchedule
from oracle import schedule_ref


def _norm(ticks):
    return [[[c.name] + ([c.buffer_id] if hasattr(c, 'buffer_id') else []) for c in t] for t in ticks]


def test_matches_reference_generator_traces(golden_dir):
    with open(os.path.join(golden_dir, 'schedule_traces.json')) as f:
        golden = json.load(f)
    assert len(golden) >= 50
    for key, ticks in golden.items():
        m, s, st = (int(x) for x in key.split(','))
        assert _norm(TrainSchedule(m, s, st).steps()) == ticks, key


def test_oracle_restatement_matches_golden(golden_dir):
    with open(os.path.join(golden_dir, 'schedule_traces.json')) as f:
        golden = json.load(f)
    for key, ticks in golden.items():
        m, s, st = (int(x) for x in key.split(','))
        assert schedule_ref.train_schedule(m, s, st) == ticks, key


def test_appendix_c_known_answer():
    # SURVEY.md Appendix C, stage 1 of (M=4, S=3)
    names = ['+'.join(c.name + str(getattr(c, 'buffer_id', '')) for c in t) for t in TrainSchedule(4, 3, 1).steps()]
    assert names[1] == 'RecvActivation0+ForwardPass0'
    assert names[4] == 'RecvGrad0+SendActivation1+BackwardPass0'
    assert names[-1] == 'SendGrad1+ReduceTiedGrads+ReduceGrads+OptimizerStep'


@pytest.mark.parametrize('m,s', [(16, 8), (16, 4), (21, 8), (64, 8), (1, 8), (3, 2)])
def test_invariants_at_scale(m, s):
    sends = {}
    for st in range(s):
        sched = TrainSchedule(m, s, st)
        ticks = sched.steps()
        assert len(ticks) == 2 * (m + s - 1)
        assert _norm(ticks) == schedule_ref.train_schedule(m, s, st)
        fwd, bwd, inflight, peak = [], [], 0, 0
        for ti, t in enumerate(ticks):
            for c in t:
                if c.name == 'ForwardPass':
                    fwd.append(c.micro_batch_id); inflight += 1; peak = max(peak, inflight)
                elif c.name == 'BackwardPass':
                    bwd.append(c.micro_batch_id); inflight -= 1
                    assert c.micro_batch_id in fwd
                elif c.name in ('SendActivation', 'RecvActivation', 'SendGrad', 'RecvGrad'):
                    sends.setdefault((c.name, st, c.micro_batch_id), ti)
        assert fwd == list(range(m)) and bwd == list(range(m))
        assert peak <= sched.num_pipe_buffers()
    # every send is matched by the neighbour's receive in the same tick
    for (name, st, mb), ti in sends.items():
        if name == 'SendActivation':
            assert sends[('RecvActivation', st + 1, mb)] == ti
        if name == 'SendGrad':
            assert sends[('RecvGrad', st - 1, mb)] == ti


def test_inference_schedule():
    for m, s in [(1, 1), (4, 2), (5, 3), (9, 8)]:
        for st in range(s):
            ticks = InferenceSchedule(m, s, st).steps()
            assert len(ticks) == m + s - 1
            assert [c.micro_batch_id for t in ticks for c in t if c.name == 'ForwardPass'] == list(range(m))
            assert _norm(ticks) == schedule_ref.inference_schedule(m, s, st)


# ---------------------------------------------------------------------------------------------------------------------
# split-backward (zero-bubble) planner: not in the reference, so the checks are structural
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('m,s,infl', [(16, 8, 16), (16, 8, 8), (16, 4, 8), (4, 2, 4), (1, 3, 2), (5, 1, 2), (21, 8, 16)])
def test_zero_bubble_order_is_complete_and_consistent(m, s, infl):
    from diffusion_pipe_b200.pipe.schedule import ZeroBubbleSchedule
    for st in range(s):
        sched = ZeroBubbleSchedule(m, s, st, (13, 17, 10), infl)
        seq = [(c.name, getattr(c, 'micro_batch_id', None)) for t in sched.steps() for c in t]
        f = [mb for n, mb in seq if n == 'ForwardPass']
        b = [mb for n, mb in seq if n == 'BackwardInput']
        w = [mb for n, mb in seq if n == 'BackwardWeight']
        assert f == b == w == list(range(m))                      # every pass once, in micro-batch order
        pos = {x: i for i, x in enumerate(seq)}
        inflight = peak = 0
        for n, mb in seq:
            if n == 'ForwardPass':
                inflight += 1
                peak = max(peak, inflight)
            if n == 'BackwardInput':
                inflight -= 1
        assert peak <= infl
        for mb in range(m):
            assert pos[('ForwardPass', mb)] < pos[('BackwardInput', mb)] < pos[('BackwardWeight', mb)]
            if st > 0:
                assert pos[('RecvActivation', mb)] < pos[('ForwardPass', mb)]
                assert pos[('SendGrad', mb)] == pos[('BackwardInput', mb)] + 1
            if st < s - 1:
                assert pos[('SendActivation', mb)] == pos[('ForwardPass', mb)] + 1
                assert pos[('RecvGrad', mb)] == pos[('BackwardInput', mb)] - 1
        assert [n for n, _ in seq[-3:]] == ['ReduceTiedGrads', 'ReduceGrads', 'OptimizerStep']
    assert ZeroBubbleSchedule(m, s, 0, (13, 17, 10), infl).simulated_makespan() > 0   # the joint order never deadlocks


def test_zero_bubble_beats_1f1b_in_the_cost_model():
    from diffusion_pipe_b200.pipe.schedule import ZeroBubbleSchedule
    m, s, costs = 16, 8, (13, 17, 10)
    work = m * sum(costs)
    zb = ZeroBubbleSchedule(m, s, 0, costs, 2 * s).simulated_makespan()
    one_f_one_b = (m + s - 1) * sum(costs)
    assert zb < 0.85 * one_f_one_b
    assert work / zb > 0.86          # (S-1)*tf of fill is the only bubble left: 640 / (640 + 91)


@pytest.mark.parametrize('m,s,weights', [(16, 8, [8, 7, 7, 7, 7, 7, 7, 7]), (16, 8, [7, 7, 7, 7, 7, 7, 7, 8]),
                                         (16, 4, [15, 14, 14, 14]), (6, 3, [1, 5, 2])])
def test_zero_bubble_with_stage_weights_is_consistent_and_bounded(m, s, weights):
    """uneven partitions: the order stays complete, dependency-correct, memory-bounded and deadlock-free"""
    from diffusion_pipe_b200.pipe.schedule import ZeroBubbleSchedule
    infl = min(m, 2 * s)
    for st in range(s):
        sched = ZeroBubbleSchedule(m, s, st, (13, 17, 10), infl, weights)
        seq = [(c.name, getattr(c, 'micro_batch_id', None)) for t in sched.steps() for c in t]
        assert [mb for n, mb in seq if n == 'ForwardPass'] == list(range(m))
        assert [mb for n, mb in seq if n == 'BackwardInput'] == list(range(m))
        assert [mb for n, mb in seq if n == 'BackwardWeight'] == list(range(m))
        held = peak = 0
        for n, mb in seq:
            held += (n == 'ForwardPass') - (n == 'BackwardWeight')
            peak = max(peak, held)
        assert peak <= infl                                   # forward done, weight-gradient pending
    ms = ZeroBubbleSchedule(m, s, 0, (13, 17, 10), infl, weights).simulated_makespan()
    assert ms >= m * 40 * max(weights)                        # no stage can beat its own work
    with pytest.raises(ValueError):
        ZeroBubbleSchedule(m, s, 0, (13, 17, 10), infl, weights[:-1])
    with pytest.raises(ValueError):
        ZeroBubbleSchedule(m, s, 0, (13, 17, 10), infl, [0] + weights[1:])


def test_zero_bubble_prefers_the_heavier_stage_first():
    """57 blocks over 8 stages, 16 micro-batches: the simulated speed-up over one stage (7.125 is the bound set by the
    8-block stage) — what bench.py's flop_balanced_split relies on"""
    from diffusion_pipe_b200.pipe.schedule import ZeroBubbleSchedule
    work = 16 * 57 * 40

    def speedup(w):
        return work / ZeroBubbleSchedule(16, 8, 0, (13, 17, 10), 16, w).simulated_makespan()
    first, last = speedup([8] + [7] * 7), speedup([7] * 7 + [8])
    assert first > 7.0 and first > last + 0.5
    assert last > 5.57 * 57 / 64                              # still better than ideal 1F1B on the same partition


def test_zero_bubble_planner_randomised_validity():
    """300 random geometries (micro-batches, stages, per-stage weights, costs, memory bound): every order is complete,
    per-kind in micro-batch order, dependency-consistent across stages (the joint replay never deadlocks) and respects
    the held-micro-batch bound"""
    import random
    from diffusion_pipe_b200.pipe.schedule import ZeroBubbleSchedule
    rng = random.Random(20240922)
    for _ in range(300):
        s = rng.randint(1, 9)
        m = rng.randint(1, 24)
        infl = rng.randint(1, max(1, min(m, 2 * s + 2)))
        costs = (rng.randint(1, 30), rng.randint(1, 40), rng.randint(1, 30))
        weights = [rng.randint(1, 9) for _ in range(s)] if rng.random() < 0.7 else None
        orders = []
        for st in range(s):
            seq = [(c.name, getattr(c, 'micro_batch_id', None)) for t in ZeroBubbleSchedule(m, s, st, costs, infl, weights).steps() for c in t]
            for kind in ('ForwardPass', 'BackwardInput', 'BackwardWeight'):
                assert [mb for n, mb in seq if n == kind] == list(range(m)), (m, s, st, kind)
            held = peak = 0
            for n, _ in seq:
                held += (n == 'ForwardPass') - (n == 'BackwardWeight')
                peak = max(peak, held)
            assert peak <= infl, (m, s, infl, weights, st, peak)
            orders.append([(n[0] if n != 'BackwardWeight' else 'W', mb) for n, mb in seq if n in ('ForwardPass', 'BackwardInput', 'BackwardWeight')])
        # joint replay in pure Python: must complete (independent of the C++ replay used by simulated_makespan)
        done_f, done_b = set(), set()
        pos = [0] * s
        progress = True
        while progress:
            progress = False
            for st in range(s):
                while pos[st] < len(orders[st]):
                    kind, mb = orders[st][pos[st]]
                    if kind == 'F':
                        ok = st == 0 or (st - 1, mb) in done_f
                    elif kind == 'B':
                        ok = ((st, mb) in done_f) if st == s - 1 else ((st + 1, mb) in done_b)
                    else:
                        ok = (st, mb) in done_b
                    if not ok:
                        break
                    if kind == 'F':
                        done_f.add((st, mb))
                    elif kind == 'B':
                        done_b.add((st, mb))
                    pos[st] += 1
                    progress = True
        assert all(pos[st] == 3 * m for st in range(s)), ('deadlock', m, s, infl, costs, weights)
        assert ZeroBubbleSchedule(m, s, 0, costs, infl, weights).simulated_makespan() > 0
