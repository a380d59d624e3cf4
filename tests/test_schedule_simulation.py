"""Executes the per-stage 1F1B programs of ``OobleckPipelineSchedule`` against each other on a small abstract machine.

deepspeed's ``TrainSchedule`` (the index math behind oobleck/execution/pipeline.py:24-84) is not in this container, so the
golden tables under tests/golden pin the restatement only against the reference's own wrapper.  This test pins what the
math must *achieve*, for every (micro-batches, stages) in a grid, whatever the formulas look like:

* with the reference's blocking, unbuffered point-to-point transfers (NCCL send / recv pairs rendezvous; pipeline.py:
  270-286) the P programs run to completion -- no deadlock, every send meets the matching receive of its neighbour in
  the same order;
* every stage forwards and backwards every micro-batch exactly once, a backward only after its own forward, stage s + 1
  sees micro-batch m's activation only after stage s computed it, stage s sees m's gradient only after stage s + 1
  computed it;
* a pipe buffer is never overwritten while the micro-batch it holds still needs it (inputs live from load / receive to
  the end of that micro-batch's backward + SendGrad; outputs from forward to backward), i.e. ``num_pipe_buffers`` is
  enough;
* the first and the last stage -- and only they -- load micro-batches (pipeline.py:158-167).
"""
import itertools

import pytest

from oobleck_b200.execution.schedule import (BackwardPass, ForwardPass, LoadMicroBatch, OobleckPipelineSchedule,
                                             RecvActivation, RecvGrad, SendActivation, SendGrad)


class Stage:
    def __init__(self, M, P, s):
        self.sched = OobleckPipelineSchedule(M, P, s)
        self.s, self.P, self.M = s, P, M
        self.cmds = [c for step in self.sched.program for c in step]
        self.pc = 0
        nbuf = self.sched.num_pipe_buffers()
        self.inputs = [None] * nbuf        # micro-batch whose stage input sits in the slot
        self.outputs = [None] * nbuf       # micro-batch whose stage output sits in the slot
        self.grads = [None] * nbuf         # micro-batch whose output gradient has been received for the slot
        self.next_load = 0                 # the data iterator hands out micro-batches in order
        self.forwarded, self.backwarded = [], []
        self.input_grad_ready = [None] * nbuf

    def done(self):
        return self.pc == len(self.cmds)

    def current(self):
        return self.cmds[self.pc]


def run(M, P):
    stages = [Stage(M, P, s) for s in range(P)]
    first, last = stages[0], stages[-1]
    while not all(st.done() for st in stages):
        progressed = False
        for st in stages:
            while not st.done():
                cmd = st.current()
                b = cmd.buffer_id
                if isinstance(cmd, LoadMicroBatch):
                    assert st is first or st is last, "only the edge stages load"
                    if st is first:
                        assert st.inputs[b] is None, f"stage {st.s}: input slot {b} overwritten while in use"
                        st.inputs[b] = st.next_load
                        st.next_load += 1
                    # the last stage's load is a no-op: its labels arrive with the activations (pipeline.py:163-167)
                elif isinstance(cmd, ForwardPass):
                    m = st.inputs[b]
                    assert m is not None, f"stage {st.s}: forward on an empty slot {b}"
                    assert m == len(st.forwarded), "micro-batches are forwarded in order"
                    st.forwarded.append(m)
                    if st is not last:
                        assert st.outputs[b] is None, f"stage {st.s}: output slot {b} overwritten while in use"
                        st.outputs[b] = m
                elif isinstance(cmd, BackwardPass):
                    m = st.inputs[b]
                    assert m is not None and m in st.forwarded, f"stage {st.s}: backward before forward"
                    assert m == len(st.backwarded), "micro-batches are backwarded in order"
                    if st is not last:
                        assert st.outputs[b] == m and st.grads[b] == m, f"stage {st.s}: backward without its gradient"
                        st.outputs[b] = None
                        st.grads[b] = None
                    st.backwarded.append(m)
                    if st is first:
                        st.inputs[b] = None            # nothing to send back: the slot is free again
                    else:
                        st.input_grad_ready[b] = m
                elif isinstance(cmd, (SendActivation, SendGrad, RecvActivation, RecvGrad)):
                    # blocking rendezvous: both sides must stand at the matching instruction
                    if isinstance(cmd, SendActivation):
                        peer, want = stages[st.s + 1], RecvActivation
                    elif isinstance(cmd, RecvActivation):
                        peer, want = stages[st.s - 1], SendActivation
                    elif isinstance(cmd, SendGrad):
                        peer, want = stages[st.s - 1], RecvGrad
                    else:
                        peer, want = stages[st.s + 1], SendGrad
                    if peer.done() or not isinstance(peer.current(), want):
                        break                                          # wait for the neighbour
                    snd, rcv = (st, peer) if isinstance(cmd, (SendActivation, SendGrad)) else (peer, st)
                    sb, rb = snd.current().buffer_id, rcv.current().buffer_id
                    if isinstance(snd.current(), SendActivation):
                        m = snd.outputs[sb]
                        assert m is not None and m in snd.forwarded
                        assert rcv.inputs[rb] is None, f"stage {rcv.s}: input slot {rb} overwritten while in use"
                        rcv.inputs[rb] = m
                    else:
                        m = snd.input_grad_ready[sb]
                        assert m is not None and snd.inputs[sb] == m, f"stage {snd.s}: SendGrad without a gradient"
                        snd.input_grad_ready[sb] = None
                        snd.inputs[sb] = None                          # send_gradients frees the slot (pipeline.py:404)
                        assert rcv.outputs[rb] == m, f"stage {rcv.s}: gradient of {m} lands on slot of {rcv.outputs[rb]}"
                        assert rcv.grads[rb] is None
                        rcv.grads[rb] = m
                    snd.pc += 1
                    rcv.pc += 1
                    progressed = True
                    continue
                else:
                    raise AssertionError(cmd)
                st.pc += 1
                progressed = True
        assert progressed, ("deadlock: " + ", ".join(
            f"stage {st.s} at {st.current() if not st.done() else 'end'}" for st in stages))
    for st in stages:
        assert st.forwarded == list(range(M)) and st.backwarded == list(range(M))
        assert all(x is None for x in st.inputs + st.outputs + st.grads + st.input_grad_ready)
    return stages


GRID = [(M, P) for P in range(1, 9) for M in (1, 2, 3, 4, 5, 7, 8, 16)] + [(64, 8), (64, 7), (32, 4), (128, 2)]


@pytest.mark.parametrize("M,P", GRID)
def test_programs_run_to_completion_with_blocking_transfers(M, P):
    run(M, P)


def test_in_flight_micro_batches_match_the_buffer_count():
    """1F1B keeps at most ``stages - stage_id`` micro-batches in flight on a stage (and at least 2 slots)."""
    for M, P in [(64, 8), (4, 4), (3, 6)]:
        for s in range(P):
            sched = OobleckPipelineSchedule(M, P, s)
            live, worst = 0, 0
            for cmd in itertools.chain.from_iterable(sched.program):
                if isinstance(cmd, ForwardPass):
                    live += 1
                elif isinstance(cmd, BackwardPass):
                    live -= 1
                worst = max(worst, live)
            assert worst == min(P - s, M)
            assert sched.num_pipe_buffers() == max(2, worst)


def test_program_length_and_idle_steps():
    for M, P in [(4, 2), (64, 8), (5, 3)]:
        for s in range(P):
            prog = OobleckPipelineSchedule(M, P, s).program
            assert len(prog) == 2 * (M + P - 1)                        # pipeline.py:36
            busy = sum(1 for step in prog if any(isinstance(c, (ForwardPass, BackwardPass)) for c in step))
            assert busy == 2 * M


def test_compute_order_is_the_canonical_1f1b():
    """An independent statement of 1F1B that IS in this image: torch.distributed.pipelining.Schedule1F1B (schedules.py,
    ``_step_microbatches``): ``warmup_chunks = min(n_microbatches, num_stages - stage_index)`` forwards, then one backward
    and one forward in turn while forwards remain, then the remaining backwards.  The per-stage order of forward / backward
    passes of ``OobleckPipelineSchedule`` (deepspeed's ``TrainSchedule`` restated) must be exactly that."""
    import inspect

    from torch.distributed.pipelining import schedules
    src = inspect.getsource(schedules.Schedule1F1B)
    assert "self._num_stages - self._stage.stage_index" in src      # the rule quoted above is the installed torch's
    for M, P in GRID:
        for s in range(P):
            got = "".join("F" if isinstance(c, ForwardPass) else "B"
                          for c in itertools.chain.from_iterable(OobleckPipelineSchedule(M, P, s).program)
                          if isinstance(c, (ForwardPass, BackwardPass)))
            warmup = min(M, P - s)
            want = "F" * warmup + "BF" * (M - warmup) + "B" * warmup
            assert got == want, (M, P, s)
