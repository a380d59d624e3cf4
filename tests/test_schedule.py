"""Product schedule (oobleck_b200/execution/schedule.py) == golden vectors from the reference's steps()."""
import json
import os

from oobleck_b200.execution.schedule import OobleckPipelineSchedule

G = os.path.join(os.path.dirname(__file__), "golden")


def test_schedule_matches_reference_golden():
    cases = json.load(open(os.path.join(G, "schedule.json")))
    assert len(cases) >= 30
    for c in cases:
        s = OobleckPipelineSchedule(c["micro_batches"], c["stages"], c["stage_id"])
        assert s.num_pipe_buffers() == c["num_pipe_buffers"]
        got = [[[type(i).__name__, i.kwargs["buffer_id"]] for i in cmds] for cmds in s]
        assert got == c["steps"], (c["micro_batches"], c["stages"], c["stage_id"])
        # re-iterable every step (deepspeed PipeSchedule.__iter__)
        assert [[type(i).__name__ for i in cmds] for cmds in s.steps()] == [[n for n, _ in cmds] for cmds in c["steps"]]
