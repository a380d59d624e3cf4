"""The reference's pipeline tests (tests/execution/test_pipeline.py) restated against this repo's classes, on CPU / gloo
with the oracle's torch layers as stage compute:

    TestSingleStagePipeline   :20-183   attributes, load_microbatch / forward / backward / optimizer_step
    TestMultiStagePipeline    :186-371  neighbours, send / recv in forward and backward, ``train()`` on 1 / 2 / 4 stages

Same call sequences and assertions; where the reference inspects torch internals that this engine replaced (``p.grad`` of
``nn.Parameter``s, ``optimizer.state[p]``), the equivalent attribute of the flat-state ``Layer`` contract is checked instead
(``flat_param.grad`` / ``flat_grad``, ``exp_avg`` / ``exp_avg_sq`` / ``opt_step``) and the substitution is noted in place.
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from test_pipeline_gloo import make_engine, run_spawn  # noqa: E402

TRAIN_BATCH_SIZE = 1          # tests/conftest.py:33
M = 4                         # GRADIENT_ACCUMULATION_STEP, tests/conftest.py:35


def _single_stage_pipeline():
    for k in ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        os.environ.pop(k, None)
    eng = make_engine(0, 1, 1, M, TRAIN_BATCH_SIZE, steps=1)
    eng.instantiate_pipelines(M)
    return eng, eng._pipeline


# ---- TestSingleStagePipeline ------------------------------------------------------------------------------------------
def test_single_stage_attributes_type():
    from oobleck_b200.execution.optimizer import WarmupLR
    eng, pipeline = _single_stage_pipeline()
    assert pipeline.communication.prev_rank is None
    assert pipeline.communication.next_rank is None
    # one rank: it executes every layer of the model (test_pipeline.py:33-34)
    assert len(pipeline.execution._layers) == len(eng._model.layers)
    assert [l.layer_id for l in pipeline.execution._layers] == list(range(len(eng._model.layers)))
    # AdamW + WarmupLR (:36-37): the stage optimizer is the layer class's AdamW; the scheduler is this repo's WarmupLR
    assert type(pipeline.execution._optimizer).__name__.endswith("AdamW")
    assert isinstance(pipeline.execution._lr_scheduler, WarmupLR)
    assert pipeline._global_step == 0
    for layer in pipeline.execution._layers:               # :40-41 (`is_cuda` there; the CPU checker's device here)
        assert layer._param_handle.flat_param.device == pipeline.device
        assert layer._group_size == 1
        assert layer._param_handle._sharding_strategy == "NO_SHARD"     # layer.py:124-125


def test_single_stage_load_microbatch():
    _, pipeline = _single_stage_pipeline()
    assert pipeline.pipe_buffers["inputs"][0] is None
    pipeline.execution.load_microbatch(buffer_id=0)
    buf = pipeline.pipe_buffers["inputs"][0]
    assert isinstance(buf, tuple) and len(buf) == 3        # input_ids, attention_mask, labels (pipeline.py:150-156)
    assert all(isinstance(t, torch.Tensor) for t in buf)
    assert all(t.shape[0] == TRAIN_BATCH_SIZE for t in buf)
    assert not any(t.requires_grad for t in buf)           # integer tensors never require grad (:143-146)


def test_single_stage_forward():
    _, pipeline = _single_stage_pipeline()
    pipeline.execution.load_microbatch(buffer_id=0)
    assert pipeline.pipe_buffers["outputs"][0] is None
    assert pipeline.execution._loss is None
    assert pipeline.execution.total_loss is None
    pipeline.execution.forward_pass(buffer_id=0)
    # last stage: no output to send; the loss and the running total are written instead (:100-106)
    assert pipeline.pipe_buffers["outputs"][0] is None
    assert pipeline.execution._loss is not None
    assert pipeline.execution.total_loss is not None
    assert float(pipeline.execution.total_loss) == pytest.approx(float(pipeline.execution._loss))


def test_single_stage_backward():
    _, pipeline = _single_stage_pipeline()
    pipeline.execution.load_microbatch(buffer_id=0)
    pipeline.execution.forward_pass(buffer_id=0)
    pipeline.pipe_buffers["outputs"][0] = torch.zeros(1)   # backward_pass must clear it (:118-121)
    # before the backward pass no gradient exists (:123-126; here: the flat gradient is still all zero)
    assert all(float(l.flat_grad.abs().max()) == 0.0 for l in pipeline.execution._layers)
    pipeline.execution.backward_pass(buffer_id=0)
    assert pipeline.pipe_buffers["outputs"][0] is None
    assert pipeline.execution._loss is None                # pipeline.py:238
    assert all(float(l.flat_grad.abs().max()) > 0.0 for l in pipeline.execution._layers)


def test_single_stage_optimizer_step():
    _, pipeline = _single_stage_pipeline()
    pipeline.execution.load_microbatch(buffer_id=0)
    pipeline.execution.forward_pass(buffer_id=0)
    pipeline.execution.backward_pass(buffer_id=0)
    layers = pipeline.execution._layers
    # the optimizer has no internal data yet (:153-155): step count 0, moments untouched
    assert all(l.opt_step == 0 for l in layers)
    assert all(float(l.exp_avg.abs().max()) == 0.0 and float(l.exp_avg_sq.abs().max()) == 0.0 for l in layers)
    before = [l.flat_param.clone() for l in layers]
    pipeline.execution.optimizer_step()
    # ... and has "step", "exp_avg", "exp_avg_sq" afterwards (:159-168)
    assert all(l.opt_step == 1 for l in layers)
    assert all(float(l.exp_avg.abs().max()) > 0.0 and float(l.exp_avg_sq.abs().max()) > 0.0 for l in layers)
    # WarmupLR as the reference constructs it publishes lr = 0 for the first optimizer step: parameters
    # do not move yet, and the gradient has been cleared for the next step's accumulation
    assert all(torch.equal(a, l.flat_param) for a, l in zip(before, layers))
    assert all(float(l.flat_grad.abs().max()) == 0.0 for l in layers)


# ---- TestMultiStagePipeline -------------------------------------------------------------------------------------------
def _four_stage_worker(rank, world, port, phases, q):
    """Every phase gets a pipeline of its own (the reference spawns a process group per case; one spawn serves all the
    cases of a world size here: starting the processes is most of a gloo test's time)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    try:
        out = {}
        for which in phases:
            out[which] = _run_phase(rank, world, which)
            if world > 1:
                dist.barrier()
        q.put((rank, out, None, None))
        if world > 1:
            dist.barrier()
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, None, None, traceback.format_exc()))
        raise


def _run_phase(rank, world, which):
    eng = make_engine(rank, world, 1, M, TRAIN_BATCH_SIZE, steps=1)
    eng.instantiate_pipelines(M)
    pipeline = eng._pipeline
    comm, ex = pipeline.communication, pipeline.execution
    last = world - 1
    out = None
    if which == "attributes":                                            # test_pipeline.py:188-216
        assert comm.prev_rank == (None if rank == 0 else rank - 1)
        assert comm.next_rank == (None if rank == last else rank + 1)
        assert len(ex._layers) < len(eng._model.layers)
        out = (len(ex._layers), len(eng._model.layers))
    elif which == "send_recv_in_forward":                                # :218-263
        assert pipeline.pipe_buffers["inputs"][0] is None
        assert pipeline.pipe_buffers["outputs"][0] is None
        assert comm.sent_activation_meta is False
        assert comm.activation_recv_buf is None
        assert comm.grad_recv_buf is None
        if rank == 0:
            ex.load_microbatch(buffer_id=0)
        else:
            comm.recv_activations(buffer_id=0)
        assert pipeline.pipe_buffers["inputs"][0] is not None
        ex.forward_pass(buffer_id=0)
        if rank < last:
            assert pipeline.pipe_buffers["outputs"][0] is not None
            comm.send_activations(buffer_id=0)
            assert ex._loss is None
            assert comm.sent_activation_meta is True
        else:
            assert pipeline.pipe_buffers["outputs"][0] is None
            assert ex._loss is not None
        if rank != 0:
            assert comm.activation_recv_buf is not None
            # the wire tuple: hidden states (fp32, requires grad) + the integer tensors that travel with them
            got = pipeline.pipe_buffers["inputs"][0]
            assert got[0].dtype == torch.float32 and got[0].requires_grad
            assert all(not t.requires_grad for t in got[1:])
    elif which == "send_recv_in_backward":                               # :265-311
        if rank == 0:
            ex.load_microbatch(buffer_id=0)
        else:
            comm.recv_activations(buffer_id=0)
        ex.forward_pass(buffer_id=0)
        if rank < last:
            comm.send_activations(buffer_id=0)
        assert comm.grad_recv_buf is None
        assert all(float(l.flat_grad.abs().max()) == 0.0 for l in ex._layers)
        if rank == last:
            ex.backward_pass(buffer_id=0)
            comm.send_gradients(buffer_id=0)
        elif rank > 0:
            comm.recv_gradients(buffer_id=0)
            assert comm.grad_recv_buf is not None
            ex.backward_pass(buffer_id=0)
            comm.send_gradients(buffer_id=0)
        else:
            comm.recv_gradients(buffer_id=0)
            assert comm.grad_recv_buf is not None
            ex.backward_pass(buffer_id=0)
        assert all(float(l.flat_grad.abs().max()) > 0.0 for l in ex._layers)
        if rank > 0:
            assert pipeline.pipe_buffers["inputs"][0] is None            # send_gradients frees the slot (:404)
    elif which == "pipeline_train":                                      # :322-342
        assert pipeline._global_step == 0
        assert ex._loss is None
        assert ex.total_loss is None
        pipeline.train()
        assert pipeline._global_step == 1
        assert ex._loss is None
        if pipeline.is_last_stage():
            assert ex.total_loss is not None
        for pipe_buffers in pipeline.pipe_buffers.values():
            assert all(x is None for x in pipe_buffers)
    else:
        raise AssertionError(which)
    return out


@pytest.mark.timeout(400)
def test_four_stages_attributes_send_recv_and_train():
    """TestMultiStagePipeline.test_attributes_type, test_distributed_execution[send_recv_in_forward / _backward] and
    test_pipeline_train[4stages] (test_pipeline.py:186-371), one after the other on the same four gloo ranks."""
    results = run_spawn(_four_stage_worker, 4, ["attributes", "send_recv_in_forward", "send_recv_in_backward",
                                                "pipeline_train"])
    assert len(results) == 4
    counts = [r[1]["attributes"] for r in results]
    assert sum(c[0] for c in counts) == counts[0][1]             # stage layer counts add up to the model (:213-216)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("num_stages", [1, 2], ids=["1stage", "2stages"])
def test_pipeline_train(num_stages):
    run_spawn(_four_stage_worker, num_stages, ["pipeline_train"])
