"""Single-GPU end-to-end: the public engine API (OobleckEngine -> OobleckPipeline.train -> optimizer_step) against the
oracle, including AdamW + WarmupLR over several steps (loss curve within 1e-4, north_star)."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
if not torch.cuda.is_available():
    pytest.skip("needs CUDA", allow_module_level=True)

from oobleck_b200.execution.dataloader import SyntheticTokenDataset  # noqa: E402
from oobleck_b200.execution.engine import JobArguments, ModelArguments, OobleckArguments, OobleckEngine  # noqa: E402
from oracle import gpt2 as og  # noqa: E402
from oracle import optim as oo  # noqa: E402


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


def test_training_loss_curve_matches_oracle():
    margs = dict(n_embd=128, n_head=2, num_hidden_layers=2, n_positions=64, vocab_size=500)
    M, mb, steps = 3, 2, 5
    args = OobleckArguments(job=JobArguments(microbatch_size=mb, global_microbatch_size=mb * M, steps=steps),
                            model=ModelArguments(model_name="gpt2", model_tag="t", model_args=margs))
    ds = SyntheticTokenDataset(num_samples=256, seq_len=64, vocab_size=500)
    eng = OobleckEngine(0, 1, 1, None, args, dataset=ds)
    eng.initialize_distributed()
    eng.instantiate_pipelines(M)
    pipe = eng._pipeline
    assert pipe.is_first_stage() and pipe.is_last_stage()
    assert len(pipe.execution._layers) == 4 and pipe.train_schedule.num_pipe_buffers() == 2

    d = og.GPT2Dims(n_embd=128, n_head=2, n_layer=2, n_positions=64, vocab_size=500)
    olayers = og.build_layers(d)
    for ol, l in zip(olayers, pipe.execution._layers):
        og.load_flat_(ol, l.flat_param.cpu())
    flats = [og.flat_params(ol).clone() for ol in olayers]
    ms = [torch.zeros_like(f) for f in flats]
    vs = [torch.zeros_like(f) for f in flats]
    lrs = oo.lr_sequence(steps, warmup_min_lr=0)
    import itertools
    batches = list(itertools.islice(iter(pipe._dataloader), steps * M))  # not exhausted: same epoch after reset
    pipe.reset_iterator()

    prev_total, ref_prev = 0.0, 0.0
    for step in range(steps):
        # oracle step
        for ol in olayers:
            ol.zero_grad()
        ref_step = 0.0
        for b in batches[step * M:(step + 1) * M]:
            x = (b["input_ids"], b["attention_mask"], b["labels"])
            for ol in olayers:
                x = ol(*x)
            x[0].backward()
            ref_step += float(x[0])
        for i, ol in enumerate(olayers):
            oo.adamw_step_(flats[i], og.flat_grads(ol), ms[i], vs[i], step + 1, lrs[step])
            og.load_flat_(ol, flats[i])
        # engine step
        eng._train_step()
        total = float(pipe.execution.total_loss)
        got_step = total - prev_total
        prev_total = total
        assert abs(got_step - ref_step) < 1e-4 * abs(ref_step), (step, got_step, ref_step)
        assert pipe._global_step == step + 1
        assert all(b is None for b in pipe.pipe_buffers["inputs"]) and all(b is None for b in pipe.pipe_buffers["outputs"])
    for i, l in enumerate(pipe.execution._layers):
        err = (l.flat_param.cpu() - flats[i]).abs().max() / flats[i].abs().max()
        assert err < 1e-4, (i, float(err))
    # test_layer.py:125-136: optimizer state keys exist after a step
    st = pipe.execution._optimizer.state[pipe.execution._layers[1].flat_param]
    assert {"step", "exp_avg", "exp_avg_sq"} <= set(st)


def test_fb_overlap_gives_the_same_training():
    """Forward of micro-batch i+1 overlapped with backward of micro-batch i (two streams) must train identically
    (up to the float atomics of the embedding scatter)."""
    from oobleck_b200.execution import pipeline as P
    margs = dict(n_embd=128, n_head=2, num_hidden_layers=3, n_positions=64, vocab_size=500)
    M, mb, steps = 4, 2, 3
    curves = {}
    for on in (False, True):
        old = P.FB_OVERLAP
        P.FB_OVERLAP = on
        try:
            args = OobleckArguments(job=JobArguments(microbatch_size=mb, global_microbatch_size=mb * M, steps=steps),
                                    model=ModelArguments(model_name="gpt2", model_tag="t", model_args=margs))
            ds = SyntheticTokenDataset(num_samples=256, seq_len=64, vocab_size=500)
            eng = OobleckEngine(0, 1, 1, None, args, dataset=ds)
            eng.initialize_distributed()
            eng.instantiate_pipelines(M)
            losses = []
            for _ in range(steps):
                eng._train_step()
                losses.append(float(eng._pipeline.execution.total_loss.item()))
            params = torch.cat([l.flat_param.flatten() for l in eng._pipeline.execution._layers]).clone()
            curves[on] = (losses, params)
        finally:
            P.FB_OVERLAP = old
    (l0, p0), (l1, p1) = curves[False], curves[True]
    for a, b in zip(l0, l1):
        assert abs(a - b) <= 1e-6 * abs(a), (l0, l1)
    assert ((p0 - p1).abs().max() / p0.abs().max()).item() < 1e-5


def test_profiler_style_layer_calls_match_oracle():
    """The reference's profiler loop (planning/profiler.py:41-91, 272-274) over ``model.layers``: init_tensors, deep copy,
    ``.to("cuda")``, ``layer(*input)`` feeding each output tuple to the next layer.  The loss that comes out of the last
    layer must be the oracle's."""
    import copy

    from oobleck_b200.execution.layer import init_tensors
    from oobleck_b200.module.model import OobleckModel
    from oracle import gpt2 as og
    margs = dict(n_embd=128, n_head=2, n_layer=2, n_positions=64, vocab_size=503)
    model = OobleckModel("gpt2", {"input_ids": None, "attention_mask": None, "labels": None}, None, "t", margs)
    device = torch.device("cuda")
    for layer in model.layers:
        init_tensors(layer, device)
    batch = og.synthetic_batch(3, 64, 503)
    inp = tuple(t.detach().clone().to("cuda") for t in (batch["input_ids"], batch["attention_mask"], batch["labels"]))
    for layer in model.layers:
        gpu_layer = copy.deepcopy(layer).to("cuda")
        with torch.no_grad():
            out = gpu_layer(*inp)
            torch.cuda.synchronize()
        assert isinstance(out, tuple)
        inp = tuple(t.detach().clone() if isinstance(t, torch.Tensor) else t for t in out)
    d = og.GPT2Dims(n_embd=128, n_head=2, n_layer=2, n_positions=64, vocab_size=503)
    olayers = og.build_layers(d)
    for ol, l in zip(olayers, model.layers):
        og.load_flat_(ol, l.init_flat())
    x = (batch["input_ids"], batch["attention_mask"], batch["labels"])
    for ol in olayers:
        x = ol(*x)
    assert abs(float(inp[0]) - float(x[0])) < 1e-5 * abs(float(x[0]))
