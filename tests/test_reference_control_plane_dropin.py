"""The drop-in boundary exercised by the reference's OWN control-plane code (INTEGRATION.md level 1).

What runs unmodified from /root/reference in every worker of this test:

* the C++ planner -- ``PipelineTemplateGenerator.create_pipeline_templates`` on a layer profile, producing the reference's own
  ``PipelineTemplate`` / ``StageExecutionResult`` objects (oracle/_ref, built by ``make -C oracle``);
* ``oobleck/planning/instantiator.py``: ``PipelineInstantiator._enumerate_instantiation_options`` (paper section 4.2.1)
  and ``HeterogeneousPipelinesExecutionPlan`` -- ``my_pipeline_index``, ``num_microbatches``, ``instantiate(model,
  dataloader, training_args, num_gpus_per_node, step=0)`` (instantiator.py:103-152), imported with exactly the import swap
  INTEGRATION.md prescribes (``oobleck.execution.pipeline`` / ``dataloader`` / ``oobleck.module.model`` resolve to this
  package) and inert stand-ins for pyomo / deepspeed, which this image lacks (only ``_distribute_batch`` needs pyomo: the
  micro-batch split comes from this engine's integer stand-in);

and then the call sequence of ``OobleckEngine.instantiate_pipelines`` (engine.py:616-643) on what ``instantiate`` returned:
``initialize_distributed_fsdp`` / ``initialize_distributed_pipeline`` on every pipeline, ``initialize_execution(model)`` on
mine, ``DataParallelEngine(engine, pipelines)``, and ``_train_step`` (:645-649).  Stage compute is the oracle's torch layers
(gloo / CPU); the trained parameters must match a single-process run of the same model.
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from test_pipeline_gloo import MARGS, reference_run, run_spawn  # noqa: E402

REF = "/root/reference"
REF_SO_DIR = os.path.join(ROOT, "oracle", "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "oobleck", "planning")) and os.path.isdir(REF_SO_DIR) and \
        any(f.startswith("pipeline_template") and f.endswith(".so") for f in os.listdir(REF_SO_DIR))


def import_reference_instantiator():
    """``oobleck.planning.instantiator`` from /root/reference, unmodified, bound to this package."""
    import importlib
    import types
    from unittest.mock import MagicMock

    import oobleck_b200.execution.dataloader as our_dataloader
    import oobleck_b200.execution.pipeline as our_pipeline
    import oobleck_b200.module.model as our_model
    sys.path.insert(0, REF_SO_DIR)
    real_planner = importlib.import_module("pipeline_template")
    sys.path.remove(REF_SO_DIR)

    def package(name, path=None):
        m = types.ModuleType(name)
        m.__path__ = [path] if path else []
        sys.modules[name] = m
        return m

    # the reference's packages: real directories where its files are to be found, no __init__ side effects
    package("oobleck", os.path.join(REF, "oobleck"))
    package("oobleck.planning", os.path.join(REF, "oobleck", "planning"))
    package("oobleck.execution")
    package("oobleck.module")
    package("oobleck.csrc")
    package("oobleck.csrc.planning")
    sys.modules["oobleck.csrc.planning.pipeline_template"] = real_planner          # the reference's own C++ module
    sys.modules["oobleck.execution.pipeline"] = our_pipeline                        # the swap of INTEGRATION.md
    sys.modules["oobleck.execution.dataloader"] = our_dataloader
    sys.modules["oobleck.module.model"] = our_model
    # third-party modules this image lacks; instantiate() and the enumeration never touch them
    pyomo = package("pyomo")
    pyomo.environ = MagicMock(name="pyomo.environ")
    sys.modules["pyomo.environ"] = pyomo.environ
    ds = package("deepspeed")
    ds.comm = types.ModuleType("deepspeed.comm")
    ds.comm.get_rank = lambda *a, **k: dist.get_rank() if dist.is_initialized() else 0
    ds.utils = types.ModuleType("deepspeed.utils")
    ds.utils.logger = MagicMock(name="logger")
    sys.modules["deepspeed.comm"], sys.modules["deepspeed.utils"] = ds.comm, ds.utils
    return importlib.import_module("oobleck.planning.instantiator"), real_planner


def worker(rank, world, port, option, M, steps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    try:
        from oracle_layer import OracleLayer

        from oobleck_b200.execution.dataloader import LoaderType, OobleckDataLoader, SyntheticTokenDataset
        from oobleck_b200.execution.engine import DataParallelEngine, OobleckEngine, layer_cost_model
        from oobleck_b200.execution.training_args import TrainingArguments
        from oobleck_b200.module.model import OobleckModel
        dist.init_process_group("gloo")
        instantiator, R = import_reference_instantiator()
        assert instantiator.__file__.startswith(REF)                       # the reference's file, not a restatement
        assert instantiator.OobleckPipeline.__module__ == "oobleck_b200.execution.pipeline"

        training_args = TrainingArguments(per_device_train_batch_size=1, max_steps=steps)
        model = OobleckModel("gpt2", {"input_ids": None, "attention_mask": None, "labels": None}, training_args, "t",
                             dict(MARGS))
        dataset = SyntheticTokenDataset(num_samples=128, seq_len=32, vocab_size=211, pin_memory=False)

        # control plane, reference code: profile -> templates -> instantiation options -> execution plan
        costs = layer_cost_model(model, 1)
        profile = R.LayerExecutionResults([
            R.LayerExecutionResult(i, c / 3e6, 2 * c / 3e6, {1: 0.0}, {n + 1: 1e-3 for n in range(world)},
                                   (4 * l.num_params, l.activation_bytes(1))) for i, (c, l) in enumerate(zip(costs, model.layers))])
        templates = R.PipelineTemplateGenerator().create_pipeline_templates(profile, (1, world), 1)
        assert [t._num_nodes for t in templates] == list(range(1, world + 1))
        options = instantiator.PipelineInstantiator()._enumerate_instantiation_options(templates, world)
        by_size = {t._num_nodes: t for t in templates}
        want = {by_size[n]: k for n, k in option.items()}                  # e.g. {2-node template: 2}
        chosen = next(o for o in options if {t: k for t, k in o.items() if k} == want)
        # batch distribution (instantiator.py:254-329) is a pyomo MINLP: this engine's integer stand-in splits M
        order = [t for t in templates if chosen.get(t)]
        flat = [t for t in order for _ in range(chosen[t])]
        per_pipeline = OobleckEngine.distribute_microbatches(None, flat, M)
        num_microbatches_set = {t: per_pipeline[flat.index(t)] for t in order}
        plan = instantiator.HeterogeneousPipelinesExecutionPlan(templates, dict(chosen), num_microbatches_set,
                                                                [dict(l._allreduce_across_nodes) for l in profile.get()])
        assert plan.total_num_microbatches == M

        # engine.py:616-643, statement for statement
        dataloader = OobleckDataLoader(args=training_args, datasets=dataset, dataloader_type=LoaderType.Training,
                                       pipeline_index=plan.my_pipeline_index, num_microbatches=plan.num_microbatches,
                                       num_iterations_done=0, epoch=0)
        my_pipeline, pipelines = plan.instantiate(model=model, dataloader=dataloader, training_args=training_args,
                                                  num_gpus_per_node=1, step=0)
        assert type(my_pipeline).__module__ == "oobleck_b200.execution.pipeline" and my_pipeline.my_pipeline
        assert type(my_pipeline._template).__module__ == "pipeline_template"        # the reference's C++ object inside
        for pipeline in pipelines:
            pipeline._layer_cls, pipeline.device = OracleLayer, torch.device("cpu")  # CPU checker instead of the CUDA layers
            pipeline.initialize_distributed_fsdp()
            pipeline.initialize_distributed_pipeline()
        my_pipeline.initialize_execution(model)
        assert my_pipeline.communication is not None and my_pipeline.execution is not None
        class EngineSide:                      # what DataParallelEngine reads of its engine (engine.py:363-372)
            _num_gpus_per_node, _comm_timeout, _pipeline = 1, None, my_pipeline
        engine = EngineSide()
        dp_engine = DataParallelEngine(engine, pipelines)

        # the wiring the reference's planner objects produced
        stage_of = {r: i for i, r in enumerate(my_pipeline._ranks)}
        me = stage_of[rank]
        owned = [l.layer_id for l in my_pipeline.execution._layers]
        assert owned == list(my_pipeline._template.get_stages()[me]._layer_indices)
        assert my_pipeline.communication.prev_rank == (my_pipeline._ranks[me - 1] if me > 0 else None)
        assert my_pipeline.communication.next_rank == (my_pipeline._ranks[me + 1] if me + 1 < len(stage_of) else None)

        for _ in range(steps):                                              # engine.py:645-649
            my_pipeline.train()
            dp_engine.do_allreduce()
            my_pipeline.execution.optimizer_step()
        out = {l.layer_id: l.flat_param.numpy().copy() for l in my_pipeline.execution._layers}
        q.put((rank, out, [p._ranks for p in pipelines], None))
        dist.barrier()
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, None, None, traceback.format_exc()))
        raise


@pytest.mark.timeout(400)
@pytest.mark.parametrize("world,option,replicas", [(4, {2: 2}, 2), (3, {1: 1, 2: 1}, 2)],
                         ids=["two-2-stage-replicas", "heterogeneous-1+2-stages"])
def test_reference_planner_and_instantiator_drive_this_pipeline(world, option, replicas):
    if not available():
        pytest.skip("needs /root/reference and oracle/_ref (make -C oracle)")
    M, steps = 4, 2
    results = run_spawn(worker, world, option, M, steps)
    ranks = results[0][2]
    assert sorted(r for p in ranks for r in p) == list(range(world)) and len(ranks) == replicas
    if len({len(p) for p in ranks}) == 1:
        flats, _, _ = reference_run(M, 1, steps, num_pipelines=replicas)
    else:
        # heterogeneous pipelines take different shares of the global batch (here 1 + 3 of 4 micro-batches): the summed
        # gradient is that of the whole global batch all the same
        flats, _, _ = reference_run(M, 1, steps, num_pipelines=1)
    seen = set()
    for _, out, _, _ in results:
        for lid, f in out.items():
            seen.add(lid)
            torch.testing.assert_close(torch.from_numpy(f), flats[lid], rtol=1e-5, atol=1e-7)
    assert seen == {0, 1, 2, 3}
