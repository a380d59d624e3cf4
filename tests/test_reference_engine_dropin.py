"""The reference's whole engine -- oobleck/execution/engine.py, unmodified -- running on this package's execution layer.

Everything the control plane does on the way to a training step is the reference's own code, imported from /root/reference
in each of two gloo workers that are started through the reference's unmodified ``worker_main`` under its fake-agent harness:

    worker.py            worker_main                                                   (elastic/worker.py:13-34)
    engine.py            OobleckEngine.__init__ / _initialize_engine (profile -> minimum node count -> template
                         generation), initialize_distributed (pipe protocol, TCPStore), instantiate_pipelines,
                         DataParallelEngine (one new_group per layer and shard column; do_allreduce), ReconfigurationEngine
                         (constructed, listener thread started), _train_step, train
    instantiator.py      PipelineInstantiator.get_best_execution_plan, _enumerate_instantiation_options,
                         HeterogeneousPipelinesExecutionPlan.instantiate
    pipeline_template    the C++ planner (oracle/_ref): get_profile_results reading the profile files this test writes in the
                         reference profiler's own format, PipelineTemplateGenerator.create_pipeline_templates, get_rank_grid
    training_util.py / message_util.py / utils/timer.py

and underneath, bound by the import swap of INTEGRATION.md: this package's ``OobleckPipeline`` (schedule interpreter, wire
protocol, ``Layer`` contract, optimizer + LR schedule), ``OobleckDataLoader`` / ``OobleckSampler``, ``OobleckModel``.

Stand-ins, each for something this image cannot provide: deepspeed (``deepspeed.comm`` -> ``torch.distributed``, logger, timer),
pyomo (``_distribute_batch``, a MINLP: an even integer split), HF ``TrainingArguments`` (transformers 5.5 wants accelerate for
it), the wikitext download (``OobleckDataset`` -> the synthetic corpus), CUDA / NCCL (``backend="nccl"`` -> gloo, the oracle's
torch layers as stage compute, ``torch.cuda.synchronize`` a no-op).  The trained parameters must match a single-process run.
"""
import os
import sys
import threading

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from test_pipeline_gloo import MARGS, reference_run  # noqa: E402

REF = "/root/reference"
REF_SO_DIR = os.path.join(ROOT, "oracle", "_ref")
M, STEPS = 4, 2
IPS = ["127.0.0.1", "127.0.0.2", "127.0.0.3", "127.0.0.4"]
TAG = "b200_dropin_test"


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "oobleck", "execution", "engine.py")) and os.path.isdir(REF_SO_DIR) and \
        any(f.startswith("pipeline_template") and f.endswith(".so") for f in os.listdir(REF_SO_DIR))


def write_profile_files(allreduce_across_nodes=None):
    """What the reference's profiler leaves under /tmp/oobleck/profiles/<model>-<tag>/ and ``get_profile_results`` (C++,
    pipeline_template.cpp:26-79) reads back -- written by this package's ``save_profile_results``."""
    from oobleck_b200.execution.engine import layer_cost_model
    from oobleck_b200.module.model import OobleckModel
    from oobleck_b200.planning.pipeline_template import LayerExecutionResult, LayerExecutionResults
    from oobleck_b200.planning.profiler import save_profile_results
    model = OobleckModel("gpt2", {"input_ids": None}, None, TAG, dict(MARGS))
    costs = layer_cost_model(model, 1)
    results = LayerExecutionResults([
        LayerExecutionResult(i, c / 3e6, 2 * c / 3e6, {g + 1: 1e-4 * (g + 1) for g in range(8)},
                             allreduce_across_nodes or {n + 1: 1e-3 * (n + 1) for n in range(64)},
                             (4 * l.num_params, l.activation_bytes(1)))
        for i, (c, l) in enumerate(zip(costs, model.layers))])
    save_profile_results(results, "gpt2", TAG, 1)


def bind_reference():
    """Install the bindings described in the module docstring; return the reference's worker_main and training_util."""
    import dataclasses
    import importlib
    import logging
    import types
    from unittest.mock import MagicMock, patch

    import torch.distributed as tdist
    from oracle_layer import OracleLayer

    import oobleck_b200.execution.dataloader as our_dataloader
    import oobleck_b200.execution.pipeline as our_pipeline
    import oobleck_b200.module.model as our_model
    from oobleck_b200.execution.dataloader import SyntheticTokenDataset
    from oobleck_b200.execution.training_args import TrainingArguments as OurTrainingArguments

    def package(name, path=None):
        m = types.ModuleType(name)
        m.__path__ = [path] if path else []
        sys.modules[name] = m
        return m

    # ---- the reference's own files -------------------------------------------------------------------------------------
    for name in ("oobleck", "oobleck.elastic", "oobleck.planning", "oobleck.utils", "oobleck.execution"):
        package(name, os.path.join(REF, *name.split(".")))
    package("oobleck.csrc")
    package("oobleck.csrc.planning")
    sys.path.insert(0, REF_SO_DIR)
    sys.modules["oobleck.csrc.planning.pipeline_template"] = importlib.import_module("pipeline_template")
    sys.path.remove(REF_SO_DIR)

    # ---- this package underneath ------------------------------------------------------------------------
    class OobleckPipeline(our_pipeline.OobleckPipeline):
        """CPU checker defaults for the keyword-only extras; everything else is the product's class."""

        def __init__(self, *a, **kw):
            kw.setdefault("layer_cls", OracleLayer)
            super().__init__(*a, **kw)

    pipeline_shim = types.ModuleType("oobleck.execution.pipeline")
    pipeline_shim.OobleckPipeline = OobleckPipeline
    sys.modules["oobleck.execution.pipeline"] = pipeline_shim
    sys.modules["oobleck.execution.dataloader"] = our_dataloader
    package("oobleck.module")
    sys.modules["oobleck.module.model"] = our_model
    dataset_shim = types.ModuleType("oobleck.execution.dataset")
    dataset_shim.OobleckDataset = lambda model_name, dataset_path, dataset_name, max_seq_length=None: \
        SyntheticTokenDataset(num_samples=128, seq_len=max_seq_length or 32, vocab_size=211, pin_memory=False)
    sys.modules["oobleck.execution.dataset"] = dataset_shim

    # ---- third parties this image lacks ------------------------------------------------------------------------------
    ds = package("deepspeed")
    comm = types.ModuleType("deepspeed.comm")
    comm.__getattr__ = lambda name: getattr(tdist, name)             # get_rank, new_group, broadcast, barrier, ...
    comm.init_distributed = lambda *a, **k: None
    # _copy_model_states issues ``broadcast(..., async_op=True)`` and never waits (engine.py:301-306): on NCCL the CUDA
    # stream orders it before everything that follows, gloo has no stream -- the stand-in completes it before returning
    comm.broadcast = lambda tensor, src, group=None, async_op=False: tdist.broadcast(tensor, src, group=group)
    comm.cdb = None
    ds.comm = comm
    sys.modules["deepspeed.comm"] = comm
    utils = package("deepspeed.utils")
    utils.logger = logging.getLogger("oobleck")
    lg = types.ModuleType("deepspeed.utils.logging")
    lg.LoggerFactory = types.SimpleNamespace(create_logger=lambda name=None, **k: logging.getLogger(str(name)))
    lg.log_dist = lambda *a, **k: None
    sys.modules["deepspeed.utils.logging"] = lg

    class SynchronizedWallClockTimer:
        class Timer:
            def start(self): pass
            def stop(self, *a, **k): pass
            def reset(self): pass

        def __call__(self, name):
            return self.Timer()

        def log(self, *a, **k): pass

        @staticmethod
        def memory_usage():
            return ""
    tm = types.ModuleType("deepspeed.utils.timer")
    tm.SynchronizedWallClockTimer = SynchronizedWallClockTimer
    sys.modules["deepspeed.utils.timer"] = tm
    sp = package("simple_parsing")
    sp.Serializable = type("Serializable", (), {})
    pyomo = package("pyomo")
    pyomo.environ = MagicMock(name="pyomo.environ")
    sys.modules["pyomo.environ"] = pyomo.environ

    @dataclasses.dataclass
    class HFTrainingArguments(OurTrainingArguments):     # engine.py:430-440 passes these two on top of ours
        no_cuda: bool = False
        log_level: str = "passive"
    package("transformers")
    ta = types.ModuleType("transformers.training_args")
    ta.TrainingArguments = HFTrainingArguments
    sys.modules["transformers.training_args"] = ta

    # ---- no GPU here -----------------------------------------------------------------------------------------------------
    real_init = tdist.init_process_group
    patch("torch.distributed.init_process_group",
          lambda backend=None, **kw: real_init(backend="gloo", **kw)).start()             # engine.py:588 says "nccl"
    patch("torch.cuda.synchronize", lambda *a, **k: None).start()                          # engine.py:309, 667
    patch("torch.cuda.get_device_properties", lambda *a, **k: types.SimpleNamespace(total_memory=180 * 2 ** 30)).start()
    patch("torch.cuda.device_count", return_value=1).start()                               # worker.py:20
    patch("torch.cuda.current_device", return_value=0).start()

    instantiator = importlib.import_module("oobleck.planning.instantiator")

    def even_split(self, global_num_microbatch, num_instances_set):
        """Stand-in for the pyomo MINLP (instantiator.py:254-329): the same number of micro-batches for every pipeline
        when that divides the global batch, None (option infeasible, like the reference's own failure path) otherwise."""
        pipelines = sum(num_instances_set.values())
        if pipelines == 0 or global_num_microbatch % pipelines:
            return None
        return {t: global_num_microbatch // pipelines for t in num_instances_set}
    instantiator.PipelineInstantiator._distribute_batch = even_split

    worker = importlib.import_module("oobleck.elastic.worker")
    tu = importlib.import_module("oobleck.elastic.training_util")
    engine = importlib.import_module("oobleck.execution.engine")
    for mod in (worker, tu, engine, instantiator):
        assert mod.__file__.startswith(REF), mod.__file__
    return worker, tu, engine


def process(rank, WORLD, pipe, q):
    torch.set_num_threads(1)
    try:
        from unittest.mock import patch
        worker, tu, engine_mod = bind_reference()
        patch("socket.gethostbyname", return_value=IPS[rank]).start()
        real_tcpstore = torch.distributed.TCPStore
        patch("torch.distributed.TCPStore", lambda host_name, *a, **kw: real_tcpstore("127.0.0.1", *a, **kw)).start()
        created = []
        real_engine_init = engine_mod.OobleckEngine.__init__

        def recording_init(self, *a, **kw):
            created.append(self)
            return real_engine_init(self, *a, **kw)
        patch.object(engine_mod.OobleckEngine, "__init__", recording_init).start()

        args = tu.OobleckArguments(
            dist=tu.DistributedArguments(master_ip="127.0.0.1", master_port=0, node_ips=list(IPS[:WORLD])),
            job=tu.JobArguments(microbatch_size=1, global_microbatch_size=M, steps=STEPS),
            model=tu.ModelArguments(model_name="gpt2", model_tag=TAG, dataset_path="synthetic", model_args=dict(MARGS)))
        worker.worker_main(0, WORLD, 1, pipe, args)
        (engine,) = created
        assert type(engine).__module__ == "oobleck.execution.engine" and type(engine).__name__ == "OobleckEngine"
        assert type(engine._dp_engine).__module__ == "oobleck.execution.engine"          # the reference's DP engine
        assert type(engine._pipeline).__mro__[1].__module__ == "oobleck_b200.execution.pipeline"
        assert [t._num_nodes for t in engine._pipeline_templates] == list(range(1, WORLD + 1))   # C++ planner, min..max nodes
        pipeline = engine._pipeline
        assert pipeline._global_step == STEPS
        out = {l.layer_id: l.flat_param.numpy().copy() for l in pipeline.execution._layers}
        q.put((rank, out, [p._ranks for p in engine._reconfiguration._pipelines], None))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, None, None, traceback.format_exc()))
        raise


@pytest.mark.timeout(300)
@pytest.mark.parametrize("WORLD", [2, 3])
def test_reference_engine_trains_on_this_execution_layer(WORLD):
    """2 workers: the reference's ``get_best_execution_plan`` picks two single-stage replicas (its DataParallelEngine reduces
    this package's layers); 3 workers with 4 micro-batches: three replicas cannot split the batch evenly, so the plan is
    either heterogeneous (1 + 2 stages) or one 3-stage pipeline -- inter-stage transfers under the reference's engine."""
    if not available():
        pytest.skip("needs /root/reference and oracle/_ref (make -C oracle)")
    from oobleck_b200.execution.engine import DistributionInfo
    write_profile_files()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    pipes = [ctx.Pipe(duplex=True) for _ in range(WORLD)]
    procs = [ctx.Process(target=process, args=(r, WORLD, pipes[r][1], q)) for r in range(WORLD)]
    for p in procs:
        p.start()

    def agent():                                                                    # test_engine.py:650-657
        for pipe, _ in pipes:
            pipe.send(DistributionInfo(list(IPS[:WORLD]), WORLD))
        port = pipes[0][0].recv()
        for pipe, _ in pipes:
            pipe.send(port)

    t = threading.Thread(target=agent, daemon=True)
    t.start()
    results = sorted((q.get(timeout=240) for _ in range(WORLD)), key=lambda r: r[0])
    t.join(timeout=30)
    for p in procs:
        p.join(timeout=60)
    for r in results:
        assert r[3] is None, r[3]
    pipelines = results[0][2]
    assert sorted(r for p in pipelines for r in p) == list(range(WORLD))
    if WORLD == 3:
        assert max(len(p) for p in pipelines) >= 2          # at least one multi-stage pipeline
    print("execution plan chosen by the reference's get_best_execution_plan:", pipelines)
    flats, _, _ = reference_run(M, 1, STEPS, num_pipelines=len(pipelines))
    covered = set()
    for _, out, _, _ in results:
        for lid, f in out.items():
            covered.add(lid)
            torch.testing.assert_close(torch.from_numpy(f), flats[lid], rtol=1e-5, atol=1e-7)
    assert covered == {0, 1, 2, 3}
