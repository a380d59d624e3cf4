"""``worker_main`` itself -- the reference's unmodified oobleck/elastic/worker.py:13-34 -- launching this package's engine.

SURVEY 8(b), first row: the elastic agent starts every worker with ``worker_main(local_rank, num_nodes, num_gpus_per_node,
pipe, args)``, which constructs ``OobleckEngine(local_rank, num_nodes, num_gpus_per_node, pipe, args)`` and calls
``initialize_distributed()``, ``instantiate_pipelines(global_num_microbatch)``, ``train()``.  Here that very function is
imported from /root/reference and run in two processes under the reference's fake-agent harness
(tests/execution/test_engine.py:650-657), with

* ``args`` an instance of the REFERENCE's own ``OobleckArguments`` dataclasses (oobleck/elastic/training_util.py, also
  imported unmodified; ``simple_parsing.Serializable`` -- absent here -- is an empty base class for the purpose);
* ``oobleck.execution.engine`` bound to this package's engine (the import swap of INTEGRATION.md level 1a).  The engine runs
  its CPU checker configuration (gloo, the oracle's torch layers, synthetic tokens) because this container has no GPU: the
  subclass below only supplies those keyword defaults, every method is the product's.

The two-stage pipeline it trains must end where a single-process oracle run ends.
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
WORLD, M, STEPS = 2, 4, 3
IPS = ["127.0.0.1", "127.0.0.2"]


def bind_reference_worker():
    """Returns (worker_main, the reference's training_util module, the list engines are recorded in)."""
    import importlib
    import types
    from unittest.mock import MagicMock

    from oracle_layer import OracleLayer

    from oobleck_b200.execution import engine as our_engine
    from oobleck_b200.execution.dataloader import SyntheticTokenDataset
    created = []

    class OobleckEngine(our_engine.OobleckEngine):
        """Same positional signature as worker.py:23 passes; CPU checker defaults, nothing overridden."""

        def __init__(self, local_rank, num_nodes, num_gpus_per_node, pipe, args):
            dataset = SyntheticTokenDataset(num_samples=128, seq_len=32, vocab_size=211, pin_memory=False)
            super().__init__(local_rank, num_nodes, num_gpus_per_node, pipe, args, dataset=dataset,
                             layer_cls=OracleLayer, backend="gloo", comm_timeout_s=60)
            created.append(self)

    def package(name, path=None):
        m = types.ModuleType(name)
        m.__path__ = [path] if path else []
        sys.modules[name] = m
        return m

    package("oobleck", os.path.join(REF, "oobleck"))
    package("oobleck.elastic", os.path.join(REF, "oobleck", "elastic"))      # worker.py, training_util.py: the real files
    package("oobleck.execution")
    shim = types.ModuleType("oobleck.execution.engine")
    shim.OobleckEngine = OobleckEngine
    sys.modules["oobleck.execution.engine"] = shim
    sp = package("simple_parsing")
    sp.Serializable = type("Serializable", (), {})
    ds = package("deepspeed")
    ds.utils = package("deepspeed.utils")
    ds.utils.logging = types.ModuleType("deepspeed.utils.logging")
    ds.utils.logging.LoggerFactory = MagicMock(name="LoggerFactory")
    sys.modules["deepspeed.utils.logging"] = ds.utils.logging
    worker = importlib.import_module("oobleck.elastic.worker")
    tu = importlib.import_module("oobleck.elastic.training_util")
    assert worker.__file__.startswith(REF) and tu.__file__.startswith(REF)
    return worker.worker_main, tu, created


def process(rank, pipe, q):
    torch.set_num_threads(1)
    try:
        from unittest.mock import patch
        worker_main, tu, created = bind_reference_worker()
        patch("socket.gethostbyname", return_value=IPS[rank]).start()              # one "node" per process
        real_tcpstore = torch.distributed.TCPStore
        patch("torch.distributed.TCPStore", lambda host_name, *a, **kw: real_tcpstore("127.0.0.1", *a, **kw)).start()
        # worker.py:20 asserts a one-GPU view of the machine (CUDA_VISIBLE_DEVICES is set by the agent); none here
        patch("torch.cuda.device_count", return_value=1).start()
        patch("torch.cuda.current_device", return_value=0).start()
        args = tu.OobleckArguments(                                                 # the reference's own dataclasses
            dist=tu.DistributedArguments(master_ip="127.0.0.1", master_port=0, node_ips=list(IPS)),
            job=tu.JobArguments(microbatch_size=1, global_microbatch_size=M, steps=STEPS),
            model=tu.ModelArguments(model_name="gpt2", model_tag="t", dataset_path="synthetic", model_args=dict(MARGS)))
        worker_main(0, WORLD, 1, pipe, args)                                        # worker.py:13-34, unmodified
        (engine,) = created
        assert type(engine._args).__module__ == "oobleck.elastic.training_util"
        pipeline = engine._pipeline
        assert pipeline._global_step == STEPS and len(engine.step_seconds) == STEPS
        out = {l.layer_id: l.flat_param.numpy().copy() for l in pipeline.execution._layers}
        q.put((rank, out, [p._ranks for p in engine._reconfiguration._pipelines], None))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, None, None, traceback.format_exc()))
        raise


@pytest.mark.timeout(300)
def test_reference_worker_main_runs_this_engine():
    if not os.path.isfile(os.path.join(REF, "oobleck", "elastic", "worker.py")):
        pytest.skip("needs /root/reference")
    from oobleck_b200.execution.engine import DistributionInfo
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    pipes = [ctx.Pipe(duplex=True) for _ in range(WORLD)]
    procs = [ctx.Process(target=process, args=(r, pipes[r][1], q)) for r in range(WORLD)]
    for p in procs:
        p.start()

    def agent():                                                                    # test_engine.py:650-657
        for pipe, _ in pipes:
            pipe.send(DistributionInfo(list(IPS), WORLD))
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
    assert results[0][2] == [[0, 1]]                       # the largest template that fits: one 2-stage pipeline
    flats, _, _ = reference_run(M, 1, STEPS)
    seen = {}
    for _, out, _, _ in results:
        seen.update(out)
    assert sorted(seen) == [0, 1, 2, 3]
    for lid, f in seen.items():
        torch.testing.assert_close(torch.from_numpy(f), flats[lid], rtol=1e-5, atol=1e-7)
