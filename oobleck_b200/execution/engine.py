"""Engine layer above the pipeline: data-parallel gradient all-reduce, reconfiguration, the training loop.

Mirrors oobleck/execution/engine.py:

    DataParallelEngine       :363-412  per-(layer, fsdp_index) groups over heterogeneous pipelines, SUM all-reduce
    ReconfigurationEngine    :39-360   rank re-assignment policy (borrow / merge) + state copy after a loss
    OobleckEngine            :415-668  construct -> initialize_distributed -> instantiate_pipelines -> train

The policy code (which ranks go where) is bit-exact with the reference (golden vectors from the reference's own
``on_reconfigure``; tests/test_engine_bookkeeping.py).  The *mechanism* is what changed:
* DP communicators are created once per distinct rank set instead of once per (layer, fsdp_index)
  (engine.py:390-392 issues O(layers) collective ``new_group`` calls on every rank);
* pipelines are rebuilt without tearing down the world process group and without per-layer groups;
* the planner (csrc/planning, planning/instantiator.py) is control plane and not part of this package: templates
  are injected (``templates=``) or derived with ``planning.pipeline_template.balanced_template``.
"""
from __future__ import annotations

import copy
import os
import queue
import socket
import threading
import time
import weakref
from collections import defaultdict
from dataclasses import dataclass, field
from typing import Any

import torch
import torch.distributed as dist

from ..module.model import OobleckModel
from ..planning.pipeline_template import PipelineTemplate, balanced_template
from .dataloader import LoaderType, OobleckDataLoader, SyntheticTokenDataset
from .pipeline import OobleckPipeline, RankGroup
from .training_args import TrainingArguments


# ---- arguments (oobleck/elastic/training_util.py:7-39, plain dataclasses) ------------------------------------------
@dataclass
class DistributedArguments:
    master_ip: str = "127.0.0.1"
    master_port: int = 0
    node_ips: list[str] = field(default_factory=list)
    node_port: int = 22
    num_workers: int = 1
    num_agents_per_node: int = 1
    username: str | None = None


@dataclass
class JobArguments:
    fault_threshold: int = 3
    microbatch_size: int = 8
    global_microbatch_size: int = 128
    steps: int = 50


@dataclass
class ModelArguments:
    model_name: str = "gpt2"
    model_tag: str = "small"
    dataset_path: str = "wikitext"
    dataset_name: str | None = "wikitext-2-raw-v1"
    model_args: dict[str, Any] = field(default_factory=dict)


@dataclass
class OobleckArguments:
    dist: DistributedArguments = field(default_factory=DistributedArguments)
    job: JobArguments = field(default_factory=JobArguments)
    model: ModelArguments = field(default_factory=ModelArguments)


@dataclass
class DistributionInfo:
    """What the agent sends once on the worker pipe (oobleck/elastic/message_util.py:11-13).  The engine only reads
    ``agent_ips`` and ``world_size``: the control plane's own class is accepted as is (duck typing)."""
    agent_ips: list[str]
    world_size: int


class PipelineAborted(RuntimeError):
    """A train step was cut short because a peer was lost (transport abort or communicator abort)."""


def _rank() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def _my_ip() -> str:
    """engine.py:555 (``socket.gethostbyname(socket.gethostname())``; the reference's tests patch it)."""
    try:
        return socket.gethostbyname(socket.gethostname())
    except OSError:
        return "127.0.0.1"


# Communicators survive pipeline rebuilds: one per rank set for the life of the process.  (torch names group-local
# groups by a hash of their ranks, so creating the same set twice would also collide in the rendezvous store.)
_COMMUNICATORS: dict[tuple[int, ...], Any] = {}


# ---- data parallel -------------------------------------------------------------------------------------------------
class DataParallelEngine:
    def __init__(self, engine, pipelines: list[OobleckPipeline], new_group=None):
        self._engine = weakref.ref(engine)
        # layer_index -> fsdp_index -> ranks (engine.py:374-384)
        ranks_grid: dict[int, dict[int, list[int]]] = defaultdict(dict)
        for pipeline in pipelines:
            for layer_index, ranks in pipeline.rank_grid.items():
                assert isinstance(ranks, list) and len(ranks) == engine._num_gpus_per_node
                for fsdp_index, rank in enumerate(ranks):
                    ranks_grid[layer_index].setdefault(fsdp_index, []).append(rank)
        self._ranks_grid = ranks_grid
        # One communicator per DISTINCT rank set, created in first-use order (identical on every rank).  The reference
        # calls new_group once per (layer, fsdp_index) (:390-392); the mapping below is otherwise the same.
        my_rank = _rank()

        def _make(ranks):
            # group-local creation: only member ranks take part, so survivors can build communicators after a loss
            # without re-creating the world group (engine.py:539-593 tears the NCCL world down and re-inits it)
            if not dist.is_initialized() or len(ranks) <= 1 or my_rank not in ranks:
                return None
            key = tuple(sorted(ranks))
            if key not in _COMMUNICATORS:
                _COMMUNICATORS[key] = dist.new_group(list(key), use_local_synchronization=True)
            return _COMMUNICATORS[key]

        make = new_group or _make
        cache: dict[tuple[int, ...], Any] = {}
        self._dp_process_groups: dict[int, dict[int, RankGroup]] = defaultdict(dict)
        self._fsdp_indices: dict[int, list[int]] = defaultdict(list)
        self.group_creation_order: list[list[int]] = []
        for layer_index, per_layer in ranks_grid.items():
            for fsdp_index, ranks in per_layer.items():
                key = tuple(ranks)
                if key not in cache:
                    cache[key] = make(list(ranks))
                    self.group_creation_order.append(list(ranks))
                self._dp_process_groups[layer_index][fsdp_index] = RankGroup(list(ranks), my_rank, cache[key])
                if my_rank in ranks:
                    self._fsdp_indices[layer_index].append(fsdp_index)

    @property
    def engine(self):
        return self._engine()

    def do_allreduce(self):
        """engine.py:404-412: every local layer reduces over the groups that contain this rank."""
        for layer in self.engine._pipeline.execution._layers:
            process_groups = {fi: pg for fi, pg in self._dp_process_groups[layer.layer_id].items()
                              if pg.rank_index() >= 0}
            if process_groups and any(pg.size() > 1 for pg in process_groups.values()):
                layer.reduce_gradients(process_groups)


# ---- reconfiguration -----------------------------------------------------------------------------------------------
class ReconfigurationEngine:
    def __init__(self, engine, pipelines: list[OobleckPipeline], start_listener: bool = True):
        self._engine = weakref.ref(engine)
        self._pipelines = pipelines
        self._num_instances_set: dict[PipelineTemplate, int] = defaultdict(int)
        for pipeline in self._pipelines:
            self._num_instances_set[pipeline._template] += 1
        t0 = engine._pipeline_templates[0]
        self._min_num_ranks = t0._num_nodes * t0._num_gpus_per_node          # engine.py:46-49
        self.last_reconfiguration_seconds: float | None = None
        self.last_notification_time: float | None = None
        # lost-rank lists received by the listener thread, applied by the training thread at its next safe point
        self._pending: "queue.Queue[tuple[list[int], float]]" = queue.Queue()
        self._reconfiguration_listener = None
        if start_listener and getattr(engine, "_agent_pipe", None) is not None and engine._listen:
            self._reconfiguration_listener = threading.Thread(target=self._reconfiguration_listener_fn, daemon=True)
            self._reconfiguration_listener.start()

    @property
    def engine(self):
        return self._engine()

    def _reconfiguration_listener_fn(self):
        """Daemon thread (engine.py:50-61).  The reference rebuilds the pipelines from this thread while the training
        thread is inside a step, with no lock; here the listener only does what cannot wait -- the pipe round trip and
        releasing every stream / communicator that is stuck on the lost ranks -- and hands the re-planning to the
        training thread (``poll``), which applies it between two steps."""
        while True:
            try:
                engine = self.engine
                if engine is None:
                    return
                lost_node: str = engine._agent_pipe.recv()
            except (EOFError, ValueError, OSError):
                return                                          # connection closed (engine.py:78-80)
            t0 = time.perf_counter()
            try:
                lost_ranks = self.remove_lost_node_from_dist_info(lost_node)
                engine.on_ranks_lost(lost_ranks)                # first: un-wedge the GPU
                engine.initialize_distributed()                 # port round trip with the agent (engine.py:572-578)
            except (EOFError, ValueError, OSError):
                return
            self._pending.put((lost_ranks, t0))

    def poll(self) -> bool:
        """Training thread: apply every reconfiguration the listener has queued.  True if the pipelines changed."""
        changed = False
        while True:
            try:
                lost_ranks, t0 = self._pending.get_nowait()
            except queue.Empty:
                return changed
            self.last_notification_time = t0
            self.on_reconfigure(lost_ranks)
            self.last_reconfiguration_seconds = time.perf_counter() - t0
            changed = True

    def _on_receive_reconfiguration_notification(self) -> bool:
        """engine.py:63-80, synchronous form (the reference's tests call it directly,
        tests/execution/test_engine.py:967-1019).  Returns False when the agent pipe is closed."""
        try:
            engine = self.engine
            lost_node: str = engine._agent_pipe.recv()
            t0 = time.perf_counter()
            lost_ranks = self.remove_lost_node_from_dist_info(lost_node)
            engine.on_ranks_lost(lost_ranks)
            engine.initialize_distributed()
            self.last_notification_time = t0
            self.on_reconfigure(lost_ranks)
            self.last_reconfiguration_seconds = time.perf_counter() - t0
            return True
        except (EOFError, ValueError, OSError):
            return False

    def remove_lost_node_from_dist_info(self, lost_node_ip: str) -> list[int]:   # engine.py:82-89
        engine = self.engine
        assert getattr(engine, "_dist_info", None) is not None, "Distributed is not initialized yet."
        engine._dist_info.agent_ips.remove(lost_node_ip)
        engine._dist_info.world_size -= engine._num_gpus_per_node
        return engine._rank_map.pop(lost_node_ip)

    # -- policy (bit-exact with engine.py:91-166, 311-360) -----------------------------------------------------------
    def plan_new_ranks(self, lost_ranks: list[int]) -> list[list[int]]:
        for pipeline in self._pipelines:
            pipeline._ranks = [rank for rank in pipeline._ranks if rank not in lost_ranks]
        need_merge = False
        new_ranks_list: list[list[int]] = []
        for pipeline in self._pipelines:
            ranks = pipeline._ranks
            if len(ranks) == 0:
                continue                              # every rank of this pipeline is gone
            if len(ranks) >= self._min_num_ranks:
                new_ranks_list.append(ranks)          # some template still fits
                continue
            while len(ranks) < self._min_num_ranks:   # borrow from the biggest pipeline
                biggest = self._find_biggest_pipeline(self._pipelines)
                if biggest is None:
                    need_merge = True
                    break
                while len(biggest._ranks) > self._min_num_ranks and len(ranks) < self._min_num_ranks:
                    ranks.append(biggest._ranks.pop())
            new_ranks_list.append(ranks)
        if need_merge:
            new_ranks_list = self._merge_pipelines(new_ranks_list)
        for ranks in new_ranks_list:
            ranks.sort()
        new_ranks_list.sort(key=lambda ranks: (len(ranks), ranks[0]))
        return new_ranks_list

    def _merge_pipelines(self, ranks_list: list[list[int]]) -> list[list[int]]:
        to_merge = [r for r in ranks_list if len(r) < self._min_num_ranks]
        results = [r for r in ranks_list if len(r) >= self._min_num_ranks]
        try:
            while to_merge:
                ranks = to_merge.pop(0)
                try:
                    while len(ranks) < self._min_num_ranks:
                        ranks.extend(to_merge.pop(0))
                except IndexError:
                    ranks.extend(results.pop(0))      # nothing small left: absorb the first healthy pipeline
                assert len(ranks) >= self._min_num_ranks
                results.append(ranks)
        except IndexError:
            raise RuntimeError("Ranks are insufficient")
        return results

    def _find_biggest_pipeline(self, pipelines):
        biggest = None
        for pipeline in pipelines:
            if biggest is None or len(pipeline._ranks) >= len(biggest._ranks):
                biggest = pipeline
        if biggest is not None and len(biggest._ranks) > self._min_num_ranks:
            return biggest
        return None

    def on_reconfigure(self, lost_ranks: list[int]):
        def get_pipeline_template(ranks, templates):
            return next((t for t in templates if t._num_nodes * t._num_gpus_per_node == len(ranks)), None)

        old_rank_grids = [copy.deepcopy(pipeline.rank_grid) for pipeline in self._pipelines]
        new_ranks_list = self.plan_new_ranks(lost_ranks)
        new_num_instances_set: dict[PipelineTemplate, int] = defaultdict(int)
        for ranks in new_ranks_list:
            new_num_instances_set[get_pipeline_template(ranks, self.engine._pipeline_templates)] += 1
        new_pipeline = self._reinstantiate(new_num_instances_set, new_ranks_list)
        new_rank_grids = []
        remaining = list(new_ranks_list)
        for template, num_instance in new_num_instances_set.items():
            for _ in range(num_instance):
                new_rank_grids.append(template.get_rank_grid(remaining.pop(0)))
        self._copy_model_states(old_rank_grids, new_rank_grids, new_pipeline)
        for layer in self.engine._pipeline.execution._layers:        # engine.py:176-178
            if all(layer is not l for l in new_pipeline.execution._layers):
                layer.remove_tensors()
        self.engine._pipeline = new_pipeline

    # -- mechanism ---------------------------------------------------------------------------------------------------
    def _reinstantiate(self, num_instances_set, new_ranks_list) -> OobleckPipeline:
        engine = self.engine
        global_num_microbatch = engine._args.job.global_microbatch_size // engine._args.job.microbatch_size
        templates = [t for t, n in num_instances_set.items() for _ in range(n)]
        num_microbatches = engine.distribute_microbatches(templates, global_num_microbatch)
        my_index = next(i for i, ranks in enumerate(new_ranks_list) if engine._rank in ranks)
        sampler = engine._pipeline._dataloader.batch_sampler
        dataloader = OobleckDataLoader(engine._hf_training_args, engine._dataset, LoaderType.Training, my_index,
                                       num_microbatches, sampler.num_iterations_done, sampler.epoch,
                                       device_resident=engine._device_resident)
        pipelines, mine = [], None
        for pid, (template, ranks) in enumerate(zip(templates, new_ranks_list)):
            p = engine.make_pipeline(pid, template, list(ranks), dataloader, engine._pipeline._global_step)
            pipelines.append(p)
            if p.my_pipeline:
                mine = p
        for p in pipelines:
            p.initialize_distributed_fsdp()
            p.initialize_distributed_pipeline()
        mine.initialize_execution(engine._model, engine._pipeline)
        engine._dp_engine = DataParallelEngine(engine, pipelines)
        self._pipelines = pipelines
        return mine

    def _copy_model_states(self, old_rank_grids, new_rank_grids, new_pipeline: OobleckPipeline):
        """engine.py:238-309: per layer, some old owner whose rank list survives unchanged sends the flat parameters
        to the new owners.  Same sender choice and the same RuntimeError; the transfer itself is a broadcast inside the
        (cached) DP communicator of that layer."""
        engine = self.engine
        works = []
        for layer_index in range(len(old_rank_grids[0])):
            old_ranks = [g[layer_index] for g in old_rank_grids]
            new_ranks = [g[layer_index] for g in new_rank_grids]
            if all(rank in old_ranks for rank in new_ranks):
                continue
            alive = [ranks for ranks in old_ranks if ranks in new_ranks]
            if not alive:
                raise RuntimeError(f"No alive ranks for the layer {layer_index}. Terminating.")
            ranks_to_send = alive[0]
            my_rank = _rank()
            for ranks_recv in new_ranks:
                if my_rank not in ranks_recv:
                    continue
                fsdp_index = ranks_recv.index(my_rank)
                dp_group = engine._dp_engine._dp_process_groups[layer_index][fsdp_index]
                new_layer = next(l for l in new_pipeline.execution._layers if l.layer_id == layer_index)
                if my_rank == ranks_to_send[fsdp_index]:
                    old_layer = next(l for l in engine._pipeline.execution._layers if l.layer_id == layer_index)
                    if new_layer is not old_layer:
                        new_layer.load_flat_(old_layer.flat_param)
                if dp_group.group is not None:
                    works.append((dist.broadcast(new_layer.flat_param, src=ranks_to_send[fsdp_index],
                                                 group=dp_group.group, async_op=True), new_layer))
        for work, layer in works:
            work.wait()
            layer.refresh_planes()
        # no world barrier (engine.py:308): a lost rank can never join it; the broadcasts above are the only ordering
        if torch.cuda.is_available():
            torch.cuda.synchronize()


# ---- engine --------------------------------------------------------------------------------------------------------
def layer_cost_model(model: OobleckModel, microbatch: int) -> list[float]:
    """Relative fwd+bwd cost per stage layer (FLOPs): stand-in for the profiler's per-layer latencies
    (planning/profiler.py:41-123), used only to balance stages when no planner template is injected."""
    costs = []
    for spec in model.layers:
        E, T, V = spec.n_embd, spec.n_positions, spec.vocab_size
        if spec.kind == "block":
            costs.append(6.0 * 12 * E * E + 12.0 * T * E * 0.5)
        elif spec.kind == "head":
            costs.append(6.0 * E * V)
        else:
            costs.append(1e-3 * E)
    return costs


class OobleckEngine:
    """engine.py:415-668.  ``pipe`` is the agent connection (may be None when launched by torchrun; rank/world then
    come from the environment)."""

    def __init__(self, local_rank: int, num_nodes: int, num_gpus_per_node: int, pipe, args: OobleckArguments, *,
                 dataset=None, templates: list[PipelineTemplate] | None = None, nsplit: int = 3, layer_cls=None,
                 transport_cls=None, device_resident: bool = False):
        self._agent_pipe = pipe
        self._args = args
        self._hf_training_args = TrainingArguments(per_device_train_batch_size=args.job.microbatch_size,
                                                   max_steps=args.job.steps)
        self._local_rank = local_rank
        self._num_nodes = num_nodes
        self._num_gpus_per_node = num_gpus_per_node
        self._nsplit = nsplit
        self._layer_cls = layer_cls
        self._transport_cls = transport_cls
        self._device_resident = device_resident
        self._rank = 0
        self._world_size = num_nodes * num_gpus_per_node
        self.step_seconds: list[float] = []

        margs = dict(args.model.model_args)
        n_positions = margs.get("n_positions", 1024)
        self._model = OobleckModel(args.model.model_name, {"input_ids": None, "attention_mask": None, "labels": None},
                                   self._hf_training_args, args.model.model_tag, margs)
        self._dataset = dataset if dataset is not None else SyntheticTokenDataset(
            seq_len=n_positions, vocab_size=self._model.model_args.vocab_size)
        if templates is None:
            costs = layer_cost_model(self._model, args.job.microbatch_size)
            templates = [balanced_template(costs, n, num_gpus_per_node)
                         for n in range(1, num_nodes + 1) if n <= len(costs)]
        self._pipeline_templates = templates

    # -- distributed -------------------------------------------------------------------------------------------------
    def initialize_distributed(self, backend: str | None = None):
        """engine.py:526-596 without the agent round trip when launched by torchrun."""
        if dist.is_initialized():
            self._rank, self._world_size = dist.get_rank(), dist.get_world_size()
            return
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world > 1:
            backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            dist.init_process_group(backend=backend)
            self._rank, self._world_size = dist.get_rank(), dist.get_world_size()
        else:
            self._rank, self._world_size = 0, 1

    def on_ranks_lost(self, lost_ranks: list[int]):
        """Hook for the transport layer to drop peers (no world process-group teardown; engine.py:532-540 destroys and
        re-creates the NCCL world here)."""

    # -- planning stand-ins ------------------------------------------------------------------------------------------
    def distribute_microbatches(self, templates: list[PipelineTemplate], global_num_microbatch: int) -> list[int]:
        """Integer stand-in for ``PipelineInstantiator._distribute_batch`` (instantiator.py:254-329, pyomo MINLP,
        control plane): proportional to each pipeline's GPU count, remainder to the largest pipelines first."""
        sizes = [t._num_nodes * t._num_gpus_per_node for t in templates]
        total = sum(sizes)
        out = [global_num_microbatch * s // total for s in sizes]
        order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
        for i in range(global_num_microbatch - sum(out)):
            out[order[i % len(order)]] += 1
        return out

    def make_pipeline(self, pipeline_id, template, ranks, dataloader, step) -> OobleckPipeline:
        return OobleckPipeline(pipeline_id=pipeline_id, pipeline_template=template, ranks=ranks, dataloader=dataloader,
                               step=step, training_args=self._hf_training_args, layer_cls=self._layer_cls,
                               transport_cls=self._transport_cls, nsplit=self._nsplit)

    def choose_plan(self) -> list[PipelineTemplate]:
        """Largest template that divides the world evenly, replicated (every BASELINE config has this shape)."""
        world_nodes = self._world_size // self._num_gpus_per_node
        for t in sorted(self._pipeline_templates, key=lambda t: -t._num_nodes):
            if world_nodes % t._num_nodes == 0:
                return [t] * (world_nodes // t._num_nodes)
        raise RuntimeError("no pipeline template fits the world size")

    def instantiate_pipelines(self, global_num_microbatch: int, plan: list[PipelineTemplate] | None = None):
        """engine.py:600-643."""
        plan = plan or self.choose_plan()
        num_microbatches = self.distribute_microbatches(plan, global_num_microbatch)
        ranks_list, used = [], 0
        for t in plan:
            n = t._num_nodes * t._num_gpus_per_node
            ranks_list.append(list(range(used, used + n)))
            used += n
        my_index = next(i for i, ranks in enumerate(ranks_list) if self._rank in ranks)
        dataloader = OobleckDataLoader(self._hf_training_args, self._dataset, LoaderType.Training, my_index,
                                       num_microbatches, 0, 0, device_resident=self._device_resident)
        pipelines = []
        self._pipeline = None
        for pid, (t, ranks) in enumerate(zip(plan, ranks_list)):     # every rank builds every pipeline (:118-149)
            p = self.make_pipeline(pid, t, ranks, dataloader, 0)
            pipelines.append(p)
            if p.my_pipeline:
                self._pipeline = p
        for p in pipelines:
            p.initialize_distributed_fsdp()
            p.initialize_distributed_pipeline()
        self._pipeline.initialize_execution(self._model)
        assert self._pipeline.communication is not None and self._pipeline.execution is not None
        self._dp_engine = DataParallelEngine(self, pipelines)
        self._reconfiguration = ReconfigurationEngine(self, pipelines)

    # -- training ----------------------------------------------------------------------------------------------------
    def _train_step(self):
        """engine.py:645-649."""
        self._pipeline.train()
        self._dp_engine.do_allreduce()
        self._pipeline.execution.optimizer_step()

    def train(self):
        assert self._hf_training_args.max_steps > 0
        for _ in range(self._hf_training_args.max_steps):
            try:
                t0 = time.perf_counter()
                self._train_step()
                self.step_seconds.append(time.perf_counter() - t0)
            except StopIteration:
                self._pipeline.reset_iterator()                      # engine.py:660-663
        if dist.is_initialized():
            dist.barrier()
        if torch.cuda.is_available():
            torch.cuda.synchronize()
