"""The reference's own dataloader tests (tests/execution/test_dataloader.py:16-254) restated against this repo's
``OobleckSampler`` / ``OobleckDataLoader``: same constants (tests/conftest.py:33-35), same assertions.  The reference runs
every case in 1 or 4 spawned processes only to obtain RANK / WORLD_SIZE; the sampler itself never communicates, so the
"ranks" are plain loop indices here.  ``test_stop_iteration`` is skipped upstream ("Too long to run" on wikitext); on the
synthetic corpus it takes milliseconds and runs.
"""
import pytest
import torch

from oobleck_b200.execution.dataloader import LoaderType, OobleckDataLoader, OobleckSampler, SyntheticTokenDataset
from oobleck_b200.execution.training_args import TrainingArguments

TRAIN_BATCH_SIZE = 1
EVAL_BATCH_SIZE = 2
GRADIENT_ACCUMULATION_STEP = 4
WORLD = 4


@pytest.fixture(scope="module")
def dataset():
    return SyntheticTokenDataset(num_samples=203, seq_len=16, vocab_size=97, pin_memory=False)


def get_dataloader(dataset, pipeline_index, num_microbatches, num_iterations=0, shuffle=False) -> OobleckDataLoader:
    # tests/conftest.py:233-251 OobleckDynamicClassFactory.get_dataloader: Training loader, epoch 0, shuffle=False
    args = TrainingArguments(per_device_train_batch_size=TRAIN_BATCH_SIZE, per_device_eval_batch_size=EVAL_BATCH_SIZE)
    return OobleckDataLoader(args, dataset, LoaderType.Training, pipeline_index, num_microbatches, num_iterations, 0,
                             shuffle=shuffle)


@pytest.mark.parametrize("num_iterations", [0, 14])
def test_attributes_type(dataset, num_iterations):
    for rank in range(WORLD):
        loader = get_dataloader(dataset, rank, [GRADIENT_ACCUMULATION_STEP] * WORLD, num_iterations)
        sampler = loader.batch_sampler
        assert isinstance(sampler, OobleckSampler)
        assert sampler.microbatch_size == TRAIN_BATCH_SIZE
        assert sampler.num_iterations_done == num_iterations
        assert len(sampler.num_microbatches) == WORLD
        assert sampler.num_microbatches[rank] == GRADIENT_ACCUMULATION_STEP


def test_eval_loader_uses_the_eval_batch_size(dataset):
    args = TrainingArguments(per_device_train_batch_size=TRAIN_BATCH_SIZE, per_device_eval_batch_size=EVAL_BATCH_SIZE)
    loader = OobleckDataLoader(args, dataset, LoaderType.Evaluation, 0, [1], 0, 0)
    assert loader.batch_sampler.microbatch_size == EVAL_BATCH_SIZE        # dataloader.py:124-129


def test_batch_samples(dataset):
    loader = get_dataloader(dataset, 0, [1])
    assert loader.batch_sampler.num_iterations_done == 0
    inputs = next(iter(loader))
    assert isinstance(inputs, dict)
    assert list(inputs) == ["input_ids", "attention_mask", "labels"]     # the field order _prepare_inputs relies on
    for tensor in inputs.values():
        assert isinstance(tensor, torch.Tensor)
        assert tensor.size(dim=0) == TRAIN_BATCH_SIZE
        assert tensor.dtype == torch.int64
    assert torch.equal(inputs["labels"], inputs["input_ids"])            # group_texts: labels = input_ids
    assert bool((inputs["attention_mask"] == 1).all())


@pytest.mark.parametrize("shuffle", [False, True])
def test_batch_deterministic(dataset, shuffle):
    sampler = get_dataloader(dataset, 0, [2], shuffle=shuffle).batch_sampler
    assert sampler.num_iterations_done == 0
    batch1 = next(iter(sampler))
    assert sampler.num_iterations_done == 0
    batch2 = next(iter(sampler))          # a re-created iterator restarts the epoch: same order
    assert sampler.num_iterations_done == 0
    assert batch1 == batch2


def test_distributed_batch_samples(dataset):
    for rank in range(WORLD):
        sampler = get_dataloader(dataset, rank, [GRADIENT_ACCUMULATION_STEP] * WORLD).batch_sampler
        assert sampler.num_iterations_done == 0
        iterator = iter(sampler)
        for _ in range(GRADIENT_ACCUMULATION_STEP):
            next(iterator)
        assert sampler.num_iterations_done == 1


HETEROGENEOUS = [[GRADIENT_ACCUMULATION_STEP] * WORLD, [GRADIENT_ACCUMULATION_STEP + i for i in range(WORLD)]]


@pytest.mark.parametrize("num_microbatches", HETEROGENEOUS, ids=["equal", "heterogeneous"])
def test_unique_batches_per_index(dataset, num_microbatches):
    results = []
    for rank in range(WORLD):
        sampler = get_dataloader(dataset, rank, num_microbatches).batch_sampler
        iterator = iter(sampler)
        for _ in range(num_microbatches[rank]):
            results.extend(next(iterator))
        assert sampler.num_iterations_done == 1
    assert len(results) == TRAIN_BATCH_SIZE * sum(num_microbatches)
    assert len(set(results)) == len(results)      # no two pipelines ever see the same sample


@pytest.mark.parametrize("num_microbatches", HETEROGENEOUS, ids=["equal", "heterogeneous"])
def test_jump_batch(dataset, num_microbatches):
    target_jump_size = TRAIN_BATCH_SIZE * sum(num_microbatches)
    for rank in range(WORLD):
        sampler = get_dataloader(dataset, rank, num_microbatches).batch_sampler
        iterator = iter(sampler)
        result = []
        for i in range(num_microbatches[rank]):
            batch = next(iterator)
            if i == 0:
                result.extend(batch)
        assert sampler.num_iterations_done == 1
        result.extend(next(iterator))
        assert len(result) == 2
        assert result[1] == result[0] + target_jump_size


def test_jump_batch_shuffled_positions(dataset):
    """Same property under the shuffled order the engine trains with: the first micro-batch of iteration 1 sits
    ``sum(num_microbatches) * microbatch_size`` positions after the first micro-batch of iteration 0."""
    num_microbatches = HETEROGENEOUS[1]
    g = torch.Generator()
    g.manual_seed(0)                                                      # dataloader.py:46-48: seed + epoch
    order = torch.randperm(len(dataset), generator=g).tolist()
    for rank in range(WORLD):
        it = iter(get_dataloader(dataset, rank, num_microbatches, shuffle=True).batch_sampler)
        first = next(it)
        for _ in range(num_microbatches[rank] - 1):
            next(it)
        second = next(it)
        assert order.index(second[0]) == order.index(first[0]) + TRAIN_BATCH_SIZE * sum(num_microbatches)


def test_stop_iteration(dataset):
    num_microbatches = [GRADIENT_ACCUMULATION_STEP]
    loader = get_dataloader(dataset, 0, num_microbatches)
    sampler = loader.batch_sampler
    iterator = iter(loader)
    assert sampler.num_iterations_done == 0
    assert sampler.epoch == 0
    assert len(loader) == len(dataset) // (TRAIN_BATCH_SIZE * GRADIENT_ACCUMULATION_STEP)
    for iter_num in range(len(loader)):
        assert sampler.num_iterations_done == iter_num
        for _ in range(num_microbatches[0]):
            next(iterator)                        # must not raise before the last full iteration
        assert sampler.num_iterations_done == iter_num + 1
    with pytest.raises(StopIteration):
        next(iterator)
    assert sampler.num_iterations_done == 0
    assert sampler.epoch == 1
    # the next epoch is shuffled with seed + epoch: a different order over the same samples
    e0 = [i for b in OobleckSampler(dataset, 1, 0, [4], 0, epoch=0) for i in b]
    e1 = [i for b in OobleckSampler(dataset, 1, 0, [4], 0, epoch=1) for i in b]
    assert e0 != e1 and len(set(e0)) == len(e0) == len(e1)
