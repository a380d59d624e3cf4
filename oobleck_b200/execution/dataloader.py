"""Micro-batch sampler and loader (oobleck/execution/dataloader.py:13-147) plus a synthetic dataset.

``OobleckSampler`` keeps the reference's deterministic index arithmetic bit for bit (it is what guarantees that
heterogeneous pipelines never see the same sample; pinned by tests/golden/sampler.json).  The loader no longer
goes through ``torch.utils.data.DataLoader`` worker machinery: batches are gathered from one pinned host
tensor per field (or from a device-resident copy for the kernel-only benchmark) so that ``load_microbatch`` is a
single async H2D per field.
"""
from __future__ import annotations

from enum import Enum
from typing import Iterator, List

import torch


class OobleckSampler:
    def __init__(self, dataset, microbatch_size: int, pipeline_index: int, num_microbatches: List[int],
                 num_iterations_done: int, epoch: int = 0, shuffle: bool = True, seed: int = 0):
        self.num_samples = len(dataset)
        self.microbatch_size = microbatch_size
        self.pipeline_index = pipeline_index
        self.num_microbatches = num_microbatches
        self.num_iterations_done = num_iterations_done
        self.epoch = epoch
        self.shuffle = shuffle
        self.seed = seed
        assert self.pipeline_index < len(self.num_microbatches)
        self.total_bucket_size = self.microbatch_size * sum(self.num_microbatches)

    def __len__(self) -> int:
        return self.num_samples // self.total_bucket_size

    def __iter__(self) -> Iterator[List[int]]:
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.seed + self.epoch)       # dataloader.py:46-48
            order = torch.randperm(self.num_samples, generator=g).tolist()
        else:
            order = list(range(self.num_samples))
        mine = self.num_microbatches[self.pipeline_index]
        jump = self.total_bucket_size                     # samples consumed by all pipelines per iteration
        offset = sum(self.num_microbatches[: self.pipeline_index]) * self.microbatch_size
        for it in range(len(self)):
            if self.num_samples - it * jump < self.total_bucket_size:
                break                                      # incomplete last iteration is dropped
            for mb in range(mine):
                if mb == mine - 1:
                    self.num_iterations_done += 1
                start = it * jump + mb * self.microbatch_size + offset
                yield order[start: start + self.microbatch_size]
        self.num_iterations_done = 0
        self.epoch += 1


class LoaderType(Enum):
    Training = (0,)
    Evaluation = (1,)


class SyntheticTokenDataset:
    """Wikitext-2-shaped synthetic corpus: ``num_samples`` blocks of ``seq_len`` token ids,
    ``attention_mask`` all ones, ``labels = input_ids`` (what ``group_texts`` produces, dataset.py:183-206)."""

    def __init__(self, num_samples: int = 2334, seq_len: int = 1024, vocab_size: int = 50257, seed: int = 0,
                 pin_memory: bool = True):
        g = torch.Generator().manual_seed(seed)
        ids = torch.randint(0, vocab_size, (num_samples, seq_len), generator=g, dtype=torch.int64)
        if pin_memory and torch.cuda.is_available():
            ids = ids.pin_memory()
        self.input_ids = ids
        self.seq_len = seq_len
        self.vocab_size = vocab_size
        self.sample = {"input_ids": ids[0], "attention_mask": torch.ones_like(ids[0]), "labels": ids[0].clone()}
        self.dataset = {"train": self, "validation": self}

    def __len__(self) -> int:
        return self.input_ids.shape[0]

    def to_device(self, device) -> "SyntheticTokenDataset":
        """Device-resident copy: the kernel-only benchmark starts with inputs already in HBM."""
        self.input_ids_device = self.input_ids.to(device)
        return self


class OobleckDataLoader:
    """Iterable over this pipeline's micro-batches: dicts of int64 [mb, T] tensors in the reference's field order
    (input_ids, attention_mask, labels)."""

    def __init__(self, args, datasets, dataloader_type: LoaderType, pipeline_index: int, num_microbatches: List[int],
                 num_iterations_done: int, epoch: int, shuffle: bool = True, device_resident: bool = False):
        dataset = datasets.dataset["train" if dataloader_type == LoaderType.Training else "validation"]
        mbs = (args.per_device_train_batch_size if dataloader_type == LoaderType.Training
               else args.per_device_eval_batch_size)
        self.dataset = dataset
        self.batch_sampler = OobleckSampler(dataset, mbs, pipeline_index, num_microbatches, num_iterations_done, epoch,
                                            shuffle)
        self.device_resident = device_resident
        self._ones = None

    def __len__(self) -> int:
        return len(self.batch_sampler)

    def __iter__(self):
        ds = self.dataset
        for idx in self.batch_sampler:
            index = torch.as_tensor(idx, dtype=torch.int64)
            if self.device_resident:
                src = ds.input_ids_device
                ids = src.index_select(0, index.to(src.device))
            else:
                ids = ds.input_ids.index_select(0, index)
                if ds.input_ids.is_pinned():
                    ids = ids.pin_memory()
            if self._ones is None or self._ones.shape != ids.shape or self._ones.device != ids.device:
                self._ones = torch.ones_like(ids)
            yield {"input_ids": ids, "attention_mask": self._ones, "labels": ids}
