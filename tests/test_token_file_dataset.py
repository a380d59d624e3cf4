"""``TokenFileDataset`` against a direct restatement of the reference's ``group_texts``
(oobleck/execution/dataset.py:183-206) applied the way ``Dataset.map(batched=True)`` applies it: 1000 documents per call."""
import os
import sys
from itertools import chain

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def group_texts(examples, max_seq_length):
    """dataset.py:183-202, statement for statement."""
    concatenated = {k: list(chain(*examples[k])) for k in examples.keys()}
    total_length = len(concatenated[list(examples.keys())[0]])
    if total_length >= max_seq_length:
        total_length = (total_length // max_seq_length) * max_seq_length
    result = {k: [t[i: i + max_seq_length] for i in range(0, total_length, max_seq_length)]
              for k, t in concatenated.items()}
    result["labels"] = result["input_ids"].copy()
    return result


def reference_blocks(docs, seq_len, batch=1000):
    ids, masks, labels = [], [], []
    for i in range(0, len(docs), batch):
        ex = {"input_ids": docs[i: i + batch], "attention_mask": [[1] * len(d) for d in docs[i: i + batch]]}
        r = group_texts(ex, seq_len)
        ids += r["input_ids"]; masks += r["attention_mask"]; labels += r["labels"]   # noqa: E702
    return ids, masks, labels


def corpus(num_docs, seed, vocab=50257, max_len=700):
    rng = np.random.default_rng(seed)
    return [rng.integers(0, vocab, size=int(rng.integers(0, max_len))).tolist() for _ in range(num_docs)]


@pytest.mark.parametrize("num_docs,seq_len,group", [(2500, 128, 1000), (37, 64, 10), (5, 32, 1000)])
def test_blocks_match_group_texts(tmp_path, num_docs, seq_len, group):
    from oobleck_b200.execution.dataloader import TokenFileDataset
    docs = corpus(num_docs, seed=num_docs)
    flat = np.fromiter(chain(*docs), dtype=np.uint16)
    offsets = np.cumsum([0] + [len(d) for d in docs[:-1]])
    path = tmp_path / "train.bin"
    flat.tofile(path)
    np.save(tmp_path / "offsets.npy", offsets)
    ds = TokenFileDataset(str(path), seq_len=seq_len, doc_offsets=str(tmp_path / "offsets.npy"), group_size=group,
                          pin_memory=False)
    ids, masks, labels = reference_blocks(docs, seq_len, group)
    full = [b for b in ids if len(b) == seq_len]        # a group shorter than one block stays ragged in the reference
    assert ds.dropped_short_groups == len(ids) - len(full)
    assert len(ds) == len(full)
    assert torch.equal(ds.input_ids, torch.tensor(full, dtype=torch.int64).reshape(len(full), seq_len))
    assert all(m == [1] * seq_len for m, b in zip(masks, ids) if len(b) == seq_len) and labels == ids
    assert ds.sample["input_ids"].tolist() == full[0] and ds.sample["labels"].tolist() == full[0]
    assert ds.sample["attention_mask"].tolist() == [1] * seq_len
    assert ds.dropped_tokens == len(flat) - seq_len * len(full)


def test_single_stream_npy_and_loader(tmp_path):
    """No document index: one group; ``.npy`` input; the loader serves micro-batches from it like from the synthetic set."""
    from types import SimpleNamespace

    from oobleck_b200.execution.dataloader import LoaderType, OobleckDataLoader, TokenFileDataset
    flat = np.arange(1000, dtype=np.int32) % 211
    np.save(tmp_path / "tok.npy", flat)
    ds = TokenFileDataset(str(tmp_path / "tok.npy"), seq_len=32, vocab_size=211, pin_memory=False)
    assert len(ds) == 1000 // 32 and ds.dropped_tokens == 1000 % 32
    assert torch.equal(ds.input_ids.flatten(), torch.from_numpy(flat[: 31 * 32].astype(np.int64)))
    args = SimpleNamespace(per_device_train_batch_size=2, per_device_eval_batch_size=2)
    loader = OobleckDataLoader(args, ds, LoaderType.Training, 0, [3], 0, 0, shuffle=False)
    batches = list(loader)
    assert len(batches) == (31 // 6) * 3
    assert torch.equal(batches[0]["input_ids"], ds.input_ids[0:2]) and batches[0]["labels"] is batches[0]["input_ids"]
    assert torch.equal(batches[4]["input_ids"], ds.input_ids[8:10])
    capped = TokenFileDataset(str(tmp_path / "tok.npy"), seq_len=32, vocab_size=211, max_samples=5, pin_memory=False)
    assert len(capped) == 5 and torch.equal(capped.input_ids, ds.input_ids[:5])
    with pytest.raises(ValueError):
        TokenFileDataset(str(tmp_path / "tok.npy"), seq_len=32, vocab_size=100, pin_memory=False)
