"""The handful of HF ``TrainingArguments`` fields the hot path reads (engine.py:430-440, pipeline.py:117-127),
with HF's defaults.  Any object exposing the same attributes (e.g. a real ``transformers.TrainingArguments``)
can be passed instead."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class TrainingArguments:
    per_device_train_batch_size: int = 8
    per_device_eval_batch_size: int = 8
    learning_rate: float = 5e-5
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_epsilon: float = 1e-8
    warmup_steps: int = 0
    warmup_ratio: float = 0.0
    max_steps: int = -1
    dataloader_num_workers: int = 0
    dataloader_pin_memory: bool = True
    output_dir: str = "/tmp/oobleck/output"

    def get_warmup_steps(self, num_training_steps: int) -> int:
        import math
        return self.warmup_steps if self.warmup_steps > 0 else math.ceil(num_training_steps * self.warmup_ratio)
