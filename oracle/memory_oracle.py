"""Oracle restatement of the AMP memories (TEST INFRASTRUCTURE -- see ``oracle/__init__.py``).

Upstream skrl (>= 1.4.3; third party, neither vendored nor installed: PARITY UNPINNED, literal restatement):
``skrl.memories.torch.base.Memory.add_samples`` / ``sample_by_index`` and ``RandomMemory.sample`` for a memory with one
tensor (``"states"``) and ``num_envs = 1`` -- how ``AMP.__init__`` builds ``motion_dataset`` and ``reply_buffer``
(reference call sites: ``train.py:131-134, 290-298``; sizes ``agents/skrl_g1_dance_amp_cfg.yaml:50-58``).
"""

from __future__ import annotations

import numpy as np
import torch


class OracleRandomMemory:
    def __init__(self, memory_size: int, width: int):
        self.memory_size, self.width = memory_size, width
        self.states = torch.zeros((memory_size, width), dtype=torch.float32)
        self.memory_index = 0
        self.filled = False

    def __len__(self):
        return self.memory_size if self.filled else self.memory_index

    def add_samples(self, states: torch.Tensor) -> None:
        """``Memory.add_samples`` with a batch of rows and ``num_envs == 1``: rows go in one after the other, the index wraps
        to 0 after the last slot and ``filled`` is raised.  (Upstream copies the batch in at most two slices; written here
        row by row, which is the same memory content for every batch length.)"""
        for row in states.reshape(-1, self.width):
            self.states[self.memory_index] = row
            self.memory_index += 1
            if self.memory_index >= self.memory_size:
                self.memory_index = 0
                self.filled = True

    def sample_by_index(self, indexes, mini_batches: int = 1):
        """``Memory.sample_by_index``: ``tensors_view[name][indexes]``; ``mini_batches > 1`` splits like ``np.array_split``."""
        idx = torch.as_tensor(indexes, dtype=torch.int64)
        if mini_batches > 1:
            return [self.states[torch.as_tensor(b, dtype=torch.int64)] for b in np.array_split(idx.numpy(), mini_batches)]
        return [self.states[idx]]
