"""``GramMatrix`` container with the reference's cell-selection and symmetrisation rules (host side).

Mirrors /root/reference/src/matchcake/ml/kernels/gram_matrix.py:12-209: float32 storage, only cells ``i > j``
(``n0 >= n1``) or ``j > i`` (``n0 < n1``) are evaluated, the other triangle of the leading square block is
mirrored and the diagonal is set to 1.  The index generator is vectorised (the reference walks ``np.ndindex``
over all cells in Python); order and content of the batches are identical.  The GPU Gram builder applies the
same rules inside the kernel (csrc/pfaffian.cu: gram_cell_active / gram_store).
"""
from __future__ import annotations

from typing import Any, Callable, Iterator, Tuple

import numpy as np
import torch


class GramMatrix:
    def __init__(self, shape: Tuple[int, ...], initial_value: float = 0.0, requires_grad: bool = False):
        self._shape = tuple(shape)
        self._requires_grad = requires_grad
        self._tensor = torch.full(self._shape, initial_value, dtype=torch.float32)

    @property
    def shape(self) -> Tuple[int, ...]:
        return self._shape

    @property
    def requires_grad(self) -> bool:
        return self._requires_grad

    def __getitem__(self, item) -> torch.Tensor:
        return self._tensor[item]

    def __setitem__(self, key, value):
        if isinstance(value, torch.Tensor):
            value = value.detach().to("cpu", torch.float32)
        else:
            value = torch.as_tensor(np.asarray(value, dtype=np.float32))
        self._tensor[key] = value

    def apply_(self, func: Callable[[np.ndarray], Any], batch_size: int = 32, symmetrize: bool = True) -> "GramMatrix":
        for indices in self.indices_batch_generator(batch_size):
            self[indices[:, 0], indices[:, 1]] = func(indices)
        if symmetrize:
            self.symmetrize_()
        return self

    def to_tensor(self) -> torch.Tensor:
        return self._tensor

    def symmetrize_(self) -> "GramMatrix":
        if self.shape[0] < self.shape[1]:
            idx = np.tril_indices(n=self.shape[0], m=self.shape[1], k=-1)
        else:
            idx = np.triu_indices(n=self.shape[0], m=self.shape[1], k=1)
        self._tensor[idx[0], idx[1]] = self._tensor[idx[1], idx[0]]
        self._tensor.fill_diagonal_(1.0)
        return self

    def indices_batch_generator(self, batch_size: int = 32) -> Iterator[np.ndarray]:
        n0, n1 = self.shape
        rows = np.arange(n0, dtype=np.int64)
        if n0 < n1:
            counts = np.maximum(n1 - 1 - rows, 0)
            starts = rows + 1
        else:
            counts = np.minimum(rows, n1)
            starts = np.zeros(n0, dtype=np.int64)
        total = int(counts.sum())
        if total == 0:
            return
        ii = np.repeat(rows, counts)
        offs = np.arange(total, dtype=np.int64) - np.repeat(np.cumsum(counts) - counts, counts)
        jj = np.repeat(starts, counts) + offs
        idx = np.stack([ii, jj], axis=1)
        for s in range(0, total, batch_size):
            yield idx[s:s + batch_size]
