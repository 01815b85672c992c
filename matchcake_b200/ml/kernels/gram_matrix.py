"""``GramMatrix`` container with the reference's cell-selection and symmetrisation rules (host side).

Mirrors /root/reference/src/matchcake/ml/kernels/gram_matrix.py:12-209: float32 storage, only cells ``i > j``
(``n0 >= n1``) or ``j > i`` (``n0 < n1``) are evaluated, the other triangle of the leading square block is
mirrored and the diagonal is set to 1.  The index generator is vectorised (the reference walks ``np.ndindex``
over all cells in Python); order and content of the batches are identical.  The GPU Gram builder applies the
same rules inside the kernel (csrc/pfaffian.cu: gram_cell_active / gram_store).
"""
from __future__ import annotations

import os
from tempfile import mkstemp
from typing import Any, Callable, Iterator, Optional, Tuple

import numpy as np
import torch


class GramMatrix:
    """``memmap=True`` keeps the float32 entries in a ``numpy.memmap`` on a temporary file like the reference does whenever
    no gradient is required (gram_matrix.py:25-47: a Gram that does not fit in host memory); the default is a host tensor,
    since the GPU builder hands over whole matrices.  ``from_device`` streams a device Gram into either store."""

    def __init__(self, shape: Tuple[int, ...], initial_value: float = 0.0, requires_grad: bool = False,
                 memmap: bool = False):
        self._shape = tuple(shape)
        self._requires_grad = requires_grad
        self._filepath: Optional[str] = None
        if memmap and not requires_grad:
            fd, self._filepath = mkstemp(suffix=".gram")
            os.close(fd)
            self._memmap = np.memmap(self._filepath, dtype="float32", mode="w+", shape=self._shape)
            self._memmap.fill(initial_value)
            self._memmap.flush()
            self._tensor = torch.from_numpy(self._memmap)  # a view: writes through the tensor land in the file's pages
        else:
            self._memmap = None
            self._tensor = torch.full(self._shape, initial_value, dtype=torch.float32)

    @classmethod
    def from_device(cls, gram: torch.Tensor, memmap: bool = False, rows_per_copy: int = 4096) -> "GramMatrix":
        """Float32 store of a finished (device) Gram, copied ``rows_per_copy`` rows at a time (flushed per block when
        memory mapped, like the reference's per-batch flush, gram_matrix.py:96-109)."""
        out = cls(tuple(gram.shape), memmap=memmap)
        for s in range(0, gram.shape[0], max(int(rows_per_copy), 1)):
            e = min(s + rows_per_copy, gram.shape[0])
            out._tensor[s:e] = gram[s:e].detach().to("cpu", torch.float32)
            out.flush()
        return out

    def flush(self) -> None:
        if self._memmap is not None:
            self._memmap.flush()

    def close(self) -> None:
        """Drops the memory map and removes its temporary file."""
        if self._memmap is not None:
            self._tensor = self._tensor.clone()
            self._memmap = None
        if self._filepath is not None and os.path.exists(self._filepath):
            os.remove(self._filepath)
        self._filepath = None

    def __del__(self):
        try:
            if getattr(self, "_filepath", None) and os.path.exists(self._filepath):
                self._memmap = None
                os.remove(self._filepath)
        except Exception:
            pass

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
            self.flush()
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
        self.flush()
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
