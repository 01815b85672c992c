"""sklearn-style kernel base class (mirrors /root/reference/src/matchcake/ml/kernels/kernel.py:15-198)."""
from __future__ import annotations

from typing import Optional, Union

import numpy as np
import torch
from sklearn.base import BaseEstimator, TransformerMixin


class Kernel(torch.nn.Module, TransformerMixin, BaseEstimator):
    DEFAULT_GRAM_BATCH_SIZE = 10_000
    DEFAULT_RANDOM_STATE = 0
    DEFAULT_ALIGNMENT = False

    def __init__(self, *, gram_batch_size: int = DEFAULT_GRAM_BATCH_SIZE, random_state: int = DEFAULT_RANDOM_STATE,
                 alignment: bool = DEFAULT_ALIGNMENT):
        super().__init__()
        self.gram_batch_size = gram_batch_size
        self.random_state = random_state
        self.alignment = alignment
        self.np_rn_gen = np.random.RandomState(seed=random_state)
        self.x_train_ = None
        self.y_train_ = None
        self.is_fitted_ = False

    def forward(self, x0, x1=None, **kwargs) -> torch.Tensor:
        raise NotImplementedError

    def transform(self, x):
        """K(x, x_train) -- numpy in -> float32 numpy out, torch in -> torch out (kernel.py:100-124)."""
        is_torch = isinstance(x, torch.Tensor)
        self.eval()
        with torch.no_grad():
            k = self(x, self.x_train_, x1_train=True)
        if is_torch:
            return k
        return k.detach().cpu().numpy().astype(np.float32)

    def fit(self, x_train, y_train=None):
        self.x_train_ = x_train
        self.y_train_ = y_train
        self.build_model()
        if self.alignment:
            raise NotImplementedError("Kernel alignment needs backward kernels (SURVEY.md 8f rank 1); "
                                      "the B200 path is inference-only in this build.")
        self.is_fitted_ = True
        return self

    def predict(self, x):
        self.eval()
        with torch.no_grad():
            return self.transform(x)

    def freeze(self):
        self.eval()
        self.requires_grad_(False)
        return self

    def build_model(self) -> "Kernel":
        assert self.x_train_ is not None, "Training data must be provided to build the model."
        with torch.no_grad():
            x = self.x_train_
            self(x[0:3] if len(x) >= 3 else x)
        return self
