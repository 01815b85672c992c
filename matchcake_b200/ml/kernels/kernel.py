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
    DEFAULT_ALIGNMENT_ITERATIONS = 100
    DEFAULT_ALIGNMENT_LEARNING_RATE = 1e-3
    DEFAULT_ALIGNMENT_EARLY_STOPPING_PATIENCE = 10
    DEFAULT_ALIGNMENT_EARLY_STOPPING_THRESHOLD = 1e-5

    def __init__(self, *, gram_batch_size: int = DEFAULT_GRAM_BATCH_SIZE, random_state: int = DEFAULT_RANDOM_STATE,
                 alignment: bool = DEFAULT_ALIGNMENT, alignment_iterations: int = DEFAULT_ALIGNMENT_ITERATIONS,
                 alignment_learning_rate: float = DEFAULT_ALIGNMENT_LEARNING_RATE,
                 alignment_early_stopping_patience: int = DEFAULT_ALIGNMENT_EARLY_STOPPING_PATIENCE,
                 alignment_early_stopping_threshold: float = DEFAULT_ALIGNMENT_EARLY_STOPPING_THRESHOLD):
        super().__init__()
        self.gram_batch_size = gram_batch_size
        self.random_state = random_state
        self.alignment = alignment
        self.alignment_iterations = alignment_iterations
        self.alignment_learning_rate = alignment_learning_rate
        self.alignment_early_stopping_patience = alignment_early_stopping_patience
        self.alignment_early_stopping_threshold = alignment_early_stopping_threshold
        self.opt_ = None
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
            self._align_kernel()
        # leave fit() in eval mode: the differentiable (unfused) route is taken only after an explicit .train()
        self.eval()
        self.is_fitted_ = True
        return self

    # ------------------------------------------------------------------ kernel-target alignment (kernel.py:200-330)
    def _create_y_kernel(self) -> torch.Tensor:
        """Ideal kernel of the labels: one-hot inner products for integer class labels, plain inner products for
        regression targets (kernel.py:285-316)."""
        y = torch.as_tensor(np.asarray(self.y_train_) if not isinstance(self.y_train_, torch.Tensor) else self.y_train_)
        y = y.to("cuda")
        if y.dim() == 1 and torch.allclose(y.double(), y.long().double()):
            one_hot = torch.nn.functional.one_hot(y.long(), num_classes=int(y.max().item()) + 1).double()
            return one_hot @ one_hot.T
        y = y.double().reshape(len(y), -1)
        return y @ y.T

    def _create_kernel_centerer(self) -> torch.Tensor:
        n = len(self.x_train_)
        return torch.eye(n, dtype=torch.float64, device="cuda") - 1.0 / n

    def kernel_alignment(self, kernel_matrix: torch.Tensor, centered_y_kernel: torch.Tensor) -> torch.Tensor:
        """<K_c, Y_c>_F / (||K_c|| ||Y_c|| + 1e-8) with K_c = C K C (kernel.py:210-258)."""
        c = self._create_kernel_centerer()
        kc = c @ kernel_matrix.to(torch.float64) @ c
        return (kc * centered_y_kernel).sum() / (torch.linalg.norm(kc) * torch.linalg.norm(centered_y_kernel) + 1e-8)

    def _align_kernel(self, verbose: bool = False) -> "Kernel":
        """Gradient ascent (AdamW) on the centred kernel-target alignment over the kernel's parameters; gradients come
        from the adjoint kernels of csrc/backward.cu.  Keeps the best parameters seen; stops early when the score has
        moved less than the threshold for ``patience`` iterations."""
        from scipy.optimize import OptimizeResult

        if self.x_train_ is None or self.y_train_ is None:
            raise ValueError("Training data must be provided before kernel alignment.")
        params = [p for p in self.parameters() if p.requires_grad]
        if not params:
            raise ValueError("The kernel has no trainable parameters to align.")
        c = self._create_kernel_centerer()
        yc = (c @ self._create_y_kernel() @ c).detach()
        self.opt_ = OptimizeResult(fun=np.inf, history=[])
        best = torch.nn.utils.parameters_to_vector(params).detach().clone()
        best_score = -np.inf
        optimizer = torch.optim.AdamW(params, lr=self.alignment_learning_rate)
        patience, thr = self.alignment_early_stopping_patience, self.alignment_early_stopping_threshold
        for _ in range(self.alignment_iterations):
            self.train()
            optimizer.zero_grad()
            score = self.kernel_alignment(self(self.x_train_, self.x_train_), yc)
            value = float(score.detach())
            self.opt_.history.append(value)
            if value >= best_score:
                best_score = value
                best = torch.nn.utils.parameters_to_vector(params).detach().clone()
            hist = self.opt_.history
            if len(hist) > patience and np.all(np.abs(np.diff(hist[-patience:])) <= thr):
                break
            (-score).backward()
            optimizer.step()
            if verbose:
                print(f"alignment {value:.6f}")
        with torch.no_grad():
            torch.nn.utils.vector_to_parameters(best, params)
        self.opt_.fun = best_score
        self.eval()
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
