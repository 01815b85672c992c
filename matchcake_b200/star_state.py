"""Star-state search: the basis state of highest probability of the evolved state.

Mirrors /root/reference/src/matchcake/devices/star_state_finding_strategies/: the ABC ``StarStateFindingStrategy``
(star_state_finding_strategy.py:9-19), ``GreedyStrategy`` (greedy_strategy.py:14-55), ``FromSamplingStrategy``
(from_sampling_strategy.py:12-46) and the name lookup ``get_star_state_finding_strategy`` (__init__.py:13-21).
A strategy is called as ``strategy(device, states_prob_func, **kwargs)`` and returns ``(star_states, star_probs)``;
``states_prob_func(states, wires)`` is the device's ``get_states_probability`` -- every prefix probability the search
asks for is a GPU read-out (covariance gather + batched Pfaffians of libmcb200).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Callable, Tuple, Union

import numpy as np
import torch


def _to_numpy(x) -> np.ndarray:
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


class StarStateFindingStrategy(ABC):
    NAME: str = "StarStateFindingStrategy"

    @abstractmethod
    def __call__(self, device, states_prob_func: Callable, **kwargs) -> Tuple[np.ndarray, np.ndarray]:
        raise NotImplementedError("This method should be implemented by the subclass.")


class GreedyStrategy(StarStateFindingStrategy):
    """Grow the bit string two wires at a time, keeping for every circuit of the batch the candidate of highest JOINT
    prefix probability.

    Like the reference, the candidates of a step are the four extensions of EVERY current prefix (the prefixes of all
    circuits of the batch), and each circuit takes the best of all of them -- not only the extensions of its own
    prefix.  Duplicate prefixes are evaluated once (the reference evaluates them ``batch`` times); the arg-max is
    unchanged because equal candidates are the same bit string."""

    NAME = "Greedy"

    def __call__(self, device, states_prob_func: Callable, **kwargs) -> Tuple[np.ndarray, np.ndarray]:
        n = device.num_wires
        k = 2
        blocks = device.states_to_binary(np.arange(2 ** k), k)
        probs = _to_numpy(states_prob_func(blocks, list(device.wires[:k])))          # (4,) or (4, B)
        star_states = blocks[np.argmax(probs, axis=0)]                                # (k,) or (B, k)
        star_probs = np.max(probs, axis=0)
        left = n % k
        steps = [k] * len(range(k, n - left, k)) + ([left] if left else [])
        for kj in steps:
            blocks = device.states_to_binary(np.arange(2 ** kj), kj)
            prefixes = np.unique(star_states.reshape(-1, star_states.shape[-1]), axis=0)
            cand = np.concatenate([np.repeat(prefixes, len(blocks), axis=0),
                                   np.tile(blocks, (len(prefixes), 1))], axis=1)      # (P * 2^kj, len + kj)
            probs = _to_numpy(states_prob_func(cand, list(device.wires[: cand.shape[-1]])))
            star_states = cand[np.argmax(probs, axis=0)]
            star_probs = np.max(probs, axis=0)
        return star_states, star_probs


class FromSamplingStrategy(StarStateFindingStrategy):
    """Most frequent sampled bit string per circuit and its relative frequency (needs shots)."""

    NAME = "FromSampling"

    def __call__(self, device, states_prob_func: Callable, **kwargs) -> Tuple[np.ndarray, np.ndarray]:
        samples = kwargs.get("samples", None)
        if samples is None:
            samples = device.samples
        if samples is None:
            samples = device.generate_samples()
        samples = _to_numpy(samples).astype(int)
        flat = samples.reshape(samples.shape[0], -1, samples.shape[-1])              # (shots, B, n)
        states, probs = [], []
        for b in range(flat.shape[1]):
            uniq, counts = np.unique(flat[:, b, :], return_counts=True, axis=0)
            best = int(np.argmax(counts))
            states.append(uniq[best])
            probs.append(counts[best] / samples.shape[0])
        return (np.stack(states, axis=0).reshape(samples.shape[1:]),
                np.asarray(probs, dtype=float).reshape(samples.shape[1:-1]))


star_state_finding_strategy_map = {c.NAME.lower().strip(): c for c in (GreedyStrategy, FromSamplingStrategy)}


def get_star_state_finding_strategy(name: Union[str, StarStateFindingStrategy]) -> StarStateFindingStrategy:
    if isinstance(name, StarStateFindingStrategy):
        return name
    key = str(name).lower().strip()
    if key not in star_state_finding_strategy_map:
        raise ValueError(f"Unknown star state finding strategy name: {name}")
    return star_state_finding_strategy_map[key]()
