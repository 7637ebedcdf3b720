"""Helpers shared by the -m gpu parity tests (CUDA path vs oracle / golden vectors)."""
import numpy as np
import torch

from oracle import truncgptq_oracle as O


def dev():
    return torch.device("cuda:0")


def to_gpu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev())
    return t if dtype is None else t.to(dtype)


def frac_equal(a: np.ndarray, b: np.ndarray) -> float:
    return float(np.mean(a == b))


def rel_fro(a: np.ndarray, b: np.ndarray) -> float:
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def oracle_quantizer(g):
    return O.Quantizer(int(g["bits"]), int(g["group"]), bool(g["sym"]))
