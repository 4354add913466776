import torch

from .tensor import Tensor, _unwrap


def lstsq(matrix, rhs, l2_regularizer=0.0, fast=True):
    """tf.linalg.lstsq default (fast=True): Cholesky solve of the normal equations, in the input dtype."""
    a, b = _unwrap(matrix), _unwrap(rhs)
    gram = a.transpose(-1, -2) @ a
    if l2_regularizer:
        gram = gram + l2_regularizer * torch.eye(gram.shape[-1], dtype=gram.dtype)
    chol = torch.linalg.cholesky(gram)
    return Tensor(torch.cholesky_solve(a.transpose(-1, -2) @ b, chol))


def eigh(matrix):
    w, v = torch.linalg.eigh(_unwrap(matrix))
    return Tensor(w), Tensor(v)
