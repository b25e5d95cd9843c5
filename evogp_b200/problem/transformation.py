"""Transformation — fitness = |Pearson correlation| between a tree's output and the labels, for feature construction
(reference: src/evogp/problem/transformation.py:11-69).  The reference takes the outputs from Forest.batch_forward, which
it implements by replicating the forest once per datapoint; here batch_forward is one fused kernel ([P, N, 1] out), the
reduction stays in torch exactly as the reference writes it (including its use of the mean over ALL outputs)."""
from typing import Optional

import torch
from torch import Tensor

from .. import _native
from ..tree import Forest
from .base import BaseProblem


class Transformation(BaseProblem):
    def __init__(self, datapoints: Optional[Tensor] = None, labels: Optional[Tensor] = None, dataset: Optional[str] = None):
        if datapoints is not None and labels is not None:
            self.datapoints, self.labels = datapoints, labels
        else:
            assert dataset is not None, "dataset must be provided when datapoints and labels are not provided"
            self.datapoints, self.labels = self.generate_data(dataset)

    def generate_data(self, dataset: str):
        if dataset != "diabetes":
            raise ValueError("Invalid dataset")
        from sklearn.datasets import load_diabetes

        X, y = load_diabetes(return_X_y=True)
        dev = _native.device()
        return torch.tensor(X, dtype=torch.float32, device=dev), torch.tensor(y, dtype=torch.float32, device=dev)

    def _outputs(self, forest: Forest) -> Tensor:
        out = forest.batch_forward(self.datapoints)            # [P, N, 1]
        return out.reshape(out.shape[0], out.shape[1])

    def evaluate(self, forest: Forest) -> Tensor:
        out = self._outputs(forest)
        out_c = out - torch.mean(out)                          # the reference demeans with the mean over all trees (:38)
        lab_c = self.labels - torch.mean(self.labels)
        corr = torch.sum(out_c * lab_c, dim=1) / torch.sqrt(torch.sum(out_c ** 2, dim=1) * torch.sum(lab_c ** 2))
        return torch.abs(corr)

    def new_feature(self, forest: Forest, n_best: int, n_features: int) -> Tensor:
        """Outputs [N, n_features] of the n_features least mutually correlated trees among the n_best fittest (:45-69)."""
        fitness = self.evaluate(forest)
        best = fitness.argsort(descending=True)[:n_best]
        outs = self._outputs(forest[best])
        corr = torch.abs(torch.corrcoef(outs))
        corr.fill_diagonal_(0)
        keep = torch.ones(best.shape[0], dtype=torch.bool, device=corr.device)
        while int(keep.sum()) > n_features:                    # drop the later tree of the most correlated pair
            flat = int(torch.argmax(corr))
            worst = max(flat // corr.shape[1], flat % corr.shape[1])
            keep[worst] = False
            corr[worst, :] = 0
            corr[:, worst] = 0
        return outs[keep].T

    @property
    def problem_dim(self):
        return self.datapoints.shape[1]

    @property
    def solution_dim(self):
        return 1
