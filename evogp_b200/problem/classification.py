"""Classification — fitness = accuracy of the forest's outputs on a labelled dataset
(reference: src/evogp/problem/classification.py:11-83), on the fused batch_forward."""
from typing import Optional

import torch
from torch import Tensor

from .. import _native
from ..tree import Forest
from .base import BaseProblem


class Classification(BaseProblem):
    def __init__(self, datapoints: Optional[Tensor] = None, labels: Optional[Tensor] = None,
                 dataset: Optional[str] = None, multi_output: bool = True):
        self.multi_output = multi_output
        if datapoints is not None and labels is not None:
            self.datapoints, self.labels = datapoints, labels
        else:
            assert dataset is not None, "dataset must be provided when datapoints and labels are not provided"
            self.datapoints, self.labels = self.generate_data(dataset)
        self.maximum = int(torch.max(self.labels))
        self.onehot_labels = torch.nn.functional.one_hot(self.labels.long(), self.maximum + 1).to(torch.float32)

    def generate_data(self, dataset: str):
        from sklearn import datasets as skd

        loaders = {"iris": skd.load_iris, "wine": skd.load_wine, "breast_cancer": skd.load_breast_cancer,
                   "digits": skd.load_digits}
        if dataset not in loaders:
            raise ValueError("Invalid dataset")
        X, y = loaders[dataset](return_X_y=True)
        dev = _native.device()
        return torch.tensor(X, dtype=torch.float32, device=dev), torch.tensor(y, dtype=torch.float32, device=dev)

    def transform(self, x: Tensor):
        return torch.clamp(torch.round(x + self.maximum / 2), 0, self.maximum).squeeze(-1)

    def evaluate(self, forest: Forest):
        outputs = forest.batch_forward(self.datapoints)   # [P, N, O]
        if self.multi_output:
            # argmax(softmax(x)) with the reference's clipping (classification.py:62-64)
            prob = torch.clip(torch.softmax(outputs, dim=2), 1e-15, 1 - 1e-15)
            pred = torch.argmax(prob, dim=2)
        else:
            pred = self.transform(outputs)
        return torch.sum(pred == self.labels, dim=1, dtype=torch.float32) / self.labels.shape[0]

    @property
    def problem_dim(self):
        return self.datapoints.shape[1]

    @property
    def solution_dim(self):
        return self.maximum + 1 if self.multi_output else 1
