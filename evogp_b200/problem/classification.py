"""Classification — fitness = accuracy of the forest's outputs on a labelled dataset
(reference: src/evogp/problem/classification.py:11-83).  The reference evaluates batch_forward ([P, N, O], by
replicating the forest N times) and reduces it with torch softmax / argmax; here the prediction and the comparison with
the label happen inside the evaluation kernel (`tree_classification_accuracy`) and one float per tree comes back."""
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
        f = forest
        dev = f.batch_node_value.device
        X = self.datapoints.to(dev, torch.float32).contiguous()
        y = self.labels.to(dev, torch.float32).contiguous()
        if self.multi_output != (f.output_len > 1):       # shapes the fused kernel does not cover: the reference's formulation
            return self.evaluate_unfused(forest)
        return torch.ops.evogp_cuda.tree_classification_accuracy(
            f.pop_size, X.shape[0], f.max_tree_len, f.input_len, f.output_len, f.batch_node_value.contiguous(),
            f.batch_node_type.contiguous(), f.batch_subtree_size.contiguous(), X, y, float(self.maximum))

    def evaluate_unfused(self, forest: Forest):
        """The reference's formulation on the fused batch_forward output (kept as the cross-check of `evaluate`)."""
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
