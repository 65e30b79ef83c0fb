"""BYOL projector / predictor heads (ref:src/models/multi_layer_heads.py:4-49).  These are tiny [B,1024] GEMMs with
BatchNorm1d; they stay stock torch.  The attribute name ``layers`` and the Sequential indices are
the checkpoint contract (``online_projector.layers.0.weight`` ...)."""
import torch
import torch.nn as nn


class ProjectionHead(nn.Module):
    """Linear - BN - ReLU - Linear - BN."""

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Linear(input_dim, hidden_dim), nn.BatchNorm1d(hidden_dim), nn.ReLU(),
            nn.Linear(hidden_dim, output_dim), nn.BatchNorm1d(output_dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.layers(x)


class PredictionHead(nn.Module):
    """Linear - BN - ReLU - Linear - BN - ReLU - Linear."""

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Linear(input_dim, hidden_dim), nn.BatchNorm1d(hidden_dim), nn.ReLU(),
            nn.Linear(hidden_dim, hidden_dim), nn.BatchNorm1d(hidden_dim), nn.ReLU(),
            nn.Linear(hidden_dim, output_dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.layers(x)
