"""Eager fp32 torch modules with the notebooks' architectures -- TEST INFRASTRUCTURE (the comparison side of the K5 /
critic parity tests); nothing in the product package imports this."""
from typing import Mapping

import torch


def reference_policy(state_dict: Mapping[str, torch.Tensor], head: int = 3) -> torch.nn.Module:
    """An eager fp32 torch module with the notebook's architecture (for tests / comparisons); head=1: the critic."""
    net = torch.nn.Sequential(
        torch.nn.Linear(15, 128), torch.nn.LayerNorm(128), torch.nn.ReLU(),
        torch.nn.Linear(128, 128), torch.nn.LayerNorm(128), torch.nn.ReLU(),
        torch.nn.Linear(128, 64), torch.nn.LayerNorm(64), torch.nn.ReLU(),
        *((torch.nn.Linear(64, 3), torch.nn.Sigmoid()) if head == 3 else (torch.nn.Linear(64, 1),)))
    net.load_state_dict({k.replace("network.", ""): v for k, v in state_dict.items()})
    return net
