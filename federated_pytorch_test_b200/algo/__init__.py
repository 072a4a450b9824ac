"""Algorithms: aggregation strategies and the block-coordinate engine."""
from .strategies import ADMM, BBConfig, FedAvg, FedProx, NoConsensus, Penalty, Strategy
from .engine import Engine, EngineConfig, Replica, Task, Visit

__all__ = ["ADMM", "BBConfig", "FedAvg", "FedProx", "NoConsensus", "Penalty", "Strategy", "Engine", "EngineConfig",
           "Replica", "Task", "Visit"]
