"""Parallel runtime: topology (worker placement) and block collectives."""
from .topology import Topology
from .collective import TorchCollective, make_collective

__all__ = ["Topology", "TorchCollective", "make_collective"]
