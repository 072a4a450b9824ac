"""Data: synthetic CIFAR10 / LOFAR sources, reference shard math, loaders."""
from .cifar import CifarData, ShardLoader, make_synthetic_cifar, normalize_batch, shard_ranges, worker_norm
from .lofar import LofarSource, get_data_minibatch

__all__ = ["CifarData", "ShardLoader", "make_synthetic_cifar", "normalize_batch", "shard_ranges", "worker_norm",
           "LofarSource", "get_data_minibatch"]
