"""Utilities: reference-named parameter plumbing, the flat arena, checkpoints, metrics."""
from .flat import FlatArena, arena_of, flatten_module
from .simple_utils import (init_weights, unfreeze_one_layer, unfreeze_all_layers, freeze_all_layers,
                           unfreeze_one_block, get_trainable_values, put_trainable_values,
                           number_of_layers, number_of_blocks)

__all__ = ["FlatArena", "arena_of", "flatten_module", "init_weights", "unfreeze_one_layer", "unfreeze_all_layers",
           "freeze_all_layers", "unfreeze_one_block", "get_trainable_values", "put_trainable_values",
           "number_of_layers", "number_of_blocks"]
