"""Entry points with the reference's script names."""
from . import (consensus_multi, federated_cpc, federated_multi, federated_vae, federated_vae_cl, fedprox_multi,
               no_consensus_multi)

__all__ = ["no_consensus_multi", "federated_multi", "fedprox_multi", "consensus_multi", "federated_vae",
           "federated_vae_cl", "federated_cpc"]
