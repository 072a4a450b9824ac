"""``no_consensus_multi`` — K independent models on 1/K of the data each, no exchange.

Reference: /root/reference/src/no_consensus_multi.py (lower/upper-bound experiment,
all parameters trainable, Adam 1e-3 recreated every epoch, accuracy after every epoch).
"""
from __future__ import annotations

from ..algo.strategies import NoConsensus
from ..config import NoConsensusConfig, parse_config
from . import common

Config = NoConsensusConfig


class _Task(common.ClassifierTask):
    def on_epoch_start(self, epoch, engine) -> None:
        engine.log("Epoch %d" % epoch, root_only=True)


def run(cfg: Config, log=print):
    topo, coll = common.setup_runtime(cfg)
    task = _Task(cfg, topo, whole_model=True)
    ecfg = common.engine_config(cfg, Nloop=1, Nadmm=1, reset_optimizer_each_epoch=True)
    engine = common.run_engine(cfg, task, topo, coll, NoConsensus(coll, topo), ecfg, log)
    common.save_legacy(cfg, engine)
    return engine


def main(argv=None):
    return run(parse_config(Config, argv, prog="no_consensus_multi"))


if __name__ == "__main__":
    main()
