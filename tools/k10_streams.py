"""K = 10 co-resident replicas on ONE GPU (the reference's default K): wall time of one FedAvg round of block 0
(10 x 40 minibatches, ResNet18, CUDA graphs) with the replicas stepped sequentially on one stream vs concurrently on
their own CUDA streams (EngineConfig.streams)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from federated_pytorch_test_b200.api import federated_multi

for streams in (False, True, False, True):
    cfg = federated_multi.Config(K=10, use_resnet=True, Nloop=1, Nadmm=3, max_minibatches=20, check_results=False, save_model=False,
                                 streams=streams, distributed=False)
    marks = []

    def log(msg):
        if msg.startswith("dual ("):
            torch.cuda.synchronize()
            marks.append(time.perf_counter())

    t0 = time.perf_counter()
    from federated_pytorch_test_b200.algo import engine as E
    stop = {"n": 0}
    orig = E.Engine._aggregate

    def agg(self, visit, *a, **k):
        r = orig(self, visit, *a, **k)
        stop["n"] += 1
        if stop["n"] >= 3:
            self.stop_requested = True
        return r

    E.Engine._aggregate = agg
    try:
        eng = federated_multi.run(cfg, log=log)
    finally:
        E.Engine._aggregate = orig
    rounds = [b - a for a, b in zip(marks[:-1], marks[1:])]
    print("streams=%s: rounds (10 replicas x 20 steps each) %s s -> %.3f ms per replica-step" % (streams, ["%.3f" % r for r in rounds], 1e3 * min(rounds) / 200), flush=True)
