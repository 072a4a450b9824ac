"""Summarise the JSONL metrics of tools/accuracy_gpu.sh: mean test accuracy per aggregation round, rounds to reach the
targets, final ordering, and the per-round performance fields (images/s, aggregation us, bus GB/s)."""
import json
import os
import sys

d = sys.argv[1]
targets = [float(t) for t in os.environ.get("TARGETS", "50,60,70,80").split(",")]
runs = [("federated_multi (FedAvg)", "fedavg"), ("consensus_multi --bb_update (adaptive ADMM)", "admm_bb"), ("fedprox_multi", "fedprox"),
        ("no_consensus_multi K=8 (stand-alone, 1/8 of the data each; round = epoch)", "standalone_k8"),
        ("no_consensus_multi K=1 (stand-alone, all data; round = epoch)", "standalone_k1")]
print("# Accuracy / rounds-to-target on B200 (ResNet18, synthetic CIFAR10-shaped data)\n")
print("| run | rounds | acc r=1 | r=5 | r=10 | r=20 | r=30 | last | " + " | ".join("rounds to %.0f %%" % t for t in targets) + " | img/s (median round) | aggregation us (median) |")
print("|---|---|---|---|---|---|---|---|" + "---|" * len(targets) + "---|---|")
final = {}
for label, stem in runs:
    p = os.path.join(d, stem + ".jsonl")
    if not os.path.exists(p):
        continue
    rows = [json.loads(l) for l in open(p) if l.strip()]
    ev = [r for r in rows if r.get("kind") == "eval"]
    rd = [r for r in rows if r.get("kind") == "round"]
    acc = [sum(r["accuracy"]) / len(r["accuracy"]) for r in ev]
    if not acc:
        continue
    at = lambda i: ("%.1f" % acc[i - 1]) if len(acc) >= i else "-"
    hit = []
    for t in targets:
        idx = next((i + 1 for i, a in enumerate(acc) if a >= t), None)
        hit.append(str(idx) if idx else "not reached")
    ips = sorted(r["images_per_s"] for r in rd if "images_per_s" in r)
    agg = sorted(r["aggregate_us"] for r in rd if "aggregate_us" in r)
    med = lambda v: ("%.0f" % v[len(v) // 2]) if v else "-"
    final[label] = acc[-1]
    print("| %s | %d | %s | %s | %s | %s | %s | %.1f | %s | %s | %s |" % (label, len(acc), at(1), at(5), at(10), at(20), at(30), acc[-1],
                                                                  " | ".join(hit), med(ips), med(agg)))
print("\nOrdering by final mean accuracy: " + " > ".join("%s (%.1f)" % kv for kv in sorted(final.items(), key=lambda kv: -kv[1])))
print("\nReference (comparison.png, README.md:28-30, real CIFAR10, Net, K=10): K=1 stand-alone > FedAvg > consensus ADMM > K=10 stand-alone.")
