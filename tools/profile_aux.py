"""Kernel-time table (torch.profiler / CUPTI) of a few optimizer steps of the VAE / VAE-CL / CPC drivers.
usage: python tools/profile_aux.py cpc|vae|vae_cl [steps]   -> prints kernels sorted by device time, share of the total."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from federated_pytorch_test_b200.algo.engine import Engine
from federated_pytorch_test_b200.algo.strategies import FedAvg
from federated_pytorch_test_b200.api import common, federated_cpc, federated_vae, federated_vae_cl

driver = sys.argv[1] if len(sys.argv) > 1 else "cpc"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if driver == "cpc":
    mod, task_cls = federated_cpc, federated_cpc.CPCTask
    cfg = mod.Config(K=1, Nloop=1, Nadmm=2, Niter=10, load_model=False, init_model=True, save_model=False, be_verbose=False,
                     check_results=False, graphs=False, seed=69)
else:
    mod = federated_vae if driver == "vae" else federated_vae_cl
    task_cls = federated_vae.VAETask if driver == "vae" else federated_vae_cl.VAECLTask
    cfg = mod.Config(K=1, Nloop=1, Nadmm=2, Nepoch=1, check_results=False, save_model=False, be_verbose=False,
                     graphs=False, max_minibatches=12, seed=69)
topo, coll = common.setup_runtime(cfg)
task = task_cls(cfg, topo)
ecfg = common.engine_config(cfg, Nepoch=1, diagnostics="pre") if driver == "cpc" else common.engine_config(cfg)
eng = Engine(task, topo, FedAvg(coll, topo), coll, ecfg, log=lambda m: None)
prof = profile(activities=[ProfilerActivity.CUDA])
first = 4
ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]


def hook(e):
    if e.steps_done == first:
        torch.cuda.synchronize()
        ev[0].record()
        prof.start()
    elif e.steps_done == first + nsteps:
        ev[1].record()
        torch.cuda.synchronize()
        prof.stop()
        e.stop_requested = True


eng.step_hook = hook
eng.run()
def _us(k):
    return float(getattr(k, "device_time_total", 0) or getattr(k, "cuda_time_total", 0) or 0)


rows = [(k.key, k.count, _us(k)) for k in prof.key_averages() if _us(k) > 0]
total = sum(r[2] for r in rows)
rows.sort(key=lambda r: -r[2])
print("driver %s: %d optimizer steps, %.2f ms wall (events, profiler on), %.2f ms of kernels, %d launches" % (
    driver, nsteps, ev[0].elapsed_time(ev[1]), total / 1e3, sum(r[1] for r in rows)))
print("| kernel | launches | sum us | share |\n|---|---|---|---|")
for name, cnt, us in rows[:45]:
    print("| `%s` | %d | %.1f | %.1f %% |" % (name[:110], cnt, us, 100.0 * us / total))
