"""The reference's ``print`` formats, verbatim, so logs can be diffed.

Sources: /root/reference/src/federated_multi.py:120-121,199-200,214;
consensus_multi.py:257,271,299; federated_cpc.py:296; no_consensus_multi.py:135.
"""
from __future__ import annotations


def accuracy_line(ck: int, total: int, correct: int) -> str:
    # the reference prints the *floored* integer percentage through %f (Q5)
    return "Accuracy of the network %d on the %d test images:%%%f" % (ck, total, 100 * correct // total)


def accuracy_exact(correct: int, total: int) -> float:
    return 100.0 * correct / max(total, 1)


def dual_line(epoch: int, nloop: int, label, nadmm: int, dual: float) -> str:
    return "dual (epoch=%d,loop=%d,block=[%d,%d],avg=%d)=%e" % (epoch, nloop, label[0], label[1], nadmm, dual)


def admm_line(label, N: int, rho_mean: float, nadmm: int, nloop: int, primal: float, dual: float) -> str:
    return "block=[%d,%d](%d,%f) ADMM=%d/%d primal=%e dual=%e" % (label[0], label[1], N, rho_mean, nadmm, nloop, primal, dual)


def minibatch_line(ck: int, label, nloop: int, N: int, i: int, epoch: int, loss: float) -> str:
    return "model=%d block=[%d,%d] %d(%d) minibatch=%d epoch=%d loss %e" % (ck, label[0], label[1], nloop, N, i, epoch, loss)


def minibatch_line_noblock(ck: int, i: int, epoch: int, loss: float) -> str:
    return "model=%d minibatch=%d epoch=%d loss %e" % (ck, i, epoch, loss)


def cpc_dual_line(N: int, niter: int, nloop: int, mdl: int, ci: int, nadmm: int, dual: float) -> str:
    return "dual (N=%d,iter=%d,loop=%d,model=%d,block=%d,avg=%d)=%e" % (N, niter, nloop, mdl, ci, nadmm, dual)


def cluster_costs_line(k: int, c1: float, c2: float, c21: float, c3: float) -> str:
    return "cluster %d costs %f,%f,%f,%f" % (k, c1, c2, c21, c3)
