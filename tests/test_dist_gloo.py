"""world_size-2 gloo run of the training harness on CPU (host logic of the multi-GPU path): gradients
are all-reduced (replicas stay identical although every rank initialises from its own seed: rank 0's parameters are
broadcast), the per-step coin / site choice / boxes come from ONE numpy stream shared by the ranks (the reference draws
them once per step for all replicas) while permutations differ per rank, timing is the max over ranks.  Legs: the
eager-PyTorch operator set under DistributedDataParallel, the package's own modules over the stand-in backend, and the
flat-buffer gradient exchange of train.GraphedStep (its CUDA-graph capture needs a GPU; the exchange does not)."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, which="eager"):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import numpy as np
    import torch.distributed as dist
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cnsn_b200.train import bench_wrn
    extra = None
    graph = None
    if which in ("eager", "graphed"):
        from oracle import eager_modules as ops
        graph = which == "graphed"
    else:                       # the package's own modules and autograd Functions over the oracle-backed stand-in backend
        import cnsn_b200._lib as L
        import cnsn_b200.cnsn as ops
        from fake_backend import OracleBackend
        fake = OracleBackend()
        L.set_backend_for_tests(fake)
    r = bench_wrn(torch.device("cpu"), world, rank, batch=4, steps=2, warmup=1, cn_prob=1.0, ops=ops, graph=graph)
    if which == "package":
        extra = (fake.calls.count("site_fwd"), fake.calls.count("site_bwd"), fake.calls.count("selfnorm_bwd"))
    q.put((rank, r["param_checksum"], r["value"], r["n_gpus"], float(np.random.rand()), extra, float(torch.rand(1))))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("which", ["eager", "package", "graphed"])
def test_two_rank_gloo_training_keeps_replicas_in_sync(which):
    """which='package': the product's modules (CNSN.forward with the fused site, the autograd Functions' parameter
    gradients) under DistributedDataParallel, over the oracle-backed stand-in backend."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, which)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=500) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, c0, v0, n0, u0, e0, t0), (_, c1, v1, n1, u1, e1, t1) = res
    if which == "package":      # 3 steps x 2 active sites, every one through the fused call; the other sites plain SelfNorm
        assert e0[0] == e0[1] == 6 and e1[0] == e1[1] == 6 and e0[2] > 0
    assert n0 == n1 == 2
    assert c0 == pytest.approx(c1, rel=1e-12)          # same initial weights + all-reduced grads -> identical replicas
    assert v0 == pytest.approx(v1, rel=1e-9)           # throughput is computed from the max-over-ranks time
    assert u0 == u1                                    # ONE numpy stream: same coin, same active sites, same boxes on every rank
    assert t0 != t1                                    # per-rank torch stream (seed + rank): permutations differ
