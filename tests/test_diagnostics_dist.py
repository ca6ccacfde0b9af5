"""The only collective of the framework (cross-chain diagnostics) on a world_size-2 gloo group (CPU)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ar1(rng, chains, n, rho):
    x = np.empty((chains, n))
    x[:, 0] = rng.standard_normal(chains)
    for t in range(1, n):
        x[:, t] = rho * x[:, t - 1] + np.sqrt(1 - rho ** 2) * rng.standard_normal(chains)
    return x


def _worker(rank, world, port, x_all, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from aux_ssm_samplers_b200 import diagnostics
    shard = torch.from_numpy(np.array_split(x_all, world, axis=0)[rank])
    out = dict(rhat=diagnostics.split_rhat(shard), ess=diagnostics.effective_sample_size(shard),
               ess_var=diagnostics.effective_sample_size(shard, var=1.0),
               summ=diagnostics.cross_chain_summary(shard[:, :2]))
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("rho", [0.0, 0.8])
def test_diagnostics_world2_matches_single_process(rho):
    from oracle import diagnostics as od
    rng = np.random.default_rng(5)
    x = _ar1(rng, 8, 400, rho)
    x[3] += 0.4  # one off-centre chain so R-hat is not trivially 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + int(rng.integers(0, 2000))
    procs = [ctx.Process(target=_worker, args=(r, 2, port, x, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_allclose(out["rhat"], od.split_rhat(x), rtol=1e-10)
    np.testing.assert_allclose(out["ess"], od.effective_sample_size(x), rtol=1e-8)
    np.testing.assert_allclose(out["ess_var"], od.effective_sample_size(x, var=1.0), rtol=1e-8)
    assert out["summ"]["n_chains"] == 8 and out["summ"]["world_size"] == 2
    np.testing.assert_allclose(out["summ"]["mean"], x[:, :2].mean(0), rtol=1e-10)


def test_single_process_matches_oracle_and_iid_sanity():
    from aux_ssm_samplers_b200 import diagnostics
    from oracle import diagnostics as od
    rng = np.random.default_rng(0)
    x = _ar1(rng, 4, 1000, 0.0)
    ess = diagnostics.effective_sample_size(torch.from_numpy(x))
    np.testing.assert_allclose(ess, od.effective_sample_size(x), rtol=1e-8)
    assert 0.7 * x.size < ess < 1.5 * x.size
    assert abs(diagnostics.split_rhat(torch.from_numpy(x)) - 1.0) < 0.02
