"""CPU tier, world_size 2 over gloo: the data-parallel recipe of the learners (pql_b200/algo/_dp.py)
- per-rank half batches, gradient SUM all-reduce, grad_scale 1/world applied before the global
clip, identical AdamW step on every rank - equals one reference update on the concatenated batch;
per-rank replay shards keep independent ring pointers."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import learner as L
from oracle import replay as R
from tests.golden import inputs

B, O, A = 64, 12, 4


def _worker(rank, world, init_file, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    from pql_b200.algo import _dp
    torch.set_num_threads(1)
    case = inputs.learner_case(99, B, O, A, False)
    half = slice(rank * B // world, (rank + 1) * B // world)
    batch = tuple(x[half] for x in case["batch"])
    # every rank starts from rank 0's parameters
    flat0 = torch.cat([t.reshape(-1) for t in L.flat([case["q1"], case["q2"]])]).clone()
    if rank != 0:
        flat0 += 1.0
    _dp.broadcast_(flat0, 0)
    v = L.VLearnerOracle(case["q1"], case["q2"], max_grad_norm=None)
    tensors = L.flat([v.q1, v.q2])
    # local gradient of the local mean loss
    obs, action, reward, next_obs, done = batch
    nobs = L.normalize(next_obs, case["norm"]); cobs = L.normalize(obs, case["norm"])
    with torch.no_grad():
        na = L.target_policy_action(nobs, case["actor"], case["noises"][0][half])
        y = reward + (1 - done) * v.gamma_n * L.q_min(nobs, na, v.tq1, v.tq2)
    c1, c2 = L.q1_q2(cobs, action, v.q1, v.q2)
    loss = torch.nn.functional.mse_loss(c1, y) + torch.nn.functional.mse_loss(c2, y)
    grads = torch.autograd.grad(loss, tensors)
    flat_g = torch.cat([g.reshape(-1) for g in grads])
    _dp.allreduce_sum_(flat_g)
    flat_g *= 1.0 / world                      # grad_scale of pqlb_adamw_polyak
    # clip on the REDUCED gradient, then AdamW: identical on every rank
    split = torch.split(flat_g, [t.numel() for t in tensors])
    gl = [g.reshape(t.shape) for g, t in zip(split, tensors)]
    gl, total = L.clip_grad_norm(gl, 0.5)
    v.opt.step(tensors, gl)
    flat_p = torch.cat([t.detach().reshape(-1) for t in tensors])
    assert _dp.params_in_sync(flat_p)
    assert _dp.world_size() == world
    # replay shards: independent pointers per rank
    ring = R.RingOracle(50, O, A)
    for i, n in enumerate([16, 30 + rank, 9]):
        ring.insert(*inputs.flat_rows(10 * rank + i, n, O, A))
    np.save(os.path.join(out_dir, f"p{rank}.npy"), flat_p.numpy())
    np.save(os.path.join(out_dir, f"g{rank}.npy"), flat_g.numpy())
    np.save(os.path.join(out_dir, f"ptr{rank}.npy"), np.array([ring.next_p, int(ring.if_full), ring.cur_capacity, float(flat0[0])]))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_step_equals_reference_step_on_concatenated_batch():
    world = 2
    with tempfile.TemporaryDirectory() as d:
        init = os.path.join(d, "init")
        mp.spawn(_worker, args=(world, init, d), nprocs=world, join=True)
        p = [np.load(os.path.join(d, f"p{r}.npy")) for r in range(world)]
        g = [np.load(os.path.join(d, f"g{r}.npy")) for r in range(world)]
        ptr = [np.load(os.path.join(d, f"ptr{r}.npy")) for r in range(world)]
    assert np.array_equal(p[0], p[1]) and np.array_equal(g[0], g[1])
    # single-process reference update on the whole batch
    case = inputs.learner_case(99, B, O, A, False)
    v = L.VLearnerOracle(case["q1"], case["q2"])
    v.learn(case["batch"], case["noises"][0], case["actor"], case["norm"])
    ref_g = np.concatenate([x.reshape(-1).numpy() for x in v.last["grads"]])
    ref_p = np.concatenate([t.detach().reshape(-1).numpy() for t in L.flat([v.q1, v.q2])])
    np.testing.assert_allclose(g[0], ref_g, rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(p[0], ref_p, rtol=1e-4, atol=1e-6)   # Adam normalises the step: re-association noise in g shows at 1e-5
    # shards: rank 0 inserted 16+30+9 = 55 > 50 -> wrapped to 5; rank 1 inserted 56 -> 6
    assert ptr[0][:3].tolist() == [5, 1, 50] and ptr[1][:3].tolist() == [6, 1, 50]
    assert ptr[0][3] == ptr[1][3]            # broadcast made rank 1 adopt rank 0's parameters
