"""Two-GPU (NCCL) test of the data-parallel hot path: start states sharded contiguously over the ranks with
``row_offset`` = the shard's first global start-state index, replicated weights, ONE flat-bucket all-reduce of the
actor + critic gradients.  The shards [0, N/2) and [N/2, N) must reproduce the single-GPU update over all N start states:
the same rollout row for row (Philox counters are global indices), the same K4 gradients up to the fp32 order of the
row sums, the same parameters after clip + AdamW.  Needs two devices; run with `gpurun --gpus 2`.
"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

M = dict(D=200, A=6, discrete=True, layer_norm=True, predict_discount=True, entropy_scale=1e-3, gamma=0.99, H=6)
N_TOTAL = 1024
SEED = 4242


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _agent_and_states(device):
    from oracle import oracle_port as orc
    from tests.test_gpu_agent import make_agent
    torch.manual_seed(0)
    agent = make_agent(M, device)
    wm, actor, critic = orc.make_params(303, D=M["D"], A=M["A"], discrete=True, layer_norm=True, predict_discount=True)
    agent.world_model.load_state_dict(wm, strict=False)
    agent.actor.load_state_dict(actor)
    agent.critic.load_state_dict(critic)
    agent.mark_weights_changed()
    h0, z0 = orc.make_start(304, N_TOTAL, M["D"])
    return agent, h0.to(device), z0.to(device)


def _one_update(agent, h0, z0, row_offset):
    from rl_sandbox.agents.dreamer.rssm import State
    n = h0.shape[0]
    init = State(h0.unsqueeze(0), torch.zeros(1, n, 32, 32, device=h0.device), z0.unsqueeze(0))
    losses, metrics = agent.behaviour_update(init, noise={"seed": SEED, "row_offset": row_offset})
    torch.cuda.synchronize()
    k1 = agent.last_rollout
    named = [("actor." + k, p) for k, p in agent.actor.actor.named_parameters()] + \
            [("critic." + k, p) for k, p in agent.critic.critic.named_parameters()]
    return {"losses": {k: float(v) for k, v in losses.items()},
            "grads": {k: p.grad.detach().cpu().clone() for k, p in named},
            "params": {k: p.detach().cpu().clone() for k, p in named},
            "determ": k1["determ"].cpu().clone(), "stoch_idx": k1["stoch_idx"].cpu().clone(),
            "actions": k1["actions"].cpu().clone(), "rewards": k1["rewards"].cpu().clone()}


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    agent, h0, z0 = _agent_and_states(dev)
    per = N_TOTAL // world
    res = _one_update(agent, h0[rank * per:(rank + 1) * per].contiguous(), z0[rank * per:(rank + 1) * per].contiguous(),
                      rank * per)
    assert agent._ac_bucket is not None and agent._ac_bucket.attached()
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


def test_two_shards_reproduce_the_single_gpu_update(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    agent, h0, z0 = _agent_and_states("cuda:0")
    single = _one_update(agent, h0, z0, 0)
    del agent
    torch.cuda.empty_cache()
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    per = N_TOTAL // world
    for r in range(world):
        sl = slice(r * per, (r + 1) * per)
        # the rollout of a shard is the corresponding rows of the full run, bit for bit
        for k in ("determ", "stoch_idx", "actions", "rewards"):
            assert torch.equal(out[r][k], single[k][:, sl]), (r, k)
    # after the all-reduce both ranks hold the same gradients = the single-GPU gradients (mean of the shard means)
    for k, g in single["grads"].items():
        g0, g1 = out[0]["grads"][k], out[1]["grads"][k]
        assert torch.equal(g0, g1), f"{k}: ranks hold different gradients after the all-reduce"
        rel = ((g0 - g).norm() / g.norm().clamp_min(1e-20)).item()
        assert rel < 2e-4, (k, rel)      # fp32 order of the row sums (and the split-count of the weight-gradient kernel)
    for k, p in single["params"].items():
        assert torch.equal(out[0]["params"][k], out[1]["params"][k]), k
        torch.testing.assert_close(out[0]["params"][k], p, rtol=0, atol=2e-6)   # AdamW moves a weight by <= lr = 1e-4
    for k, v in single["losses"].items():
        mean = 0.5 * (out[0]["losses"][k] + out[1]["losses"][k])
        assert abs(mean - v) <= 1e-5 * abs(v) + 1e-6, (k, mean, v)
