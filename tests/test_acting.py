"""The acting path (DreamerV2.get_action, agents/dreamer_v2.py:139-154: preprocess, conv encoder on one frame, one RSSM.forward
observe step, actor, action draw) against the REFERENCE's own get_action (fixture tests/golden/acting.npz, written by
oracle/gen_golden.py::run_acting from the unmodified reference with the torch CPU generator seeded).

CPU: the host mirror draws from the same generator in the same order, so every step — recurrent state, sampled latent, actor
probabilities, returned action — must reproduce the reference.  GPU (CUDA-graph replay): the CUDA generator draws other numbers,
so the steps are teacher-forced from the reference's states: everything that does not depend on the step's own draw (h, posterior
logits) and the actor's probabilities on the reference's state must match.
"""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import oracle_port as orc
from tests.test_gpu_agent import make_agent

GOLD = Path(__file__).resolve().parent / "golden" / "acting.npz"


def _load():
    z = np.load(GOLD)
    meta = json.loads(str(z["meta"]))
    g = torch.Generator().manual_seed(meta["frame_seed"])
    frames = torch.randint(0, 256, (meta["steps"], 64, 64, 3), generator=g, dtype=torch.uint8)
    return meta, {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}, frames


def _agent(meta, device):
    m = dict(meta, entropy_scale=1e-5, gamma=0.99, H=3)
    agent = make_agent(m, device)
    wm, actor, critic = orc.make_params(meta["param_seed"], D=meta["D"], A=meta["A"], discrete=meta["discrete"],
                                        layer_norm=meta["layer_norm"], predict_discount=meta["predict_discount"])
    agent.world_model.load_state_dict(wm, strict=False)
    agent.actor.load_state_dict(actor)
    agent.critic.load_state_dict(critic)
    agent.world_model.encoder.load_state_dict(orc.seeded_module_params(agent.world_model.encoder, meta["enc_seed"]))
    agent.mark_weights_changed()
    return agent


def test_get_action_reproduces_the_reference_on_cpu():
    meta, gold, frames = _load()
    agent = _agent(meta, "cpu")
    agent.reset()
    torch.manual_seed(meta["torch_seed"])
    with torch.no_grad():
        for t, f in enumerate(frames):
            a = agent.get_action(f.numpy())
            st = agent._state
            assert int(a) == int(gold["action"][t]), t
            assert torch.equal(st.stoch.reshape(32, 32).argmax(-1), gold["stoch_idx"][t].long()), t
            assert torch.allclose(st.determ.reshape(-1), gold["determ"][t], rtol=1e-5, atol=1e-6), t
            assert torch.allclose(st.stoch_logits.reshape(-1), gold["post_logits"][t], rtol=1e-5, atol=1e-5), t
            assert torch.allclose(agent.actor.get_action(st).probs.reshape(-1), gold["probs"][t], rtol=1e-5, atol=1e-6), t
    assert torch.allclose(agent._action_probs.reshape(-1), gold["action_probs_sum"], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_graphed_get_action_matches_the_reference_teacher_forced(cuda):
    """the CUDA-graph replay of the acting step, started at every step from the reference's previous state and action"""
    from rl_sandbox_b200.agents.dreamer.rssm import State
    meta, gold, frames = _load()
    agent = _agent(meta, cuda)
    assert agent.cuda_graph_act
    A, D = meta["A"], meta["D"]
    for t, f in enumerate(frames):
        agent.reset()
        if t > 0:   # the reference's state and action after step t - 1
            z = torch.nn.functional.one_hot(gold["stoch_idx"][t - 1].long(), 32).float()
            agent._state = State(gold["determ"][t - 1].view(1, 1, D).to(cuda), gold["post_logits"][t - 1].view(1, 1, 32, 32).to(cuda),
                                 z.view(1, 1, 1024).to(cuda))
            agent._last_action = torch.nn.functional.one_hot(gold["action"][t - 1].long(), A).float().view(1, 1, A).to(cuda)
        a = agent.get_action(f.numpy())
        st = agent._state
        assert 0 <= int(a) < A
        e_h = (st.determ.reshape(-1).cpu() - gold["determ"][t]).abs().max().item()
        e_l = (st.stoch_logits.reshape(-1).cpu() - gold["post_logits"][t]).abs().max().item()
        print(f"[parity] acting step {t}: |determ - ref| max {e_h:.2e}, |posterior logits - ref| max {e_l:.2e}")
        assert e_h < 2e-3 and e_l < 2e-2          # TF32 / cuDNN conv against the CPU's fp32
        # the actor on the reference's own state of this step
        ref_state = State(gold["determ"][t].view(1, 1, D).to(cuda), gold["post_logits"][t].view(1, 1, 32, 32).to(cuda),
                          torch.nn.functional.one_hot(gold["stoch_idx"][t].long(), 32).float().view(1, 1, 1024).to(cuda))
        p = agent.actor.get_action(ref_state).probs.reshape(-1).cpu()
        assert torch.allclose(p, gold["probs"][t], rtol=2e-2, atol=2e-3), t
