"""oracle_port.py — TEST INFRASTRUCTURE ONLY.  Never imported by rl_sandbox_b200 (the product).

CPU restatement (fp32, torch-CPU / numpy tensor arithmetic, explicit noise) of the DreamerV2
imagination + lambda-return + actor-critic hot path of Midren/rl_sandbox.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

Every function names the reference code it restates (paths relative to the reference's
``rl_sandbox/`` package).  The restatement is pinned against the reference itself: the golden
vectors under tests/golden/ were produced by importing the real reference modules in the build
container (oracle/gen_golden.py) and tests/test_oracle.py checks this port against them.

``bf16=True`` rounds the operands of every Linear to bfloat16 (fp32 accumulate) — the arithmetic
the CUDA path performs on the tensor cores — so that algorithmic differences can be told apart
from precision differences.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from pathlib import Path

import numpy as np
import torch

_HERE = Path(__file__).resolve().parent
_clib = None


def clib():
    """C oracle (oracle/rlsb_oracle.c): deterministic log / Gumbel / Philox / lambda-return."""
    global _clib
    if _clib is None:
        so = _HERE / "liboracle.so"
        if not so.exists():
            import subprocess
            subprocess.check_call(["make", "-C", os.fspath(_HERE)], stdout=subprocess.DEVNULL)
        _clib = C.CDLL(os.fspath(so))
        _clib.orc_logf.restype = C.c_float
        _clib.orc_logf.argtypes = [C.c_float]
    return _clib


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def gumbel(u) -> torch.Tensor:
    """g(u) = -log(-log u) with the bit-reproducible log (see rlsb_oracle.c)."""
    un = np.ascontiguousarray(torch.as_tensor(u).detach().cpu().numpy(), dtype=np.float32)
    g = np.empty_like(un)
    clib().orc_gumbel_array(_fp(un), _fp(g), C.c_int64(un.size))
    return torch.from_numpy(g)


def det_logf(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    clib().orc_logf_array(_fp(x), _fp(y), C.c_int64(x.size))
    return y


def philox_uniform(seed: int, n0: int, t: int, stream: int, per_row: int, rows: int) -> np.ndarray:
    out = np.empty((rows, per_row), dtype=np.float32)
    clib().orc_philox_uniform(C.c_uint64(seed), C.c_uint32(n0), C.c_uint32(t), C.c_uint32(stream),
                              C.c_int(per_row), C.c_int64(out.size), _fp(out))
    return out


def sample_categorical(logits, uniforms) -> torch.Tensor:
    """OneHotCategorical.sample as Gumbel-max: utils/dists.py:177-179, agents/dreamer/rssm.py:34-37.
    idx = argmax_k(logits_k + g(u_k)); ties -> lowest index (torch.argmax)."""
    lg = np.ascontiguousarray(torch.as_tensor(logits).detach().cpu().numpy(), dtype=np.float32)
    un = np.ascontiguousarray(torch.as_tensor(uniforms).detach().cpu().numpy(), dtype=np.float32)
    classes = lg.shape[-1]
    rows = lg.size // classes
    idx = np.empty(rows, dtype=np.int32)
    clib().orc_sample_categorical(_fp(lg), _fp(un), C.c_int64(rows), C.c_int(classes), _fp(idx))
    return torch.from_numpy(idx.reshape(lg.shape[:-1]).astype(np.int64))


def lambda_return_c(r, v, d, lam: float):
    """C oracle of ac.py:52-62 + dreamer_v2.py:192-197 + ac.py:118 on (T, N) arrays."""
    r = np.ascontiguousarray(r, dtype=np.float32)
    v = np.ascontiguousarray(v, dtype=np.float32)
    d = np.ascontiguousarray(d, dtype=np.float32)
    T, N = r.shape
    vs = np.empty((T - 1, N), np.float32)
    w = np.empty((T, N), np.float32)
    adv = np.empty((max(T - 2, 0), N), np.float32)
    clib().orc_lambda_return(_fp(r), _fp(v), _fp(d), C.c_int(T), C.c_int64(N), C.c_double(lam), _fp(vs), _fp(w),
                             _fp(adv))
    return vs, w, adv


def lambda_return_loop(vs: torch.Tensor, rs: torch.Tensor, ds: torch.Tensor, lam: float) -> torch.Tensor:
    """ImaginativeCritic._lambda_return, agents/dreamer/ac.py:52-62 (vs has one more row than rs)."""
    out = [vs[-1]]
    for i in range(rs.shape[0] - 1, -1, -1):
        out.append(rs[i] + ds[i] * ((1 - lam) * vs[i + 1] + lam * out[-1]))
    return torch.stack(out[::-1])[:-1]


# ------------------------------------------------------------------------------------------------
# layers
# ------------------------------------------------------------------------------------------------
def _r(x: torch.Tensor, bf16: bool) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32) if bf16 else x


def linear(x, w, b, bf16=False):
    """nn.Linear: x W^T + b."""
    y = _r(x, bf16) @ _r(w, bf16).t()
    return y + b if b is not None else y


def layer_norm(x, g, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def elu(x):
    return torch.where(x > 0, x, torch.expm1(x))


def mlp(x, sd: dict, prefix: str, bf16=False):
    """utils/fc_nn.py:4-23 with num_layers=5: Linear 0,3,6,9,12; LN 1 always, 4/7/10 iff present."""
    for li in (0, 3, 6, 9):
        x = linear(x, sd[f"{prefix}{li}.weight"], sd[f"{prefix}{li}.bias"], bf16)
        if f"{prefix}{li + 1}.weight" in sd:
            x = layer_norm(x, sd[f"{prefix}{li + 1}.weight"], sd[f"{prefix}{li + 1}.bias"])
        x = elu(x)
    return linear(x, sd[f"{prefix}12.weight"], sd[f"{prefix}12.bias"], bf16)


def gru_cell(x, h, sd, prefix, bf16=False):
    """GRUCell.forward, agents/dreamer/common.py:69-81 (LayerNorm over all 3D, update bias -1)."""
    parts = linear(torch.cat([x, h], -1), sd[prefix + "_layer.weight"], sd[prefix + "_layer.bias"], bf16)
    parts = layer_norm(parts, sd[prefix + "_norm.weight"], sd[prefix + "_norm.bias"])
    reset, cand, update = parts.chunk(3, -1)
    reset = torch.sigmoid(reset)
    cand = torch.tanh(reset * cand)
    update = torch.sigmoid(update - 1.0)
    return update * cand + (1 - update) * h


def rssm_predict_next(h, z, a, sd, rp="recurrent_model.", bf16=False):
    """RSSM.predict_next, agents/dreamer/rssm.py:176-193 (discrete_rssm = false)."""
    x = linear(torch.cat([z, a], -1), sd[rp + "pre_determ_recurrent.0.weight"], sd[rp + "pre_determ_recurrent.0.bias"], bf16)
    if rp + "pre_determ_recurrent.1.weight" in sd:
        x = layer_norm(x, sd[rp + "pre_determ_recurrent.1.weight"], sd[rp + "pre_determ_recurrent.1.bias"])
    x = elu(x)
    h2 = gru_cell(x, h, sd, rp + "determ_recurrent.", bf16)
    y = linear(h2, sd[rp + "ensemble_prior_estimator.0.weight"], sd[rp + "ensemble_prior_estimator.0.bias"], bf16)
    if rp + "ensemble_prior_estimator.1.weight" in sd:
        y = layer_norm(y, sd[rp + "ensemble_prior_estimator.1.weight"], sd[rp + "ensemble_prior_estimator.1.bias"])
    y = elu(y)
    logits = linear(y, sd[rp + "ensemble_prior_estimator.3.weight"], sd[rp + "ensemble_prior_estimator.3.bias"], bf16)
    return h2, logits


def bernoulli_mode(logit):
    """torch Bernoulli(logits).mode as used at world_model.py:137: (p >= .5), NaN where p == .5."""
    p = torch.sigmoid(logit)
    m = (p >= 0.5).to(p.dtype)
    m[p == 0.5] = float("nan")
    return m


def imagine(wm_sd, actor_sd, critic_sd, h0, z0, *, H, A, discrete, predict_discount, latent_uniforms,
            action_noise, logits0=None, precomp_actions=None, bf16=False, teacher=None, groups=32, classes=32,
            target_prefix="target_critic."):
    """DreamerV2.imagine_trajectory (agents/dreamer_v2.py:68-96) + WorldModel.predict_next
    (agents/dreamer/world_model.py:131-140) + the target-critic read of lambda_return (ac.py:65).

    h0 (N,D), z0 (N,1024) one-hot.  latent_uniforms (H,N,1024), action_noise (H,N,A) (uniforms for a
    discrete actor, standard normals otherwise).  ``teacher``: optional dict with 'determ' / 'stoch'
    (H+1,N,.) — when given, step t starts from the teacher's state t (teacher forcing) so one-step
    quantities can be compared without error accumulation."""
    N = h0.shape[0]
    S = groups * classes
    h, z = h0.clone(), z0.clone()
    out = {k: [] for k in ("determ", "logits", "stoch", "stoch_idx", "actions", "rewards", "discounts", "values",
                           "actor_raw")}
    out["determ"].append(h)
    out["logits"].append(logits0 if logits0 is not None else torch.zeros(N, S))
    out["stoch"].append(z)
    out["stoch_idx"].append(z.view(N, groups, classes).argmax(-1))
    out["actions"].append(torch.zeros(N, A))
    for t in range(H + 1):
        if teacher is not None:
            h, z = teacher["determ"][t], teacher["stoch"][t]
        s = torch.cat([h, z], -1)
        out["rewards"].append(mlp(s, wm_sd, "reward_predictor.", bf16).squeeze(-1))
        if t == 0 or not predict_discount:
            out["discounts"].append(torch.ones(N))
        else:
            out["discounts"].append(bernoulli_mode(mlp(s, wm_sd, "discount_predictor.", bf16).squeeze(-1)))
        if critic_sd is not None:
            out["values"].append(mlp(s, critic_sd, target_prefix, bf16).squeeze(-1))
        if t == H:
            break
        raw = mlp(s, actor_sd, "actor.", bf16)
        out["actor_raw"].append(raw)
        if precomp_actions is not None:
            a = precomp_actions[t]
        elif discrete:
            idx = sample_categorical(raw, action_noise[t])
            a = torch.nn.functional.one_hot(idx, A).float()
        else:
            mu, sd_ = raw.chunk(2, -1)
            a = torch.tanh(mu) + (2 * torch.sigmoid(sd_ / 2) + 0.1) * action_noise[t]  # unclamped rsample
        h, logits = rssm_predict_next(h, z, a, wm_sd, bf16=bf16)
        idx = sample_categorical(logits.view(N, groups, classes), latent_uniforms[t].view(N, groups, classes))
        z = torch.nn.functional.one_hot(idx, classes).float().view(N, S)
        out["determ"].append(h); out["logits"].append(logits); out["stoch"].append(z)
        out["stoch_idx"].append(idx); out["actions"].append(a)
    return {k: torch.stack(v) for k, v in out.items() if len(v)}


# ------------------------------------------------------------------------------------------------
# losses (agents/dreamer/ac.py:68-81, 113-146; agents/dreamer_v2.py:184-207)
# ------------------------------------------------------------------------------------------------
LOG_SQRT_2PI = 0.5 * math.log(2 * math.pi)


def normal_logprob(x, mean, std=None):
    if std is None:
        return -0.5 * (x - mean) ** 2 - LOG_SQRT_2PI
    return -((x - mean) ** 2) / (2 * std ** 2) - torch.log(std) - LOG_SQRT_2PI


def ac_losses(traj, actor_sd, critic_sd, *, lam, discrete, rho, eta, bf16=False):
    """Returns dict with vs, w, loss_critic, loss_actor (+ parts) for a trajectory dict from imagine()."""
    zs = torch.cat([traj["determ"], traj["stoch"]], -1)          # (H+1, N, Z)
    r, v, d = traj["rewards"], traj["values"], traj["discounts"]  # (H+1, N)
    vs = lambda_return_loop(v, r[:-1], d, lam)                    # (H, N)
    w = torch.cumprod(torch.cat([torch.ones_like(d[:1]), d[:-1]], 0), 0)
    w = w.detach()
    pred = mlp(zs[:-1].detach(), critic_sd, "critic.", bf16).squeeze(-1)          # ac.py:70
    loss_critic = -(normal_logprob(vs.detach(), pred) * w[:-1]).mean()            # ac.py:73-74
    raw = mlp(zs[:-2].detach(), actor_sd, "actor.", bf16)                         # ac.py:117
    adv = (vs[1:] - v[:-2]).detach()                                              # ac.py:118
    acts = traj["actions"][1:-1]
    if discrete:
        logp_all = torch.log_softmax(raw, -1)
        logp = (logp_all * acts.detach()).sum(-1)
        ent = -(logp_all.exp() * logp_all).sum(-1)
    else:
        mu, s_ = raw.chunk(2, -1)
        loc, scale = torch.tanh(mu), 2 * torch.sigmoid(s_ / 2) + 0.1
        logp = normal_logprob(acts.detach(), loc, scale).sum(-1)
        ent = (0.5 + LOG_SQRT_2PI + torch.log(scale)).sum(-1)
    l_reinforce = -(rho * logp * w[:-2] * adv).mean()
    l_dyn = -((1 - rho) * vs[1:] * w[:-2]).mean() if rho != 1.0 else torch.tensor(0.0)
    l_ent = -(eta * ent * w[:-2]).mean()
    return dict(vs=vs, w=w, adv=adv, loss_critic=loss_critic, loss_actor=l_reinforce + l_dyn + l_ent,
                loss_actor_reinforce=l_reinforce, loss_actor_dynamics_backprop=l_dyn, loss_actor_entropy=l_ent)


# ------------------------------------------------------------------------------------------------
# world-model observe loop (agents/dreamer/world_model.py:187-202 -> rssm.py:176-209), differentiable
# ------------------------------------------------------------------------------------------------
OBSERVE_PARAM_KEYS = [
    "pre_determ_recurrent.0.weight", "pre_determ_recurrent.0.bias", "pre_determ_recurrent.1.weight", "pre_determ_recurrent.1.bias",
    "determ_recurrent._layer.weight", "determ_recurrent._layer.bias", "determ_recurrent._norm.weight", "determ_recurrent._norm.bias",
    "ensemble_prior_estimator.0.weight", "ensemble_prior_estimator.0.bias", "ensemble_prior_estimator.1.weight",
    "ensemble_prior_estimator.1.bias", "ensemble_prior_estimator.3.weight", "ensemble_prior_estimator.3.bias",
    "stoch_net.0.weight", "stoch_net.0.bias", "stoch_net.1.weight", "stoch_net.1.bias", "stoch_net.3.weight", "stoch_net.3.bias"]


def observe_scan(wm_sd, embed, actions, uniforms, *, bf16=False, rp="recurrent_model."):
    """embed (T,B,E), actions (T,B,A) (already masked by is_first), uniforms (T,B,1024) for the posterior draws.
    Returns prior_logits, post_logits (T,B,1024), determ (T,B,D), stoch (T,B,1024) straight-through, stoch_idx."""
    T, B = embed.shape[:2]
    D = wm_sd[rp + "ensemble_prior_estimator.0.weight"].shape[0]
    h, z = torch.zeros(B, D), torch.zeros(B, 1024)
    outs = {k: [] for k in ("prior_logits", "post_logits", "determ", "stoch", "stoch_idx")}
    for t in range(T):
        h, prior_logits = rssm_predict_next(h, z, actions[t], wm_sd, rp, bf16)
        y = linear(torch.cat([h, embed[t]], -1), wm_sd[rp + "stoch_net.0.weight"], wm_sd[rp + "stoch_net.0.bias"], bf16)
        if rp + "stoch_net.1.weight" in wm_sd:
            y = layer_norm(y, wm_sd[rp + "stoch_net.1.weight"], wm_sd[rp + "stoch_net.1.bias"])
        post_logits = linear(elu(y), wm_sd[rp + "stoch_net.3.weight"], wm_sd[rp + "stoch_net.3.bias"], bf16)
        lg = post_logits.view(B, 32, 32)
        idx = sample_categorical(lg.detach(), uniforms[t].view(B, 32, 32))
        probs = torch.softmax(lg, -1)
        z = (torch.nn.functional.one_hot(idx, 32).float() + probs - probs.detach()).view(B, 1024)
        for k, v in (("prior_logits", prior_logits), ("post_logits", post_logits), ("determ", h), ("stoch", z), ("stoch_idx", idx)):
            outs[k].append(v)
    return {k: torch.stack(v) for k, v in outs.items()}


def observe_probe_loss(o, weights):
    """a scalar that touches every output of the observe scan (used to compare gradients): weights = dict of tensors"""
    return sum((o[k] * weights[k]).sum() for k in ("prior_logits", "post_logits", "determ", "stoch"))


# ------------------------------------------------------------------------------------------------
# slotted RSSM (agents/dreamer/rssm_slots_attention.py:166-209, world_model_slots_attention.py:199-207)
# ------------------------------------------------------------------------------------------------
def position_encoding(seq_len, d, n=10000):
    """agents/dreamer/common.py:8-15 (sin on even, cos on odd columns)."""
    k = torch.arange(seq_len, dtype=torch.float64)[:, None]
    i = torch.arange(d // 2, dtype=torch.float64)[None, :]
    ang = k / torch.pow(torch.tensor(float(n), dtype=torch.float64), 2 * i / d)
    P = torch.zeros(seq_len, d, dtype=torch.float64)
    P[:, 0:2 * (d // 2):2] = torch.sin(ang)
    P[:, 1:2 * (d // 2):2] = torch.cos(ang)
    return P.float()


def slot_mixer(h, sd, rp="recurrent_model.", blocks=3, coeff=1.0, symmetric_qk=False, bf16=False):
    """rssm_slots_attention.py:186-203 on h (N, K, D): returns determ_post."""
    D = h.shape[-1]
    eye = torch.eye(h.shape[-2])
    for _ in range(blocks):
        x = layer_norm(h, sd[rp + "pre_norm.weight"], sd[rp + "pre_norm.bias"])
        q, k, v = linear(x, sd[rp + "hidden_attention_proj.weight"], None, bf16).chunk(3, -1)
        if symmetric_qk:
            k = q
        qk = torch.einsum('bih,bjh->bij', q, k)
        attn = torch.softmax(D ** -0.5 * qk, -1) + 1e-8
        attn = attn / attn.sum(-1, keepdim=True)
        attn = coeff * attn + (1 - coeff) * eye
        upd = torch.einsum('bjd,bij->bid', v, attn)
        h = h + linear(layer_norm(upd, sd[rp + "fc_norm.weight"], sd[rp + "fc_norm.bias"]), sd[rp + "fc.weight"],
                       sd[rp + "fc.bias"], bf16)
    return h


def imagine_slotted(wm_sd, actor_sd, critic_sd, h0, z0, *, H, A, K, discrete, predict_discount, latent_uniforms,
                    action_noise, blocks=3, coeff=1.0, bf16=False, target_prefix="target_critic."):
    """DreamerV2.imagine_trajectory (dreamer_v2.py:68-96) over the slotted world model: h0 (N,K,D), z0 (N,K,1024),
    latent_uniforms (H,N,K,1024), action_noise (H,N,A).  Heads see cat_k([h_k, z_k] + pos_enc_k) with the UN-mixed h;
    the mixer only shapes the prior logits (rssm_slots_attention.py:205-208)."""
    N, D = h0.shape[0], h0.shape[-1]
    S = 1024
    pos = wm_sd["pos_enc"]
    rp = "recurrent_model."
    h, z = h0.clone(), z0.clone()
    out = {k: [] for k in ("determ", "logits", "stoch_idx", "actions", "rewards", "discounts", "values")}
    out["determ"].append(h); out["logits"].append(torch.zeros(N, K, S)); out["actions"].append(torch.zeros(N, A))
    out["stoch_idx"].append(z.view(N, K, 32, 32).argmax(-1))
    for t in range(H + 1):
        s = (torch.cat([h, z], -1) + pos).flatten(1, 2)
        out["rewards"].append(mlp(s, wm_sd, "reward_predictor.", bf16).squeeze(-1))
        if t == 0 or not predict_discount:
            out["discounts"].append(torch.ones(N))
        else:
            out["discounts"].append(bernoulli_mode(mlp(s, wm_sd, "discount_predictor.", bf16).squeeze(-1)))
        out["values"].append(mlp(s, critic_sd, target_prefix, bf16).squeeze(-1))
        if t == H:
            break
        raw = mlp(s, actor_sd, "actor.", bf16)
        if discrete:
            a = torch.nn.functional.one_hot(sample_categorical(raw, action_noise[t]), A).float()
        else:
            mu, sd_ = raw.chunk(2, -1)
            a = torch.tanh(mu) + (2 * torch.sigmoid(sd_ / 2) + 0.1) * action_noise[t]
        za = torch.cat([z, a.unsqueeze(1).expand(N, K, A)], -1).reshape(N * K, S + A)
        x = linear(za, wm_sd[rp + "pre_determ_recurrent.0.weight"], wm_sd[rp + "pre_determ_recurrent.0.bias"], bf16)
        if rp + "pre_determ_recurrent.1.weight" in wm_sd:
            x = layer_norm(x, wm_sd[rp + "pre_determ_recurrent.1.weight"], wm_sd[rp + "pre_determ_recurrent.1.bias"])
        x = elu(x)
        h = gru_cell(x, h.reshape(N * K, D), wm_sd, rp + "determ_recurrent.", bf16).reshape(N, K, D)
        hp = slot_mixer(h, wm_sd, rp, blocks, coeff, bf16=bf16).reshape(N * K, D)
        y = linear(hp, wm_sd[rp + "ensemble_prior_estimator.0.weight"], wm_sd[rp + "ensemble_prior_estimator.0.bias"], bf16)
        if rp + "ensemble_prior_estimator.1.weight" in wm_sd:
            y = layer_norm(y, wm_sd[rp + "ensemble_prior_estimator.1.weight"], wm_sd[rp + "ensemble_prior_estimator.1.bias"])
        logits = linear(elu(y), wm_sd[rp + "ensemble_prior_estimator.3.weight"],
                        wm_sd[rp + "ensemble_prior_estimator.3.bias"], bf16).reshape(N, K, S)
        idx = sample_categorical(logits.view(N, K, 32, 32), latent_uniforms[t].view(N, K, 32, 32))
        z = torch.nn.functional.one_hot(idx, 32).float().view(N, K, S)
        out["determ"].append(h); out["logits"].append(logits); out["stoch_idx"].append(idx); out["actions"].append(a)
    return {k: torch.stack(v) for k, v in out.items()}


def make_params_slotted(seed, *, D, A, K, discrete, layer_norm, predict_discount, hidden=400, S=1024):
    """Random parameters of the slotted world model's hot-path modules + actor / critic on K*(D+S) inputs."""
    gen = torch.Generator().manual_seed(seed)
    wm, actor, critic = make_params(seed, D=D, A=A, discrete=discrete, layer_norm=layer_norm,
                                    predict_discount=predict_discount, hidden=hidden, S=S)
    rp = "recurrent_model."
    wm = {k: v for k, v in wm.items() if k.startswith(rp) and "stoch_net" not in k}   # slotted posterior net: other width
    wm[rp + "hidden_attention_proj.weight"] = _lin(gen, 3 * D, D)[0]
    wm[rp + "fc.weight"], wm[rp + "fc.bias"] = _lin(gen, D, D)
    for name in ("pre_norm", "fc_norm"):
        wm[rp + name + ".weight"] = 1 + 0.1 * torch.randn(D, generator=gen)
        wm[rp + name + ".bias"] = 0.1 * torch.randn(D, generator=gen)
    wm["pos_enc"] = position_encoding(K, D + S)
    Z = K * (D + S)
    wm.update(make_mlp_sd(gen, "reward_predictor.", Z, 1, hidden, layer_norm))
    if predict_discount:
        wm.update(make_mlp_sd(gen, "discount_predictor.", Z, 1, hidden, layer_norm))
    actor = make_mlp_sd(gen, "actor.", Z, A if discrete else 2 * A, hidden, layer_norm)
    critic = make_mlp_sd(gen, "critic.", Z, 1, hidden, layer_norm)
    critic.update(make_mlp_sd(gen, "target_critic.", Z, 1, hidden, layer_norm))
    return wm, actor, critic


def imagine_st(wm, actor, critic, h0, z0, lat, act, *, H, A, bf16=False, keep_action_grads=False):
    """Differentiable rollout for a continuous actor (rho != 1): the autograd graph DreamerV2.imagine_trajectory
    builds (dreamer_v2.py:83-91) — actor on the DETACHED state, unclamped rsample, straight-through latents
    (rssm.py:34-37), reward head and target critic on every state.  lat: uniforms (H,N,1024), act: normals (H,N,A)."""
    N = h0.shape[0]
    h, z = h0, z0
    out = {k: [] for k in ("determ", "stoch", "actions", "rewards", "discounts", "values", "stoch_idx")}
    out["determ"].append(h); out["stoch"].append(z); out["actions"].append(torch.zeros(N, A))
    out["stoch_idx"].append(z.view(N, 32, 32).argmax(-1))
    step_actions = []
    for t in range(H + 1):
        s = torch.cat([h, z], -1)
        out["rewards"].append(mlp(s, wm, "reward_predictor.", bf16).squeeze(-1))
        out["discounts"].append(torch.ones(N))
        out["values"].append(mlp(s, critic, "target_critic.", bf16).squeeze(-1))
        if t == H:
            break
        mu, sd_ = mlp(s.detach(), actor, "actor.", bf16).chunk(2, -1)
        a = torch.tanh(mu) + (2 * torch.sigmoid(sd_ / 2) + 0.1) * act[t]
        if keep_action_grads and a.requires_grad:
            a.retain_grad()
        step_actions.append(a)
        h, logits = rssm_predict_next(h, z, a, wm, bf16=bf16)
        lg = logits.view(N, 32, 32)
        idx = sample_categorical(lg.detach(), lat[t].view(N, 32, 32))
        probs = torch.softmax(lg, -1)
        z = (torch.nn.functional.one_hot(idx, 32).float() + probs - probs.detach()).view(N, 1024)
        out["determ"].append(h); out["stoch"].append(z); out["actions"].append(a); out["stoch_idx"].append(idx)
    traj = {k: torch.stack(v) for k, v in out.items()}
    traj["_step_actions"] = step_actions
    return traj


def continuous_update_grads(wm, actor, critic, h0, z0, lat, act, *, H, A, lam=0.95, rho=0.0, eta=1e-5, bf16=False):
    """loss_actor.backward() / loss_critic.backward() of the continuous-actor hot path (dreamer_v2.py:182-207,
    optimizer.py:55-57): returns losses, d loss_actor / d a_t (H,N,A) and the parameter gradients."""
    actor = {k: v.detach().clone().requires_grad_() for k, v in actor.items()}
    critic = {k: (v.detach().clone().requires_grad_() if k.startswith("critic.") else v.detach()) for k, v in critic.items()}
    traj = imagine_st(wm, actor, critic, h0, z0, lat, act, H=H, A=A, bf16=bf16, keep_action_grads=True)
    losses = ac_losses(traj, actor, critic, lam=lam, discrete=False, rho=rho, eta=eta, bf16=bf16)
    losses["loss_actor"].backward(retain_graph=True)
    g_actions = torch.stack([a.grad if a.grad is not None else torch.zeros_like(a) for a in traj["_step_actions"]])
    losses["loss_critic"].backward()
    grads = {k: v.grad for k, v in actor.items()}
    grads |= {k: v.grad for k, v in critic.items() if k.startswith("critic.")}
    return dict(losses={k: v.detach() for k, v in losses.items()}, g_actions=g_actions.detach(), grads=grads,
                traj={k: (v.detach() if torch.is_tensor(v) else v) for k, v in traj.items() if k != "_step_actions"})


# ------------------------------------------------------------------------------------------------
# synthetic parameters with the reference's state-dict names and default nn.Linear init
# ------------------------------------------------------------------------------------------------
def _lin(gen, out_f, in_f):
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=gen) * 2 - 1) * bound
    return w, b


def make_mlp_sd(gen, prefix, n_in, n_out, hidden, layer_norm, randomize_ln=True):
    sd = {}
    dims = [n_in, hidden, hidden, hidden, hidden, n_out]
    for i, li in enumerate((0, 3, 6, 9, 12)):
        sd[f"{prefix}{li}.weight"], sd[f"{prefix}{li}.bias"] = _lin(gen, dims[i + 1], dims[i])
        if li != 12 and (li == 0 or layer_norm):
            g = 1 + 0.1 * torch.randn(hidden, generator=gen) if randomize_ln else torch.ones(hidden)
            b = 0.1 * torch.randn(hidden, generator=gen) if randomize_ln else torch.zeros(hidden)
            sd[f"{prefix}{li + 1}.weight"], sd[f"{prefix}{li + 1}.bias"] = g, b
    return sd


def seeded_module_params(module: torch.nn.Module, seed: int, prefix: str = "") -> dict:
    """Seeded parameters for a module whose state-dict names / shapes both sides share (the conv encoder of the acting
    fixture): matrices / kernels ~ N(0, 1 / fan_in), norm gains 1 + 0.1 N, everything else 0.1 N; state-dict order."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in module.state_dict().items():
        if not v.dtype.is_floating_point:
            sd[prefix + k] = v.clone()
        elif v.dim() > 1:
            sd[prefix + k] = torch.randn(v.shape, generator=gen) / float(v[0].numel()) ** 0.5
        elif k.endswith("weight"):
            sd[prefix + k] = 1 + 0.1 * torch.randn(v.shape, generator=gen)
        else:
            sd[prefix + k] = 0.1 * torch.randn(v.shape, generator=gen)
    return sd


def make_params(seed, *, D, A, discrete, layer_norm, predict_discount, hidden=400, S=1024):
    """Random parameters under the reference's names."""
    gen = torch.Generator().manual_seed(seed)
    wm, rp = {}, "recurrent_model."
    wm[rp + "pre_determ_recurrent.0.weight"], wm[rp + "pre_determ_recurrent.0.bias"] = _lin(gen, D, S + A)
    wm[rp + "determ_recurrent._layer.weight"], wm[rp + "determ_recurrent._layer.bias"] = _lin(gen, 3 * D, 2 * D)
    wm[rp + "determ_recurrent._norm.weight"] = 1 + 0.1 * torch.randn(3 * D, generator=gen)
    wm[rp + "determ_recurrent._norm.bias"] = 0.1 * torch.randn(3 * D, generator=gen)
    wm[rp + "ensemble_prior_estimator.0.weight"], wm[rp + "ensemble_prior_estimator.0.bias"] = _lin(gen, D, D)
    wm[rp + "ensemble_prior_estimator.3.weight"], wm[rp + "ensemble_prior_estimator.3.bias"] = _lin(gen, S, D)
    if layer_norm:
        for name in ("pre_determ_recurrent.1", "ensemble_prior_estimator.1"):
            wm[rp + name + ".weight"] = 1 + 0.1 * torch.randn(D, generator=gen)
            wm[rp + name + ".bias"] = 0.1 * torch.randn(D, generator=gen)
    # posterior net (rssm.py:156-165), drawn from its own generator so that older fixtures keep their parameters
    gen_post = torch.Generator().manual_seed(seed + 7919)
    wm[rp + "stoch_net.0.weight"], wm[rp + "stoch_net.0.bias"] = _lin(gen_post, D, D + 1536)
    wm[rp + "stoch_net.3.weight"], wm[rp + "stoch_net.3.bias"] = _lin(gen_post, S, D)
    if layer_norm:
        wm[rp + "stoch_net.1.weight"] = 1 + 0.1 * torch.randn(D, generator=gen_post)
        wm[rp + "stoch_net.1.bias"] = 0.1 * torch.randn(D, generator=gen_post)
    wm.update(make_mlp_sd(gen, "reward_predictor.", D + S, 1, hidden, layer_norm))
    if predict_discount:
        wm.update(make_mlp_sd(gen, "discount_predictor.", D + S, 1, hidden, layer_norm))
    actor = make_mlp_sd(gen, "actor.", D + S, A if discrete else 2 * A, hidden, layer_norm)
    critic = make_mlp_sd(gen, "critic.", D + S, 1, hidden, layer_norm)
    critic.update(make_mlp_sd(gen, "target_critic.", D + S, 1, hidden, layer_norm))
    return wm, actor, critic


def make_start(seed, N, D, groups=32, classes=32):
    """determ ~ 0.5 N(0,1); stoch = one-hot of randint(classes) per group."""
    gen = torch.Generator().manual_seed(seed)
    h0 = 0.5 * torch.randn(N, D, generator=gen)
    idx = torch.randint(0, classes, (N, groups), generator=gen)
    z0 = torch.nn.functional.one_hot(idx, classes).float().view(N, groups * classes)
    return h0, z0


# ------------------------------------------------------------------------------------------------
# whole hot path on the CPU (bench.py cpu_baseline / --impl reference): imagine -> lambda-return ->
# critic / actor losses -> backward -> clip -> AdamW x2 -> target update
# (agents/dreamer_v2.py:179-211, utils/optimizer.py:43-71)
# ------------------------------------------------------------------------------------------------
class HotPathCPU:
    def __init__(self, *, D, A, discrete, layer_norm, predict_discount, H=15, lam=0.95, eta=3e-3, lr=1e-4,
                 seed=0, metrics_samples=128):
        self.cfg = dict(D=D, A=A, discrete=discrete, layer_norm=layer_norm, predict_discount=predict_discount)
        self.H, self.lam, self.eta, self.A, self.discrete = H, lam, eta, A, discrete
        self.rho = 1.0 if discrete else 0.0
        self.metrics_samples = metrics_samples
        self.wm, self.actor, self.critic = make_params(seed, **self.cfg)
        self.actor = {k: v.requires_grad_() for k, v in self.actor.items()}
        train_c = {k: v.requires_grad_() for k, v in self.critic.items() if k.startswith("critic.")}
        self.critic.update(train_c)
        mk = lambda ps: torch.optim.AdamW(ps, lr=lr, eps=1e-5, weight_decay=1e-6)
        self.opt_a, self.opt_c = mk(list(self.actor.values())), mk(list(train_c.values()))
        self._updates = 0

    def step(self, h0, z0, gen):
        H, N, A = self.H, h0.shape[0], self.A
        lat = torch.rand(H, N, 1024, generator=gen)
        act = torch.rand(H, N, A, generator=gen) if self.discrete else torch.randn(H, N, A, generator=gen)
        if self.discrete:      # rho == 1: nothing differentiates through the rollout (SURVEY hard part 3)
            with torch.no_grad():
                traj = imagine(self.wm, self.actor, self.critic, h0, z0, H=H, A=A, discrete=True,
                               predict_discount=self.cfg["predict_discount"], latent_uniforms=lat, action_noise=act)
        else:
            traj = self._imagine_st(h0, z0, lat, act)
        losses = ac_losses(traj, self.actor, self.critic, lam=self.lam, discrete=self.discrete, rho=self.rho,
                           eta=self.eta)
        with torch.no_grad():  # the 128-draw action statistics of ac.py:137-143
            raw = mlp(torch.cat([traj["determ"], traj["stoch"]], -1)[:-2].detach(), self.actor, "actor.")
            if self.discrete:
                p = torch.softmax(raw, -1).reshape(-1, A)
                s = torch.nn.functional.one_hot(torch.multinomial(p, self.metrics_samples, True, generator=gen), A).float()
                stats = (s.mean(), s.var())
            else:
                mu, sd_ = raw.chunk(2, -1)
                s = torch.tanh(mu) + (2 * torch.sigmoid(sd_ / 2) + 0.1) * torch.randn((self.metrics_samples,) + mu.shape, generator=gen)
                stats = (s.mean(), s.var())
        for opt, loss in ((self.opt_a, losses["loss_actor"]), (self.opt_c, losses["loss_critic"])):
            opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_([p for g in opt.param_groups for p in g["params"]], 100)
            opt.step()
        if self._updates % 100 == 0:
            with torch.no_grad():
                for k in list(self.critic):
                    if k.startswith("target_critic."):
                        self.critic[k] = self.critic["critic." + k[len("target_critic."):]].detach().clone()
        self._updates += 1
        return {k: float(v) for k, v in losses.items() if v.ndim == 0}

    def _imagine_st(self, h0, z0, lat, act):
        """continuous actor: rollout with straight-through latents so that the dynamics loss reaches the actor"""
        return imagine_st(self.wm, self.actor, self.critic, h0, z0, lat, act, H=self.H, A=self.A)


# ------------------------------------------------------------------------------------------------
# K3 oracle: SlotAttention.forward, rl_sandbox/vision/slot_attention.py:52-77 (explicit prev_slots)
# ------------------------------------------------------------------------------------------------
def slot_attention(X, slots, sd, n_iter, bf16=False, prefix=""):
    """X (B,T,dim), slots (B,K,dim); sd holds the reference's parameter names.  Returns (slots, attn).
    bf16=True rounds the operands of the contractions that the CUDA path runs on tensor cores
    (k/v projection, q, GRU, MLP) and stores k, v in bf16, like the kernel."""
    g = lambda k: sd[prefix + k]
    dim = X.shape[-1]
    B, K = slots.shape[0], slots.shape[1]
    ln = lambda x, n: layer_norm(x, g(n + ".weight"), g(n + ".bias"))
    kv = linear(ln(X, "inputs_norm"), g("inputs_proj.weight"), None, bf16)
    if bf16:
        kv = _r(kv, True)
    k, v = kv.chunk(2, -1)
    attn = None
    for _ in range(n_iter):
        prev = slots
        q = linear(ln(slots, "slots_norm"), g("slots_proj.weight"), None, bf16)
        logits = dim ** -0.5 * torch.einsum("bik,bjk->bij", q, k)
        attn = torch.softmax(logits, dim=1) + 1e-8                      # softmax over SLOTS
        attn = attn / attn.sum(-1, keepdim=True)                         # renormalise over tokens
        upd = torch.einsum("bjd,bij->bid", v, attn).reshape(B * K, dim)
        h = prev.reshape(B * K, dim)
        gi = linear(upd, g("slots_reccur.weight_ih"), g("slots_reccur.bias_ih"), bf16)
        gh = linear(h, g("slots_reccur.weight_hh"), g("slots_reccur.bias_hh"), bf16)
        ir, iz, in_ = gi.chunk(3, -1)
        hr, hz, hn = gh.chunk(3, -1)
        r, z = torch.sigmoid(ir + hr), torch.sigmoid(iz + hz)
        n = torch.tanh(in_ + r * hn)
        s = ((1 - z) * n + z * h).reshape(B, K, dim)
        hid = torch.relu(linear(ln(s, "slots_norm_2"), g("slots_proj_2.0.weight"), g("slots_proj_2.0.bias"), bf16))
        slots = s + linear(hid, g("slots_proj_2.2.weight"), g("slots_proj_2.2.bias"), bf16)
    return slots, attn


def make_slot_params(seed, dim=384, slots=4):
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for name in ("inputs_norm", "slots_norm", "slots_norm_2"):
        sd[name + ".weight"] = 1 + 0.1 * torch.randn(dim, generator=gen)
        sd[name + ".bias"] = 0.1 * torch.randn(dim, generator=gen)
    sd["inputs_proj.weight"], _ = _lin(gen, 2 * dim, dim)
    sd["slots_proj.weight"], _ = _lin(gen, dim, dim)
    sd["slots_reccur.weight_ih"], sd["slots_reccur.bias_ih"] = _lin(gen, 3 * dim, dim)
    sd["slots_reccur.weight_hh"], sd["slots_reccur.bias_hh"] = _lin(gen, 3 * dim, dim)
    sd["slots_proj_2.0.weight"], sd["slots_proj_2.0.bias"] = _lin(gen, 4 * dim, dim)
    sd["slots_proj_2.2.weight"], sd["slots_proj_2.2.bias"] = _lin(gen, dim, 4 * dim)
    sd["slots_mu"] = torch.randn(1, slots, dim, generator=gen)
    sd["slots_logsigma"] = 0.1 * torch.randn(1, slots, dim, generator=gen)
    return sd
