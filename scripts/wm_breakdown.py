"""Where the world-model half of train() spends its time (torch ops): encoder / observe loop / heads+decoder / backward."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dims = bench.DIMS["crafter"]
agent = bench.build_agent(dims, 15, "cuda", 128)
wm = agent.world_model
B, T = 16, 50
N = B * T
g = torch.Generator().manual_seed(0)
obs = (torch.randint(0, 255, (N, 3, 64, 64), generator=g).float() / 255 - 0.5).cuda()
a = torch.nn.functional.one_hot(torch.randint(0, 17, (N,), generator=g), 17).float().cuda()
r = torch.randn(N, generator=g).cuda()
disc = 0.999 * torch.ones(N).cuda()
first = torch.zeros(N).cuda()
ev = lambda: torch.cuda.Event(enable_timing=True)
def run():
    e = [ev() for _ in range(5)]
    e[0].record()
    embed = wm.encoder(obs).reshape(B, T, -1)
    e[1].record()
    a_c = a.reshape(B, T, -1); first_c = first.reshape(B, T, 1)
    priors, posts = [], []
    state = wm.get_initial_state(B)
    for step in range(T):
        a_t = (a_c[:, step] * (1 - first_c[:, step])).unsqueeze(0)
        prior, post, _ = wm.recurrent_model.forward(state, embed[:, step].unsqueeze(0), a_t)
        priors.append(prior); posts.append(post); state = post
    from rl_sandbox_b200.agents.dreamer.rssm import State
    posterior, prior = State.stack(posts), State.stack(priors)
    e[2].record()
    feat = posterior.combined.transpose(0, 1)
    flat = feat.flatten(0, 1)
    loss = -wm.image_predictor(flat).log_prob(obs).float().mean() - wm.reward_predictor(feat).log_prob(r.reshape(B, T, 1)).float().mean() \
        - wm.discount_predictor(feat).log_prob(disc.reshape(B, T, 1)).float().mean() + 2 * wm._kl(prior.stoch_logits, posterior.stoch_logits)
    e[3].record()
    for p in wm.parameters(): p.grad = None
    loss.backward()
    e[4].record()
    torch.cuda.synchronize()
    return [e[i].elapsed_time(e[i + 1]) for i in range(4)]
for _ in range(3): run()
ts = [run() for _ in range(5)]
import statistics
med = [statistics.median(t[i] for t in ts) for i in range(4)]
print("encoder %.2f ms | observe loop fwd %.2f ms | heads+decoder+losses fwd %.2f ms | backward (all) %.2f ms" % tuple(med))
# backward of the observe loop alone: loss that only touches the RSSM outputs
def run2():
    e0, e1 = ev(), ev()
    with torch.no_grad():
        embed = wm.encoder(obs).reshape(B, T, -1)
    embed.requires_grad_()
    a_c = a.reshape(B, T, -1)
    state = wm.get_initial_state(B)
    acc = 0
    posts = []
    for step in range(T):
        prior, post, _ = wm.recurrent_model.forward(state, embed[:, step].unsqueeze(0), a_c[:, step].unsqueeze(0))
        state = post; posts.append(post)
        acc = acc + prior.stoch_logits.square().mean() + post.stoch_logits.square().mean() + post.determ.square().mean() + post.stoch.mean()
    e0.record(); acc.backward(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for _ in range(2): run2()
print("observe loop backward alone %.2f ms" % statistics.median(run2() for _ in range(5)))

# ---- the same with the observe loop in librlsb (K5) ----
from rl_sandbox_b200.agents.dreamer.rssm import State
def run3():
    e = [ev() for _ in range(5)]
    e[0].record()
    embed = wm.encoder(obs).reshape(B, T, -1)
    e[1].record()
    posterior, prior = wm._observe_scan(embed, a.reshape(B, T, -1))
    e[2].record()
    feat = posterior.combined.transpose(0, 1)
    flat = feat.flatten(0, 1)
    loss = -wm.image_predictor(flat).log_prob(obs).float().mean() - wm.reward_predictor(feat).log_prob(r.reshape(B, T, 1)).float().mean() \
        - wm.discount_predictor(feat).log_prob(disc.reshape(B, T, 1)).float().mean() + 2 * wm._kl(prior.stoch_logits, posterior.stoch_logits)
    e[3].record()
    for p in wm.parameters(): p.grad = None
    loss.backward()
    e[4].record()
    torch.cuda.synchronize()
    return [e[i].elapsed_time(e[i + 1]) for i in range(4)]
for _ in range(3): run3()
ts = [run3() for _ in range(5)]
med = [statistics.median(t[i] for t in ts) for i in range(4)]
print("K5 path: encoder %.2f ms | observe scan fwd %.2f ms | heads+decoder+losses fwd %.2f ms | backward (all) %.2f ms" % tuple(med))
eng = wm._observe_engine
emb = wm.encoder(obs).reshape(B, T, -1).transpose(0, 1).contiguous().detach()
act = a.reshape(B, T, -1).transpose(0, 1).contiguous()
def t_fwd():
    e0, e1 = ev(), ev(); e0.record(); out = eng.forward(emb, act, seed=1); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1), out
def t_bwd(out):
    gp = torch.randn_like(out["prior_logits"]); gd = torch.randn_like(out["determ"])
    e0, e1 = ev(), ev(); e0.record(); eng.backward(out, gp, gp, gd, gp); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
for _ in range(2): t_bwd(t_fwd()[1])
f = [t_fwd() for _ in range(5)]
print("K5 alone: fwd %.2f ms, bwd %.2f ms" % (statistics.median(x[0] for x in f), statistics.median(t_bwd(x[1]) for x in f)))
