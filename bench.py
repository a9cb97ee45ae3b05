#!/usr/bin/env python
"""bench.py — imagined RSSM steps / second of the DreamerV2 hot path on B200.

A "step" is one pass of the hot path (agents/dreamer_v2.py:179-211 of the reference) over one batch
of synthetic start states: imagination rollout (K1) -> lambda-return / weights / advantage (K2) ->
critic + actor losses -> backward -> [gradient all-reduce] -> clip -> AdamW x2 -> target update.

  python bench.py --gpus N --steps K --warmup W            # the B200 arm (this repo)
  python bench.py --impl reference --gpus N ...            # the reference's own CPU path (the unmodified reference
                                                           # staged under oracle/_ref/; the oracle port if it is absent)

Workloads (config.workload):
  sweep    configs[3] of BASELINE.json: config-1 dims (D=1024, 32x32 latents, A=17 discrete,
           layer_norm, discount head), `--rows` start states PER GPU (weak scaling; default 32768,
           i.e. 262144 = the top of the 16k-256k sweep at 8 GPUs), H=15.  This is the default: it is
           the configuration the metric ("at 1/2/4/8 B200") is quoted on.
  config1  the reference shape N = 16x50 = 800 (launch/latency bound; SURVEY hard part 8)
  dino     config-2 dims (D=200, continuous A=12, rho = 0: dynamics back-propagation through K1 backward), N = 800
  crafter  configs[4]: the FULL DreamerV2.train() on a Crafter-shaped batch (16 x 50 frames of 64x64x3 uint8 per GPU):
           world-model observe + losses + AdamW (torch / cuDNN, out of kernel scope) followed by the hot path
  slotted  configs[2]: the full train() of config_slotted (slot-attention encoder, slotted RSSM, DINO-feature targets
           supplied as synthetic d_features); imagination of this config runs K1 with slots = 4 under no_grad callers,
           its (continuous-actor) training differentiates through the torch replay
One JSON line is printed by rank 0.  Besides the contract's keys it carries `roofline`, `cpu_baseline`, and (sweep /
config1 / dino workloads) `sweep_strong` (BASELINE configs[3]: N_total in {16 384, 65 536, 262 144} start states split
over the launched GPUs), `torch_gpu_baseline` (the reference's eager PyTorch path on the same B200, TF32 as its
train.py:40) and `parity_mode` (cost of the split-operand contraction mode the 1e-3 parity tests run in).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from functools import partial

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_STEP = {"c1": 26.72e6, "c2": 5.28e6}   # reference-equivalent MFLOP / start state / step
GRU_FLOP = {"c1": 2 * 3072 * 2048, "c2": 2 * 600 * 400}

DIMS = {
    "sweep": dict(D=1024, A=17, discrete=True, layer_norm=True, predict_discount=True, eta=3e-3, lr=1e-4, fl="c1"),
    "config1": dict(D=1024, A=17, discrete=True, layer_norm=True, predict_discount=True, eta=3e-3, lr=1e-4, fl="c1"),
    "dino": dict(D=200, A=12, discrete=False, layer_norm=False, predict_discount=False, eta=1e-5, lr=8e-5, fl="c2"),
    "crafter": dict(D=1024, A=17, discrete=True, layer_norm=True, predict_discount=True, eta=3e-3, lr=1e-4, fl="c1"),
    "slotted": dict(D=200, A=1, discrete=False, layer_norm=True, predict_discount=False, eta=1e-4, lr=8e-5, fl="c2",
                    slots=4),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sweep", choices=list(DIMS))
    ap.add_argument("--rows", type=int, default=None, help="start states per GPU")
    ap.add_argument("--horizon", type=int, default=15)
    ap.add_argument("--cpu-rows", type=int, default=None, help="start states of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--metrics-samples", type=int, default=128)
    ap.add_argument("--no-extras", action="store_true",
                    help="skip sweep_strong / torch_gpu_baseline / parity_mode (the default line's keys are unchanged)")
    ap.add_argument("--strong-totals", default="16384,65536,262144", help="N_total values of the strong-scaling sweep")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, source="fallback (B200_PROFILING.md)")


def init_nccl_quietly(dist, device):
    """stdout carries the one JSON line only: NCCL prints its version banner to the C-level stdout when the first
    communicator comes up, so file descriptor 1 points at stderr while the process group and its communicator are created."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=torch.device(device))
        t = torch.zeros(1, device=device)
        dist.all_reduce(t)
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self._stop = index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "samples": len(self.samples), "reasons": reasons}


# ------------------------------------------------------------------------------------------------
def reference_available() -> bool:
    """the unmodified reference: /root/reference (build container) or the copy oracle/make_ref.py stages in oracle/_ref/"""
    try:
        from oracle import ref_runner
        return ref_runner.available()
    except Exception:
        return False


def cpu_hot_path(dims, H, rows, steps, warmup, metrics_samples, device="cpu"):
    """The hot path on the host cores (or, device='cuda', the reference's eager PyTorch path on the GPU).
    kind 'reference': the UNMODIFIED reference's modules executing agents/dreamer_v2.py:179-217 (oracle/ref_runner.py);
    kind 'port': oracle/oracle_port.py::HotPathCPU when no copy of the reference is present (host only).
    Returns (steps/s, seconds per step, threads, kind)."""
    from oracle import oracle_port as orc
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    h0, z0 = orc.make_start(1, rows, dims["D"])
    if reference_available():
        from oracle import ref_runner
        hp = ref_runner.ReferenceHotPath(D=dims["D"], A=dims["A"], discrete=dims["discrete"], layer_norm=dims["layer_norm"],
                                         predict_discount=dims["predict_discount"], H=H, eta=dims["eta"], lr=dims["lr"],
                                         device=device)
        state = hp.state(h0, z0)
        run = lambda: hp.step(state)
        kind = "reference"
    else:
        if device != "cpu":
            raise RuntimeError("torch-on-GPU baseline needs the staged reference (oracle/_ref)")
        hp = orc.HotPathCPU(D=dims["D"], A=dims["A"], discrete=dims["discrete"], layer_norm=dims["layer_norm"],
                            predict_discount=dims["predict_discount"], H=H, eta=dims["eta"], lr=dims["lr"],
                            metrics_samples=metrics_samples)
        gen = torch.Generator().manual_seed(2)
        run = lambda: hp.step(h0, z0, gen)
        kind = "port"
    sync = (lambda: torch.cuda.synchronize()) if device != "cpu" else (lambda: None)
    for _ in range(warmup):
        run()
    ts = []
    for _ in range(steps):
        sync()
        t0 = time.perf_counter()
        run()            # ends with the host read of every loss / metric (dreamer_v2.py:216-217)
        sync()
        ts.append(time.perf_counter() - t0)
    sec = statistics.median(ts)
    return rows * H / sec, sec, threads, kind


def workload_name(args, dims, rows):
    if args.workload in ("crafter", "slotted"):
        return f"{args.workload}: full DreamerV2.train() on 16 x 50 frames of 64x64x3 uint8 per GPU, H={args.horizon}"
    return (f"{args.workload}: {rows} start states/GPU x H={args.horizon}, D={dims['D']}, 32x32 latents, A={dims['A']} "
            f"{'discrete' if dims['discrete'] else 'continuous'}, layer_norm={dims['layer_norm']}")


def reference_arm(args, dims):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = args.cpu_rows or 256
    val, sec, threads, kind = cpu_hot_path(dims, args.horizon, rows, args.steps, max(1, min(args.warmup, 2)),
                                           args.metrics_samples)
    # the same path at 8x the rows: steps/s of the host path must be (about) flat in the number of start states for the
    # bounded sample to stand for the b200 arm's workload
    rows_big = 8 * rows
    val_big, sec_big, _, _ = cpu_hot_path(dims, args.horizon, rows_big, min(args.steps, 2), 1, args.metrics_samples)
    cpu_model = "unknown"
    try:
        cpu_model = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    note = ("the UNMODIFIED reference (Midren/rl_sandbox) executing agents/dreamer_v2.py:179-217 through oracle/ref_runner.py "
            "(imagine_trajectory + lambda_return + calculate_loss x2 + Optimizer.step x2 + update_target + host read)"
            if kind == "reference" else
            "reference's PyTorch-CPU hot path restated in oracle/oracle_port.py (HotPathCPU); no copy of the reference present")
    line = {
        "impl": "reference", "metric": "imagined_rssm_steps_per_sec", "value": val, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the b200 arm's workload (same dims, same hot path); each timed step is a bounded sample of it
        "config": {"workload": workload_name(args, dims, args.rows or (32768 if args.workload == "sweep" else 800)),
                   "sample": f"bounded sample: {rows} start states x H={args.horizon} per step on the host cores",
                   "impl_note": note},
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": threads, "kind": kind,
                         "sample": f"{rows} start states x H={args.horizon}, median of {args.steps} steps", "cpu": cpu_model,
                         "flatness": {"rows": rows_big, "value": val_big, "ms_per_step": sec_big * 1e3,
                                      "note": f"same path at {rows_big} start states: steps/s vs the {rows}-row sample"}},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def build_agent(dims, H, device, metrics_samples):
    from rl_sandbox_b200.agents.dreamer_v2 import DreamerV2
    from rl_sandbox_b200.agents.dreamer.ac import ImaginativeActor, ImaginativeCritic
    from rl_sandbox_b200.agents.dreamer.world_model import WorldModel
    from rl_sandbox_b200.utils.optimizer import Optimizer
    ln = dims["layer_norm"]
    torch.manual_seed(0)
    opt = partial(Optimizer, lr=dims["lr"], eps=1e-5, weight_decay=1e-6, clip=100)
    if dims.get("slots"):
        from rl_sandbox_b200.agents.dreamer.world_model_slots_attention import WorldModel as SlotWM
        wm = partial(SlotWM, batch_cluster_size=50, latent_dim=32, latent_classes=32, rssm_dim=dims["D"],
                     slots_num=dims["slots"], slots_iter_num=2, kl_loss_scale=1000, kl_loss_balancing=0.8,
                     kl_free_nats=0.0005, discrete_rssm=False, decode_vit=True, vit_l2_ratio=0.75, use_prev_slots=False,
                     encode_vit=False, predict_discount=False, layer_norm=ln, discount_loss_scale=1.0, vit_img_size=224)
    else:
        wm = partial(WorldModel, batch_cluster_size=50, latent_dim=32, latent_classes=32, rssm_dim=dims["D"],
                     discount_loss_scale=1.0, kl_loss_scale=2, kl_loss_balancing=0.8, kl_free_nats=1.0,
                     discrete_rssm=False, predict_discount=dims["predict_discount"], layer_norm=ln,
                     encode_vit=False, decode_vit=False, vit_l2_ratio=0.5, vit_img_size=224)
    agent = DreamerV2(
        obs_space_num=[64, 64, 3], clip_rewards="tanh", actions_num=dims["A"],
        world_model=wm,
        actor=partial(ImaginativeActor, layer_norm=ln, reinforce_fraction=None, entropy_scale=dims["eta"]),
        critic=partial(ImaginativeCritic, discount_factor=0.999, update_interval=100, soft_update_fraction=1,
                       value_target_lambda=0.95, layer_norm=ln),
        action_type="discrete" if dims["discrete"] else "continuous", imagination_horizon=H,
        wm_optim=opt, actor_optim=opt, critic_optim=opt, layer_norm=ln, batch_cluster_size=50,
        f16_precision=False, device_type=device)
    agent.metrics_samples = metrics_samples
    return agent


def full_train_main(args, dims):
    """configs[4] / configs[2]: DreamerV2.train(RolloutChunks) end to end, host uint8 frames in, loss dict out."""
    import torch.distributed as dist
    from rl_sandbox_b200 import _lib
    from rl_sandbox_b200.utils.replay_buffer import RolloutChunks
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    _lib.require_device()
    if world > 1:
        init_nccl_quietly(dist, device)
    torch.backends.cuda.matmul.allow_tf32 = True   # reference train.py:40
    lib = _lib.load()
    H, B, T = args.horizon, 16, 50
    N = B * T
    agent = build_agent(dims, H, device, args.metrics_samples)
    g = torch.Generator().manual_seed(1 + rank)
    obs_host = torch.randint(0, 255, (N, 64, 64, 3), dtype=torch.uint8, generator=g).pin_memory()
    A = dims["A"]
    act_host = (torch.randint(0, A, (N, 1), generator=g) if dims["discrete"] else torch.randn(N, A, generator=g)).pin_memory()
    rew_host = torch.randn(N, generator=g).pin_memory()
    first = torch.zeros(N)
    first[::T] = 1
    first_host = first.pin_memory()
    dfeat_host = torch.randn(N, 384, 196, generator=g).pin_memory() if dims.get("slots") else None

    def step():
        obs = agent.preprocess_obs(obs_host.to(device, non_blocking=True))
        add = {"d_features": dfeat_host.to(device, non_blocking=True)} if dfeat_host is not None else {}
        chunks = RolloutChunks(obs=obs, actions=act_host.to(device, non_blocking=True),
                               rewards=torch.tanh(rew_host.to(device, non_blocking=True)),
                               is_finished=torch.zeros(N, device=device), is_first=first_host.to(device, non_blocking=True),
                               additional_data=add)
        return agent.train(chunks)   # ends with the device->host read of every loss (dreamer_v2.py:216-217)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        step()
    barrier()
    lib.rlsb_launch_count(1)
    flush = torch.empty(256 << 20, device=device, dtype=torch.uint8)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local) as clk:
        barrier()
        for e0, e1 in evs:
            flush.zero_()          # > 126 MB L2
            e0.record()
            out = step()
            e1.record()
        barrier()
    launches = lib.rlsb_launch_count(0)
    ms = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in evs)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / args.steps
    value = world * N * H / (ms_per_step * 1e-3)
    if rank == 0:
        h2d = obs_host.numel() + act_host.numel() * act_host.element_size() + 2 * N * 4 + (dfeat_host.numel() * 4 if dfeat_host is not None else 0)
        line = {
            "metric": "imagined_rssm_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.workload}: full DreamerV2.train() on {B} x {T} frames of 64x64x3 uint8 per GPU, H={H}, "
                                   f"D={dims['D']}, A={A} {'discrete' if dims['discrete'] else 'continuous'}"
                                   + (f", {dims['slots']} slots, slot attention 2 iterations, DINO-feature targets" if dims.get("slots") else ""),
                       "step": "world-model observe + losses + AdamW (torch/cuDNN) -> imagination + lambda-return + actor-critic update",
                       "l2": "256 MB flush buffer written between timed steps (per-step events)",
                       "metrics_samples": args.metrics_samples},
            "clocks": clk.summary(),
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4 * len(out),
                    "ms_per_step": ms_per_step, "note": "the timed call takes pinned host buffers and returns host numpy losses"},
            "gpu_launches": int(launches),
            "losses": {k: float(out[k].reshape(-1)[0]) for k in ("loss_wm", "loss_actor", "loss_critic") if k in out},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    dims = DIMS[args.workload]
    if args.impl == "reference":
        return reference_arm(args, dims)
    if args.workload in ("crafter", "slotted"):
        return full_train_main(args, dims)

    import torch.distributed as dist
    from rl_sandbox_b200 import _lib, ops
    from rl_sandbox_b200.agents.dreamer.rssm import State
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    _lib.require_device()
    if world > 1:
        init_nccl_quietly(dist, device)
    torch.backends.cuda.matmul.allow_tf32 = True   # as the reference's train.py:40 (loss MLPs run in torch)
    lib = _lib.load()

    H = args.horizon
    N = args.rows or (32768 if args.workload == "sweep" else 800)
    agent = build_agent(dims, H, device, args.metrics_samples)
    D = dims["D"]
    g = torch.Generator(device=device).manual_seed(1 + rank)
    h0 = 0.5 * torch.randn(N, D, device=device, generator=g)
    idx0 = torch.randint(0, 32, (N, 32), device=device, generator=g)
    z0 = torch.nn.functional.one_hot(idx0, 32).float().view(N, 1024)
    logits0 = torch.zeros(1, N, 32, 32, device=device)

    def make_state(h, z):
        return State(h.unsqueeze(0), logits0, z.unsqueeze(0))

    is_train = True   # discrete: K1 -> K2 -> K4; continuous: K1 (+tape) -> K2 -> K2 bwd -> K1 bwd -> K4

    def step(state, it):
        noise = {"seed": 1000 + it, "row_offset": rank * N}
        if is_train:
            losses, metrics = agent.behaviour_update(state, noise=noise)
            # every loss and metric the reference's train() hands to the host (dreamer_v2.py:216-217)
            return torch.stack([v.reshape(-1)[0].float() for v in list(losses.values()) + list(metrics.values())])
        with torch.no_grad():
            states, actions, rewards, discounts = agent.imagine_trajectory(state, noise=noise)
            vs, w, adv = ops.lambda_return(rewards, agent.last_rollout["values"].unsqueeze(-1), discounts, 0.95)
        return vs.mean()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    state = make_state(h0, z0)
    for i in range(args.warmup):
        step(state, i)
    barrier()
    lib.rlsb_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # inputs + per-step outputs exceed the 126 MB L2 from N = 8192 start states up; below that an explicit flush
    # buffer is written between timed steps (per-step events, so the flush itself is not timed)
    flush = torch.empty(256 << 20, device=device, dtype=torch.uint8) if N < 8192 else None
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local) as clk:
        barrier()
        for i, (e0, e1) in enumerate(evs):
            if flush is not None:
                flush.zero_()
            e0.record()
            out = step(state, args.warmup + i)
            e1.record()
        barrier()
    launches = lib.rlsb_launch_count(0)
    ms = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in evs)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / args.steps
    value = world * N * H / (ms_per_step * 1e-3)

    # ---- imagination-only timing (K1 alone): median of per-call CUDA-event times after 3 warm-up calls -------------
    with torch.no_grad():
        for i in range(3):
            agent.imagine_trajectory(state, noise={"seed": 1 + i})
        torch.cuda.synchronize()
        k1_evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(max(3, args.steps))]
        for i, (e0, e1) in enumerate(k1_evs):
            e0.record()
            agent.imagine_trajectory(state, noise={"seed": 10 + i})
            e1.record()
        torch.cuda.synchronize()
    k1_ms = statistics.median(e0.elapsed_time(e1) for e0, e1 in k1_evs)
    pk = peaks()
    k1_tflops = N * H * FLOP_PER_STEP[dims["fl"]] / (k1_ms * 1e-3) / 1e12

    # ---- end to end through the public API with HOST buffers ------------------------------------
    h_host = h0.cpu().pin_memory()
    i_host = idx0.to(torch.uint8).cpu().pin_memory()
    # Every step uploads its start states from pinned host memory and its result is read back on the host.  The upload
    # of step i+1 runs on a copy stream while step i computes (two device buffer sets), and the host reads the result
    # of step i-1 after it has enqueued step i — the input pipeline a training loop with a prefetching loader has.
    copy_stream = torch.cuda.Stream()
    dbuf = [(torch.empty_like(h0), torch.empty((N, 32), device=device, dtype=torch.uint8)) for _ in range(2)]
    up_ev = [torch.cuda.Event() for _ in range(2)]
    done_ev = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done_ev[i % 2])      # the step that last read this buffer set has finished
            dbuf[i % 2][0].copy_(h_host, non_blocking=True)
            dbuf[i % 2][1].copy_(i_host, non_blocking=True)
            up_ev[i % 2].record(copy_stream)

    barrier()
    ev0.record()
    upload(0)
    prev = None
    for i in range(args.steps):
        torch.cuda.current_stream().wait_event(up_ev[i % 2])
        if i + 1 < args.steps:
            upload(i + 1)
        h_dev, i_dev = dbuf[i % 2]
        z_dev = torch.nn.functional.one_hot(i_dev.long(), 32).float().view(N, 1024)
        res = step(make_state(h_dev, z_dev), 5000 + i)
        done_ev[i % 2].record()
        if prev is not None:
            res_host = prev.detach().cpu()   # device->host read of the previous step's result (blocks the host only)
        prev = res
    res_host = prev.detach().cpu()
    ev1.record()
    barrier()
    e2e_ms = torch.tensor([ev0.elapsed_time(ev1)], device=device)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * N * H / (e2e_ms.item() / args.steps * 1e-3)
    h2d = h_host.numel() * 4 + i_host.numel()
    d2h = res_host.numel() * 4

    # ---- extras: every rank takes part in the strong sweep (its all-reduce is collective); the rest is rank 0's ------
    extras = {}
    if not args.no_extras:
        extras = run_extras(args, dims, agent, world, rank, device, H, N, k1_ms, barrier, make_state)

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel: the GRU contraction (tcgen05 GEMM) — with the fused RSSM epilogues (default,
        #      D % 64 == 0, no tape) the whole GRU cell in ONE launch (EPI_GRU: LayerNorm, gates and the convex update in the
        #      contraction's epilogue), else the contraction writing fp32 pre-activations + statistics (EPI_STATS) --------
        fl = dims["fl"]
        Kg, Ng = (2048, 3072) if fl == "c1" else (512, 600)   # packed K of cat[x, h]
        from rl_sandbox_b200 import _lib as _rl
        gru_fused = fl == "c1" and _rl.load().rlsb_set_fused_rssm(-1) == 1
        xa = torch.randn(N, Kg, device=device)
        wa = torch.randn(Ng, Kg, device=device) / Kg ** 0.5
        if gru_fused:
            cell = ops.GRUCellOp(D, D).pack(wa, None, None, None)
            xp, hp = ops.pack_rows(xa[:, :D].contiguous()), ops.pack_rows(xa[:, D:].contiguous())
            h_prev = xa[:, D:].contiguous()
            del xa
            bufs = dict(h_next=torch.empty((N, D), device=device),
                        h_next_packed=torch.empty(ops.round_up(N, 128) * D, device=device, dtype=torch.bfloat16))
            run_gru = lambda: cell.forward_packed(xp, hp, h_prev, N, **bufs)
        else:
            rb, nb = ops.plan_blocks(Ng)
            xp = ops.pack_rows(xa)
            wp = ops.pack_rows(wa, row_block=rb, rows_pad=rb * nb, k_pad=Kg)
            del xa
            m_pad = ops.round_up(N, 128)
            bufs = dict(out=torch.empty((m_pad, Ng), device=device), stats=torch.empty((nb, m_pad, 2), device=device),
                        bias_p=torch.zeros(rb * nb, device=device))
            run_gru = lambda: ops.gemm_bias(xp, Kg, wp, rb, nb, None, N, Ng, want_stats=True, **bufs)
        for _ in range(3):
            run_gru()
        torch.cuda.synchronize()
        reps = 20
        ev0.record()
        for _ in range(reps):
            run_gru()
        ev1.record()
        torch.cuda.synchronize()
        gemm_ms = ev0.elapsed_time(ev1) / reps
        # beside the fused cell: the plain contraction (EPI_STATS, 256-column blocks) the unfused path runs before its gate kernel
        contraction_only = None
        if gru_fused:
            rb, nb = ops.plan_blocks(Ng)
            wp = ops.pack_rows(wa, row_block=rb, rows_pad=rb * nb, k_pad=Kg)
            xp2 = ops.pack_rows(torch.randn(N, Kg, device=device))
            m_pad = ops.round_up(N, 128)
            b2 = dict(out=torch.empty((m_pad, Ng), device=device), stats=torch.empty((nb, m_pad, 2), device=device),
                      bias_p=torch.zeros(rb * nb, device=device))
            for _ in range(3):
                ops.gemm_bias(xp2, Kg, wp, rb, nb, None, N, Ng, want_stats=True, **b2)
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(reps):
                ops.gemm_bias(xp2, Kg, wp, rb, nb, None, N, Ng, want_stats=True, **b2)
            ev1.record()
            torch.cuda.synchronize()
            c_ms = ev0.elapsed_time(ev1) / reps
            c_tf = N * GRU_FLOP[fl] / (c_ms * 1e-3) / 1e12
            contraction_only = {"kernel": "gemm_kernel<EPI_STATS> (the contraction alone: fp32 pre-activations + statistics out, "
                                          "gates in a second kernel)", "ms_per_launch": c_ms, "achieved": c_tf,
                                "frac": c_tf / pk["tf_burst"], "frac_of_sustained": c_tf / pk["tf_sus"]}
            del b2, xp2, wp
        # algorithmic FLOPs of the GRU contraction as the reference executes it (2*3D*2D per row)
        achieved = N * GRU_FLOP[fl] / (gemm_ms * 1e-3) / 1e12
        # the kernel is timed alone (20 back-to-back launches): the burst bf16 figure is its denominator; the sustained
        # figure (a kernel inside a long step) is quoted beside it.  traffic: DRAM bytes per launch of this kernel from the
        # committed ncu --set full capture (profiles/roofline_traffic.json), valid for the sweep shape only
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("rows") == N and tj.get("D") == D and bool(tj.get("gru_fused", False)) == gru_fused:
                traffic = tj["dram_bytes_per_launch"]
        roofline = {"bound": "tensor",
                    "kernel": ("gemm_kernel<EPI_GRU> (GRU cell in one launch: contraction cat[x,h] -> 3D with LayerNorm, gates and "
                               "the convex update in its epilogue)" if gru_fused else
                               "gemm_kernel<EPI_STATS> (GRU contraction cat[x,h] -> 3D)"),
                    "achieved": achieved, "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": achieved / pk["tf_burst"],
                    "peak_source": pk["source"] + ", burst bf16 (kernel timed alone)",
                    "frac_of_sustained": achieved / pk["tf_sus"], "peak_sustained": pk["tf_sus"],
                    "ms_per_launch": gemm_ms, "traffic": traffic, "contraction_only": contraction_only,
                    # the GRU CELL's bytes (common.py:69-81): x, h in (bf16 operands) + h in fp32 (the convex update's operand)
                    # + h' out (fp32 + packed bf16) + the weights; the unfused contraction instead writes its fp32
                    # pre-activations + statistics for the gate kernel, which reads them back
                    "algorithmic_bytes": N * (2 * D * 2 + D * 4 + D * 4 + D * 2) + 3 * D * 2 * D * 2,
                    "executed_bytes_unfused": N * (2 * D * 2 + 3 * D * 4) + 3 * D * 2 * D * 2,
                    "whole_rollout": {"tflops": k1_tflops, "frac": k1_tflops / pk["tf_sus"], "ms": k1_ms,
                                      "note": "reference-equivalent FLOPs of all layers / K1 time, vs sustained bf16"}}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rows = args.cpu_rows or 256
            v, sec, threads, kind = cpu_hot_path(dims, H, rows, 3, 1, args.metrics_samples)
            cpu = {"value": v, "unit": "steps/s", "cores": threads, "kind": kind,
                   "sample": f"{rows} start states x H={H} (same dims, same hot path), median of 3 steps, {sec:.2f} s/step"}
        line = {
            "metric": "imagined_rssm_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(args, dims, N),
                       "step": "imagine(K1) + lambda-return(K2) + fused critic/actor loss fwd+bwd(K4) + allreduce + AdamW x2"
                               if dims["discrete"] else
                               "imagine(K1, tape) + lambda-return(K2) + K2 bwd + rollout backward(K1 bwd) + K4 + allreduce + AdamW x2",
                       "l2": ("256 MB flush buffer written between timed steps (per-step events)" if flush is not None else
                              "per-step working set (>= 2 GB of rollout outputs) exceeds the 126 MB L2; no explicit flush"),
                       "metrics_samples": args.metrics_samples, "noise": "Philox4x32-10 on device",
                       "rollout_kernel": (f"persistent: rlsb_rollout_fwd, one launch per rollout, thread-block clusters of "
                                          f"{agent._get_engine().rollout_cluster_for(N)} per 128 start states"
                                          if getattr(agent._get_engine(), "last_rollout_persistent", False) else
                                          "chained: rlsb_imagine_fwd, one tcgen05 GEMM launch per layer (CTA pairs)")},
            "clocks": clk.summary(),
            "e2e": {"note": "pinned host start states uploaded on a copy stream one step ahead, result read back every step", "value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms.item() / args.steps},
            "gpu_launches": int(launches),
            "imagination_only": {"steps_per_sec": world * N * H / (k1_ms * 1e-3), "ms": k1_ms},
            "roofline": roofline,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_extras(args, dims, agent, world, rank, device, H, N, k1_ms, barrier, make_state):
    """sweep_strong / torch_gpu_baseline / parity_mode — measured after the headline numbers, never inside their timed
    regions.  Failures are reported as strings: an extra must not take the default line down."""
    import torch.distributed as dist
    from rl_sandbox_b200 import ops
    out = {}
    D = dims["D"]

    # ---- strong scaling (BASELINE configs[3]): N_total start states split into contiguous shards of N_total / G ------
    try:
        totals = [int(x) for x in args.strong_totals.split(",") if x]
        rows_list = []
        for n_total in totals:
            rows = n_total // world
            if rows < 128 or rows * world != n_total:
                rows_list.append(None)
                continue
            rows_list.append(rows)
        strong = []
        agent.max_rows_per_pass = 65536            # bounded HBM footprint for the 262 144-row point on one GPU
        for n_total, rows in zip(totals, rows_list):
            if rows is None:
                strong.append({"n_total": n_total, "skipped": "not divisible into shards of >= 128 rows"})
                continue
            g = torch.Generator(device=device).manual_seed(77 + rank)
            h = 0.5 * torch.randn(rows, D, device=device, generator=g)
            z = torch.nn.functional.one_hot(torch.randint(0, 32, (rows, 32), device=device, generator=g), 32).float().view(rows, 1024)
            st = type(make_state(h[:1], z[:1]))(h.unsqueeze(0), torch.zeros(1, rows, 32, 32, device=device), z.unsqueeze(0))
            steps = 3
            for i in range(2):
                agent.behaviour_update(st, noise={"seed": 7000 + i, "row_offset": rank * rows})
            barrier()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for i, (e0, e1) in enumerate(evs):
                e0.record()
                agent.behaviour_update(st, noise={"seed": 7100 + i, "row_offset": rank * rows})
                e1.record()
            barrier()
            ms = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in evs) / steps], device=device)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            strong.append({"n_total": n_total, "rows_per_gpu": rows, "ms_per_step": ms.item(),
                           "steps_per_sec": n_total * H / (ms.item() * 1e-3)})
            del st, h, z
        out["sweep_strong"] = {"n_gpus": world, "points": strong,
                               "note": "same step as `value`; total start states fixed, contiguous shards of N_total / n_gpus; "
                                       "efficiency = steps_per_sec(G) / (G x steps_per_sec(1)) across runs of this bench"}
    except Exception as exc:   # noqa: BLE001
        out["sweep_strong"] = {"error": f"{type(exc).__name__}: {exc}"}
    if rank != 0:
        return out

    # ---- cost of the split-operand ("bf16 x 3") contraction mode the 1e-3 parity tests run in (K1 only) --------------
    try:
        cfg = ops.ImagineConfig(D=D, A=dims["A"], discrete=dims["discrete"], layer_norm=dims["layer_norm"],
                                predict_discount=dims["predict_discount"], H=H, parity=True)
        eng = ops.ImaginationEngine(cfg, device=device)
        eng.pack(agent.world_model.state_dict(), agent.actor.state_dict(), agent.critic.state_dict())
        g = torch.Generator(device=device).manual_seed(5)
        h = 0.5 * torch.randn(N, D, device=device, generator=g)
        z = torch.nn.functional.one_hot(torch.randint(0, 32, (N, 32), device=device, generator=g), 32).float().view(N, 1024)
        res = eng.rollout(h, z, None, None, None, seed=1, want_stoch=False)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
        for i, (e0, e1) in enumerate(evs):
            e0.record()
            eng.rollout(h, z, None, None, None, seed=2 + i, want_stoch=False, out=res)
            e1.record()
        torch.cuda.synchronize()
        pm = statistics.median(e0.elapsed_time(e1) for e0, e1 in evs)
        out["parity_mode"] = {"k1_ms": pm, "k1_fast_ms": k1_ms, "cost_ratio": pm / k1_ms,
                              "note": "imagination rollout with every contraction as hi.Whi + hi.Wlo + lo.Whi (fp32-grade: "
                                      "<= 2e-5 of the reference on the golden fixtures, tests/test_gpu_parity_mode.py)"}
        del eng, res, h, z
    except Exception as exc:   # noqa: BLE001
        out["parity_mode"] = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- the reference's eager PyTorch path on this B200 (TF32 allowed, reference train.py:40) -----------------------
    try:
        if reference_available():
            torch.cuda.empty_cache()
            rows = min(N, 32768)
            v, sec, _, kind = cpu_hot_path(dims, H, rows, 3, 1, args.metrics_samples, device=device)
            out["torch_gpu_baseline"] = {"value": v, "unit": "steps/s", "ms_per_step": sec * 1e3, "rows": rows, "kind": kind,
                                         "note": "the unmodified reference's modules (oracle/ref_runner.py) on cuda: eager PyTorch, "
                                                 "TF32 matmuls, host read of the losses every step"}
        else:
            out["torch_gpu_baseline"] = {"unavailable": "no staged copy of the reference (oracle/_ref) on this box"}
    except Exception as exc:   # noqa: BLE001
        out["torch_gpu_baseline"] = {"error": f"{type(exc).__name__}: {exc}"}
    return out


if __name__ == "__main__":
    main()
