"""Torch-facing wrappers over the C ABI (include/rlsb.h).

torch supplies device memory and the current stream; all arithmetic happens in librlsb.so.
Every function raises if the library or a B200 device is missing — there is no fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import (AcCfg, ActorSlots, ImagineCfg, ImagineOut, ImagineParams, MlpGrads, MlpParams, Noise, SlotCfg,
                   SlotGrads, SlotParams, check)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_cuda:
        raise _lib.RlsbError(f"expected a CUDA float32 tensor, got {t.dtype} on {t.device}")
    # the kernels are enqueued on the CURRENT device's stream (_stream) with per-device function attributes: a tensor
    # living on another GPU of the process would be a silent cross-device launch
    if t.device.index is not None and t.device.index != torch.cuda.current_device():
        raise _lib.RlsbError(f"tensor on {t.device} while the current CUDA device is cuda:{torch.cuda.current_device()}: "
                             f"wrap the call in torch.cuda.device({t.device.index})")
    return t.contiguous()


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# ------------------------------------------------------------------------------------------------
# K2
# ------------------------------------------------------------------------------------------------
def lambda_return(r: torch.Tensor, v: torch.Tensor, d: torch.Tensor, lambda_: float,
                  batch_major: bool = False, want_weights: bool = True, want_adv: bool = True):
    """lambda-return + cumprod weights + advantage (ac.py:52-66, dreamer_v2.py:192-197, ac.py:118).

    time-major: r, v, d are (T, N[, 1]); returns vs (T-1, N), w (T, N), adv (T-2, N).
    """
    _lib.require_device()
    lib = _lib.load()
    r, v, d = _f32c(r), _f32c(v), _f32c(d)
    if batch_major:
        N, T = v.shape[0], v.shape[1]
        if r.shape != v.shape or d.shape != v.shape:
            raise _lib.RlsbError("lambda_return(batch_major): r, v, d must all be (N, T)")
        vs = torch.empty((N, T - 1), device=r.device, dtype=torch.float32)
        w = torch.empty((N, T), device=r.device, dtype=torch.float32) if want_weights else None
        adv = torch.empty((N, T - 2), device=r.device, dtype=torch.float32) if want_adv else None
    else:
        # v has T = H+1 rows; r and d may have H or H+1 rows (only rows 0..H-1 are read, ac.py:57-58)
        T = v.shape[0]
        N = v[0].numel()
        for name, x in (("r", r), ("d", d)):
            if x.shape[0] not in (T - 1, T) or x[0].numel() != N:
                raise _lib.RlsbError(f"lambda_return: {name} has shape {tuple(x.shape)}, v {tuple(v.shape)}")
        vs = torch.empty((T - 1,) + tuple(v.shape[1:]), device=r.device, dtype=torch.float32)
        w = torch.empty_like(v) if want_weights else None
        adv = torch.empty((T - 2,) + tuple(v.shape[1:]), device=r.device, dtype=torch.float32) if want_adv else None
    check(lib.rlsb_lambda_return_fwd(r.data_ptr(), v.data_ptr(), d.data_ptr(), T, N, float(lambda_),
                                     vs.data_ptr(), _ptr(w), _ptr(adv), int(batch_major), _stream()),
          "rlsb_lambda_return_fwd")
    return vs, w, adv


def lambda_return_bwd(g_vs, v, d, vs, lambda_):
    _lib.require_device()
    lib = _lib.load()
    g_vs, v, d, vs = _f32c(g_vs), _f32c(v), _f32c(d), _f32c(vs)
    T = v.shape[0]
    N = v[0].numel()
    g_r, g_v, g_d = torch.empty_like(v), torch.empty_like(v), torch.empty_like(v)
    check(lib.rlsb_lambda_return_bwd(g_vs.data_ptr(), v.data_ptr(), d.data_ptr(), vs.data_ptr(), T, N,
                                     float(lambda_), g_r.data_ptr(), g_v.data_ptr(), g_d.data_ptr(), _stream()),
          "rlsb_lambda_return_bwd")
    return g_r, g_v, g_d


class LambdaReturnFn(torch.autograd.Function):
    """Differentiable K2 (needed when rho != 1: dynamics back-propagation, ac.py:121-123)."""

    @staticmethod
    def forward(ctx, r, v, d, lambda_):
        vs, _, _ = lambda_return(r, v, d, lambda_, want_weights=False, want_adv=False)
        ctx.save_for_backward(v, d, vs)
        ctx.lambda_ = lambda_
        ctx.rows = (r.shape[0], d.shape[0])
        return vs

    @staticmethod
    def backward(ctx, g_vs):
        v, d, vs = ctx.saved_tensors
        g_r, g_v, g_d = lambda_return_bwd(g_vs.contiguous(), v, d, vs, ctx.lambda_)
        return g_r[:ctx.rows[0]], g_v, g_d[:ctx.rows[1]], None


# ------------------------------------------------------------------------------------------------
# sampler / RNG
# ------------------------------------------------------------------------------------------------
def sample_categorical(logits: torch.Tensor, uniforms: torch.Tensor) -> torch.Tensor:
    """idx = argmax_k(logits + gumbel(u)) over the last axis (dists.py:177-179); int32 indices."""
    _lib.require_device()
    lib = _lib.load()
    logits, uniforms = _f32c(logits), _f32c(uniforms)
    classes = logits.shape[-1]
    rows = logits.numel() // classes
    idx = torch.empty(logits.shape[:-1], device=logits.device, dtype=torch.int32)
    check(lib.rlsb_sample_categorical(logits.data_ptr(), uniforms.data_ptr(), rows, classes, idx.data_ptr(),
                                      _stream()), "rlsb_sample_categorical")
    return idx


def sample_latent(logits: torch.Tensor, uniforms: Optional[torch.Tensor] = None, seed: int = 0, row_offset: int = 0,
                  step: int = 0, want_onehot: bool = False):
    """The rollout's latent draw on its own (rssm.py:34-37): logits (rows, groups, 32) -> uint8 indices (rows, groups)
    [, fp32 one-hot (rows, groups*32)].  uniforms=None draws Philox noise exactly as K1 does."""
    _lib.require_device()
    lib = _lib.load()
    logits = _f32c(logits)
    rows, groups, classes = logits.shape
    assert classes == 32
    if uniforms is not None:
        uniforms = _f32c(uniforms)
        assert uniforms.numel() == logits.numel()
    idx = torch.empty((rows, groups), device=logits.device, dtype=torch.uint8)
    onehot = torch.empty((rows, groups * 32), device=logits.device) if want_onehot else None
    check(lib.rlsb_sample_latent(logits.data_ptr(), rows, groups, uniforms.data_ptr() if uniforms is not None else None,
                                 seed, row_offset, step, idx.data_ptr(), onehot.data_ptr() if want_onehot else None,
                                 _stream()), "rlsb_sample_latent")
    return (idx, onehot) if want_onehot else idx


def philox_uniform(seed: int, n0: int, t: int, stream_id: int, per_row: int, rows: int,
                   device="cuda") -> torch.Tensor:
    _lib.require_device()
    lib = _lib.load()
    out = torch.empty((rows, per_row), device=device, dtype=torch.float32)
    check(lib.rlsb_philox_uniform(seed, n0, t, stream_id, per_row, out.numel(), out.data_ptr(), _stream()),
          "rlsb_philox_uniform")
    return out


# ------------------------------------------------------------------------------------------------
# packed operands + GEMM (test surface for the kernel every layer uses)
# ------------------------------------------------------------------------------------------------
def pack_rows(x: torch.Tensor, row_block: int = 128, rows_pad: Optional[int] = None,
              k_pad: Optional[int] = None) -> torch.Tensor:
    _lib.require_device()
    lib = _lib.load()
    x = _f32c(x)
    rows, cols = x.shape
    rows_pad = rows_pad or round_up(rows, row_block)
    k_pad = k_pad or round_up(cols, 64)
    out = torch.empty(rows_pad * k_pad, device=x.device, dtype=torch.bfloat16)
    check(lib.rlsb_pack_rows(x.data_ptr(), cols, rows, out.data_ptr(), row_block, rows_pad, k_pad, 0, 0, cols,
                             _stream()), "rlsb_pack_rows")
    return out


def unpack_rows(packed: torch.Tensor, rows: int, cols: int, row_block: int = 128,
                k_pad: Optional[int] = None) -> torch.Tensor:
    """Inverse of the packed layout (host-side index math; tests only)."""
    k_pad = k_pad or round_up(cols, 64)
    dev = packed.device
    r = torch.arange(rows, device=dev).view(-1, 1)
    k = torch.arange(cols, device=dev).view(1, -1)
    rb, rr = r // row_block, r % row_block
    kt, kk = k // 64, k % 64
    chunk = (kk // 8) ^ (rr % 8)
    idx = ((rb * (k_pad // 64) + kt) * row_block + rr) * 64 + chunk * 8 + (kk % 8)
    return packed[idx.reshape(-1)].view(rows, cols).float()


def plan_blocks(n: int) -> tuple[int, int]:
    """(row_block, n_blocks) the library uses for an output width n (see rlsb_imagine.cu::plan_nb)."""
    if round_up(n, 32) <= 512:
        return round_up(n, 32), 1
    nb = (n + 255) // 256
    return round_up((n + nb - 1) // nb, 32), nb


def gemm_bias(a_packed, k_pad, w_packed, rb, nb, bias, M, N, want_stats=False, out=None, stats=None, bias_p=None):
    """out, stats, bias_p may be pre-allocated by the caller (benchmark loops)."""
    _lib.require_device()
    lib = _lib.load()
    m_pad = round_up(M, 128)
    if out is None:
        out = torch.zeros((m_pad, N), device=a_packed.device, dtype=torch.float32)
    if want_stats and stats is None:
        stats = torch.zeros((nb, m_pad, 2), device=a_packed.device, dtype=torch.float32)
    if bias_p is None:
        bias_p = torch.zeros(rb * nb, device=a_packed.device, dtype=torch.float32)
        if bias is not None:
            bias_p[:N] = bias
    check(lib.rlsb_gemm_bias(a_packed.data_ptr(), k_pad, w_packed.data_ptr(), rb, nb, bias_p.data_ptr(), M, N,
                             out.data_ptr(), N, _ptr(stats) if want_stats else None, _stream()), "rlsb_gemm_bias")
    return out[:M], stats


def gemm_ln_act(a_packed, k_pad, w_packed, rb, bias, M, N, gamma, beta, eps, act, out_kpad=None):
    _lib.require_device()
    lib = _lib.load()
    m_pad = round_up(M, 128)
    out_kpad = out_kpad or round_up(N, 64)
    out = torch.empty(m_pad * out_kpad, device=a_packed.device, dtype=torch.bfloat16)
    pad = lambda t, fill: None if t is None else torch.cat(
        [t.float(), torch.full((rb - N,), fill, device=t.device)]).contiguous()
    bias_p, g_p, b_p = pad(bias, 0.0), pad(gamma, 1.0), pad(beta, 0.0)
    check(lib.rlsb_gemm_ln_act(a_packed.data_ptr(), k_pad, w_packed.data_ptr(), rb, _ptr(bias_p), M, N,
                               _ptr(g_p), _ptr(b_p), float(eps), int(act), out.data_ptr(), out_kpad, _stream()),
          "rlsb_gemm_ln_act")
    return out


class GRUCellOp:
    """`GRUCell` of the reference (agents/dreamer/common.py:58-81, norm=True) as ONE launch: the tcgen05 contraction over
    cat[x, h] with LayerNorm, gates and the convex update in its epilogue (csrc/rlsb_gemm.cu, EPI_GRU) — the kernel the chained
    rollout runs per imagined step.  `pack(weight, bias, ln_weight, ln_bias)` takes the module's parameters
    (`_layer.weight` [3D, Dx + D], `_layer.bias`, `_norm.weight`, `_norm.bias`); `forward(x, h)` returns h' (fp32) and keeps the
    packed bf16 image of h' in `self.h_packed`."""

    def __init__(self, Dx: int, D: int, update_bias: float = -1.0, eps: float = 1e-5):
        _lib.require_device()
        self.lib = _lib.load()
        self.Dx, self.D, self.update_bias, self.eps = int(Dx), int(D), float(update_bias), float(eps)
        nbytes = self.lib.rlsb_gru_cell_packed_bytes(self.Dx, self.D)
        if nbytes == 0:
            raise ValueError(f"GRUCellOp: unsupported sizes Dx={Dx}, D={D} (D % 64 == 0 and 3 D > 512 required)")
        self.packed = torch.empty(nbytes, device="cuda", dtype=torch.uint8)
        self._ws = None
        self.h_packed = None

    def pack(self, weight, bias, ln_weight, ln_bias):
        w = _f32c(weight)
        assert tuple(w.shape) == (3 * self.D, self.Dx + self.D)
        keep = [w] + [None if t is None else _f32c(t) for t in (bias, ln_weight, ln_bias)]
        check(self.lib.rlsb_gru_cell_pack(w.data_ptr(), _ptr(keep[1]), _ptr(keep[2]), _ptr(keep[3]), self.Dx, self.D,
                                          self.packed.data_ptr(), _stream()), "rlsb_gru_cell_pack")
        return self

    def forward_packed(self, x_packed, h_packed, h_prev, M, h_next=None, h_next_packed=None):
        """operands already in the packed bf16 layout (benchmark loops: no re-pack inside)"""
        need = self.lib.rlsb_gru_cell_workspace_bytes(self.D, M)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, device=h_prev.device, dtype=torch.uint8)
        if h_next is None:
            h_next = torch.empty((M, self.D), device=h_prev.device, dtype=torch.float32)
        if h_next_packed is None:
            h_next_packed = torch.empty(round_up(M, 128) * self.D, device=h_prev.device, dtype=torch.bfloat16)
        check(self.lib.rlsb_gru_cell_fwd(self.packed.data_ptr(), self.Dx, self.D, x_packed.data_ptr(), h_packed.data_ptr(),
                                         h_prev.data_ptr(), M, self.update_bias, self.eps, h_next.data_ptr(),
                                         h_next_packed.data_ptr(), self._ws.data_ptr(), _stream()), "rlsb_gru_cell_fwd")
        self.h_packed = h_next_packed
        return h_next

    def forward(self, x, h):
        x, h = _f32c(x), _f32c(h)
        return self.forward_packed(pack_rows(x), pack_rows(h), h, x.shape[0])


def gemm_wgrad(dy_packed, n_pad, x_packed, k_pad, M):
    """out[n_pad, k_pad] = dY^T X from two packed images (weight-gradient contraction; test surface)."""
    _lib.require_device()
    lib = _lib.load()
    ws = torch.empty(lib.rlsb_gemm_wgrad_workspace_bytes(n_pad, k_pad, M), device=dy_packed.device, dtype=torch.uint8)
    out = torch.zeros((n_pad, k_pad), device=dy_packed.device, dtype=torch.float32)
    check(lib.rlsb_gemm_wgrad(dy_packed.data_ptr(), n_pad, x_packed.data_ptr(), k_pad, M, out.data_ptr(),
                              ws.data_ptr(), _stream()), "rlsb_gemm_wgrad")
    return out


# ------------------------------------------------------------------------------------------------
# K1
# ------------------------------------------------------------------------------------------------
@dataclass
class ImagineConfig:
    D: int
    A: int
    discrete: bool
    layer_norm: bool
    predict_discount: bool
    H: int = 15
    groups: int = 32
    classes: int = 32
    hidden: int = 400
    with_critic: bool = True
    discount_nan_on_tie: bool = True   # reference-exact Bernoulli.mode (NaN at p == 0.5)
    with_backward: bool = False        # also pack the transposed images rlsb_imagine_bwd needs (D <= 512)
    slots: int = 1                     # > 1: slotted RSSM (rssm_slots_attention.py), rows ordered (n, slot)
    attention_blocks: int = 3
    symmetric_qk: bool = False
    mixer_coeff: float = 1.0           # attention_scheduler.val
    parity: bool = False               # split-operand ("bf16 x 3") contractions: fp32-grade results, 3x tensor work
    last_step_value_only: bool = False  # training path: at step H only the target critic's head runs (values[H])
    rollout_cluster: int = 0           # persistent rollout: CTAs per cluster (4 | 8 | 16); 0 = chosen per call from the rows

    def to_c(self) -> ImagineCfg:
        return ImagineCfg(self.D, self.groups, self.classes, self.A, self.hidden, int(self.discrete),
                          int(self.layer_norm), int(self.predict_discount), int(self.with_critic), self.H,
                          int(self.discount_nan_on_tie), int(self.with_backward), int(self.slots),
                          int(self.attention_blocks), int(self.symmetric_qk), float(self.mixer_coeff), int(self.parity),
                          int(self.last_step_value_only), int(self.rollout_cluster))


def _mlp_params(sd: dict, prefix: str, keep: list) -> MlpParams:
    """fc_nn.py Sequential indices: Linear 0,3,6,9,12; LayerNorm 1,4,7,10."""
    mp = MlpParams()
    for i, li in enumerate((0, 3, 6, 9, 12)):
        w, b = _f32c(sd[f"{prefix}{li}.weight"]), _f32c(sd[f"{prefix}{li}.bias"])
        keep += [w, b]
        mp.w[i], mp.b[i] = w.data_ptr(), b.data_ptr()
    for i, li in enumerate((1, 4, 7, 10)):
        key = f"{prefix}{li}.weight"
        if key in sd:
            g, b = _f32c(sd[key]), _f32c(sd[f"{prefix}{li}.bias"])
            keep += [g, b]
            mp.ln_g[i], mp.ln_b[i] = g.data_ptr(), b.data_ptr()
    return mp


class ImaginationEngine:
    """Owns the packed bf16 weight images and the activation workspace of K1.

    ``pack`` must be called whenever the fp32 parameters change (once per optimizer step);
    ``rollout`` replaces DreamerV2.imagine_trajectory (dreamer_v2.py:68-96) for N start states.
    """

    def __init__(self, cfg: ImagineConfig, device="cuda"):
        _lib.require_device()
        self.lib = _lib.load()
        self.cfg = cfg
        self.ccfg = cfg.to_c()
        self.device = torch.device(device)
        nbytes = self.lib.rlsb_imagine_packed_bytes(C.byref(self.ccfg))
        if nbytes == 0:
            raise _lib.RlsbError(f"unsupported imagination config {cfg}")
        self.packed = torch.zeros(nbytes, device=self.device, dtype=torch.uint8)
        # The persistent rollout kernel (rlsb_rollout_fwd: the whole H-step rollout in ONE launch, a thread-block cluster
        # per 128 start states) takes over below `persistent_max_rows` start states, where the chained rollout is bound
        # by the latency of its ~230 dependent launches.  RLSB_PERSISTENT=0 disables it, =1 forces it for any size.
        env = os.environ.get("RLSB_PERSISTENT", "")
        self.persistent_max_rows = {"0": 0, "1": 1 << 30}.get(env, 2048)
        self._max_clusters = {}   # cluster size -> clusters the device keeps resident (rlsb_rollout_max_clusters)
        ro_bytes = self.lib.rlsb_rollout_packed_bytes(C.byref(self.ccfg)) if self.persistent_max_rows > 0 else 0
        # its weights are re-ordered per CTA of the cluster, i.e. they depend on the cluster size: packed on first use
        # after every `pack` (one blob per cluster size seen)
        self.packed_ro: Optional[dict] = {} if ro_bytes else None
        self._ro_version: dict = {}
        self._pack_version = 0
        self._chained_version = None   # (pack version, fused-epilogue switch) the chained blob was built for
        self._params = None
        self._ws = None
        self._ws_bytes = 0
        # workspaces owned by a captured CUDA graph (key = the graph's identity): a graph replays raw pointers, so its
        # buffers are allocated once per key and never resized or freed while the engine lives; eager calls use the
        # growable buffers above, which no graph ever references
        self._pinned: dict = {}

    # state-dict keys follow SURVEY Appendix A.1 / A.2 (reference module attribute names)
    def pack(self, wm_sd: dict, actor_sd: dict, critic_sd: Optional[dict],
             rssm_prefix="recurrent_model.", target_prefix="target_critic.") -> None:
        keep: list = []
        p = ImagineParams()

        def take(sd, key):
            if key not in sd:
                return None
            t = _f32c(sd[key])
            keep.append(t)
            return t.data_ptr()

        rp = rssm_prefix
        p.img_in_w, p.img_in_b = take(wm_sd, rp + "pre_determ_recurrent.0.weight"), take(wm_sd, rp + "pre_determ_recurrent.0.bias")
        p.img_in_ln_g, p.img_in_ln_b = take(wm_sd, rp + "pre_determ_recurrent.1.weight"), take(wm_sd, rp + "pre_determ_recurrent.1.bias")
        p.gru_w, p.gru_b = take(wm_sd, rp + "determ_recurrent._layer.weight"), take(wm_sd, rp + "determ_recurrent._layer.bias")
        p.gru_ln_g, p.gru_ln_b = take(wm_sd, rp + "determ_recurrent._norm.weight"), take(wm_sd, rp + "determ_recurrent._norm.bias")
        p.prior1_w, p.prior1_b = take(wm_sd, rp + "ensemble_prior_estimator.0.weight"), take(wm_sd, rp + "ensemble_prior_estimator.0.bias")
        p.prior1_ln_g, p.prior1_ln_b = take(wm_sd, rp + "ensemble_prior_estimator.1.weight"), take(wm_sd, rp + "ensemble_prior_estimator.1.bias")
        p.prior2_w, p.prior2_b = take(wm_sd, rp + "ensemble_prior_estimator.3.weight"), take(wm_sd, rp + "ensemble_prior_estimator.3.bias")
        if self.cfg.slots > 1:   # slot mixer + positional encoding (rssm_slots_attention.py:141-145, world model pos_enc)
            p.mix_qkv_w = take(wm_sd, rp + "hidden_attention_proj.weight")
            p.mix_pre_norm_g, p.mix_pre_norm_b = take(wm_sd, rp + "pre_norm.weight"), take(wm_sd, rp + "pre_norm.bias")
            p.mix_fc_w, p.mix_fc_b = take(wm_sd, rp + "fc.weight"), take(wm_sd, rp + "fc.bias")
            p.mix_fc_norm_g, p.mix_fc_norm_b = take(wm_sd, rp + "fc_norm.weight"), take(wm_sd, rp + "fc_norm.bias")
            p.pos_enc = take(wm_sd, "pos_enc")
        p.actor = _mlp_params(actor_sd, "actor.", keep)
        p.reward = _mlp_params(wm_sd, "reward_predictor.", keep)
        if self.cfg.predict_discount:
            p.discount = _mlp_params(wm_sd, "discount_predictor.", keep)
        if self.cfg.with_critic:
            p.critic = _mlp_params(critic_sd, target_prefix, keep)
        # the blobs are (re-)built lazily, by the kernel that is about to read them (_packed_chained / _packed_rollout): at the
        # configured 800 start states only the persistent kernels run, and packing the chained rollout's image as well cost
        # ~ 60 us of every 3 ms step
        self._params = p
        self._pack_version += 1
        # `keep` tensors must outlive the enqueued pack kernels: stream-ordered frees make that safe
        self._keep = keep

    def rollout_cluster_for(self, n: int) -> int:
        """CTAs per cluster of the persistent rollout for n start states: the configured / RLSB_ROLLOUT_CLUSTER value, else
        16 while the device keeps all row blocks' clusters of 16 resident at once (rlsb_rollout_max_clusters: a cluster lives
        inside one GPC — 7 of 16 on the pool's B200s, not 148 / 16 = 9), else 8."""
        if self.cfg.rollout_cluster:
            return int(self.cfg.rollout_cluster)
        if os.environ.get("RLSB_ROLLOUT_CLUSTER"):
            return int(self.lib.rlsb_rollout_cluster_size())
        return 16 if (n + 127) // 128 <= self.rollout_max_clusters(16) else 8

    def rollout_max_clusters(self, c: int) -> int:
        v = self._max_clusters.get(c)
        if v is None:
            v = self._max_clusters[c] = int(self.lib.rlsb_rollout_max_clusters(int(c)))
        return v

    def would_run_persistent(self, n: int) -> bool:
        """the engine's own choice for a rollout of n start states without actor slots: the persistent kernel while all row
        blocks run in one wave — at config-1 dims (D = 1024) only with clusters of 16 (measured at 1024 / 1536 / 1920 start
        states: with clusters of 8 the chained rollout and its fused epilogues are 2-10 % faster per step; at D = 200 the
        persistent forward + backward win by 9-14 % up to 1920)"""
        if self.packed_ro is None or n > self.persistent_max_rows:
            return False
        if self.persistent_max_rows >= (1 << 30):
            return True
        return self.rollout_fits_one_wave(n) and (self.cfg.D <= 512 or self.rollout_cluster_for(n) == 16)

    def rollout_fits_one_wave(self, n: int) -> bool:
        """all row blocks of n start states run concurrently in the persistent kernels (a second wave doubles their time:
        the chained rollout is faster then)"""
        return (n + 127) // 128 <= self.rollout_max_clusters(self.rollout_cluster_for(n))

    def _packed_chained(self) -> torch.Tensor:
        """the chained kernels' weight blob (rlsb_imagine_fwd / rlsb_imagine_bwd), re-packed when `pack` ran since it was made
        or the fused-epilogue switch (which decides the GRU rows' order inside the blob) changed"""
        if self._params is None:
            raise _lib.RlsbError("ImaginationEngine.rollout before pack")
        fused = int(self.lib.rlsb_set_fused_rssm(-1))
        if self._chained_version != (self._pack_version, fused):
            check(self.lib.rlsb_imagine_pack(C.byref(self.ccfg), C.byref(self._params), self.packed.data_ptr(), _stream()),
                  "rlsb_imagine_pack")
            self._chained_version = (self._pack_version, fused)
        return self.packed

    def _packed_rollout(self, ccfg) -> torch.Tensor:
        """the persistent kernel's weight blob for ccfg.rollout_cluster, re-packed when `pack` ran since it was made"""
        if self._params is None:
            raise _lib.RlsbError("ImaginationEngine.rollout before pack")
        c = int(ccfg.rollout_cluster)
        blob = self.packed_ro.get(c)
        if blob is None:
            nbytes = self.lib.rlsb_rollout_packed_bytes(C.byref(ccfg))
            if nbytes == 0:
                raise _lib.RlsbError(f"persistent rollout: unsupported config {self.cfg} with a cluster of {c}")
            blob = self.packed_ro[c] = torch.zeros(nbytes, device=self.device, dtype=torch.uint8)
        if self._ro_version.get(c) != self._pack_version:
            pc = self.cfg.to_c()
            pc.rollout_cluster = c
            check(self.lib.rlsb_rollout_pack(C.byref(pc), C.byref(self._params), blob.data_ptr(), _stream()), "rlsb_rollout_pack")
            self._ro_version[c] = self._pack_version
        return blob

    def workspace(self, n: int, pin=None) -> torch.Tensor:
        nbytes = self.lib.rlsb_imagine_workspace_bytes(C.byref(self.ccfg), n)
        if pin is not None:
            key = ("ws", pin)
            buf = self._pinned.get(key)
            if buf is None:
                buf = self._pinned[key] = torch.zeros(nbytes, device=self.device, dtype=torch.uint8)
            elif buf.numel() < nbytes:
                raise _lib.RlsbError("a pinned (graph-owned) workspace cannot grow: capture a new graph for the new shape")
            return buf
        if self._ws is None or self._ws_bytes < nbytes:
            self._ws = torch.zeros(nbytes, device=self.device, dtype=torch.uint8)
            self._ws_bytes = nbytes
        return self._ws

    def rollout(self, h0: torch.Tensor, z0: torch.Tensor, logits0: Optional[torch.Tensor] = None,
                latent_uniforms: Optional[torch.Tensor] = None, action_noise: Optional[torch.Tensor] = None,
                seed: int = 0, row_offset: int = 0, precomp_actions: Optional[torch.Tensor] = None,
                horizon: Optional[int] = None, want_stoch: bool = True, want_actor_raw: bool = False,
                out: Optional[dict] = None, keep_packed: bool = False, tape: bool = False,
                seed_device: Optional[torch.Tensor] = None, actor_slots=None, pin=None,
                last_step_value_only: bool = False, persistent: Optional[bool] = None) -> dict:
        """``pin``: identity of the CUDA graph this call is captured into (see ``workspace``).  ``actor_slots`` (``ACUpdateEngine.actor_slots(n)``): the actor head's activations of steps 0..H-1 are written
        into the update's workspace, so that ``ACUpdateEngine.update(..., actor_forward_done=True)`` skips that forward."""
        cfg = self.cfg
        H = horizon if horizon is not None else cfg.H
        ccfg = cfg.to_c()
        ccfg.H = H
        ccfg.last_step_value_only = int(last_step_value_only or cfg.last_step_value_only)
        h0, z0 = _f32c(h0), _f32c(z0)
        S = cfg.groups * cfg.classes
        K = max(1, cfg.slots)
        # slotted: h0 (n, K, D) / (n*K, D), rows ordered (start state, slot)
        h0, z0 = h0.reshape(-1, cfg.D), z0.reshape(-1, S)
        if h0.shape[0] % K or z0.shape[0] != h0.shape[0]:
            raise _lib.RlsbError(f"rollout: h0 {tuple(h0.shape)} / z0 {tuple(z0.shape)} for slots={K}")
        n = h0.shape[0] // K
        dev = h0.device
        sh = (n,) if K == 1 else (n, K)
        if out is None:
            out = {
                "determ": torch.empty((H + 1,) + sh + (cfg.D,), device=dev, dtype=torch.float32),
                "logits": torch.empty((H + 1,) + sh + (S,), device=dev, dtype=torch.float32),
                "stoch_idx": torch.empty((H + 1,) + sh + (cfg.groups,), device=dev, dtype=torch.uint8),
                "stoch": torch.empty((H + 1,) + sh + (S,), device=dev, dtype=torch.float32) if want_stoch else None,
                "actions": torch.empty((H + 1, n, cfg.A), device=dev, dtype=torch.float32),
                "rewards": torch.empty((H + 1, n), device=dev, dtype=torch.float32),
                "discounts": torch.empty((H + 1, n), device=dev, dtype=torch.float32),
                "values": torch.empty((H + 1, n), device=dev, dtype=torch.float32) if cfg.with_critic else None,
                "actor_raw": torch.empty((H, n, cfg.A if cfg.discrete else 2 * cfg.A), device=dev,
                                         dtype=torch.float32) if want_actor_raw else None,
            }
        if keep_packed and out.get("determ_packed") is None:
            # packed bf16 state images, one slot per step, kept for the actor-critic update (K4)
            rows = round_up(n, 128)
            out["determ_packed"] = torch.empty((H + 1, rows, round_up(cfg.D, 64)), device=dev, dtype=torch.bfloat16)
            out["stoch_packed"] = torch.empty((H + 1, rows, round_up(S, 64)), device=dev, dtype=torch.bfloat16)
        if tape and out.get("tape") is None:
            nbytes = self.lib.rlsb_imagine_tape_bytes(C.byref(ccfg), n)
            if nbytes == 0:
                raise _lib.RlsbError("rollout(tape=True) needs ImagineConfig(with_backward=True) and D <= 512")
            out["tape"] = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        co = ImagineOut(*[_ptr(out.get(k)) for k in ("determ", "logits", "stoch_idx", "stoch", "actions",
                                                      "rewards", "discounts", "values", "actor_raw",
                                                      "determ_packed", "stoch_packed", "tape")])
        if actor_slots is not None:
            co.actor_slots = C.pointer(actor_slots)
        nz = Noise(_ptr(None if latent_uniforms is None else _f32c(latent_uniforms)),
                   _ptr(None if action_noise is None else _f32c(action_noise)), seed, row_offset,
                   _ptr(None if precomp_actions is None else _f32c(precomp_actions)), _ptr(seed_device))
        ws = self.workspace(n, pin)
        if persistent is None:
            persistent = actor_slots is None and self.would_run_persistent(n)
        elif persistent and (self.packed_ro is None or actor_slots is not None):
            raise _lib.RlsbError("rollout(persistent=True): flat RSSM without parity mode / actor_slots only")
        self.last_rollout_persistent = bool(persistent)
        if persistent:
            ccfg.rollout_cluster = self.rollout_cluster_for(n)
            blob = self._packed_rollout(ccfg)
            check(self.lib.rlsb_rollout_fwd(C.byref(ccfg), blob.data_ptr(), n, h0.data_ptr(), z0.data_ptr(),
                                            _ptr(None if logits0 is None else _f32c(logits0)), C.byref(nz),
                                            C.byref(co), ws.data_ptr(), _stream()), "rlsb_rollout_fwd")
            return out
        check(self.lib.rlsb_imagine_fwd(C.byref(ccfg), self._packed_chained().data_ptr(), n, h0.data_ptr(), z0.data_ptr(),
                                        _ptr(None if logits0 is None else _f32c(logits0)), C.byref(nz),
                                        C.byref(co), ws.data_ptr(), _stream()), "rlsb_imagine_fwd")
        return out

    def backward(self, out: dict, g_rewards: torch.Tensor, g_values: torch.Tensor, pin=None,
                 persistent: Optional[bool] = None) -> torch.Tensor:
        """d loss / d actions (H, N, A) from d loss / d rewards, d loss / d values (each (H+1, N)) through the
        rollout recorded in ``out`` (made with tape=True): rlsb_imagine_bwd (one launch per layer) or, up to
        ``persistent_max_rows`` start states, rlsb_rollout_bwd (ONE persistent kernel; ``persistent`` forces either)."""
        if out.get("tape") is None:
            raise _lib.RlsbError("ImaginationEngine.backward needs a rollout made with tape=True")
        H = out["determ"].shape[0] - 1
        ccfg = self.cfg.to_c()
        ccfg.H = H
        n = out["determ"].shape[1]
        g_rewards, g_values = _f32c(g_rewards), _f32c(g_values)
        if g_rewards.numel() != (H + 1) * n or g_values.numel() != (H + 1) * n:
            raise _lib.RlsbError("backward: g_rewards / g_values must be (H+1, N)")
        nbytes = self.lib.rlsb_imagine_bwd_workspace_bytes(C.byref(ccfg), n)
        if pin is not None:
            bws = self._pinned.get(("bws", pin))
            if bws is None or bws.numel() < nbytes:
                if bws is not None:
                    raise _lib.RlsbError("a pinned (graph-owned) backward workspace cannot grow")
                bws = self._pinned[("bws", pin)] = torch.empty(nbytes, device=self.device, dtype=torch.uint8)
        else:
            if getattr(self, "_bws", None) is None or self._bws.numel() < nbytes:
                self._bws = torch.empty(nbytes, device=self.device, dtype=torch.uint8)
            bws = self._bws
        g_actions = torch.empty((H, n, self.cfg.A), device=self.device, dtype=torch.float32)
        co = ImagineOut(*[_ptr(out.get(k)) for k in ("determ", "logits", "stoch_idx", "stoch", "actions",
                                                      "rewards", "discounts", "values", "actor_raw",
                                                      "determ_packed", "stoch_packed", "tape")])
        if persistent is None:
            persistent = self.would_run_persistent(n) and os.environ.get("RLSB_PERSISTENT_BWD", "1") != "0"
            if persistent:   # cluster sizes whose slices do not fit the epilogue's register plan fall back to the chain
                probe = self.cfg.to_c()
                probe.rollout_cluster = self.rollout_cluster_for(n)
                persistent = self.lib.rlsb_rollout_bwd_supported(C.byref(probe)) == 1
        elif persistent and self.packed_ro is None:
            raise _lib.RlsbError("backward(persistent=True): flat RSSM without parity mode only")
        self.last_backward_persistent = bool(persistent)
        if persistent:
            ccfg.rollout_cluster = self.rollout_cluster_for(n)
            blob = self._packed_rollout(ccfg)
            check(self.lib.rlsb_rollout_bwd(C.byref(ccfg), blob.data_ptr(), n, C.byref(co), g_rewards.data_ptr(),
                                            g_values.data_ptr(), g_actions.data_ptr(), bws.data_ptr(), _stream()),
                  "rlsb_rollout_bwd")
            return g_actions
        check(self.lib.rlsb_imagine_bwd(C.byref(ccfg), self._packed_chained().data_ptr(), n, C.byref(co), g_rewards.data_ptr(),
                                        g_values.data_ptr(), g_actions.data_ptr(), bws.data_ptr(), _stream()),
              "rlsb_imagine_bwd")
        return g_actions


# ------------------------------------------------------------------------------------------------
# K4
# ------------------------------------------------------------------------------------------------
def _mlp_grads(module_seq, keep: list) -> MlpGrads:
    """Gradient pointers of an fc_nn Sequential (Linear at 0,3,6,9,12; LayerNorm|Identity at 1,4,7,10);
    .grad tensors are allocated here when missing — the kernels overwrite every element."""
    mg = MlpGrads()

    def grad_of(p):
        if p.grad is None or p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
            p.grad = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
        keep.append(p.grad)
        return p.grad.data_ptr()

    for i, li in enumerate((0, 3, 6, 9, 12)):
        mg.w[i], mg.b[i] = grad_of(module_seq[li].weight), grad_of(module_seq[li].bias)
    for i, li in enumerate((1, 4, 7, 10)):
        m = module_seq[li]
        if isinstance(m, torch.nn.LayerNorm):
            mg.ln_g[i], mg.ln_b[i] = grad_of(m.weight), grad_of(m.bias)
    return mg


class ACUpdateEngine:
    """Packed weights + workspace of K4 (critic / actor losses and their backward pass, ac.py:68-81,113-146)."""

    def __init__(self, cfg: ImagineConfig, rho: float, eta: float, metrics_samples: int = 128, device="cuda"):
        _lib.require_device()
        self.lib = _lib.load()
        self.cfg = cfg
        self.ccfg = AcCfg(cfg.D, cfg.groups, cfg.classes, cfg.A, cfg.hidden, int(cfg.discrete), int(cfg.layer_norm),
                          cfg.H, float(rho), float(eta), int(metrics_samples))
        self.device = torch.device(device)
        nbytes = self.lib.rlsb_ac_packed_bytes(C.byref(self.ccfg))
        if nbytes == 0:
            raise _lib.RlsbError(f"unsupported actor-critic update config {cfg}")
        self.packed = torch.zeros(nbytes, device=self.device, dtype=torch.uint8)
        self._ws, self._ws_bytes = None, 0
        self._pinned: dict = {}   # graph-owned workspaces, see ImaginationEngine.workspace
        self.scalars = torch.zeros(_lib.AC_SCALARS, device=self.device, dtype=torch.float32)

    def pack(self, actor_sd: dict, critic_sd: dict, actor_prefix="actor.", critic_prefix="critic.") -> None:
        keep: list = []
        a = _mlp_params(actor_sd, actor_prefix, keep)
        c = _mlp_params(critic_sd, critic_prefix, keep)
        check(self.lib.rlsb_ac_pack(C.byref(self.ccfg), C.byref(a), C.byref(c), self.packed.data_ptr(), _stream()),
              "rlsb_ac_pack")
        self._keep = keep

    def _workspace(self, ccfg, n: int, pin=None) -> torch.Tensor:
        """The size depends on the rows AND the horizon (rlsb_ac_workspace_bytes): tracked in bytes."""
        nbytes = self.lib.rlsb_ac_workspace_bytes(C.byref(ccfg), n)
        if pin is not None:
            buf = self._pinned.get(pin)
            if buf is None:
                buf = self._pinned[pin] = torch.empty(nbytes, device=self.device, dtype=torch.uint8)
            elif buf.numel() < nbytes:
                raise _lib.RlsbError("a pinned (graph-owned) workspace cannot grow: capture a new graph for the new shape")
            return buf
        if self._ws is None or self._ws_bytes < nbytes:
            self._ws = torch.empty(nbytes, device=self.device, dtype=torch.uint8)
            self._ws_bytes = nbytes
            self._ws_layout = None
        return self._ws

    def actor_slots(self, n: int, horizon: Optional[int] = None, pin=None) -> ActorSlots:
        """Where ``ImaginationEngine.rollout(..., actor_slots=...)`` leaves the actor's activations for ``update``."""
        ccfg = AcCfg.from_buffer_copy(self.ccfg)
        ccfg.H = horizon if horizon is not None else self.cfg.H
        ws = self._workspace(ccfg, n, pin)   # the slices are laid out for (n rows, H); update(n, H) uses the same layout
        if pin is None:
            self._ws_layout = (n, ccfg.H, ws.data_ptr())
        slots = ActorSlots()
        check(self.lib.rlsb_ac_actor_slots(C.byref(ccfg), n, ws.data_ptr(), C.byref(slots)), "rlsb_ac_actor_slots")
        return slots

    def update(self, rollout: dict, vs: torch.Tensor, w: torch.Tensor, actor_seq, critic_seq, seed: int = 0,
               horizon: Optional[int] = None, g_actions: Optional[torch.Tensor] = None,
               seed_device: Optional[torch.Tensor] = None, actor_forward_done: bool = False, pin=None) -> torch.Tensor:
        """Writes .grad of every parameter of ``actor_seq`` / ``critic_seq`` (fc_nn Sequentials) and returns the
        RLSB_AC_SCALARS loss / metric vector (device tensor, see _lib.AC_SCALAR_NAMES)."""
        if rollout.get("determ_packed") is None:
            raise _lib.RlsbError("ACUpdateEngine.update needs a rollout made with keep_packed=True")
        n = rollout["determ"].shape[1]
        H = horizon if horizon is not None else self.cfg.H
        ccfg = AcCfg.from_buffer_copy(self.ccfg)
        ccfg.H = H
        ccfg.actor_fwd_in_rollout = int(actor_forward_done)
        if actor_forward_done and pin is None and getattr(self, "_ws_layout", None) is None:
            raise _lib.RlsbError("update(actor_forward_done=True): the rollout must have been given actor_slots(n) of this engine")
        if actor_forward_done and pin is not None and pin not in self._pinned:
            raise _lib.RlsbError("update(actor_forward_done=True, pin=...): actor_slots(n, pin=...) was not called for this graph")
        ws = self._workspace(ccfg, n, pin)
        if actor_forward_done and pin is None and self._ws_layout != (n, H, ws.data_ptr()):
            raise _lib.RlsbError("update(actor_forward_done=True): the workspace changed since actor_slots(n, H) was taken")
        keep: list = []
        ga, gc = _mlp_grads(actor_seq, keep), _mlp_grads(critic_seq, keep)
        vs, w = _f32c(vs), _f32c(w)
        values, actions = _f32c(rollout["values"]), _f32c(rollout["actions"])
        if vs.numel() != H * n or w.numel() != (H + 1) * n:
            raise _lib.RlsbError(f"ACUpdateEngine.update: vs {tuple(vs.shape)} / w {tuple(w.shape)} for H={H}, N={n}")
        check(self.lib.rlsb_ac_update(C.byref(ccfg), self.packed.data_ptr(), n, rollout["determ_packed"].data_ptr(),
                                      rollout["stoch_packed"].data_ptr(), vs.data_ptr(), w.data_ptr(),
                                      values.data_ptr(), actions.data_ptr(),
                                      _ptr(None if g_actions is None else _f32c(g_actions)), seed, _ptr(seed_device),
                                      C.byref(ga), C.byref(gc),
                                      self.scalars.data_ptr(), ws.data_ptr(), _stream()), "rlsb_ac_update")
        return self.scalars


    def losses_from_heads(self, actor_raw: torch.Tensor, critic_values: torch.Tensor, vs: torch.Tensor, w: torch.Tensor,
                          values: torch.Tensor, actions: torch.Tensor, seed: int = 0,
                          horizon: Optional[int] = None) -> torch.Tensor:
        """rlsb_ac_losses: the loss kernel of the update on caller-supplied head outputs — ``actor_raw`` (H, N, A or 2A)
        the actor's raw outputs and ``critic_values`` (H, N) the critic's values on states 0..H-1 (e.g. from a rollout in
        the split-operand mode, ``ImagineConfig(parity=True)``); vs (H, N), w / values (H+1, N), actions (H+1, N, A).
        Returns a copy of the RLSB_AC_SCALARS vector."""
        H = horizon if horizon is not None else self.cfg.H
        ccfg = AcCfg.from_buffer_copy(self.ccfg)
        ccfg.H = H
        n = critic_values.shape[1]
        rows = round_up(n, 128)
        actor_raw, critic_values = _f32c(actor_raw), _f32c(critic_values)
        if actor_raw.shape[:2] != (H, n) or critic_values.shape[0] != H:
            raise _lib.RlsbError(f"losses_from_heads: actor_raw {tuple(actor_raw.shape)} / critic_values "
                                 f"{tuple(critic_values.shape)} for H={H}")
        head = torch.zeros((2, H, rows, 32), device=self.device, dtype=torch.float32)
        head[0, :, :n, :actor_raw.shape[-1]] = actor_raw
        head[1, :, :n, 0] = critic_values.reshape(H, n)
        vs, w, values, actions = _f32c(vs), _f32c(w), _f32c(values), _f32c(actions)
        if vs.numel() != H * n or w.numel() != (H + 1) * n or values.numel() != (H + 1) * n:
            raise _lib.RlsbError("losses_from_heads: vs must be (H, N), w / values (H+1, N)")
        ws = torch.empty(self.lib.rlsb_ac_workspace_bytes(C.byref(ccfg), n), device=self.device, dtype=torch.uint8)
        scal = torch.zeros(_lib.AC_SCALARS, device=self.device, dtype=torch.float32)
        check(self.lib.rlsb_ac_losses(C.byref(ccfg), n, head.data_ptr(), vs.data_ptr(), w.data_ptr(), values.data_ptr(),
                                      actions.data_ptr(), seed, scal.data_ptr(), ws.data_ptr(), _stream()),
              "rlsb_ac_losses")
        return scal


# ------------------------------------------------------------------------------------------------
# K3
# ------------------------------------------------------------------------------------------------
class SlotAttentionEngine:
    """Packed weights + workspace of K3 (SlotAttention.forward, vision/slot_attention.py:52-77)."""

    KEYS = {"inputs_norm_g": "inputs_norm.weight", "inputs_norm_b": "inputs_norm.bias",
            "inputs_proj_w": "inputs_proj.weight", "slots_norm_g": "slots_norm.weight",
            "slots_norm_b": "slots_norm.bias", "slots_proj_w": "slots_proj.weight",
            "gru_w_ih": "slots_reccur.weight_ih", "gru_w_hh": "slots_reccur.weight_hh",
            "gru_b_ih": "slots_reccur.bias_ih", "gru_b_hh": "slots_reccur.bias_hh",
            "slots_norm2_g": "slots_norm_2.weight", "slots_norm2_b": "slots_norm_2.bias",
            "mlp_w1": "slots_proj_2.0.weight", "mlp_b1": "slots_proj_2.0.bias",
            "mlp_w2": "slots_proj_2.2.weight", "mlp_b2": "slots_proj_2.2.bias"}

    def __init__(self, slots: int, dim: int, tokens: int, iters: int, device="cuda"):
        _lib.require_device()
        self.lib = _lib.load()
        self.cfg = SlotCfg(slots, dim, tokens, iters)
        self.slots, self.dim, self.tokens, self.iters = slots, dim, tokens, iters
        self.device = torch.device(device)
        nbytes = self.lib.rlsb_slot_attention_packed_bytes(C.byref(self.cfg))
        if nbytes == 0:
            raise _lib.RlsbError(f"unsupported slot-attention config slots={slots} dim={dim} tokens={tokens}")
        self.packed = torch.zeros(nbytes, device=self.device, dtype=torch.uint8)
        self._ws, self._ws_b = None, 0

    def pack(self, sd: dict, prefix: str = "") -> None:
        p, keep = SlotParams(), []
        for field, key in self.KEYS.items():
            t = _f32c(sd[prefix + key])
            keep.append(t)
            setattr(p, field, t.data_ptr())
        check(self.lib.rlsb_slot_attention_pack(C.byref(self.cfg), C.byref(p), self.packed.data_ptr(), _stream()),
              "rlsb_slot_attention_pack")
        self._keep = keep

    def forward(self, X: torch.Tensor, prev_slots: torch.Tensor, want_attn: bool = True):
        X, prev_slots = _f32c(X), _f32c(prev_slots)
        B = X.shape[0]
        if X.shape[1:] != (self.tokens, self.dim) or prev_slots.shape != (B, self.slots, self.dim):
            raise _lib.RlsbError(f"slot attention: X {tuple(X.shape)} prev_slots {tuple(prev_slots.shape)}")
        if self._ws is None or self._ws_b < B:
            self._ws = torch.zeros(self.lib.rlsb_slot_attention_workspace_bytes(C.byref(self.cfg), B),
                                   device=self.device, dtype=torch.uint8)
            self._ws_b = B
        out = torch.empty_like(prev_slots)
        attn = torch.empty((B, self.slots, self.tokens), device=X.device, dtype=torch.float32) if want_attn else None
        check(self.lib.rlsb_slot_attention_fwd(C.byref(self.cfg), self.packed.data_ptr(), B, X.data_ptr(),
                                               prev_slots.data_ptr(), out.data_ptr(), _ptr(attn), self._ws.data_ptr(),
                                               _stream()), "rlsb_slot_attention_fwd")
        return out, attn


    # ---- training: forward with an activation tape + backward (rlsb_slot_attention_fwd_tape / _bwd) ----
    def forward_tape(self, X: torch.Tensor, prev_slots: torch.Tensor):
        X, prev_slots = _f32c(X), _f32c(prev_slots)
        B = X.shape[0]
        if X.shape[1:] != (self.tokens, self.dim) or prev_slots.shape != (B, self.slots, self.dim):
            raise _lib.RlsbError(f"slot attention: X {tuple(X.shape)} prev_slots {tuple(prev_slots.shape)}")
        if self._ws is None or self._ws_b < B:
            self._ws = torch.zeros(self.lib.rlsb_slot_attention_workspace_bytes(C.byref(self.cfg), B),
                                   device=self.device, dtype=torch.uint8)
            self._ws_b = B
        # zero-initialised: padding rows of the operand images on the tape enter weight-gradient contractions
        tape = torch.zeros(self.lib.rlsb_slot_attention_tape_bytes(C.byref(self.cfg), B), device=self.device,
                           dtype=torch.uint8)
        out = torch.empty_like(prev_slots)
        attn = torch.empty((B, self.slots, self.tokens), device=X.device, dtype=torch.float32)
        check(self.lib.rlsb_slot_attention_fwd_tape(C.byref(self.cfg), self.packed.data_ptr(), B, X.data_ptr(),
                                                    prev_slots.data_ptr(), out.data_ptr(), attn.data_ptr(),
                                                    tape.data_ptr(), self._ws.data_ptr(), _stream()),
              "rlsb_slot_attention_fwd_tape")
        return out, attn, tape

    def backward(self, X: torch.Tensor, tape: torch.Tensor, d_out: torch.Tensor):
        """-> (dX, d_prev_slots, {state-dict key: gradient})"""
        X, d_out = _f32c(X), _f32c(d_out)
        B = X.shape[0]
        nbytes = self.lib.rlsb_slot_attention_bwd_workspace_bytes(C.byref(self.cfg), B)
        if getattr(self, "_bws", None) is None or self._bws.numel() < nbytes:
            self._bws = torch.empty(nbytes, device=self.device, dtype=torch.uint8)
        d, K = self.dim, self.slots
        shapes = {"inputs_norm_g": (d,), "inputs_norm_b": (d,), "inputs_proj_w": (2 * d, d), "slots_norm_g": (d,),
                  "slots_norm_b": (d,), "slots_proj_w": (d, d), "gru_w_ih": (3 * d, d), "gru_w_hh": (3 * d, d),
                  "gru_b_ih": (3 * d,), "gru_b_hh": (3 * d,), "slots_norm2_g": (d,), "slots_norm2_b": (d,),
                  "mlp_w1": (4 * d, d), "mlp_b1": (4 * d,), "mlp_w2": (d, 4 * d), "mlp_b2": (d,)}
        g, grads = SlotGrads(), {}
        for field, shp in shapes.items():
            t = torch.zeros(shp, device=self.device, dtype=torch.float32)
            grads[self.KEYS[field]] = t
            setattr(g, field, t.data_ptr())
        dX = torch.empty_like(X)
        dprev = torch.empty((B, K, d), device=self.device, dtype=torch.float32)
        check(self.lib.rlsb_slot_attention_bwd(C.byref(self.cfg), self.packed.data_ptr(), B, X.data_ptr(), tape.data_ptr(),
                                               d_out.data_ptr(), C.byref(g), dX.data_ptr(), dprev.data_ptr(),
                                               self._bws.data_ptr(), _stream()), "rlsb_slot_attention_bwd")
        return dX, dprev, grads


class SlotAttentionFn(torch.autograd.Function):
    """SlotAttention.forward (vision/slot_attention.py:52-77) with both directions in librlsb (K3)."""

    @staticmethod
    def forward(ctx, engine, names, X, prev_slots, *params):
        out, attn, tape = engine.forward_tape(X.detach(), prev_slots.detach())
        ctx.engine, ctx.names = engine, names
        ctx.save_for_backward(X.detach(), tape)
        ctx.mark_non_differentiable(attn)
        return out, attn

    @staticmethod
    def backward(ctx, d_out, _d_attn):
        X, tape = ctx.saved_tensors
        dX, dprev, grads = ctx.engine.backward(X, tape, d_out.contiguous())
        return (None, None, dX, dprev) + tuple(grads[n] for n in ctx.names)


# ------------------------------------------------------------------------------------------------
# K5: world-model observe scan
# ------------------------------------------------------------------------------------------------
class ObserveEngine:
    """Packed RSSM weights of K5 (the T-step observe loop of WorldModel.calculate_loss, world_model.py:187-202)."""

    def __init__(self, D: int, A: int, E: int, layer_norm: bool, T: int, groups: int = 32, classes: int = 32, device="cuda"):
        _lib.require_device()
        self.lib = _lib.load()
        self.cfg = _lib.ObserveCfg(D, groups, classes, A, E, int(layer_norm), T)
        self.D, self.A, self.E, self.T, self.S, self.groups, self.layer_norm = D, A, E, T, groups * classes, groups, layer_norm
        self.device = torch.device(device)
        nbytes = self.lib.rlsb_observe_packed_bytes(C.byref(self.cfg))
        if nbytes == 0:
            raise _lib.RlsbError(f"unsupported observe config D={D} A={A} E={E} T={T}")
        self.packed = torch.zeros(nbytes, device=self.device, dtype=torch.uint8)
        self._bws = None

    def names(self) -> list:
        """state-dict keys (below `recurrent_model.`) in the order of rlsb_observe_params, LayerNorm entries only if present"""
        return [k for f, k in _lib.OBS_KEYS.items() if self.layer_norm or not f.endswith(("1_ln_g", "1_ln_b", "in_ln_g", "in_ln_b"))]

    def pack(self, rssm_sd: dict) -> None:
        p, keep = _lib.ObserveParams(), []
        for field, key in _lib.OBS_KEYS.items():
            if key in rssm_sd and key in self.names():
                t = _f32c(rssm_sd[key])
                keep.append(t)
                setattr(p, field, t.data_ptr())
        check(self.lib.rlsb_observe_pack(C.byref(self.cfg), C.byref(p), self.packed.data_ptr(), _stream()), "rlsb_observe_pack")
        self._keep = keep

    def forward(self, embed: torch.Tensor, actions: torch.Tensor, latent_uniforms: Optional[torch.Tensor] = None,
                seed: int = 0, row_offset: int = 0, seed_device: Optional[torch.Tensor] = None):
        """embed (T, B, E), actions (T, B, A) -> dict(prior_logits, post_logits, determ, stoch_idx, stoch, tape).
        seed_device: int64 device tensor holding the Philox key (read by the kernels; lets a CUDA graph change it)."""
        embed, actions = _f32c(embed), _f32c(actions)
        T, B = embed.shape[0], embed.shape[1]
        if T != self.T or embed.shape[2] != self.E or actions.shape != (T, B, self.A):
            raise _lib.RlsbError(f"observe: embed {tuple(embed.shape)} actions {tuple(actions.shape)} for T={self.T}")
        dev = embed.device
        out = {"prior_logits": torch.empty((T, B, self.S), device=dev), "post_logits": torch.empty((T, B, self.S), device=dev),
               "determ": torch.empty((T, B, self.D), device=dev),
               "stoch_idx": torch.empty((T, B, self.groups), device=dev, dtype=torch.uint8),
               "stoch": torch.empty((T, B, self.S), device=dev)}
        # zero-initialised: slot 0 of the state images is the zero initial state, padding rows enter weight gradients
        out["tape"] = torch.zeros(self.lib.rlsb_observe_tape_bytes(C.byref(self.cfg), B), device=dev, dtype=torch.uint8)
        co = _lib.ObserveOut(*[out[k].data_ptr() for k in ("prior_logits", "post_logits", "determ", "stoch_idx", "stoch")])
        nz = Noise(_ptr(None if latent_uniforms is None else _f32c(latent_uniforms)), None, seed, row_offset, None,
                   None if seed_device is None else seed_device.data_ptr())
        check(self.lib.rlsb_observe_fwd(C.byref(self.cfg), self.packed.data_ptr(), B, embed.data_ptr(), actions.data_ptr(),
                                        C.byref(nz), C.byref(co), out["tape"].data_ptr(), _stream()), "rlsb_observe_fwd")
        return out

    def backward(self, out: dict, g_prior, g_post, g_determ, g_stoch):
        """-> (g_embed (T, B, E), {state-dict key: gradient})"""
        T, B = out["determ"].shape[0], out["determ"].shape[1]
        nbytes = self.lib.rlsb_observe_bwd_workspace_bytes(C.byref(self.cfg), B)
        if self._bws is None or self._bws.numel() < nbytes:
            self._bws = torch.empty(nbytes, device=self.device, dtype=torch.uint8)
        D, S, A, E = self.D, self.S, self.A, self.E
        shapes = {"img_in_w": (D, S + A), "img_in_b": (D,), "img_in_ln_g": (D,), "img_in_ln_b": (D,),
                  "gru_w": (3 * D, 2 * D), "gru_b": (3 * D,), "gru_ln_g": (3 * D,), "gru_ln_b": (3 * D,),
                  "prior1_w": (D, D), "prior1_b": (D,), "prior1_ln_g": (D,), "prior1_ln_b": (D,),
                  "prior2_w": (S, D), "prior2_b": (S,), "post1_w": (D, D + E), "post1_b": (D,),
                  "post1_ln_g": (D,), "post1_ln_b": (D,), "post2_w": (S, D), "post2_b": (S,)}
        g, grads = _lib.ObserveGrads(), {}
        names = set(self.names())
        for field, key in _lib.OBS_KEYS.items():
            if key in names:
                t = torch.zeros(shapes[field], device=self.device, dtype=torch.float32)
                grads[key] = t
                setattr(g, field, t.data_ptr())
        g_embed = torch.empty((T, B, E), device=self.device, dtype=torch.float32)
        co = _lib.ObserveOut(*[out[k].data_ptr() for k in ("prior_logits", "post_logits", "determ", "stoch_idx", "stoch")])
        f = lambda t: None if t is None else _f32c(t).data_ptr()
        check(self.lib.rlsb_observe_bwd(C.byref(self.cfg), self.packed.data_ptr(), B, out["tape"].data_ptr(), C.byref(co),
                                        f(g_prior), f(g_post), f(g_determ), f(g_stoch), C.byref(g), g_embed.data_ptr(),
                                        self._bws.data_ptr(), _stream()), "rlsb_observe_bwd")
        return g_embed, grads


class ObserveScanFn(torch.autograd.Function):
    """The observe loop under torch autograd: forward = rlsb_observe_fwd, backward = rlsb_observe_bwd."""

    @staticmethod
    def forward(ctx, engine, names, noise, embed, actions, *params):
        out = engine.forward(embed.detach(), actions.detach(), **noise)
        ctx.engine, ctx.names, ctx.out = engine, names, out
        ctx.mark_non_differentiable(out["stoch_idx"])
        return out["prior_logits"], out["post_logits"], out["determ"], out["stoch"], out["stoch_idx"]

    @staticmethod
    def backward(ctx, g_prior, g_post, g_determ, g_stoch, _g_idx):
        c = lambda t: None if t is None else t.contiguous()
        g_embed, grads = ctx.engine.backward(ctx.out, c(g_prior), c(g_post), c(g_determ), c(g_stoch))
        ctx.out = None
        return (None, None, None, g_embed, None) + tuple(grads[n] for n in ctx.names)
