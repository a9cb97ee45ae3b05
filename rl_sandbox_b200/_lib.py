"""ctypes binding of librlsb.so (the C ABI declared in include/rlsb.h).

There is deliberately NO fallback: if the shared library is missing, or a kernel entry point is
called without a B200 (sm_100) device, an exception is raised.  PyTorch is used only for device
memory and streams; every pointer crossing this boundary is a raw device address.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "librlsb.so"

c_float_p = C.POINTER(C.c_float)


class RlsbError(RuntimeError):
    pass


class ImagineCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "D", "groups", "classes", "A", "hidden", "discrete", "layer_norm", "predict_discount",
        "with_critic", "H", "discount_nan_on_tie", "with_backward", "slots", "attention_blocks",
        "symmetric_qk")] + [("mixer_coeff", C.c_float), ("parity", C.c_int32), ("last_step_value_only", C.c_int32),
                                  ("rollout_cluster", C.c_int32)]


class MlpParams(C.Structure):
    _fields_ = [("w", C.c_void_p * 5), ("b", C.c_void_p * 5), ("ln_g", C.c_void_p * 4),
                ("ln_b", C.c_void_p * 4)]


class ImagineParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "img_in_w", "img_in_b", "img_in_ln_g", "img_in_ln_b", "gru_w", "gru_b", "gru_ln_g", "gru_ln_b",
        "prior1_w", "prior1_b", "prior1_ln_g", "prior1_ln_b", "prior2_w", "prior2_b")] + [
        ("actor", MlpParams), ("reward", MlpParams), ("discount", MlpParams), ("critic", MlpParams)] + [
        (n, C.c_void_p) for n in ("mix_qkv_w", "mix_pre_norm_g", "mix_pre_norm_b", "mix_fc_w", "mix_fc_b",
                                  "mix_fc_norm_g", "mix_fc_norm_b", "pos_enc")]


class Noise(C.Structure):
    _fields_ = [("latent_uniforms", C.c_void_p), ("action_noise", C.c_void_p), ("seed", C.c_uint64),
                ("row_offset", C.c_uint32), ("precomp_actions", C.c_void_p), ("seed_device", C.c_void_p)]


class ActorSlots(C.Structure):
    """rlsb_actor_slots: the actor's slices of the rlsb_ac_update workspace that rlsb_imagine_fwd fills."""
    _fields_ = [("x", C.c_void_p * 4), ("pre", C.c_void_p * 4), ("rstd", C.c_void_p * 4), ("m_pad", C.c_int64),
                ("Hp", C.c_int32), ("steps", C.c_int32)]


class ImagineOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "determ", "logits", "stoch_idx", "stoch", "actions", "rewards", "discounts", "values",
        "actor_raw", "determ_packed", "stoch_packed", "tape")] + [("actor_slots", C.POINTER(ActorSlots))]


class AcCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("D", "groups", "classes", "A", "hidden", "discrete", "layer_norm", "H")] + [
        ("rho", C.c_float), ("eta", C.c_float), ("metrics_samples", C.c_int32), ("actor_fwd_in_rollout", C.c_int32)]


class MlpGrads(C.Structure):
    _fields_ = [("w", C.c_void_p * 5), ("b", C.c_void_p * 5), ("ln_g", C.c_void_p * 4),
                ("ln_b", C.c_void_p * 4)]


AC_SCALAR_NAMES = {
    "loss_critic": 0, "loss_actor_reinforce": 1, "loss_actor_dynamics_backprop": 2, "loss_actor_entropy": 3,
    "loss_actor": 4, "critic/avg_target_value": 5, "critic/avg_lambda_value": 6, "critic/avg_predicted_value": 7,
    "actor/avg_val": 8, "actor/mean_val": 9, "actor/avg_sd": 10, "actor/min_val": 11, "actor/max_val": 12}
AC_SCALARS = 16
ABI_VERSION = 6


class SlotCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("slots", "dim", "tokens", "iters")]


class SlotParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "inputs_norm_g", "inputs_norm_b", "inputs_proj_w", "slots_norm_g", "slots_norm_b", "slots_proj_w",
        "gru_w_ih", "gru_w_hh", "gru_b_ih", "gru_b_hh", "slots_norm2_g", "slots_norm2_b",
        "mlp_w1", "mlp_b1", "mlp_w2", "mlp_b2")]


class ObserveCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("D", "groups", "classes", "A", "E", "layer_norm", "T")]


_OBS_FIELDS = ("img_in_w", "img_in_b", "img_in_ln_g", "img_in_ln_b", "gru_w", "gru_b", "gru_ln_g", "gru_ln_b",
               "prior1_w", "prior1_b", "prior1_ln_g", "prior1_ln_b", "prior2_w", "prior2_b",
               "post1_w", "post1_b", "post1_ln_g", "post1_ln_b", "post2_w", "post2_b")


class ObserveParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _OBS_FIELDS]


class ObserveGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _OBS_FIELDS]


class ObserveOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("prior_logits", "post_logits", "determ", "stoch_idx", "stoch")]


# rlsb_observe_params field -> state-dict key below `recurrent_model.` (reference: agents/dreamer/rssm.py:136-165)
OBS_KEYS = {"img_in_w": "pre_determ_recurrent.0.weight", "img_in_b": "pre_determ_recurrent.0.bias",
            "img_in_ln_g": "pre_determ_recurrent.1.weight", "img_in_ln_b": "pre_determ_recurrent.1.bias",
            "gru_w": "determ_recurrent._layer.weight", "gru_b": "determ_recurrent._layer.bias",
            "gru_ln_g": "determ_recurrent._norm.weight", "gru_ln_b": "determ_recurrent._norm.bias",
            "prior1_w": "ensemble_prior_estimator.0.weight", "prior1_b": "ensemble_prior_estimator.0.bias",
            "prior1_ln_g": "ensemble_prior_estimator.1.weight", "prior1_ln_b": "ensemble_prior_estimator.1.bias",
            "prior2_w": "ensemble_prior_estimator.3.weight", "prior2_b": "ensemble_prior_estimator.3.bias",
            "post1_w": "stoch_net.0.weight", "post1_b": "stoch_net.0.bias",
            "post1_ln_g": "stoch_net.1.weight", "post1_ln_b": "stoch_net.1.bias",
            "post2_w": "stoch_net.3.weight", "post2_b": "stoch_net.3.bias"}


class SlotGrads(C.Structure):
    _fields_ = SlotParams._fields_


_lib = None


def load() -> C.CDLL:
    """Load librlsb.so; raise loudly when it has not been built (``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RlsbError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or `make -C {_HERE / 'csrc'}`).  rl_sandbox_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(os.fspath(LIB_PATH))
    vp, i32, i64, f32, u32, u64, sz = (C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint32,
                                       C.c_uint64, C.c_size_t)
    sig = {
        "rlsb_abi_version": (C.c_int, []),
        "rlsb_check_device": (C.c_int, []),
        "rlsb_error_string": (C.c_char_p, [i32]),
        "rlsb_launch_count": (C.c_longlong, [i32]),
        "rlsb_launch_count_add": (C.c_longlong, [C.c_longlong]),
        "rlsb_set_cluster_size": (C.c_int, [i32]),
        "rlsb_set_staged_output": (C.c_int, [i32]),
        "rlsb_set_fused_rssm": (C.c_int, [i32]),
        "rlsb_gemm_set_trace": (None, [vp]),
        "rlsb_lambda_return_fwd": (C.c_int, [vp, vp, vp, i32, i64, C.c_double, vp, vp, vp, i32, vp]),
        "rlsb_lambda_return_bwd": (C.c_int, [vp, vp, vp, vp, i32, i64, C.c_double, vp, vp, vp, vp]),
        "rlsb_sample_categorical": (C.c_int, [vp, vp, i64, i32, vp, vp]),
        "rlsb_philox_uniform": (C.c_int, [u64, u32, u32, u32, i32, i64, vp, vp]),
        "rlsb_sample_latent": (C.c_int, [vp, i64, i32, vp, u64, u32, u32, vp, vp, vp]),
        "rlsb_pack_rows": (C.c_int, [vp, i64, i32, vp, i32, i32, i32, i32, i32, i32, vp]),
        "rlsb_gru_cell_packed_bytes": (sz, [i32, i32]),
        "rlsb_gru_cell_workspace_bytes": (sz, [i32, i32]),
        "rlsb_gru_cell_pack": (C.c_int, [vp, vp, vp, vp, i32, i32, vp, vp]),
        "rlsb_gru_cell_fwd": (C.c_int, [vp, i32, i32, vp, vp, vp, i32, f32, f32, vp, vp, vp, vp]),
        "rlsb_gemm_bias": (C.c_int, [vp, i32, vp, i32, i32, vp, i32, i32, vp, i64, vp, vp]),
        "rlsb_gemm_ln_act": (C.c_int, [vp, i32, vp, i32, vp, i32, i32, vp, vp, f32, i32, vp, i32, vp]),
        "rlsb_imagine_packed_bytes": (sz, [C.POINTER(ImagineCfg)]),
        "rlsb_imagine_workspace_bytes": (sz, [C.POINTER(ImagineCfg), i64]),
        "rlsb_imagine_pack": (C.c_int, [C.POINTER(ImagineCfg), C.POINTER(ImagineParams), vp, vp]),
        "rlsb_imagine_fwd": (C.c_int, [C.POINTER(ImagineCfg), vp, i64, vp, vp, vp, C.POINTER(Noise),
                                       C.POINTER(ImagineOut), vp, vp]),
        "rlsb_rollout_cluster_size": (C.c_int, []),
        "rlsb_rollout_max_clusters": (C.c_int, [i32]),
        "rlsb_rollout_set_trace": (None, [vp]),
        "rlsb_rollout_packed_bytes": (sz, [C.POINTER(ImagineCfg)]),
        "rlsb_rollout_pack": (C.c_int, [C.POINTER(ImagineCfg), C.POINTER(ImagineParams), vp, vp]),
        "rlsb_rollout_fwd": (C.c_int, [C.POINTER(ImagineCfg), vp, i64, vp, vp, vp, C.POINTER(Noise),
                                       C.POINTER(ImagineOut), vp, vp]),
        "rlsb_rollout_bwd_supported": (C.c_int, [C.POINTER(ImagineCfg)]),
        "rlsb_rollout_bwd": (C.c_int, [C.POINTER(ImagineCfg), vp, i64, C.POINTER(ImagineOut), vp, vp, vp, vp, vp]),
    }
    sig.update({
        "rlsb_slot_attention_packed_bytes": (sz, [C.POINTER(SlotCfg)]),
        "rlsb_slot_attention_workspace_bytes": (sz, [C.POINTER(SlotCfg), i64]),
        "rlsb_slot_attention_pack": (C.c_int, [C.POINTER(SlotCfg), C.POINTER(SlotParams), vp, vp]),
        "rlsb_slot_attention_fwd": (C.c_int, [C.POINTER(SlotCfg), vp, i64, vp, vp, vp, vp, vp, vp]),
        "rlsb_slot_attention_tape_bytes": (sz, [C.POINTER(SlotCfg), i64]),
        "rlsb_slot_attention_bwd_workspace_bytes": (sz, [C.POINTER(SlotCfg), i64]),
        "rlsb_slot_attention_fwd_tape": (C.c_int, [C.POINTER(SlotCfg), vp, i64, vp, vp, vp, vp, vp, vp, vp]),
        "rlsb_slot_attention_bwd": (C.c_int, [C.POINTER(SlotCfg), vp, i64, vp, vp, vp, C.POINTER(SlotGrads), vp, vp, vp, vp]),
    })
    sig.update({
        "rlsb_packed_rows": (sz, [i64]),
        "rlsb_gemm_wgrad_workspace_bytes": (sz, [i32, i32, i32]),
        "rlsb_gemm_wgrad": (C.c_int, [vp, i32, vp, i32, i32, vp, vp, vp]),
        "rlsb_ac_packed_bytes": (sz, [C.POINTER(AcCfg)]),
        "rlsb_ac_workspace_bytes": (sz, [C.POINTER(AcCfg), i64]),
        "rlsb_ac_pack": (C.c_int, [C.POINTER(AcCfg), C.POINTER(MlpParams), C.POINTER(MlpParams), vp, vp]),
        "rlsb_ac_actor_slots": (C.c_int, [C.POINTER(AcCfg), i64, vp, C.POINTER(ActorSlots)]),
        "rlsb_ac_update": (C.c_int, [C.POINTER(AcCfg), vp, i64, vp, vp, vp, vp, vp, vp, vp, u64, vp, C.POINTER(MlpGrads),
                                     C.POINTER(MlpGrads), vp, vp, vp]),
        "rlsb_ac_losses": (C.c_int, [C.POINTER(AcCfg), i64, vp, vp, vp, vp, vp, u64, vp, vp, vp]),
        "rlsb_imagine_tape_bytes": (sz, [C.POINTER(ImagineCfg), i64]),
        "rlsb_imagine_bwd_workspace_bytes": (sz, [C.POINTER(ImagineCfg), i64]),
        "rlsb_imagine_bwd": (C.c_int, [C.POINTER(ImagineCfg), vp, i64, C.POINTER(ImagineOut), vp, vp, vp, vp, vp]),
    })
    sig.update({
        "rlsb_observe_packed_bytes": (sz, [C.POINTER(ObserveCfg)]),
        "rlsb_observe_tape_bytes": (sz, [C.POINTER(ObserveCfg), i64]),
        "rlsb_observe_bwd_workspace_bytes": (sz, [C.POINTER(ObserveCfg), i64]),
        "rlsb_observe_pack": (C.c_int, [C.POINTER(ObserveCfg), C.POINTER(ObserveParams), vp, vp]),
        "rlsb_observe_fwd": (C.c_int, [C.POINTER(ObserveCfg), vp, i64, vp, vp, C.POINTER(Noise), C.POINTER(ObserveOut), vp, vp]),
        "rlsb_observe_bwd": (C.c_int, [C.POINTER(ObserveCfg), vp, i64, vp, C.POINTER(ObserveOut), vp, vp, vp, vp,
                                       C.POINTER(ObserveGrads), vp, vp, vp]),
    })
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.rlsb_abi_version() != ABI_VERSION:
        raise RlsbError("librlsb.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def exported_symbols() -> list[str]:
    """Names declared in include/rlsb.h (used by the CPU test that checks the export table)."""
    import re
    hdr = (_HERE.parent / "include" / "rlsb.h").read_text()
    return sorted(set(re.findall(r"\b(rlsb_[a-z0-9_]+)\s*\(", hdr)))


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = load().rlsb_error_string(code).decode()
        raise RlsbError(f"librlsb {what} failed with code {code}: {msg}")


def require_device() -> None:
    import torch
    if not torch.cuda.is_available():
        raise RlsbError("rl_sandbox_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")
    check(load().rlsb_check_device(), "rlsb_check_device")
