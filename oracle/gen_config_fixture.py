"""gen_config_fixture.py — TEST INFRASTRUCTURE (build container only: reads /root/reference).

Resolves the reference's Hydra agent configs (rl_sandbox/config/agent/*.yaml with their `defaults:` chains
and `${..x}` interpolations) into plain nested dicts and writes them, together with the three top-level
configs BASELINE.json names (config.yaml's overrides, config_dino.yaml, config_slotted.yaml: which agent /
env / training group each selects), to tests/golden/agent_configs.json.  The tests instantiate the
`_target_` dotted paths from this fixture through the alias package `rl_sandbox.*` — the reference's real
drop-in seam — and, where /root/reference exists, re-resolve the YAMLs and compare.

Run:  python -m oracle.gen_config_fixture
"""
import json
from pathlib import Path

import yaml

from .ref_harness import REFERENCE_ROOT

CONFIG_DIR = REFERENCE_ROOT / "rl_sandbox" / "config"
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "agent_configs.json"

# top-level config -> the groups it selects (config*.yaml `defaults:` lists) and the env facts DreamerV2's
# constructor receives from train.py:54-63 (obs_space_num, actions_num, action_type come from the env)
TOP = {
    "config": "config.yaml",
    "config_default": "config_default.yaml",
    "config_dino": "config_dino.yaml",
    "config_slotted": "config_slotted.yaml",
}
ENV_FACTS = {   # (actions_num, action_type) of the env each top-level config names
    "dm_cartpole": (1, "continuous"),
    "dm_quadruped": (12, "continuous"),
    "crafter": (17, "discrete"),
}


def _merge(base, over):
    out = dict(base)
    for k, v in over.items():
        out[k] = _merge(out[k], v) if isinstance(v, dict) and isinstance(out.get(k), dict) else v
    return out


def load_agent(name):
    """agent/<name>.yaml with its `defaults:` chain merged (Hydra group-local defaults, `_self_` last)."""
    raw = yaml.safe_load((CONFIG_DIR / "agent" / f"{name}.yaml").read_text())
    cfg = {}
    for d in raw.pop("defaults", []):
        if d != "_self_":
            cfg = _merge(cfg, load_agent(d))
    return _merge(cfg, raw)


def resolve(node, parents=()):
    """`${..key}` = sibling of the parent node (OmegaConf relative interpolation)."""
    if isinstance(node, dict):
        return {k: resolve(v, parents + (node,)) for k, v in node.items()}
    if isinstance(node, list):
        return [resolve(v, parents) for v in node]
    if isinstance(node, str) and node.startswith("${") and node.endswith("}"):
        ref = node[2:-1]
        up = len(ref) - len(ref.lstrip("."))
        return resolve(parents[-up][ref.lstrip(".")], parents[:len(parents) - up + 1])
    if isinstance(node, str):
        try:
            return float(node) if any(c in node for c in "eE.") and node.replace("-", "").replace("+", "").replace(
                ".", "").replace("e", "").replace("E", "").isdigit() else node
        except ValueError:
            return node
    return node


def build():
    out = {"agents": {}, "top": {}}
    for top, fname in TOP.items():
        raw = yaml.safe_load((CONFIG_DIR / fname).read_text())
        groups = {}
        for d in raw["defaults"]:
            if isinstance(d, dict):
                groups.update({k: v for k, v in d.items() if not k.startswith("override")})
        agent_name = groups["agent"]
        agent = _merge(load_agent(agent_name), raw.get("agent") or {})
        training = yaml.safe_load((CONFIG_DIR / "training" / f"{groups['training']}.yaml").read_text())
        training = _merge(training, {k: v for k, v in (raw.get("training") or {}).items()})
        out["agents"][top] = resolve(agent)
        actions_num, action_type = ENV_FACTS[groups["env"]]
        out["top"][top] = {"agent": agent_name, "env": groups["env"], "training": groups["training"],
                           "batch_size": training["batch_size"], "f16_precision": bool(training["f16_precision"]),
                           "actions_num": actions_num, "action_type": action_type, "obs_space_num": [64, 64, 3]}
    return out


if __name__ == "__main__":
    OUT.write_text(json.dumps(build(), indent=1, sort_keys=True) + "\n")
    print(f"wrote {OUT}")
