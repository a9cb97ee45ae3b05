"""What `hydra.utils.instantiate` does with the reference's agent configs (hydra is not installed here):
`_target_` dotted paths are imported, `_partial_: true` nodes become functools.partial, nested nodes are
instantiated first, extra keyword arguments override (train.py:54-63)."""
import importlib
from functools import partial


def locate(path: str):
    mod, _, attr = path.rpartition(".")
    return getattr(importlib.import_module(mod), attr)


def instantiate(node, **overrides):
    if isinstance(node, list):
        return [instantiate(v) for v in node]
    if not isinstance(node, dict):
        return node
    if "_target_" not in node:
        return {k: instantiate(v) for k, v in node.items()}
    kwargs = {k: instantiate(v) for k, v in node.items() if k not in ("_target_", "_partial_")}
    kwargs.update(overrides)
    target = locate(node["_target_"])
    return partial(target, **kwargs) if node.get("_partial_") else target(**kwargs)


def build_agent(fixture: dict, name: str, device: str, **overrides):
    """The agent of top-level config `name` exactly as train.py builds it (env facts from the fixture)."""
    top = fixture["top"][name]
    kw = dict(obs_space_num=top["obs_space_num"], actions_num=top["actions_num"], action_type=top["action_type"],
              device_type=device, f16_precision=top["f16_precision"], logger=None)
    kw.update(overrides)
    return instantiate(fixture["agents"][name], **kw)
