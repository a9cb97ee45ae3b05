"""Helpers shared by the parity tests: load a golden case produced by oracle/gen_golden.py and
regenerate its (seeded) parameters, start states and noise."""
import json
from pathlib import Path

import numpy as np
import torch

from oracle import oracle_port as orc
from oracle.gen_golden import CASES, make_noise

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_case(name):
    z = np.load(GOLDEN / f"imagine_{name}.npz")
    gold = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files if k != "meta"}
    meta = json.loads(str(z["meta"]))
    assert {k: meta[k] for k in CASES[name]} == CASES[name], "fixture is stale: re-run python -m oracle.gen_golden"
    wm, actor, critic = orc.make_params(meta["param_seed"], D=meta["D"], A=meta["A"], discrete=meta["discrete"],
                                        layer_norm=meta["layer_norm"], predict_discount=meta["predict_discount"])
    h0, z0 = orc.make_start(meta["start_seed"], meta["N"], meta["D"])
    lat, act = make_noise(meta)
    return dict(meta=meta, gold=gold, wm=wm, actor=actor, critic=critic, h0=h0, z0=z0, lat=lat, act=act)


def known_answers():
    return json.loads((GOLDEN / "lambda_known_answers.json").read_text())
