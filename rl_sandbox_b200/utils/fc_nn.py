"""MLP builder with the reference's layer indexing (reference: rl_sandbox/utils/fc_nn.py:4-23).

Sequential indices are part of the checkpoint format (``actor.0.weight`` ... ``actor.12.bias``):
Linear at 0,3,6,..., a LayerNorm ALWAYS at 1 (even when layer_norm=False, fc_nn.py:15),
LayerNorm|Identity at 4,7,..., activation at 2,5,8,..., final Linear, then the output layer.
"""
import typing as t

from torch import nn


def fc_nn_generator(input_num: int, output_num: int, hidden_size: int, num_layers: int,
                    intermediate_activation: t.Type[nn.Module] = nn.ReLU,
                    final_activation: nn.Module = nn.Identity(), layer_norm: bool = False):
    assert num_layers >= 3
    widths = [input_num] + [hidden_size] * (num_layers - 1)
    mods: list[nn.Module] = []
    for i in range(num_layers - 1):
        mods.append(nn.Linear(widths[i], widths[i + 1]))
        mods.append(nn.LayerNorm(hidden_size) if (i == 0 or layer_norm) else nn.Identity())
        mods.append(intermediate_activation(inplace=True))
    mods += [nn.Linear(hidden_size, output_num), final_activation]
    return nn.Sequential(*mods)
