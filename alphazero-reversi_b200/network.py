"""K4 host side: the reference network's architecture (for random-init weights and as the fp32
torch reference of the bf16 kernels), the state_dict packer, and the `RvsNetwork` evaluator handle.

reference: src/model/network.py:14-117 (ResBlock, AlphaZeroNetwork), checkpoint key handling
src/mcts/mcts.py:459-479 (`_script_module.` duplicates after TorchScript compilation).
"""
from typing import Dict, List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L


class ResBlock(nn.Module):
    def __init__(self, num_filters: int):
        super().__init__()
        self.conv1 = nn.Conv2d(num_filters, num_filters, kernel_size=3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(num_filters)
        self.conv2 = nn.Conv2d(num_filters, num_filters, kernel_size=3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(num_filters)

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return F.relu(out + x)


class AlphaZeroNetwork(nn.Module):
    """Same modules, names, creation order and initialisation as the reference network, so that
    `torch.manual_seed(s); AlphaZeroNetwork(8, nb, nf)` yields the reference's weights and
    reference checkpoints load with `load_state_dict` (no TorchScript wrapper here)."""

    def __init__(self, board_size: int = 8, num_res_blocks: int = 5, num_filters: int = 128):
        super().__init__()
        self.board_size = board_size
        self.num_filters = num_filters
        self.conv = nn.Conv2d(3, num_filters, kernel_size=3, padding=1, bias=False)
        self.bn = nn.BatchNorm2d(num_filters)
        self.res_blocks = nn.ModuleList([ResBlock(num_filters) for _ in range(num_res_blocks)])
        self.policy_conv = nn.Conv2d(num_filters, 2, kernel_size=1, bias=False)
        self.policy_bn = nn.BatchNorm2d(2)
        self.policy_fc = nn.Linear(2 * board_size * board_size, board_size * board_size + 1)
        self.value_conv = nn.Conv2d(num_filters, 1, kernel_size=1, bias=False)
        self.value_bn = nn.BatchNorm2d(1)
        self.value_fc1 = nn.Linear(board_size * board_size, 256)
        self.value_fc2 = nn.Linear(256, 1)
        for m in self.modules():  # network.py:71-78
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        x = F.relu(self.bn(self.conv(x)))
        for blk in self.res_blocks:
            x = blk(x)
        p = F.relu(self.policy_bn(self.policy_conv(x)))
        p = self.policy_fc(p.contiguous().view(x.size(0), -1))
        v = F.relu(self.value_bn(self.value_conv(x)))
        v = F.relu(self.value_fc1(v.contiguous().view(x.size(0), -1)))
        v = torch.tanh(self.value_fc2(v))
        return p, v.squeeze(1)

    def predict(self, board_state, valid_moves=None):
        if board_state.dim() == 3:
            board_state = board_state.unsqueeze(0)
        return self.forward(board_state)


def canonical_keys(num_res_blocks: int) -> List[str]:
    """state_dict keys in the order rvs_engine_load_weights expects (num_batches_tracked excluded)"""
    def bn(p):
        return [f"{p}.weight", f"{p}.bias", f"{p}.running_mean", f"{p}.running_var"]
    keys = ["conv.weight"] + bn("bn")
    for i in range(num_res_blocks):
        keys += [f"res_blocks.{i}.conv1.weight"] + bn(f"res_blocks.{i}.bn1")
        keys += [f"res_blocks.{i}.conv2.weight"] + bn(f"res_blocks.{i}.bn2")
    keys += ["policy_conv.weight"] + bn("policy_bn") + ["policy_fc.weight", "policy_fc.bias"]
    keys += ["value_conv.weight"] + bn("value_bn") + ["value_fc1.weight", "value_fc1.bias", "value_fc2.weight", "value_fc2.bias"]
    return keys


def pack_state_dict(state_dict: Dict[str, torch.Tensor]):
    """flat f32 tensor + (blocks, filters) from a reference state_dict; accepts the 168-key form
    with `_script_module.` duplicates (pipeline.py:410-418)"""
    sd = {}
    for k, v in state_dict.items():
        sd[k[len("_script_module."):] if k.startswith("_script_module.") else k] = v
    blocks = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("res_blocks."))
    filters = sd["conv.weight"].shape[0]
    flat = torch.cat([sd[k].detach().to(torch.float32).reshape(-1).cpu() for k in canonical_keys(blocks)])
    return flat.contiguous(), blocks, filters


class RvsNetwork:
    """Evaluator handle for the built-in bf16 tensor-core network (RVS_EVAL_NN).  Pass it where the
    reference takes `model`: MCTS(RvsNetwork.from_module(net), ...) / SelfPlay(...)."""
    evaluator = L.EVAL_NN

    def __init__(self, flat: torch.Tensor, blocks: int, filters: int):
        self.flat, self.net_blocks, self.net_filters = flat, blocks, filters

    @classmethod
    def from_module(cls, module: nn.Module) -> "RvsNetwork":
        return cls(*pack_state_dict(module.state_dict()))

    @classmethod
    def from_state_dict(cls, sd) -> "RvsNetwork":
        return cls(*pack_state_dict(sd))

    @classmethod
    def from_checkpoint(cls, path: str, trusted: bool = False) -> "RvsNetwork":
        """a reference checkpoint file: `checkpoint_XXXX.pth` (dict with 'model_state_dict',
        pipeline.py:463-480) or `best_model.pth` (bare state_dict, pipeline.py:482-485), with or
        without the `_script_module.` duplicates of a TorchScript-compiled model (mcts.py:459-479)"""
        # weights_only=True: a checkpoint is data, not code.  Reference checkpoints (state_dict, or a dict of
        # state_dicts + primitives) load with the safe unpickler; only a caller who vouches for the file
        # (trusted=True) gets the arbitrary-code pickle path (e.g. a whole pickled nn.Module).
        obj = torch.load(path, map_location="cpu", weights_only=not trusted)
        if isinstance(obj, dict) and "model_state_dict" in obj:
            obj = obj["model_state_dict"]
        if hasattr(obj, "state_dict"):
            obj = obj.state_dict()
        return cls(*pack_state_dict(obj))

    def attach(self, engine) -> None:
        engine.load_weights(self.flat)
