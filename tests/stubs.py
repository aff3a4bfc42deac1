"""Deterministic stub models (the reference's model duck type, mcts.py:211,235,501) restating
oracle/gen_golden.py:StubModel so that the GPU-box tests need no /root/reference."""
import numpy as np
import torch

import orc


def planes_to_bits(x):
    w = (1 << np.arange(64, dtype=np.uint64))
    own = int(((x[0].reshape(-1) > 0.5).astype(np.uint64) * w).sum())
    opp = int(((x[1].reshape(-1) > 0.5).astype(np.uint64) * w).sum())
    return own, opp


class StubModel:
    def __init__(self, kind, table=None, device="cpu"):
        self.kind = kind
        self.table = table
        self._p = torch.nn.Parameter(torch.zeros(1, device=device))
        self.calls = 0

    def parameters(self):
        return iter([self._p])

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def predict(self, x):
        dev = x.device
        xs = x.detach().cpu().numpy()
        B = xs.shape[0]
        self.calls += 1
        logits = np.zeros((B, 65), dtype=np.float32)
        values = np.zeros((B,), dtype=np.float32)
        for b in range(B):
            own, opp = planes_to_bits(xs[b])
            if self.kind == "E0":
                values[b] = np.float32(bin(own).count("1") - bin(opp).count("1")) / np.float32(64)
                continue
            h = orc.mix64((own * 0x9E3779B97F4A7C15) ^ orc.mix64(opp))
            values[b] = np.float32(((h >> 20) & 0xFFFF) - 32768) / np.float32(32768)
            if self.kind == "T1":
                sub = orc.mix64(h ^ 0xC2B2AE3D27D4EB4F)
                for i in range(64):
                    logits[b, i] = 0.0 if (sub >> i) & 1 else -np.inf
            else:
                raise ValueError("T2 priors are table-driven: use the Engine external path")
        return torch.from_numpy(logits).to(dev), torch.from_numpy(values).to(dev)


def perturb_bn(model, seed):
    """restates oracle/gen_golden.py:perturb_bn (non-trivial BN statistics, unsaturated value head)"""
    gen = torch.Generator().manual_seed(seed)
    for name, m in model.named_modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=gen) * 0.2)
            m.running_var.copy_(torch.rand(m.num_features, generator=gen) * 1.5 + 0.25)
            m.weight.data.copy_(torch.rand(m.num_features, generator=gen) * 0.5 + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=gen) * 0.1)
    model.value_fc2.weight.data.mul_(0.05)


def tame(model):
    """restates oracle/gen_golden.py:tame (damped residual branches + policy head for the deep-tower goldens)"""
    for blk in model.res_blocks:
        blk.bn2.weight.data.mul_(0.1)
        blk.bn2.bias.data.mul_(0.1)
    model.policy_fc.weight.data.mul_(0.1)
