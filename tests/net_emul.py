"""Torch restatement of the K4 kernels' ARITHMETIC (not of their code): eval-mode BatchNorm folded into the
convolution weights, tower weights and activations rounded to bf16, f32 accumulation, f32 heads -- the
"plain PyTorch fp32 reference of the same op" for the tcgen05 tower (alphazero-reversi_b200/csrc/
rvs_conv_tc.cu, rvs_net.cu).  Against this the tensor-core path may differ only by the f32 summation order
(and by bf16 roundings that such last-bit differences flip).  Runs on the CPU so that no TF32 / cuDNN
algorithm choice enters.  Reference network: src/model/network.py:80-117."""
import torch
import torch.nn.functional as F


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _fold(conv_w, bn):
    scale = bn.weight / torch.sqrt(bn.running_var + 1e-5)  # rvs_net.cu: fold_conv3x3_kernel
    return conv_w * scale.view(-1, 1, 1, 1), bn.bias - bn.running_mean * scale


@torch.no_grad()
def emulate(net, planes):
    """(logits [n,65], value [n]) of `net` (AlphaZeroNetwork mirror, eval mode) on planes [n,3,8,8] f32"""
    net = net.cpu().eval()
    x = planes.cpu().float()
    C = net.num_filters
    w, b = _fold(net.conv.weight, net.bn)
    if C == 128:  # first layer on the tensor cores: bf16 weights (inputs are exact 0/1)
        w = _bf16(w)
    x = _bf16(F.relu(F.conv2d(x, w, padding=1) + b.view(1, -1, 1, 1)))
    for blk in net.res_blocks:
        w1, b1 = _fold(blk.conv1.weight, blk.bn1)
        w2, b2 = _fold(blk.conv2.weight, blk.bn2)
        t = _bf16(F.relu(F.conv2d(x, _bf16(w1), padding=1) + b1.view(1, -1, 1, 1)))
        x = _bf16(F.relu(F.conv2d(t, _bf16(w2), padding=1) + b2.view(1, -1, 1, 1) + x))
    pw, pb = _fold(net.policy_conv.weight, net.policy_bn)   # heads in f32 on the bf16-rounded activations
    p = F.relu(F.conv2d(x, pw) + pb.view(1, -1, 1, 1))
    logits = net.policy_fc(p.reshape(x.size(0), -1))
    vw, vb = _fold(net.value_conv.weight, net.value_bn)
    v = F.relu(F.conv2d(x, vw) + vb.view(1, -1, 1, 1))
    v = F.relu(net.value_fc1(v.reshape(x.size(0), -1)))
    return logits, torch.tanh(net.value_fc2(v)).squeeze(1)
