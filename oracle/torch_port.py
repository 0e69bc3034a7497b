"""
torch-CPU restatement of the reference nn.Modules and losses  --  TEST INFRASTRUCTURE ONLY.

Used for (a) the CPU baseline leg of bench.py (`cpu_baseline`, `--impl reference`): the reference's hot
path on CPU *is* torch eager (MKL sgemm + aten elementwise), so the faithful "port" is the same op sequence
issued through torch.nn.functional; (b) the training oracle: loss + gradients via torch autograd, pinned
against the live-reference fixtures tests/golden/ref_train_*.npz.  Never imported by monoloco_b200/.

Reference: monoloco/network/architectures.py:48-71, 88-102, 135-145; monoloco/train/losses.py:28-73, 104-142;
monoloco/network/process.py:231-254, 293-304.
"""
import torch
import torch.nn.functional as F


def to_torch(sd, requires_grad=False):
    out = {}
    for k, v in sd.items():
        t = torch.as_tensor(v).clone()
        if requires_grad and t.is_floating_point() and 'running' not in k:
            t.requires_grad_(True)
        out[k] = t
    return out


def _bn(x, sd, name, training, momentum=0.1):
    return F.batch_norm(x, sd[name + '.running_mean'], sd[name + '.running_var'], sd[name + '.weight'],
                        sd[name + '.bias'], training=training, momentum=momentum, eps=1e-5)


def _block(x, sd, lin, bn, training, p, relu=True, masks=None):
    y = F.linear(x, sd[lin + '.weight'], sd[lin + '.bias'])
    if bn is not None:
        y = _bn(y, sd, bn, training)
    if relu:
        y = F.relu(y)
    if training and p > 0:
        if masks is not None:  # explicit keep masks (one per BatchNorm block, forward order): nn.Dropout semantics
            y = y * masks.pop(0).to(y.dtype) / (1.0 - p)
        else:
            y = F.dropout(y, p, True)
    return y


def model_forward(sd, x, training=False, p_dropout=0.0, masks=None):
    """LocoModel (architectures.py:48-71) or MonolocoModel (:135-145), picked from the checkpoint keys."""
    masks = list(masks) if masks is not None else None
    y = _block(x, sd, 'w1', 'batch_norm1', training, p_dropout, masks=masks)
    i = 0
    while 'linear_stages.%d.w1.weight' % i in sd:
        p = 'linear_stages.%d' % i
        z = _block(y, sd, p + '.w1', p + '.batch_norm1', training, p_dropout, masks=masks)
        z = _block(z, sd, p + '.w2', p + '.batch_norm2', training, p_dropout, masks=masks)
        y = y + z
        i += 1
    if 'w_fin.weight' in sd:
        y = F.linear(y, sd['w2.weight'], sd['w2.bias'])
        aux = F.linear(y, sd['w_aux.weight'], sd['w_aux.bias'])
        y = _block(y, sd, 'w3', 'batch_norm3', training, p_dropout, masks=masks)
        y = F.linear(y, sd['w_fin.weight'], sd['w_fin.bias'])
        return torch.cat((y, aux), dim=1)
    return F.linear(y, sd['w2.weight'], sd['w2.bias'])


def multi_task_loss(outputs, labels, tasks, lambdas=None, log_sigmas=None):
    """MultiTaskLoss (losses.py:59-73) / AutoTuneMultiTaskLoss (:28-43), phase='train'."""
    cols = {'x': slice(0, 1), 'y': slice(1, 2), 'd': slice(2, 4), 'h': slice(4, 5), 'w': slice(5, 6),
            'l': slice(6, 7), 'ori': slice(7, 9), 'aux': slice(9, 10)}
    gts = {'x': slice(0, 1), 'y': slice(1, 2), 'd': slice(3, 4), 'h': slice(4, 5), 'w': slice(5, 6),
           'l': slice(6, 7), 'ori': slice(7, 9), 'aux': slice(10, 11)}
    lambdas = lambdas if lambdas is not None else (1,) * len(tasks)
    vals = []
    for i, t in enumerate(tasks):
        o, g = outputs[:, cols[t]], labels[:, gts[t]]
        if t == 'd':
            mu, si = o[:, 0:1], o[:, 1:2]
            v = (torch.abs(1 - mu / g) * torch.exp(-si) + 0.01 + si + 2).mean()  # losses.py:121-131,139
        elif t == 'aux':
            v = F.binary_cross_entropy_with_logits(o, g)
        else:
            v = F.l1_loss(o, g)
        v = lambdas[i] * v
        if log_sigmas is not None:
            v = v / (2.0 * (log_sigmas[i].exp() ** 2))
        vals.append(v)
    loss = sum(vals)
    if log_sigmas is not None:
        loss = loss + log_sigmas.sum()
    return loss, vals
