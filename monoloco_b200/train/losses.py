"""Multi-task losses with the reference's class names, constructor arguments and return values
(monoloco/train/losses.py:17-142, 241-254).  These operate on the [B, 9|10] network outputs (a few KB) and stay
device-side torch ops -- unlike the reference's LaplacianLoss they do not pull tensors to the host every call
(losses.py:127 `.cpu().detach().numpy()`).  The fully fused alternative (forward + this loss + backward in one
kernel launch) is `monoloco_b200.train.fused.train_step`."""
import torch
from torch import nn

from ..network.process import extract_labels, extract_labels_aux, extract_outputs


class LaplacianLoss(nn.Module):
    """losses.py:104-142 (size_average / reduce / evaluate keep their meaning)."""

    def __init__(self, size_average=True, reduce=True, evaluate=False):
        super().__init__()
        self.size_average, self.reduce, self.evaluate = size_average, reduce, evaluate

    def laplacian_1d(self, mu_si, xx):
        mu, si = mu_si[:, 0:1], mu_si[:, 1:2]
        norm = 1 - mu / xx
        if self.evaluate:
            return float(torch.mean(torch.abs(norm))), float(torch.mean(torch.exp(si)))  # losses.py:127-130
        return torch.abs(norm) * torch.exp(-si) + 0.01 + si + 2

    def forward(self, outputs, targets):
        values = self.laplacian_1d(outputs, targets)
        if not self.reduce or self.evaluate:
            return values
        return torch.mean(values) if self.size_average else torch.sum(values)


def angle_loss(orient, gt_orient):
    """losses.py:241-248 (evaluation only)."""
    angles = torch.atan2(orient[:, 0], orient[:, 1])
    gt_angles = torch.atan2(gt_orient[:, 0], gt_orient[:, 1])
    return torch.mean(torch.abs(angles - gt_angles)) * 180 / 3.14


def l1_loss_from_laplace(out, gt_out):
    """losses.py:251-254 (evaluation only)."""
    return torch.mean(torch.abs(out[:, 0:1] - gt_out))


class CompositeLoss(nn.Module):
    """losses.py:76-101."""

    def __init__(self, tasks):
        super().__init__()
        self.tasks = tasks
        self.multi_loss_tr = {task: (LaplacianLoss() if task == 'd' else
                                     (nn.BCEWithLogitsLoss() if task in ('aux',) else nn.L1Loss())) for task in tasks}
        self.multi_loss_val = {}
        for task in tasks:
            if task == 'd':
                loss = l1_loss_from_laplace
            elif task == 'ori':
                loss = angle_loss
            elif task in ('aux',):
                loss = nn.BCEWithLogitsLoss()
            else:
                loss = nn.L1Loss()
            self.multi_loss_val[task] = loss

    def forward(self):
        return [self.multi_loss_tr[t] for t in self.tasks], [self.multi_loss_val[t] for t in self.tasks]


class MultiTaskLoss(nn.Module):
    """losses.py:46-73."""

    def __init__(self, losses_tr, losses_val, lambdas, tasks):
        super().__init__()
        self.losses = nn.ModuleList(losses_tr)
        self.losses_val = losses_val
        self.lambdas = lambdas
        self.tasks = tasks
        self.flag_aux = len(self.tasks) == 1 and self.tasks[0] == 'aux'

    def forward(self, outputs, labels, phase='train'):
        assert phase in ('train', 'val')
        out = extract_outputs(outputs, tasks=self.tasks)
        gt_out = extract_labels_aux(labels, tasks=self.tasks) if self.flag_aux else extract_labels(labels, tasks=self.tasks)
        loss_values = [lam * l(o, g) for lam, l, o, g in zip(self.lambdas, self.losses, out, gt_out)]
        loss = sum(loss_values)
        if phase == 'val':
            return loss, [l(o, g) for l, o, g in zip(self.losses_val, out, gt_out)]
        return loss, loss_values


class AutoTuneMultiTaskLoss(nn.Module):
    """losses.py:17-43."""

    def __init__(self, losses_tr, losses_val, lambdas, tasks):
        super().__init__()
        assert all(l in (0.0, 1.0) for l in lambdas)
        self.losses = nn.ModuleList(losses_tr)
        self.losses_val = losses_val
        self.lambdas = lambdas
        self.tasks = tasks
        self.log_sigmas = nn.Parameter(torch.zeros((len(lambdas),), dtype=torch.float32), requires_grad=True)

    def forward(self, outputs, labels, phase='train'):
        assert phase in ('train', 'val')
        out = extract_outputs(outputs, tasks=self.tasks)
        gt_out = extract_labels(labels, tasks=self.tasks)
        loss_values = [lam * l(o, g) / (2.0 * (log_sigma.exp() ** 2))
                       for lam, log_sigma, l, o, g in zip(self.lambdas, self.log_sigmas, self.losses, out, gt_out)]
        loss = sum(loss_values) + sum(log_sigma for log_sigma in self.log_sigmas)
        if phase == 'val':
            vals = [l(o, g) for l, o, g in zip(self.losses_val, out, gt_out)]
            vals.extend([s.exp() for s in self.log_sigmas])
            return loss, vals
        return loss, loss_values
