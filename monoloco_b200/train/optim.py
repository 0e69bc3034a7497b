"""`FusedClipAdam`: torch.nn.utils.clip_grad_norm_ + torch.optim.Adam.step() (trainer.py:159-160) in two multi-tensor
kernel launches (csrc/optim.cu), no host synchronisation.  Same hyper-parameters and state semantics as
torch.optim.Adam(params, lr) without amsgrad; `param_groups[0]['lr']` is honoured so lr_scheduler.StepLR keeps working.

Differences from the two reference calls, by design: the clip coefficient is applied inside the update, `p.grad` itself is
left UN-clipped (torch's clip_grad_norm_ rescales it in place -- read the gradients before `step()` if you log them); one
parameter group only; Adam's `step` lives in `state[p]['step']` exactly like torch's, so state_dict()/load_state_dict()
round-trips keep the bias correction."""
import ctypes as C

import torch

from .. import _lib as L_


class FusedClipAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=3.0, clip_params=None):
        """params: iterable of tensors (model + loss parameters); clip_params: the subset whose gradients are norm-clipped
        (default: all) -- the reference clips model.parameters() and optimises chain(model, mt_loss) (trainer.py:128-129,159)."""
        super().__init__(list(params), dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise ValueError("FusedClipAdam takes ONE parameter group (what trainer.py:128-129 builds); per-group "
                             "hyper-parameters are not supported")
        self.max_norm = float(max_norm)
        self._clip_ids = None if clip_params is None else {id(p) for p in clip_params}
        self._lib = L_.lib()
        self._scratch = None

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is None
        group = self.param_groups[0]
        ps = [p for p in group['params'] if p.grad is not None]
        if not ps:
            return None
        dev = ps[0].device
        if dev.type != 'cuda':
            raise RuntimeError("FusedClipAdam runs on CUDA tensors only")
        if self._scratch is None:
            self._scratch = torch.zeros(1, dtype=torch.float64, device=dev)
        for p in ps:
            st = self.state[p]
            if not st:
                st['step'] = torch.tensor(0.0)  # per-parameter like torch.optim.Adam: survives state_dict round trips
                st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st['step'] += 1
        step = int(self.state[ps[0]]['step'].item())
        if any(int(self.state[p]['step'].item()) != step for p in ps):
            raise RuntimeError("FusedClipAdam: parameters with different step counts in one multi-tensor launch")
        n = len(ps)
        arr = lambda vals: (C.c_void_p * n)(*vals)  # noqa: E731
        grads = [p.grad.contiguous() for p in ps]
        sizes = (C.c_int64 * n)(*[p.numel() for p in ps])
        clip = (C.c_int32 * n)(*[1 if (self._clip_ids is None or id(p) in self._clip_ids) else 0 for p in ps])
        b1, b2 = group['betas']
        L_.check(self._lib.mlb_adam_clip_step(
            n, arr([p.data_ptr() for p in ps]), arr([g.data_ptr() for g in grads]),
            arr([self.state[p]['exp_avg'].data_ptr() for p in ps]), arr([self.state[p]['exp_avg_sq'].data_ptr() for p in ps]),
            sizes, clip, self.max_norm, float(group['lr']), float(b1), float(b2), float(group['eps']),
            float(group['weight_decay']), step, self._scratch.data_ptr(),
            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), 'mlb_adam_clip_step')
        # the in-place update happened outside autograd's view: bump the version counters (no kernel) so that cached
        # packed copies of the weights (eval-mode engine) are refreshed on next use
        torch._C._autograd._unsafe_set_version_counter(ps, [p._version + 1 for p in ps])
        return None
