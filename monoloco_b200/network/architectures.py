"""
LocoModel / MonolocoModel with the reference's constructor, sub-module names (=> identical state_dict keys, the
checkpoint ABI of SURVEY.md §8b) and nn.Module behaviour, but whose forward is ONE fused CUDA kernel
(monoloco_b200/csrc/forward.cu) instead of ~31 eager launches.

Reference: monoloco/network/architectures.py:6-71 (LocoModel), 74-102 (MyLinearSimple), 105-145 (MonolocoModel),
148-176 (MyLinear).

* eval mode (`model.eval()`): fused inference kernel; `model.dropout.training = True` (the MC-dropout poke of
  net.py:141) switches the two top-level dropout sites on inside the kernel.
* train mode: fused train step (forward + backward kernels) through `monoloco_b200.train.fused` (autograd.Function).
There is no eager fallback: on a machine without the CUDA library / a B200 forward raises.
"""
import torch
from torch import nn

from ..engine import LocoEngine


class _Stage(nn.Module):
    """MyLinearSimple / MyLinear parameter container (architectures.py:74-86, 148-160)."""

    def __init__(self, linear_size, p_dropout=0.5):
        super().__init__()
        self.l_size = linear_size
        self.relu = nn.ReLU(inplace=True)
        self.dropout = nn.Dropout(p_dropout)
        self.w1 = nn.Linear(self.l_size, self.l_size)
        self.batch_norm1 = nn.BatchNorm1d(self.l_size)
        self.w2 = nn.Linear(self.l_size, self.l_size)
        self.batch_norm2 = nn.BatchNorm1d(self.l_size)

    def forward(self, x):  # pragma: no cover - the parent runs the whole network in one kernel
        raise RuntimeError("stages are executed inside the fused network kernel; call the parent model")


MyLinearSimple = _Stage
MyLinear = _Stage


class _FusedModel(nn.Module):
    _engine = None
    _engine_key = None

    def __getstate__(self):
        # the engine is a ctypes handle to device memory: never copied or pickled (copy.deepcopy(model) for a best-model
        # snapshot, torch.save(model)); the copy rebuilds its own engine lazily on first eval forward
        state = self.__dict__.copy()
        state.pop('_engine', None)
        state.pop('_engine_key', None)
        return state

    def _state_key(self):
        return tuple((id(t), t._version) for t in self.state_dict(keep_vars=True).values())

    def engine(self):
        """Device-resident packed copy of the current parameters; re-packed when any tensor changed."""
        key = self._state_key()
        if self._engine is None or key != self._engine_key:
            sd = {k: v.detach().cpu().numpy() for k, v in self.state_dict().items()}
            dev = next(self.parameters()).device
            if dev.type != 'cuda':
                if not torch.cuda.is_available():
                    raise RuntimeError("monoloco_b200: the model forward runs on a B200 only (no CPU fallback)")
                dev = torch.device('cuda', torch.cuda.current_device())
            if self._engine is None:
                object.__setattr__(self, '_engine', LocoEngine(sd, p_dropout=self.p_dropout, device=dev))
            else:
                self._engine.update_weights(sd)
            object.__setattr__(self, '_engine_key', key)
        return self._engine

    def forward(self, x):
        if self.training:
            from ..train.fused import fused_train_forward
            return fused_train_forward(self, x)
        if torch.is_grad_enabled() and x.requires_grad:
            raise RuntimeError("monoloco_b200: eval-mode forward does not build an autograd graph")
        eng = self.engine()
        xin = x.detach()
        xin = (xin if xin.is_cuda else xin.to(eng.device)).float()
        out = eng.forward(xin, want_dec=False, dropout=bool(self.dropout.training))['raw']
        return out if x.is_cuda else out.to(x.device)


class LocoModel(_FusedModel):
    """architectures.py:6-46 (same arguments, same attribute names)."""

    def __init__(self, input_size, output_size=2, linear_size=512, p_dropout=0.2, num_stage=3, device='cuda'):
        super().__init__()
        self.num_stage = num_stage
        self.stereo_size = input_size
        self.mono_size = int(input_size / 2)
        self.output_size = output_size - 1
        self.linear_size = linear_size
        self.p_dropout = p_dropout
        self.device = device
        self.w1 = nn.Linear(self.stereo_size, self.linear_size)
        self.batch_norm1 = nn.BatchNorm1d(self.linear_size)
        self.linear_stages = nn.ModuleList([MyLinearSimple(self.linear_size, self.p_dropout) for _ in range(num_stage)])
        self.w2 = nn.Linear(self.linear_size, self.linear_size)
        self.w3 = nn.Linear(self.linear_size, self.linear_size)
        self.batch_norm3 = nn.BatchNorm1d(self.linear_size)
        self.w_aux = nn.Linear(self.linear_size, 1)
        self.w_fin = nn.Linear(self.linear_size, self.output_size)
        self.relu = nn.ReLU(inplace=True)
        self.dropout = nn.Dropout(self.p_dropout)


class MonolocoModel(_FusedModel):
    """architectures.py:105-133."""

    def __init__(self, input_size, output_size=2, linear_size=256, p_dropout=0.2, num_stage=3):
        super().__init__()
        self.input_size = input_size
        self.output_size = output_size
        self.linear_size = linear_size
        self.p_dropout = p_dropout
        self.num_stage = num_stage
        self.w1 = nn.Linear(self.input_size, self.linear_size)
        self.batch_norm1 = nn.BatchNorm1d(self.linear_size)
        self.linear_stages = nn.ModuleList([MyLinear(self.linear_size, self.p_dropout) for _ in range(num_stage)])
        self.w2 = nn.Linear(self.linear_size, self.output_size)
        self.relu = nn.ReLU(inplace=True)
        self.dropout = nn.Dropout(self.p_dropout)
