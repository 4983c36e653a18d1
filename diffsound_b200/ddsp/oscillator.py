"""Modal synthesis modules: damped-sinusoid banks driven by modal frequencies.

API mirror of src/ddsp/oscillator.py:10-324 (`WeightedParam`, `WeightedSum`,
`DirectValue`, `DampedOscillator`, `GTDampedOscillator`,
`TraditionalDampedOscillator`, `init_damps`): same constructor arguments,
attributes (`damped_freq`, `alpha`, `beta`, `amp`, `forces`, ...) and output
shapes/dtypes.

What differs underneath: the reference materialises (audio_num, mode_num,
sample_num) tensors and runs cumsum/exp/sin/sum over them (oscillator.py:128-140,
:230-242, :297-304).  Damping and damped frequency are per mode in all of these
modules, so here the (mode_num,) vectors go to one CUDA kernel
(`ds_modal_synth_fwd` / `_bwd`, csrc/synth.cu) through `ModalSynth`, the force FIR to
`ds_force_fir` and the noise branch of `GTDampedOscillator` to `ds_filtered_noise_*`
(csrc/noise.cu); only the small reparameterisations stay in torch.  `damped_freq`, which the
reference exposes as a broadcast (B, k, T) tensor and the training scripts read
as `damped_freq[:, :, 0]` (material_sync_train.py:156-159), is exposed as an
expanded (stride-0) view of the (1, k, 1) values: same shape, no storage.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import native
from ..diffelastic.material_model import Material, MatSet  # noqa: F401  (re-exported like the reference)
from .filtered_noise import FilteredNoise
from .utils import modifed_sigmoid


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("diffsound_b200 needs a CUDA device (there is no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


class WeightedParam(nn.Module):
    """Scalar = convex combination (softplus-normalised) of `values_list` (oscillator.py:10-21).
    `values_list` is a plain attribute, not a buffer, exactly like the reference: the module
    (and therefore E, nu of `TrainableLinear`) lives on the CPU unless moved by hand."""

    def __init__(self, values_list: torch.Tensor):
        super().__init__()
        self.values_list = values_list
        self.probablity = nn.Parameter(torch.zeros(len(values_list)))
        self.probablity.data.uniform_(-1, 1)

    def forward(self):
        p = F.softplus(self.probablity)
        p = p / p.sum()
        return (self.values_list * p).sum()


class WeightedSum(nn.Module):
    """Tensor of shape `dims`, each entry a softplus-normalised mix of `vlist` (oscillator.py:23-35)."""

    def __init__(self, dims: list, vlist: list):
        super().__init__()
        self.values_list = torch.tensor([float(v) for v in vlist], dtype=torch.float32, device=_device())
        self.params = nn.Parameter(torch.zeros(*dims, len(self.values_list), device=_device()))
        self.params.data.uniform_(-4, 4)

    def forward(self):
        x = F.softplus(self.params)
        x = x / x.sum(dim=-1).unsqueeze(-1)
        return (self.values_list * x).sum(dim=-1)


class DirectValue(nn.Module):
    def __init__(self, dims: list):
        super().__init__()
        self.value = nn.Parameter(torch.zeros(*dims, device=_device()))
        self.value.data.uniform_(0, 0.04)

    def forward(self):
        return modifed_sigmoid(self.value)


class ModalSynth(torch.autograd.Function):
    """y[b,t] = sum_m amp[b,m] exp(-damp[m] (t+1)/sr) sin(2 pi freq[m] (t+1)/sr); both passes native."""

    @staticmethod
    def forward(ctx, amp, damp, freq, sample_num, sr):
        amp32 = amp.detach().to(torch.float32).contiguous()
        d32 = damp.detach().to(torch.float32).contiguous()
        f32 = freq.detach().to(torch.float32).contiguous()
        ctx.save_for_backward(amp32, d32, f32)
        ctx.sr = sr
        ctx.in_dtypes = (amp.dtype, damp.dtype, freq.dtype)
        return native.modal_synth_fwd(amp32, d32, f32, int(sample_num), float(sr))

    @staticmethod
    def backward(ctx, gy):
        amp32, d32, f32 = ctx.saved_tensors
        ga, gd, gf = native.modal_synth_bwd(amp32, d32, f32, gy.to(torch.float32).contiguous(), ctx.sr)
        ta, td, tf = ctx.in_dtypes
        return ga.to(ta), gd.to(td), gf.to(tf), None, None


def modal_synth(amp, damp, freq, sample_num, sr):
    """amp (B,k), damp (k,), freq (k,) [Hz, damped] -> (B, sample_num) fp32."""
    return ModalSynth.apply(amp, damp, freq, sample_num, sr)


class ForceFIR(torch.autograd.Function):
    """out[b,t] = sum_i force[b,i] signal[b,t-i]: the reference's conv1d with the flipped force, groups = audio_num,
    padding F-1, cropped to sample_num (oscillator.py:305-309), as the native kernel ds_force_fir; the backward
    pass w.r.t. the signal is its adjoint (the forces are recorded data, not parameters)."""

    @staticmethod
    def forward(ctx, signal, force):
        ctx.save_for_backward(force)
        return native.force_fir(signal.to(torch.float32).contiguous(), force)

    @staticmethod
    def backward(ctx, g):
        (force,) = ctx.saved_tensors
        return native.force_fir(g.to(torch.float32).contiguous(), force, reverse=True), None


def _apply_force(signal, forces_natural):
    """Causal FIR with the force.  A unit impulse -- what the shipped experiments use
    (material_sync_train.py:103-104) -- is the identity and is skipped by the caller."""
    return ForceFIR.apply(signal, forces_natural)


def _is_unit_impulse(flipped_forces):
    f = flipped_forces.reshape(flipped_forces.shape[0], -1)
    return bool((f[:, -1] == 1).all()) and bool((f[:, :-1] == 0).all())


class _OscBase(nn.Module):
    def _setup_forces(self, forces, audio_num):
        self.forces = torch.flip(forces.reshape(audio_num, 1, -1), [-1]).to(_device())      # reference attribute (flipped)
        self._forces_natural = forces.reshape(audio_num, -1).to(_device()).to(torch.float32).contiguous()
        self.force_frame_num = forces.shape[-1]
        self._impulse = _is_unit_impulse(self.forces)

    def _render(self, amp, damp, freq_d):
        y = modal_synth(amp, damp.reshape(-1), freq_d.reshape(-1), self.sample_num, self.sr)
        if self._impulse:
            return y
        return _apply_force(y, self._forces_natural)


def _rayleigh(freq_linear, alpha, beta):
    """lambda = (2 pi f)^2; d = (alpha + beta lambda)/2; f_d = sqrt(lambda - d^2)/2pi (oscillator.py:287-292)."""
    lbd = (freq_linear * 2 * np.pi) ** 2
    damp = 0.5 * (alpha + beta * lbd)
    return damp, (lbd - damp ** 2) ** 0.5 / (2 * np.pi)


class DampedOscillator(_OscBase):
    def __init__(self, forces, audio_num, mode_num, sample_num, sr, f_range: list, mat: Material):
        super().__init__()
        self.audio_num, self.sr, self.sample_num, self.mode_num = audio_num, sr, sample_num, mode_num
        bin_num = 64
        self.alpha_list = torch.exp(torch.linspace(np.log(mat.alpha / 10), np.log(mat.alpha * 10), bin_num))
        self.alpha = WeightedSum([1, mode_num, 1], list(self.alpha_list))
        self.beta_list = torch.exp(torch.linspace(np.log(mat.beta / 10), np.log(mat.beta * 10), bin_num))
        self.mat = mat
        self.beta = WeightedSum([1, mode_num, 1], list(self.beta_list))
        self.amp = DirectValue([audio_num, mode_num, 1])
        self.noise = FilteredNoise(audio_num, 8000)     # "not used, just for load" (oscillator.py:79): state-dict parity
        self._setup_forces(forces, audio_num)

    def forward(self, freq_linear, non_linear_rate=0.0, noise_rate=0.0):
        # both arguments are accepted and ignored, as in the reference: its only uses are commented out
        # (oscillator.py:119,126,141)
        amp = self.amp().reshape(self.audio_num, self.mode_num)
        f = torch.reshape(freq_linear, (1, self.mode_num, 1))
        damp, fd = _rayleigh(f, self.alpha(), self.beta())
        self.damped_freq = fd.expand(self.audio_num, self.mode_num, self.sample_num)
        return self._render(amp, damp, fd)

    def _curve(self, freq_linear, damping_curve):
        freq = freq_linear.detach().cpu().numpy().reshape(-1)
        damp = torch.tensor([float(damping_curve(x)) for x in freq], dtype=torch.float32,
                            device=freq_linear.device).reshape(1, self.mode_num, 1)
        f = freq_linear.reshape(1, self.mode_num, 1)
        lbd = (f * 2 * np.pi) ** 2
        fd = (lbd - damp ** 2) ** 0.5 / (2 * np.pi)
        self.damped_freq = fd
        amp = torch.ones(self.audio_num, self.mode_num, dtype=torch.float32, device=freq_linear.device)
        return self._render(amp, damp, fd)

    def early(self, freq_linear, damping_curve):
        return self._curve(freq_linear, damping_curve)

    def forward_curve(self, freq_linear, damping_curve):
        signal = self._curve(freq_linear, damping_curve)
        return signal / torch.max(torch.abs(signal), dim=1, keepdim=True)[0]


class GTDampedOscillator(_OscBase):
    def __init__(self, forces, audio_num, mode_num, sample_num, sr, f_range: list, mat: Material):
        super().__init__()
        self.audio_num, self.sr, self.sample_num, self.mode_num = audio_num, sr, sample_num, mode_num
        self.freq_linear = WeightedSum([1, mode_num, 1], f_range)
        # (audio_num, mode_num, sample_num, bins) parameters (oscillator.py:187-188): part of the reference's
        # state dict and optimiser groups; they only enter the signal with non_linear_rate != 0
        self.freq_nonlinear = WeightedSum([audio_num, mode_num, sample_num], f_range)
        bin_num = 64
        self.alpha_list = torch.exp(torch.linspace(np.log(mat.alpha / 10), np.log(mat.alpha * 100), bin_num))
        self.alpha = WeightedSum([1, mode_num, 1], list(self.alpha_list))
        self.beta_list = torch.exp(torch.linspace(np.log(mat.beta / 10), np.log(mat.beta * 100), bin_num))
        self.mat = mat
        self.beta = WeightedSum([1, mode_num, 1], list(self.beta_list))
        self.amp = DirectValue([audio_num, mode_num, 1])
        self.noise = FilteredNoise(audio_num, sample_num)
        self._setup_forces(forces, audio_num)

    def damping(self):
        lbd_linear = (self.freq_linear() * 2 * np.pi) ** 2
        return 0.5 * (self.alpha() + self.beta() * lbd_linear)

    def forward(self, non_linear_rate=0.0, noise_rate=0.0):
        if non_linear_rate != 0.0:
            raise NotImplementedError("non_linear_rate != 0 (per-sample frequency offsets, oscillator.py:220) is not used by "
                                      "any experiment of the reference and has no kernel here")
        amp = self.amp().reshape(self.audio_num, self.mode_num)
        damp, fd = _rayleigh(self.freq_linear(), self.alpha(), self.beta())
        noise = self.noise()                     # drawn every call like the reference (oscillator.py:226)
        self.undamped_freq = ((2 * np.pi * fd) ** 2 + damp ** 2) ** 0.5 / (2 * np.pi)
        return self._render(amp, damp, fd) + noise * noise_rate


class TraditionalDampedOscillator(_OscBase):
    def __init__(self, forces, audio_num, mode_num, sample_num, sr, mat: Material):
        super().__init__()
        self.audio_num, self.sr, self.sample_num, self.mode_num = audio_num, sr, sample_num, mode_num
        self.alpha = mat.alpha
        self.beta = mat.beta
        self.mat = mat
        self._setup_forces(forces, audio_num)

    def forward(self, freq_linear):
        f = torch.reshape(freq_linear, (1, self.mode_num, 1))
        damp, fd = _rayleigh(f, self.alpha, self.beta)
        self.damped_freq = fd.expand(self.audio_num, self.mode_num, self.sample_num)
        self.undamped_freq = (((2 * np.pi * fd) ** 2 + damp ** 2) ** 0.5 / (2 * np.pi)).expand(
            self.audio_num, self.mode_num, self.sample_num)
        amp = torch.ones(self.audio_num, self.mode_num, dtype=torch.float32, device=f.device)
        return self._render(amp, damp, fd)


def init_damps(osc):
    """Pre-train alpha/beta towards the material table (oscillator.py:314-324)."""
    optimizer = torch.optim.Adam(list(osc.alpha.parameters()) + list(osc.beta.parameters()), lr=0.01)
    for _ in range(2000):
        optimizer.zero_grad()
        loss = (osc.alpha() - osc.mat.alpha) ** 2 / osc.mat.alpha ** 2 + (osc.beta() - osc.mat.beta) ** 2 / osc.mat.beta ** 2
        loss = loss.mean()
        loss.backward()
        optimizer.step()
