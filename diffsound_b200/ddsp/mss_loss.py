"""Multi-scale spectral loss.  API mirror of src/ddsp/mss_loss.py:70-147 (`SSSLoss`, `MSSLoss`): same
constructor arguments, `n_ffts`, `losses[i].log_spec`, `forward(x_pred, x_true, freq=None, scale=1.0)`.

The reference computes four torchaudio spectrograms per scale and differentiates through cuFFT; here
each scale is one fused kernel pass forward and one backward (`ds_mss_loss_fwd` / `_bwd`, csrc/mss.cu:
the predicted and target frame share one complex FFT in shared memory and the loss terms are reduced per
frame, no spectrogram is materialised).  `type='geomloss'` (Sinkhorn divergence on spectrogram point
clouds, mss_loss.py:19-52,107-115) needs the third-party `geomloss` package, which is not part of this
path: it raises NotImplementedError."""
import numpy as np
import torch
import torch.nn as nn

from .. import native

_MODES = {"l1_loss": 0, "rmse_loss": 1}


def clip_spec(x, scale):
    freq_length = x.shape[-2]
    return x[..., :int(freq_length * scale), :]


class _SpectralLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_pred, x_true, n_fft, hop, mode, alpha, eps):
        xp = x_pred.detach().to(torch.float32).contiguous()
        xt = x_true.detach().to(torch.float32).contiguous()
        loss, scratch = native.mss_loss_fwd(xp, xt, n_fft, hop, mode, alpha, eps)
        ctx.save_for_backward(xp, xt, loss, scratch)
        ctx.cfg = (n_fft, hop, mode, alpha, eps)
        ctx.in_dtype = x_pred.dtype
        return loss.to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        xp, xt, loss, scratch = ctx.saved_tensors
        n_fft, hop, mode, alpha, eps = ctx.cfg
        gx = torch.empty_like(xp)
        native.mss_loss_bwd(xp, xt, n_fft, hop, mode, alpha, eps, loss, float(g), scratch, gx, False)
        return gx.to(ctx.in_dtype), None, None, None, None, None, None


class SSSLoss(nn.Module):
    """Single-scale spectral loss (mss_loss.py:70-121)."""

    def __init__(self, n_fft, sample_rate, alpha=1.0, overlap=0.75, eps=1e-7, type="geomloss"):
        super().__init__()
        self.n_fft = n_fft
        self.alpha = alpha
        self.eps = eps
        self.hop_length = int(n_fft * (1 - overlap))
        self.loss_type = type
        self.sample_rate = sample_rate

    def spec(self, x):
        """Power spectrogram (..., n_fft/2+1, frames), torchaudio.transforms.Spectrogram(n_fft, hop) semantics."""
        x2 = x.reshape(-1, x.shape[-1]).to(torch.float32).contiguous()
        S = native.stft_power(x2, self.n_fft, self.hop_length)
        return S.reshape(*x.shape[:-1], S.shape[-2], S.shape[-1])

    def log_func(self, x):
        return (x + self.eps).log2() - np.log2(self.eps)

    def log_spec(self, x, scale=1.0):
        return self.log_func(clip_spec(self.spec(x), scale))

    def forward(self, x_pred, x_true, freq=None, scale=1.0):
        if self.loss_type not in _MODES:
            raise NotImplementedError(
                f"SSSLoss type '{self.loss_type}': the Sinkhorn variant needs the geomloss package, which is not part "
                "of this path; use 'l1_loss' or 'rmse_loss'")
        if self.loss_type == "rmse_loss" and scale != 1.0:
            raise NotImplementedError("rmse_loss with a clipped spectrum (scale != 1) is not used by the reference's experiments")
        xp = x_pred if x_pred.dim() == 2 else x_pred.reshape(-1, x_pred.shape[-1])
        xt = x_true if x_true.dim() == 2 else x_true.reshape(-1, x_true.shape[-1])
        return _SpectralLoss.apply(xp, xt, self.n_fft, self.hop_length, _MODES[self.loss_type], float(self.alpha),
                                   float(self.eps))


class MSSLoss(nn.Module):
    """Multi-scale spectral loss: the sum of the single-scale losses (mss_loss.py:125-147)."""

    def __init__(self, n_ffts: list, sample_rate, alpha=1.0, overlap=0.75, eps=1e-7, type="geomloss"):
        super().__init__()
        self.n_ffts = n_ffts
        self.losses = nn.ModuleList([SSSLoss(n_fft, sample_rate, alpha, overlap, eps, type) for n_fft in n_ffts])

    def forward(self, x_pred, x_true, freq=None, scale=1.0):
        losses = [loss(x_pred, x_true, freq, scale) for loss in self.losses]
        return sum(losses).sum()
