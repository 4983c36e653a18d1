"""Modal synthesis sharded over the batch axis (BASELINE.json configs[4]).

The reference renders every audio of a batch in one process (oscillator.py:282-310).  The batch rows are independent
given the per-mode damping / frequency vectors, so rank r renders the contiguous slice `batch_slice(B, r, world)` of
the amplitudes with the same kernels (`ds_modal_synth_fwd` / `_bwd`); nothing is exchanged in the forward pass.  In the
backward pass the amplitude gradient is local to the slice, while damping and frequency are shared by the whole batch:
their gradients are summed over ranks with one all-reduce of 2 x mode_num floats (the only collective of the path).
"""
import torch
import torch.distributed as dist


def batch_slice(n_batch: int, rank: int, world: int) -> slice:
    """Contiguous share of rank `rank`: the first n_batch % world ranks get one extra row."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_batch, world)
    lo = rank * base + min(rank, extra)
    return slice(lo, lo + base + (1 if rank < extra else 0))


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allreduce_shared_grads(gdamp: torch.Tensor, gfreq: torch.Tensor):
    """Sum the gradients of the per-mode (batch-shared) parameters over all ranks, in place; one fused buffer."""
    rank, world = _world()
    if world == 1:
        return gdamp, gfreq
    buf = torch.cat([gdamp.reshape(-1), gfreq.reshape(-1)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    k = gdamp.numel()
    gdamp.copy_(buf[:k].reshape(gdamp.shape))
    gfreq.copy_(buf[k:].reshape(gfreq.shape))
    return gdamp, gfreq


class ShardedModalSynth(torch.autograd.Function):
    """y_local = synth(amp_local, damp, freq); backward all-reduces d/d(damp), d/d(freq).  `render` / `render_bwd` are the
    native wrappers (injected so that the plumbing can be tested on CPU ranks with a stand-in)."""

    @staticmethod
    def forward(ctx, amp_local, damp, freq, sample_num, sr, render, render_bwd):
        a32 = amp_local.detach().to(torch.float32).contiguous()
        d32 = damp.detach().to(torch.float32).contiguous()
        f32 = freq.detach().to(torch.float32).contiguous()
        ctx.save_for_backward(a32, d32, f32)
        ctx.sr, ctx.render_bwd = sr, render_bwd
        ctx.dtypes = (amp_local.dtype, damp.dtype, freq.dtype)
        return render(a32, d32, f32, int(sample_num), float(sr))

    @staticmethod
    def backward(ctx, gy):
        a32, d32, f32 = ctx.saved_tensors
        ga, gd, gf = ctx.render_bwd(a32, d32, f32, gy.to(torch.float32).contiguous(), ctx.sr)
        allreduce_shared_grads(gd, gf)
        ta, td, tf = ctx.dtypes
        return ga.to(ta), gd.to(td), gf.to(tf), None, None, None, None


def sharded_modal_synth(amp, damp, freq, sample_num, sr, rank=None, world=None):
    """amp (B, k) -- the FULL batch on every rank (or already the local slice when `rank` is None and the caller
    sharded it); returns the local (B_local, sample_num) audio.  Gradients: amp -> local rows, damp / freq -> summed."""
    from .. import native
    r, w = _world()
    rank = r if rank is None else rank
    world = w if world is None else world
    sl = batch_slice(amp.shape[0], rank, world)
    return ShardedModalSynth.apply(amp[sl], damp, freq, sample_num, sr, native.modal_synth_fwd, native.modal_synth_bwd), sl
