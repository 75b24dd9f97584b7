"""The Agent's tanh MLPs on the hand-written tcgen05 GEMM (csrc/tc_gemm.cu).

`TCMlp.apply(x, W0, b0, ..., W4, b4)` computes the reference's
`Linear-Tanh-Linear-Tanh-Linear-Tanh-Linear-Tanh-Linear` stack
(ppo_continuous_action_isaacgym.py:130-152) and its backward pass with:
  forward  4 x  h = tanh(h W^T + b)            one GEMM each, bias+tanh fused in the TMEM epilogue
  dgrad    3 x  dZ_prev = (dZ W) * (1 - h^2)    one GEMM each, tanh' fused in the epilogue
  wgrad    4 x  dW = dZ^T h                     one split-K GEMM each on MN-major operands (no transposes)
Operands are bf16 (activations and a per-call bf16 copy of the fp32 master weights), accumulation
is fp32 in TMEM. The 256 -> {1,2,6} head (forward, and backward fused with tanh'), the bias
gradients (column sums) and the pad/convert of the observations are small CUDA-core kernels.
"""
import torch

from .engine import (EPI_ATOMIC_F32, EPI_BIAS_TANH_BF16, EPI_DTANH_BF16, colsum_bf16, convert_bf16_batch,
                     gather_pad_bf16, gemm_bf16, head_backward, head_forward)

_SM_TARGET = 296  # ~2 CTAs' worth of split-K work per SM for the wgrad grids


def _splits(n_out, k_in, batch):
    tiles = (n_out // 128) * max(1, k_in // (128 if k_in % 128 == 0 else 64))
    # floor, not ceil: with more tiles than resident CTAs (2 per SM) a few CTAs would run a second tile
    # alone and the launch would take two tile-times
    target = _SM_TARGET // tiles
    # Every split adds n_out x k_in fp32 atomics on the same addresses; with few output tiles the atomics,
    # not the operand stream, bound the launch. Measured at batch 131072 (profiles/wgrad_splits_bench.py, us):
    # 256x64: 148 splits 37.3, 49 splits 25.0; 512x256: 37 -> 67.5, 27 -> 60.2; 256x512: 37 -> 67.6, 27 -> 61.9;
    # 512x512: 18 -> 80.0 (best).
    if tiles <= 2:
        target //= 3
    elif tiles <= 8:
        target = target * 3 // 4
    return max(1, min((batch + 63) // 64, target))


class TCMlp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, *params):
        assert len(params) == 10 and x.is_cuda and x.dtype == torch.float32
        ws, bs = params[0::2], params[1::2]
        M, n_in = x.shape
        k0 = (n_in + 63) // 64 * 64
        x16 = gather_pad_bf16(x, None, k0)
        w0 = torch.zeros((ws[0].shape[0], k0), device=x.device, dtype=torch.bfloat16)
        w0[:, :n_in] = ws[0]
        w16 = [w0] + [w.to(torch.bfloat16) for w in ws[1:4]]
        hs = [x16]
        for l in range(4):
            h = torch.empty((M, w16[l].shape[0]), device=x.device, dtype=torch.bfloat16)
            gemm_bf16(hs[-1], w16[l], h, EPI_BIAS_TANH_BF16, bias=bs[l].contiguous())
            hs.append(h)
        out = head_forward(hs[4], ws[4], bs[4])
        ctx.save_for_backward(*hs, *ws)
        ctx.n_in = n_in
        return out

    @staticmethod
    def backward(ctx, dout):
        saved = ctx.saved_tensors
        hs, ws = saved[:5], saved[5:]
        M = dout.shape[0]
        dev = dout.device
        grads = [None] * 10
        dz, grads[8], grads[9] = head_backward(dout, hs[4], ws[4])   # dZ of hidden layer 3, dW4, db4
        for l in (3, 2, 1, 0):
            n_out, k_in = dz.shape[1], hs[l].shape[1]
            dw = torch.zeros((n_out, k_in), device=dev, dtype=torch.float32)
            gemm_bf16(dz, hs[l], dw, EPI_ATOMIC_F32, splits=_splits(n_out, k_in, M), mn_major=True)
            grads[2 * l] = dw[:, :ctx.n_in] if l == 0 else dw
            grads[2 * l + 1] = colsum_bf16(dz)
            if l > 0:
                wt = ws[l].t().contiguous().to(torch.bfloat16)          # [k_in, n_out]
                dz_prev = torch.empty((M, k_in), device=dev, dtype=torch.bfloat16)
                gemm_bf16(dz, wt, dz_prev, EPI_DTANH_BF16, aux=hs[l])
                dz = dz_prev
        return (None, *grads)


def mlp_params(seq):
    """(W0, b0, ..., W4, b4) of an nn.Sequential(Linear, Tanh, ..., Linear)."""
    out = []
    for i in (0, 2, 4, 6, 8):
        out += [seq[i].weight, seq[i].bias]
    return out


def mlp_forward(seq, x):
    return TCMlp.apply(x, *mlp_params(seq))


# ---- explicit (autograd-free) training path used by ppo.train -------------------------------
class MlpWeights:
    """bf16 operand copies of one MLP's fp32 master weights: W_l (K-padded for the first layer) for
    the forward GEMMs and W_l^T for the dgrad GEMMs. `refresh()` rewrites all seven in ONE launch;
    the PPO loop calls it once per rollout and once per minibatch instead of converting per call."""

    def __init__(self, seq):
        self.seq = seq
        self.ws = [seq[i].weight for i in (0, 2, 4, 6)]
        self.bs = [seq[i].bias for i in (0, 2, 4, 6)]
        self.head_w, self.head_b = seq[8].weight, seq[8].bias
        dev = self.ws[0].device
        self.n_in = self.ws[0].shape[1]
        self.k0 = (self.n_in + 63) // 64 * 64
        bf = torch.bfloat16
        self.w16 = [torch.zeros((self.ws[0].shape[0], self.k0), device=dev, dtype=bf)]
        self.w16 += [torch.empty(tuple(w.shape), device=dev, dtype=bf) for w in self.ws[1:]]
        self.wt16 = [None] + [torch.empty((w.shape[1], w.shape[0]), device=dev, dtype=bf) for w in self.ws[1:]]
        self.dw0 = torch.zeros((self.ws[0].shape[0], self.k0), device=dev, dtype=torch.float32)
        self.refresh()

    def refresh(self):
        jobs = [(w.detach(), w16, False) for w, w16 in zip(self.ws, self.w16)]
        jobs += [(w.detach(), wt, True) for w, wt in zip(self.ws[1:], self.wt16[1:])]
        convert_bf16_batch(jobs)


def forward_explicit(mw, x16, out=None):
    """(out [M,n_out] f32, activations) for x16 [M,k0] bf16 (already gathered / padded)."""
    hs = [x16]
    for l in range(4):
        h = torch.empty((x16.shape[0], mw.w16[l].shape[0]), device=x16.device, dtype=torch.bfloat16)
        gemm_bf16(hs[-1], mw.w16[l], h, EPI_BIAS_TANH_BF16, bias=mw.bs[l])
        hs.append(h)
    return head_forward(hs[4], mw.head_w, mw.head_b, out=out), hs


def backward_explicit(mw, hs, dout):
    """Accumulates d loss / d parameters into the parameters' .grad buffers (contiguous f32, e.g.
    views of the flat gradient) given dout = d loss / d out. Same kernels as TCMlp.backward, but the
    split-K wgrad atomics and the head kernel write straight into .grad, and the bias gradients come
    out of the dgrad / head kernels' epilogues (no separate column-sum pass over dZ)."""
    M = dout.shape[0]
    # every bias gradient (column sums of dZ_l) is accumulated by the kernel that PRODUCES dZ_l
    dz = head_backward(dout, hs[4], mw.head_w, dW=mw.head_w.grad, db=mw.head_b.grad, dz_colsum=mw.bs[3].grad)[0]
    for l in (3, 2, 1, 0):
        n_out, k_in = dz.shape[1], hs[l].shape[1]
        if l == 0:
            mw.dw0.zero_()
            gemm_bf16(dz, hs[0], mw.dw0, EPI_ATOMIC_F32, splits=_splits(n_out, k_in, M), mn_major=True)
            mw.ws[0].grad.add_(mw.dw0[:, :mw.n_in])
        else:
            gemm_bf16(dz, hs[l], mw.ws[l].grad, EPI_ATOMIC_F32, splits=_splits(n_out, k_in, M), mn_major=True)
        if l > 0:
            dz_prev = torch.empty((M, k_in), device=dz.device, dtype=torch.bfloat16)
            gemm_bf16(dz, mw.wt16[l], dz_prev, EPI_DTANH_BF16, aux=hs[l], colsum=mw.bs[l - 1].grad)
            dz = dz_prev
