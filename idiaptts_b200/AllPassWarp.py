"""Neural-VTLN all-pass warp with the interface of the reference's
idiaptts/src/neural_networks/pytorch/layers/AllPassWarp.py (forward :148-173, combine_warping_parameters :176-184) and
AllPassWarpLayer.py (forward_fixed_alphas :141-150, forward_sample :125-139, get_alpha :164-177, _normalise/_denormalise
:186-200), backed by the CUDA kernels in csrc/vtln.cu through a torch.autograd.Function.

The reference materialises a [T*B, n, n] warp matrix per call from a float32 polynomial tensor that overflows for
n >= ~35; here the warp is applied by the all-pass recursion itself, valid for any
|alpha| < 1 and n <= 128, with analytic gradients w.r.t. the input and alpha."""
from functools import reduce

import numpy as np
import torch
from torch import nn

from . import ops


class _AllPassWarpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, alpha1d, n, mean, std_dev):
        y = ops.allpass_forward(x2d, alpha1d, n, mean, std_dev)
        ctx.save_for_backward(x2d, alpha1d, mean if mean is not None else torch.empty(0), std_dev if std_dev is not None else torch.empty(0))
        ctx.n = n
        ctx.has_mean = mean is not None
        ctx.has_std = std_dev is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x2d, alpha1d, mean, std_dev = ctx.saved_tensors
        gx, ga = ops.allpass_backward(gy.contiguous(), x2d, alpha1d, ctx.n, mean if ctx.has_mean else None,
                                      std_dev if ctx.has_std else None)
        return gx, ga, None, None, None


def _flat_norm(t, width, device):
    if t is None:
        return None
    t = t.to(device=device, dtype=torch.float32).reshape(-1)
    if t.numel() != width:
        raise ValueError("mean / std_dev have {} entries, features have {}".format(t.numel(), width))
    return t.contiguous()


class AllPassWarp(nn.Module):
    def __init__(self, warp_matrix_size):
        super().__init__()
        self.warp_matrix_size = warp_matrix_size

    def init_hidden(self, batch_size=1):
        return None

    def forward(self, in_tensor, alphas, out_tensor=None, mean=None, std_dev=None):
        """in_tensor [T, B, n * blocks] (or [B, T, ...]); alphas [T, B, 1] or a list of them -> (out, combined_alphas).
        Unlike the reference the input is NOT modified in place."""
        combined = AllPassWarp.combine_warping_parameters(alphas)
        n = self.warp_matrix_size
        width = in_tensor.shape[-1]
        if width % n != 0:
            raise ValueError("feature width {} is not a multiple of warp_matrix_size {}".format(width, n))
        lead = in_tensor.shape[:-1]
        x2d = in_tensor.reshape(-1, width).float().contiguous()
        a1d = combined.expand(*lead, 1).reshape(-1).float().contiguous()
        y = _AllPassWarpFn.apply(x2d, a1d, n, _flat_norm(mean, width, x2d.device), _flat_norm(std_dev, width, x2d.device))
        y = y.reshape(*lead, width)
        if out_tensor is not None:
            out_tensor.copy_(y)
            y = out_tensor
        return y, combined

    @staticmethod
    def combine_warping_parameters(alphas):
        if type(alphas) in [list, tuple]:
            return reduce(AllPassWarp._add_warping_parameters, alphas)
        return alphas

    @staticmethod
    def _add_warping_parameters(alpha_1, alpha_2):
        return (alpha_1 + alpha_2) / (1 + alpha_1 * alpha_2)


class AllPassWarpLayer(nn.Module):
    class Config:
        def __init__(self, alpha_layer_in_dims, alpha_ranges, batch_first, warp_matrix_size, gradient_scaling=None, mean=None,
                     n_frames_per_step=1, std_dev=None, **kwargs):
            if alpha_layer_in_dims and alpha_ranges is not None:
                assert len(alpha_layer_in_dims) == len(alpha_ranges), "Number of alpha_layer_dims has to match alpha_ranges."
            assert warp_matrix_size > 0, "warp_matrix_size must be greater than 0."
            self.alpha_layer_dims = alpha_layer_in_dims
            self.alpha_ranges = alpha_ranges
            self.batch_first = batch_first
            self.warp_matrix_size = warp_matrix_size
            self.gradient_scaling = gradient_scaling
            self.n_frames_per_step = n_frames_per_step

            def conv(v, what):
                if v is None:
                    return None
                if isinstance(v, np.ndarray):
                    v = torch.from_numpy(v)
                elif not isinstance(v, torch.Tensor):
                    raise TypeError(what + " has to be of type numpy.ndarray or torch.Tensor.")
                return v.float()
            self.mean = conv(mean, "mean")
            self.std_dev = conv(std_dev, "std_dev")

        def create_model(self):
            return AllPassWarpLayer(self)

    def __init__(self, config):
        super().__init__()
        self.dim_in = config.alpha_layer_dims
        self.warp_matrix_size = config.warp_matrix_size
        self.n_frames_per_step = config.n_frames_per_step
        self.gradient_scaling = config.gradient_scaling
        self.register_buffer("mean", config.mean)
        self.register_buffer("std_dev", config.std_dev)
        self.batch_first = config.batch_first
        self.batch_dim = 0 if config.batch_first else 1
        self.time_dim = 1 if config.batch_first else 0
        if config.alpha_layer_dims is not None:
            self.alpha_layers = nn.ModuleList([nn.Linear(d, self.n_frames_per_step) for d in config.alpha_layer_dims])
        self.alpha_ranges = config.alpha_ranges
        self.all_pass_warp = AllPassWarp(config.warp_matrix_size)

    def init_hidden(self, batch_size=1):
        return None

    def forward_sample(self, in_tensor, alphas):
        if not torch.cuda.is_available():
            raise RuntimeError("idiaptts_b200 needs a CUDA device; there is no CPU fallback")
        if isinstance(in_tensor, np.ndarray):
            in_tensor = torch.from_numpy(in_tensor)
        dev = torch.device("cuda", torch.cuda.current_device())
        in_tensor = in_tensor.unsqueeze(0 if self.batch_first else 1).float().to(dev)
        if type(alphas) not in [list, tuple]:
            alphas = (alphas,)
        alphas = [torch.from_numpy(a) if isinstance(a, np.ndarray) else a for a in alphas]
        alphas = [a.unsqueeze(0 if self.batch_first else 1).float().to(dev) for a in alphas]
        return self.forward_fixed_alphas(in_tensor, alphas=alphas)

    def forward_fixed_alphas(self, input_, alphas):
        assert alphas is not None, "This forward call requires alphas."
        # de-normalise -> warp -> normalise, fused into the kernel (x * std + mean, (y - mean) / std)
        return self.all_pass_warp(input_, alphas, mean=self.mean, std_dev=self.std_dev)

    def forward(self, inputs, lengths, max_lengths, **kwargs):
        inputs, *alpha_layers_inputs = inputs
        alphas = self.get_alphas(*alpha_layers_inputs)
        return [*self.forward_fixed_alphas(inputs, alphas), *alphas], {"lengths": lengths, "max_lengths": max_lengths}

    def get_alphas(self, *alpha_layer_inputs):
        return [self.get_alpha(alpha_layer_inputs[idx], idx) for idx in range(len(self.alpha_layers))]

    def get_alpha(self, alpha_layers_input, alpha_layer_idx):
        alphas = self.alpha_layers[alpha_layer_idx](alpha_layers_input)
        scaled = torch.tanh(alphas) * self.alpha_ranges[alpha_layer_idx]
        if self.gradient_scaling is not None:
            scaled = _GradScale.apply(scaled, self.gradient_scaling)
        B = scaled.shape[self.batch_dim]
        T = scaled.shape[self.time_dim]
        if self.batch_first:
            return scaled.view(B, T * self.n_frames_per_step, 1)
        return scaled.transpose(0, 1).contiguous().view(B, T * self.n_frames_per_step, 1).transpose(0, 1)

    def set_norm_params(self, mean, std_dev):
        mean = torch.from_numpy(mean) if isinstance(mean, np.ndarray) else mean
        std_dev = torch.from_numpy(std_dev) if isinstance(std_dev, np.ndarray) else std_dev
        dev = next(self.parameters()).device
        self.mean = mean.type(torch.float32).to(dev)
        self.std_dev = std_dev.type(torch.float32).to(dev)


class _GradScale(torch.autograd.Function):
    """GradientScaling of the reference (identity forward, gradient multiplied by a constant)."""
    @staticmethod
    def forward(ctx, x, lambda_):
        ctx.lambda_ = lambda_
        return x.view_as(x)

    @staticmethod
    def backward(ctx, grad):
        return grad * ctx.lambda_, None
