"""NormalizingFlow container.  Mirrors NF/normflows/core.py:10-230 including the
fork's edits: sample() returns z only, forward_kld omits the
base log-probability, reverse_kld returns (loss, z) and calls p._energy.

In eval mode the layer loops collapse into ONE C-ABI call per pass
(fs_flow_inverse / fs_flow_forward over all K layers); train mode loops over
the layers' autograd paths like the reference.
"""
import torch
import torch.nn as nn

from .flows import CircularCoupledRationalQuadraticSpline


class NormalizingFlow(nn.Module):
    def __init__(self, q0, flows, p=None):
        super().__init__()
        self.q0 = q0
        self.flows = nn.ModuleList(flows)
        self.p = p
        self._pack = None
        # "auto": tcgen05 tensor cores (TF32 operands, FP32 accumulation; log-density within ~5e-6 of float64)
        # when the flow shape has the tensor path, else the FP32 CUDA-core path | "fp32" | "tf32"
        self.precision = "auto"
        # "auto" | "prefer" | "never": whether eval-mode passes launch all K layers at once (FlowPack.set_layer_parallel);
        # "prefer" when nothing else runs beside the flow's passes (the Algorithm-2 driver sets it)
        self.layer_parallel = "auto"

    # -- packing ----------------------------------------------------------
    def _fusable(self):
        return (not self.training and len(self.flows) > 0
                and all(isinstance(f, CircularCoupledRationalQuadraticSpline) for f in self.flows))

    def _cuda_pack(self):
        from ._pack import FlowPack
        layers = list(self.flows)
        if self._pack is None or not self._pack.same_shape(layers):
            self._pack = FlowPack(layers)
        elif not self._pack.matches(layers):
            # parameters changed (optimizer step, load_state_dict): refresh the packed weights on the device
            # (fs_flow_update) instead of packing on the host again - Algorithm 2 does this every cycle
            self._pack.update(layers)
        self._pack.precision = self.precision
        self._pack.set_layer_parallel(self.layer_parallel)
        return self._pack

    def repack(self):
        """Mark the packed inference weights stale (call after editing parameters in a way that does not bump tensor
        versions); the next eval-mode call refreshes them on the device."""
        if self._pack is not None:
            self._pack._sig = None
        for f in self.flows:
            if getattr(f, "_pack", None) is not None:
                f._pack._sig = None

    def train(self, mode=True):
        # the pack survives train(): the next eval-mode call sees the bumped tensor versions and refreshes it in place
        return super().train(mode)

    def load_state_dict(self, *a, **k):
        return super().load_state_dict(*a, **k)    # in-place copies bump the tensor versions -> refreshed on next use

    def _apply(self, fn, *a, **k):
        self.repack()
        return super()._apply(fn, *a, **k)

    # -- reference API ----------------------------------------------------
    def forward(self, z):
        if self._fusable():
            return self._cuda_pack().forward(z, want_logdet=False)[0]
        for flow in self.flows:
            z, _ = flow(z)
        return z

    def forward_and_log_det(self, z):
        if self._fusable():
            return self._cuda_pack().forward(z, want_logdet=True)
        log_det = torch.zeros(len(z), device=z.device)
        for flow in self.flows:
            z, ld = flow(z)
            log_det = log_det + ld
        return z, log_det

    def inverse(self, x):
        if self._fusable():
            return self._cuda_pack().inverse(x)[0]
        for i in range(len(self.flows) - 1, -1, -1):
            x, _ = self.flows[i].inverse(x)
        return x

    def inverse_and_log_det(self, x):
        if self._fusable():
            z, ld, _ = self._cuda_pack().inverse(x)
            return z, ld
        log_det = torch.zeros(len(x), device=x.device)
        for i in range(len(self.flows) - 1, -1, -1):
            x, ld = self.flows[i].inverse(x)
            log_det = log_det + ld
        return x, log_det

    def forward_kld(self, x):
        log_q = torch.zeros(len(x), device=x.device)
        z = x
        for i in range(len(self.flows) - 1, -1, -1):
            z, log_det = self.flows[i].inverse(z)
            log_q = log_q + log_det
        return -torch.mean(log_q)

    def reverse_kld(self, num_samples=1, beta=1.0, score_fn=True):
        z = self.q0(num_samples).to(next(self.parameters()).device)
        log_q = torch.zeros(len(z), device=z.device)
        for flow in self.flows:
            z, log_det = flow(z)
            log_q = log_q - log_det
        energy = self.p._energy(z)
        return torch.mean(energy) + torch.mean(log_q), z

    def sample(self, num_samples=1):
        z = self.q0(num_samples)
        return self.forward(z)

    def sample_with_log_prob(self, num_samples=1):
        """(x, log q(x)) from ONE sampling pass: x = f(z), log q(x) = log q0(z) - log|det J_f(z)| - what upstream
        normflows' sample() returns (NF/normflows/core.py:178-196 forms log_q the same way; the fork's sample() keeps
        x only).  MonteCarlo.nf_big_move can take it as the proposal's log-density instead of
        running the proposal through the inverse pass again (logq_new=, MCMC/batched.py); the two values differ by the
        round-off of inverting the flow (tests/test_gpu_flow.py::test_sampling_pass_log_prob_equals_inverse_pass)."""
        z = self.q0(num_samples)
        x, log_det = self.forward_and_log_det(z)
        return x, self.q0.log_prob(z) - log_det

    def log_prob(self, x):
        if self._fusable() and hasattr(self.q0, "bound"):
            return self._cuda_pack().inverse(x, want_logq=True)[2]
        z, log_q = self.inverse_and_log_det(x)
        return log_q + self.q0.log_prob(z)

    def save(self, path):
        torch.save(self.state_dict(), path)

    def load(self, path):
        self.load_state_dict(torch.load(path))
