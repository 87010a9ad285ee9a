"""Base distribution of the flow: NF/normflows/Energy/Uniform.py:4-74."""
import torch
import torch.nn as nn


class UniformParticle(nn.Module):
    def __init__(self, n_particles, n_dimension, bound, device="cpu"):
        super().__init__()
        self.n_particles = n_particles
        self.n_dimension = n_dimension
        self.bound = bound
        self.device = device

    def sample(self, n_sample):
        z = torch.empty((n_sample, self.n_particles, self.n_dimension), dtype=torch.float32,
                        device=self.device).uniform_(-self.bound, self.bound)
        return z.reshape(n_sample, self.n_particles * self.n_dimension)

    def forward(self, n_sample):
        return self.sample(n_sample)

    def log_prob(self, z):
        inside = ((z >= -self.bound) & (z <= self.bound)).all(dim=1)
        d = self.n_particles * self.n_dimension
        c = -d * torch.log(torch.tensor(2 * self.bound))
        out = torch.full((z.size(0),), float(c), device=z.device, dtype=z.dtype)
        out[~inside] = -float("inf")
        return out
