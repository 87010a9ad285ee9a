"""Oracle (TEST INFRASTRUCTURE): Philox4x32-10 counter-based generator in numpy.

The reference draws from numpy's PCG64 (MCMC/monte_carlo.py:92-95, 153, 161, 215); the throughput
kernels of the build use a counter-based stream instead so that results do not depend on how chains
are sharded: one Philox4x32-10 block per (seed, global chain id, step id = attempts counter),
words = {particle index, u1, u2, accept uniform}.  This file restates the published algorithm
(Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123 known answers in
tests/test_oracle_golden.py) and the mapping of the four words onto the reference's draw order
 so the parity tests can replay the device's stream through the oracle chain.
"""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Counter words and key words as (arrays of) ints below 2^32 -> four uint64 arrays below 2^32."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    m = np.uint64(MASK)
    s32 = np.uint64(32)
    for _ in range(rounds):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        h0, l0 = p0 >> s32, p0 & m
        h1, l1 = p1 >> s32, p1 & m
        c0, c1, c2, c3 = h1 ^ c1 ^ k0, l1, h0 ^ c3 ^ k1, l0
        k0 = (k0 + np.uint64(W0)) & m
        k1 = (k1 + np.uint64(W1)) & m
    return c0, c1, c2, c3


def step_draws(seed, chain_id, step0, steps, n_particles):
    """The draws of `steps` consecutive local moves of one chain: (p [steps] int, u [steps, 3] float64).
    p = floor(word0 * N / 2^32), u_k = word_k / 2^32."""
    sid = np.arange(step0, step0 + steps, dtype=np.uint64)
    cid = np.uint64(chain_id)
    x, y, z, w = philox4x32(sid & np.uint64(MASK), sid >> np.uint64(32), np.full(steps, cid & np.uint64(MASK)),
                            np.full(steps, cid >> np.uint64(32)), seed & MASK, (seed >> 32) & MASK)
    p = ((x * np.uint64(n_particles)) >> np.uint64(32)).astype(np.int64)
    u = np.stack([y, z, w], axis=1).astype(np.float64) / 4294967296.0
    return p, u


class StepRNG:
    """numpy-Generator look-alike over step_draws for oracle.mc_ref.ChainRef: per local step integers(N) returns the
    particle index, random(2) the displacement uniforms, random() the accept uniform (consumed only when drawn)."""

    def __init__(self, seed, chain_id, step0, steps, n_particles):
        self.p, self.u = step_draws(seed, chain_id, step0, steps, n_particles)
        self.s = -1

    def integers(self, n):
        self.s += 1
        return int(self.p[self.s])

    def random(self, k=None):
        if k is None:
            return float(self.u[self.s, 2])
        assert k == 2
        return self.u[self.s, :2].copy()
