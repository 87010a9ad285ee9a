"""One hybrid NF-MCMC round of Algorithm 1's testing phase as a unit the GPU can run without the host
(hybrid_NF_MCMC/main_algorithm_1.py:381-395):

    for step in range(BIG_MOVE_INTERVAL): mc_run.particle_displacement()
    accepted = mc_run.nf_big_move(test_configs[...])

for every chain at once, with the proposals of the NEXT round sampled from the flow on a side stream while the current
round runs (they do not depend on the chains: the reference draws its whole pool up front, :340-343).

`HybridRound.step()` issues one round.  With `use_graph=True` the round - the forked sampling pass, the local sweep, the
two log-densities, the fused energy + acceptance kernel, the join - is captured ONCE per proposal buffer into a CUDA
graph and replayed: a round of the N = 32 configuration is ~45 kernel launches, which the Python host enqueues in about
the 2.3 ms the GPU needs to run them, so the eager loop is host-bound and the two streams only overlap once the host
has worked up a backlog; a graph replay is one launch.  Nothing in a round depends on host state: the sweep's and the
acceptance kernel's Philox counters live in device memory (per-chain attempt counts), the base noise comes from torch's
graph-safe generator (the PCG64 emulation keeps its state in device memory as well; only recorded replay streams need
the eager path).
"""
from .. import _lib
import numpy as np
import torch


class HybridRound:
    def __init__(self, eng, model, local_steps, use_graph=True):
        if eng.nf_model is None:
            eng.set_nf_model(model)
        elif eng.nf_model is not model:
            raise ValueError("HybridRound: the engine already carries a different flow (set_nf_model)")
        self.eng, self.model, self.local_steps = eng, model, int(local_steps)
        self.use_graph = bool(use_graph)
        self.device = eng.device
        self.B, self.n = eng.B, eng.num_particles
        self.half = np.float32(eng.half_width)
        self.side = torch.cuda.Stream(device=self.device)
        self.cfg = [torch.empty(self.B, self.n, 2, dtype=torch.float32, device=self.device) for _ in range(2)]
        self.graphs = [None, None]
        self.masks = [None, None]
        self.cur = 0
        self.rounds = 0
        self.launches_per_round = None
        self._sample_into(self.cfg[0])                     # proposals of the first round

    # -- pieces ---------------------------------------------------------------
    def _sample_into(self, out):
        z = self.model.q0(self.B)                          # UniformParticle.sample (Energy/Uniform.py:18-22)
        x = self.model.forward(z)
        torch.add(x.reshape(self.B, self.n, 2), self.half, out=out)   # centred -> MC-box coordinates

    def _body(self, cur):
        """Round on proposal buffer `cur`; fills buffer 1 - cur for the next one.  Returns the accept mask."""
        main = torch.cuda.current_stream(self.device)
        self.side.wait_stream(main)                        # fork
        with torch.cuda.stream(self.side):
            self._sample_into(self.cfg[1 - cur])
        self.eng.particle_displacement(self.local_steps)
        mask = self.eng.nf_big_move(self.cfg[cur])
        main.wait_stream(self.side)                        # join
        return mask

    def _capture(self, cur):
        l0 = _lib.lib().fs_launch_count()
        g = torch.cuda.CUDAGraph()
        pool = self.graphs[1 - cur].pool() if self.graphs[1 - cur] is not None else None
        with torch.cuda.graph(g, pool=pool):
            self.masks[cur] = self._body(cur)
        self.graphs[cur] = g
        self.launches_per_round = int(_lib.lib().fs_launch_count() - l0)   # kernels of the library inside one replay

    # -- one round ------------------------------------------------------------
    def step(self):
        """Issues one round (asynchronously) and returns its uint8 accept mask [B] (with graphs: a buffer that the
        second-next step overwrites).  The first round on each of the two proposal buffers runs eagerly (every workspace
        of this shape gets allocated) and is then captured; from the third round on a step is one graph launch."""
        cur = self.cur
        if self.use_graph and self.graphs[cur] is not None:
            self.graphs[cur].replay()
            mask = self.masks[cur]
        else:
            mask = self._body(cur)
            if self.use_graph:
                torch.cuda.synchronize(self.device)
                self._capture(cur)
        self.cur = 1 - cur
        self.rounds += 1
        return mask
