"""CUDA-graph replay of ``GraphCNF.forward`` (log-likelihood direction, evaluation).

At the per-GPU batch of BASELINE config 4 (64 molecules) the forward pass is ~1000 kernel launches of a few microseconds each:
the host cannot issue them as fast as the GPU retires them.  ``GraphedLogLikelihood`` captures the whole pass once per
(batch shape, padded pair-row counts) into a CUDA graph and replays it for every further batch:

* inputs live in static buffers (tokens, adjacency, lengths, the two pair masks, the uniform noise of the three encodings);
* the compact pair rows of the two Edge-GNN masks (bonds for step 2, all valid pairs for step 3) are padded up to a multiple
  of ``bucket`` rows (``StaticPairContext``), so every tensor shape inside the pass is fixed; counting the rows is the only host
  synchronisation per batch and happens before the replay;
* the noise is drawn by torch's generator outside the graph and handed to the encodings through their ``u_noise`` hooks, so
  replays do not reuse the Philox offsets that a capture would have frozen.

Results are those of ``model(x, adjacency=..., length=..., u_noise=...)`` on the same noise (padding rows carry zeros, land in a
spare slot and receive zero gradients).

``GraphedTrainingStep`` does the same for one training step - forward in training mode, loss, ``loss.backward()`` through the
backward kernels - and leaves the gradients in ``p.grad`` (static tensors that every replay rewrites), ready for the
optimiser of the caller's training loop (general/train.py:148-160 of the reference).
"""
import torch

from ... import ops
from ...layers.networks.graph_layers import StaticPairContext
from .mutils import adjacency2pairs


class GraphedLogLikelihood:

    def __init__(self, model, bucket=2048):
        self.model, self.bucket = model, int(bucket)
        self.graphs = {}        # (rows_bond, rows_all) -> (CUDAGraph, z_out, ldj_out)
        self.ctx = {}           # ("bond" | "all", rows) -> StaticPairContext
        self.shape = None
        self.captures = 0
        self._tensors = list(model.parameters()) + list(model.buffers())
        self._fingerprint = None

    def _check_parameters(self):
        """Evaluation graphs hold pointers to tensors derived from the parameters (built 1x1-conv matrices, fused / split
        projection weights): drop them when the parameters changed (``ops.param_fingerprint``)."""
        fp = ops.param_fingerprint(self._tensors)
        if fp != self._fingerprint:
            self.graphs.clear()
            self._fingerprint = fp

    def _setup(self, x, adjacency, length):
        dev = x.device
        B, N = x.shape
        self.shape = (B, N, dev)
        self.x = torch.zeros(B, N, dtype=torch.int64, device=dev)
        self.adjacency = torch.zeros(B, N, N, dtype=torch.int64, device=dev)
        self.length = torch.zeros(B, dtype=length.dtype, device=dev)
        P = N * (N - 1) // 2
        self.z_edges_disc = torch.zeros(B, P, dtype=torch.int64, device=dev)
        self.mask_all = torch.zeros(B, P, dtype=torch.float32, device=dev)
        self.mask_bond = torch.zeros(B, P, dtype=torch.float32, device=dev)
        m = self.model
        self.u_nodes = torch.empty(B, N, m.node_encoding.D, device=dev)
        self.u_edges = torch.empty(B, P, m.edge_attr_encoding.D, device=dev)
        self.u_virtual = torch.empty(B, P, m.edge_virtual_encoding.D, device=dev)
        self.graphs.clear()
        self.ctx.clear()

    def _context(self, kind, rows, x_indices):
        key = (kind, rows)
        c = self.ctx.get(key)
        if c is None:
            B, N, dev = self.shape
            c = self.ctx[key] = StaticPairContext(x_indices, B, N, rows, dev)
        return c

    def _run(self, x_indices):
        return self.model(self.x, adjacency=self.adjacency, length=self.length, u_noise=self.u_nodes,
                          u_noise_edges=self.u_edges, u_noise_virtual=self.u_virtual,
                          cnf_edge_masks=(self.z_edges_disc, x_indices, self.mask_all, self.mask_bond))

    def _load(self, x, adjacency, length, u_noise, u_noise_edges, u_noise_virtual):
        """Inputs -> static buffers, pair contexts for this batch.  Returns (x_indices, key) with key = None when a mask has
        no valid pair (nothing to pad with: the caller runs the plain pass)."""
        if self.shape != (x.shape[0], x.shape[1], x.device):
            self._setup(x, adjacency, length)
        self.x.copy_(x)
        self.adjacency.copy_(adjacency)
        self.length.copy_(length)
        edge_pairs, x_indices, mask_all = adjacency2pairs(adjacency=self.adjacency, length=self.length)
        self.z_edges_disc.copy_(edge_pairs)
        self.mask_all.copy_(mask_all)
        self.mask_bond.copy_(mask_all * (edge_pairs != 0).to(mask_all.dtype))
        for buf, given in ((self.u_nodes, u_noise), (self.u_edges, u_noise_edges), (self.u_virtual, u_noise_virtual)):
            if given is None:
                buf.uniform_()
            else:
                buf.copy_(given.reshape(buf.shape))
        r_bond, r_all = StaticPairContext.count(self.mask_bond), StaticPairContext.count(self.mask_all)
        if r_bond == 0 or r_all == 0:
            return x_indices, None
        up = lambda r: -(-r // self.bucket) * self.bucket
        key = (up(r_bond), up(r_all))
        c_bond, c_all = self._context("bond", key[0], x_indices), self._context("all", key[1], x_indices)
        c_bond.load(self.mask_bond)
        c_all.load(self.mask_all)
        c_bond.attach(self.mask_bond)
        c_all.attach(self.mask_all)
        return x_indices, key

    @torch.no_grad()
    def __call__(self, x, adjacency, length, u_noise=None, u_noise_edges=None, u_noise_virtual=None):
        """-> (z_nodes [B,N,D], ldj [B]) like ``model(x, adjacency=adjacency, length=length)`` in eval mode."""
        if self.model.training:
            raise RuntimeError("GraphedLogLikelihood replays the evaluation pass: call model.eval() first")
        self._check_parameters()
        x_indices, key = self._load(x, adjacency, length, u_noise, u_noise_edges, u_noise_virtual)
        if key is None:      # nothing to pad with: plain pass
            return self._run(x_indices)
        hit = self.graphs.get(key)
        if hit is None:
            # warm-up on a side stream (fills every host-side cache and cudaFuncSetAttribute outside the capture), then capture
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):
                self._run(x_indices)
            torch.cuda.current_stream(x.device).wait_stream(side)
            torch.cuda.synchronize(x.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):   # other threads (NCCL watchdog) may touch CUDA meanwhile
                z, ldj = self._run(x_indices)
            hit = self.graphs[key] = (g, z, ldj)
            self.captures += 1
        g, z, ldj = hit
        g.replay()
        return z.clone(), ldj.clone()


class GraphedTrainingStep(GraphedLogLikelihood):
    """One training step of GraphCNF replayed from a CUDA graph: ``loss = loss_fn(z, ldj, length)`` (default: the mean
    negative log-likelihood per node, the reference's objective up to its bits-per-dimension scale) and ``loss.backward()``.
    Calling it returns the loss (a 0-d tensor) and leaves the gradient of every parameter in ``p.grad``; the caller then runs
    its optimiser.  Dropout must be 0 (the flows' default) and the data-dependent initialisation done.  No autograd graph
    of an eager pass over the same parameters may be alive at the first call (its gradient accumulators are bound to the
    stream it ran on, and the capture may not touch another stream)."""

    def __init__(self, model, loss_fn=None, bucket=2048):
        super().__init__(model, bucket)
        self.loss_fn = loss_fn or (lambda z, ldj, length: -(ldj / length.to(ldj.dtype)).mean())
        self.params = [p for p in model.parameters() if p.requires_grad]

    def _step(self, x_indices):
        z, ldj = self._run(x_indices)
        loss = self.loss_fn(z, ldj, self.length)
        loss.backward()
        return loss

    def __call__(self, x, adjacency, length, u_noise=None, u_noise_edges=None, u_noise_virtual=None):
        if not self.model.training:
            raise RuntimeError("GraphedTrainingStep replays the training pass: call model.train() first")
        with torch.no_grad():
            x_indices, key = self._load(x, adjacency, length, u_noise, u_noise_edges, u_noise_virtual)
        if key is None:
            for p in self.params:
                p.grad = None
            return self._step(x_indices).detach()
        hit = self.graphs.get(key)
        if hit is None:
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    for p in self.params:
                        p.grad = None
                    self._step(x_indices)
            torch.cuda.current_stream(x.device).wait_stream(side)
            torch.cuda.synchronize(x.device)
            for p in self.params:
                p.grad = None
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):   # other threads (NCCL watchdog) may touch CUDA meanwhile
                loss = self._step(x_indices)
            grads = [p.grad for p in self.params]      # allocated in the graph's pool: every replay rewrites them
            hit = self.graphs[key] = (g, loss, grads)
            self.captures += 1
        g, loss, grads = hit
        g.replay()
        for p, gr in zip(self.params, grads):
            p.grad = gr
        return loss.detach().clone()
