"""CUDA-graph replay of ``FlowModel.forward`` for fixed-shape batches (evaluation).

A forward pass of the LM flow is ~25 kernel launches; replaying them from one CUDA graph removes the per-launch host work and
lets the host run ahead of the device.  ``GraphedFlowForward`` keeps the inputs in static buffers (tokens, optional tensor
keyword arguments such as ``length`` / ``channel_padding_mask``, the uniform noise of the categorical encoding, drawn outside
the graph so that replays do not reuse frozen Philox offsets) and captures ``model(...)`` - plus the prior log-likelihood when
``log_prior`` is given - once per input signature.  A captured graph holds pointers to tensors DERIVED from the parameters
(built 1x1-convolution matrices, fused / split projection weights), so all graphs are dropped and re-captured when the
parameters change: ``ops.param_fingerprint`` (version counters, storage addresses and the generation counter that every
optimiser step, ``train()`` / ``eval()`` switch and ``ops.invalidate_caches()`` advance).  The numerical-health word of the kernels is not read inside the graph
(``check_nan=False``): read it with ``ops.check_status`` when the results are consumed.
"""
import torch

from ... import ops


class GraphedFlowForward:

    def __init__(self, model, log_prior=None, log_likelihood=None):
        """``log_prior(z, channel_padding_mask | None) -> [B]`` per-sample prior log-likelihood, e.g.
        ``lambda z, pad: ops.logistic_logprob(z, pad=pad)[0]``; None: the third return value is None.
        ``log_likelihood(z, ldj, channel_padding_mask | None) -> [B]`` instead finishes the log-likelihood itself, e.g.
        ``lambda z, ldj, pad: ops.logistic_logprob(z, pad=pad, add=ldj, total=pair)[0]`` (one kernel, which also leaves the
        (sum, count) pair of the step's all-reduce in the static tensor ``pair``)."""
        self.model, self.log_prior, self.log_likelihood = model, log_prior, log_likelihood
        self.graphs = {}
        self.captures = 0
        self._tensors = list(model.parameters()) + list(model.buffers())
        self._fingerprint = None

    def _signature(self, x, kwargs):
        sig = [tuple(x.shape), x.dtype, x.device]
        for k in sorted(kwargs):
            v = kwargs[k]
            sig.append((k, (tuple(v.shape), v.dtype)) if isinstance(v, torch.Tensor) else (k, v))
        return tuple(sig)

    def _run(self, st):
        extra = {} if st["noise"] is None else {"u_noise": st["noise"]}
        z, ldj = self.model(st["x"], check_nan=False, **extra, **st["kwargs"])[:2]
        if self.log_likelihood is not None:
            return z, ldj, self.log_likelihood(z, ldj, st["kwargs"].get("channel_padding_mask"))
        if self.log_prior is None:
            return z, ldj, None
        return z, ldj, ldj + self.log_prior(z, st["kwargs"].get("channel_padding_mask"))

    @torch.no_grad()
    def __call__(self, x, u_noise=None, **kwargs):
        """-> (z, ldj, log_likelihood | None): static output tensors, overwritten by the next call with the same signature."""
        if self.model.training:
            raise RuntimeError("GraphedFlowForward replays the evaluation pass: call model.eval() first")
        fp = ops.param_fingerprint(self._tensors)
        if fp != self._fingerprint:          # parameters changed since the captures: their derived tensors are stale
            self.graphs.clear()
            self._fingerprint = fp
        key = self._signature(x, kwargs)
        st = self.graphs.get(key)
        if st is None:
            first = self.model.flow_layers[0]
            D = getattr(first, "D", None) if x.dtype in (torch.int64, torch.int32) else None
            st = {"x": torch.empty_like(x),
                  "kwargs": {k: (torch.empty_like(v) if isinstance(v, torch.Tensor) else v) for k, v in kwargs.items()},
                  "noise": None if D is None else torch.empty(x.shape + (D,), dtype=torch.float32, device=x.device)}
        st["x"].copy_(x, non_blocking=True)
        for k, v in kwargs.items():
            if isinstance(v, torch.Tensor):
                st["kwargs"][k].copy_(v, non_blocking=True)
        if st["noise"] is not None:
            if u_noise is None:
                st["noise"].uniform_()
            else:
                st["noise"].copy_(u_noise.reshape(st["noise"].shape))
        if "graph" not in st:
            dev = x.device
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                self._run(st)                       # fills host-side caches, sets kernel attributes outside the capture
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):   # other threads (NCCL watchdog) may touch CUDA meanwhile
                st["out"] = self._run(st)
            st["graph"] = g
            self.graphs[key] = st
            self.captures += 1
        st["graph"].replay()
        return st["out"]
