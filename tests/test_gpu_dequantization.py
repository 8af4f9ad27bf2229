"""GPU parity of SigmoidFlow and VariationalDequantization (the remaining encodings of
layers/categorical_encoding) against the reference's golden outputs, against the oracle at the LM size, and of the backward
kernel against autograd through the oracle."""
import pytest
import torch
import torch.nn as nn

from conftest import assert_close, load_golden
from oracle import cnf_oracle as O

pytestmark = pytest.mark.gpu


class ExampleNetwork(nn.Module):
    """The coupling network of the reference's usage example (variational_dequantization.py:118-131)."""

    def __init__(self, c_out, hidden, embed):
        super().__init__()
        self.inp_layer = nn.Linear(1, hidden)
        self.main_net = nn.Sequential(nn.Linear(hidden + embed, hidden), nn.ReLU(), nn.Linear(hidden, c_out))

    def forward(self, x, ext_input, **kwargs):
        return self.main_net(torch.cat([self.inp_layer(x), ext_input], dim=-1))


def _dequant(V, E, H, num_flows):
    from categoricalnf_b200.layers.categorical_encoding import VariationalDequantization
    return VariationalDequantization(flow_config={"num_flows": num_flows, "model_func": lambda c_out: ExampleNetwork(c_out, H, E),
                                                  "block_type": "Linear"}, vocab_size=V, default_embed_layer_dims=E)


def test_sigmoid_flow_golden():
    from categoricalnf_b200.layers.flows import SigmoidFlow
    g = load_golden("dequantization")
    with torch.no_grad():
        ldj0 = g.ldj0.cuda()
        z, ldj = SigmoidFlow()(g.sig_in.cuda(), ldj=ldj0)
        assert_close(z, g.sig_z, rtol=1e-5, atol=1e-7, what="sigmoid z")
        assert_close(ldj, g.sig_ldj, rtol=1e-5, atol=2e-5, what="sigmoid ldj")
        assert torch.equal(ldj0.cpu(), g.ldj0), "the caller's ldj must not be modified (ldj = ldj + ..., :44)"
        _, elem = SigmoidFlow(reverse=True)(g.sig_in.cuda(), reverse=True, sum_ldj=False)
        assert_close(elem, g.sig_elem, rtol=1e-5, atol=1e-6, what="sigmoid element ldj")
        z, ldj = SigmoidFlow()(g.logit_in.cuda(), ldj=ldj0, reverse=True)
        assert_close(z, g.logit_z, rtol=1e-5, atol=1e-5, what="logit z")
        assert_close(ldj, g.logit_ldj, rtol=1e-5, atol=2e-5, what="logit ldj")
        _, elem = SigmoidFlow(reverse=True)(g.logit_in.cuda(), sum_ldj=False)
        assert_close(elem, g.logit_elem, rtol=1e-5, atol=1e-5, what="logit element ldj")


def test_variational_dequantization_golden():
    g = load_golden("dequantization")
    deq = _dequant(g.V, 12, 20, g.num_flows)
    deq.load_state_dict({k[len("sd__"):]: v for k, v in g.items() if k.startswith("sd__")}, strict=True)
    deq = deq.cuda().eval()
    with torch.no_grad():
        z, ldj = deq(g.x.cuda(), reverse=False, u_noise=g.u.cuda())
        assert_close(z, g.z_cont, what="dequantised z")
        assert_close(ldj, g.ldj, rtol=1e-4, atol=2e-4, what="ldj")
        x_rec, _ = deq(g.z_cont.cuda(), reverse=True)
    assert x_rec.dtype == torch.int64 and torch.equal(x_rec.cpu(), g.x_rec)


def test_variational_dequantization_lm_size_vs_oracle():
    """[256, 256] tokens over 51 classes: forward against the oracle, round trip, noise inside [0,1], internal RNG path."""
    from categoricalnf_b200 import ops
    torch.manual_seed(5)
    V, E, H, NF, B, S = 51, 16, 32, 4, 256, 256
    deq = _dequant(V, E, H, NF)
    with torch.no_grad():
        for p in deq.parameters():
            p.add_(0.2 * torch.randn_like(p))
    x = torch.randint(0, V, (B, S))
    u = torch.rand(B, S)
    sd = {k: v.clone() for k, v in deq.state_dict().items()}
    z_ref, ldj_ref = O.variational_dequantization(sd, x, u, NF)
    deq = deq.cuda().eval()
    with torch.no_grad():
        z, ldj = deq(x.cuda(), u_noise=u.cuda())
        assert_close(z, z_ref, what="z")
        assert_close(ldj, ldj_ref, rtol=1e-4, atol=2e-4, what="ldj")
        assert torch.equal(deq(z, reverse=True)[0].cpu(), torch.floor(z_ref).clamp(0, V - 1).long().squeeze(-1))
        z2, ldj2 = deq(x.cuda())                               # internal noise
        noise = z2.squeeze(-1) - x.cuda().float()
        assert (noise >= 0).all() and (noise <= 1).all() and torch.isfinite(ldj2).all()
    ops.check_status(z.device, "dequantization")


@pytest.mark.parametrize("reverse", [False, True])
@pytest.mark.parametrize("sum_ldj", [True, False])
def test_sigmoid_flow_backward_vs_oracle_autograd(reverse, sum_ldj):
    from categoricalnf_b200 import functional as CF
    g = torch.Generator().manual_seed(11)
    z = (torch.rand(6, 9, 2, generator=g) * 0.98 + 0.01) if reverse else torch.randn(6, 9, 2, generator=g) * 3
    ldj0 = torch.randn(6, generator=g)
    gz, gl = torch.randn(6, 9, 2, generator=g), (torch.randn(6, generator=g) if sum_ldj else torch.randn(6, 9, 2, generator=g))
    zr = z.double().requires_grad_(True)
    lr = ldj0.double().requires_grad_(True)
    o, l = O.sigmoid_flow(zr, lr, reverse=reverse, sum_ldj=sum_ldj)
    ((o * gz).sum() + (l * gl).sum()).backward()
    zc, lc = z.cuda().requires_grad_(True), ldj0.cuda().requires_grad_(True)
    o2, l2 = CF.sigmoid_flow(zc, lc, reverse=reverse, sum_ldj=sum_ldj)
    ((o2 * gz.cuda()).sum() + (l2 * gl.cuda()).sum()).backward()
    assert_close(zc.grad, zr.grad, rtol=1e-4, atol=1e-5, what="grad z")
    if sum_ldj:
        assert_close(lc.grad, lr.grad, rtol=1e-6, atol=1e-6, what="grad ldj")


def test_dequantization_training_step():
    """Gradients reach the embedding, ActNorm and coupling-network parameters through the backward kernels and match autograd
    through the oracle."""
    torch.manual_seed(3)
    V, E, H, NF, B, S = 7, 8, 16, 2, 12, 10
    deq = _dequant(V, E, H, NF)
    with torch.no_grad():
        for p in deq.parameters():
            p.add_(0.2 * torch.randn_like(p))
    x, u = torch.randint(0, V, (B, S)), torch.rand(B, S)
    sd = {k: v.clone().double().requires_grad_(v.is_floating_point()) if v.is_floating_point() else v.clone()
          for k, v in deq.state_dict().items()}
    z_ref, ldj_ref = O.variational_dequantization(sd, x, u.double(), NF)
    w = torch.randn(B, S, 1, dtype=torch.float64)
    ((z_ref * w).sum() - ldj_ref.sum()).backward()
    deq = deq.cuda().train()
    z, ldj = deq(x.cuda(), u_noise=u.cuda())
    ((z * w.float().cuda()).sum() - ldj.sum()).backward()
    for name, p in deq.named_parameters():
        ref = sd[name].grad
        assert p.grad is not None, name
        assert_close(p.grad, ref, rtol=2e-3, atol=2e-4, what="grad " + name)
