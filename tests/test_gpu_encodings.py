"""GPU parity of the encodings beyond the plain mixture model (SURVEY 8a rows a12 with linear flows, a13) and of
BASELINE config 1 (set modeling: 2 categories, d=4, encoder with 4 x [ExtActNorm, InvConv, affine coupling], batch 256):
drop-in modules loaded from the reference's state dict against its golden outputs, and against the oracle at the
config's size."""
import pytest
import torch

from conftest import assert_close, load_golden
from oracle import cnf_oracle as O

pytestmark = pytest.mark.gpu


def _encoder(V, D, num_flows, hidden, sd=None):
    from categoricalnf_b200.layers.categorical_encoding import LinearCategoricalEncoding
    enc = LinearCategoricalEncoding(num_dimensions=D, flow_config={"num_flows": num_flows, "hidden_layers": 2, "hidden_size": hidden},
                                    vocab_size=V)
    if sd is not None:
        enc.load_state_dict(sd, strict=True)
    return enc.cuda().eval()


def test_linear_flow_encoding_golden():
    g = load_golden("encoding_variants")
    sd = {k[len("sd_flows__"):]: v for k, v in g.items() if k.startswith("sd_flows__")}
    enc = _encoder(g.V, g.D, 4, 32, sd)
    with torch.no_grad():
        z, ldj, _ = enc(g.x.cuda(), reverse=False, channel_padding_mask=g.pad.cuda(), beta=0.8, u_noise=g.flows_u.cuda())
        assert_close(z, g.flows_z, what="z")
        assert_close(ldj, g.flows_ldj, rtol=1e-4, atol=2e-4, what="ldj")
        x_dec = enc(g.flows_z.cuda(), reverse=True, channel_padding_mask=g.pad.cuda())[0]
    assert torch.equal(x_dec.cpu(), g.flows_x_dec)


def test_decoder_linear_golden():
    import numpy as np
    from categoricalnf_b200.layers.categorical_encoding.decoder import DecoderLinear
    g = load_golden("encoding_variants")
    dec = DecoderLinear(num_categories=g.V, embed_dim=g.D, hidden_size=24, num_layers=2,
                        class_prior_log=np.log(np.array([0.4, 0.3, 0.2, 0.1], dtype=np.float32)))
    dec.load_state_dict({k[len("sd_dec__"):]: v for k, v in g.items() if k.startswith("sd_dec__")}, strict=True)
    with torch.no_grad():
        out = dec.cuda()(g.dec_z.cuda())
    assert_close(out, g.dec_log_probs, rtol=1e-4, atol=1e-5, what="decoder log-probs")


def test_set_modeling_config_encoder_vs_oracle():
    """BASELINE config 1 size: batch 256 sets of size 2 over 2 categories, d=4, 4 linear flows in the encoder
    (hidden 128) - forward against the oracle, decode recovers the tokens, training-mode gradients flow."""
    torch.manual_seed(1)
    gen = torch.Generator().manual_seed(1)
    B, S, V, D = 256, 2, 2, 4
    enc = _encoder(V, D, 4, 128)
    with torch.no_grad():
        for p in enc.parameters():                      # fresh modules start at zero scale predictors: perturb
            p.add_(torch.randn(p.shape, generator=gen).cuda() * 0.05)
    sd = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
    x = torch.stack([torch.randperm(S, generator=gen) for _ in range(B)])
    u = torch.rand(B * S, 1, D, generator=gen)
    z_ref, ldj_ref = O.categ_encode_flows(sd, x, u, num_flows=4)
    with torch.no_grad():
        z, ldj, _ = enc(x.cuda(), u_noise=u.cuda())
        assert_close(z, z_ref, what="z")
        assert_close(ldj, ldj_ref, rtol=1e-4, atol=2e-4, what="ldj")
        x_dec = enc(z, reverse=True)[0]
    assert (x_dec.cpu() == O.categ_decode_flows(sd, z_ref, num_flows=4)).float().mean() > 0.995
    enc.train()
    z, ldj, stats = enc(x.cuda(), u_noise=u.cuda())
    (-ldj.mean()).backward()
    grads = [p.grad for p in enc.parameters() if p.requires_grad]
    assert all(gr is not None and torch.isfinite(gr).all() for gr in grads)
    assert "avg_token_prob" in stats
