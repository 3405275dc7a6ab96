"""GPU parity of the tensor-core decoder (tcgen05.mma kind::tf32 with split operands, csrc/decoder_tc.cuh)
against a plain torch reference of Decoder.mlp (model/decoder.py:58-82) and its input gradient
(utils/tools.py:298-311), evaluated in fp64 on the same fp32 parameters."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _decoder(leaky=False, bias=True, seed=0):
    from clid_slam_b200.config import ncd128
    from clid_slam_b200.model.decoder import Decoder

    torch.manual_seed(seed)
    cfg = ncd128()
    cfg.device = "cuda"
    cfg.mlp_leaky_relu = leaky
    cfg.mlp_bias_on = bias
    return Decoder(cfg, 64, 1, 1)


def _reference(dec, z):
    W0 = dec.layers[0].weight.detach().double()
    b0 = dec.layers[0].bias.detach().double() if dec.layers[0].bias is not None else 0.0
    wo = dec.lout.weight.detach().double()[0]
    bo = dec.lout.bias.detach().double()[0] if dec.lout.bias is not None else 0.0
    slope = 0.01 if dec.use_leaky_relu else 0.0
    pre = z.double() @ W0.T + b0
    d = torch.where(pre > 0, torch.ones_like(pre), torch.full_like(pre, slope))
    out = (pre * d) @ wo + bo
    a = (d * wo) @ W0
    return pre, out, a


@pytest.mark.parametrize("leaky,bias", [(False, True), (True, True), (False, False)])
@pytest.mark.parametrize("n", [1, 127, 128, 4099, 200_000])
def test_decoder_eval_matches_torch(leaky, bias, n):
    from clid_slam_b200.ops.query import decoder_eval

    dec = _decoder(leaky, bias, seed=n % 7)
    gen = torch.Generator(device="cuda").manual_seed(n)
    z = torch.randn(n, 11, device="cuda", generator=gen) * torch.tensor([0.05] * 8 + [0.1] * 3, device="cuda")
    out, a, mask = decoder_eval(dec, z, want_grad=True, want_mask=True)
    pre, out_ref, a_ref = _reference(dec, z)
    scale = out_ref.abs().max().item()
    assert (out.double() - out_ref).abs().max().item() <= 2e-6 * max(scale, 1e-3), "logit"
    # masks: identical wherever the pre-activation is not within fp32 rounding of zero
    bits = torch.arange(32, device="cuda")
    on = torch.cat([(mask[:, 0:1] >> bits) & 1, (mask[:, 1:2] >> bits) & 1], dim=1).bool()
    sure = pre.abs() > 1e-6 * (z.abs().max(dim=1, keepdim=True).values.double() + 1e-3)
    assert torch.equal(on[sure], (pre > 0)[sure]), "activation masks"
    same = (on == (pre > 0)).all(dim=1)
    assert same.float().mean().item() > 0.999
    err = (a.double() - a_ref).abs()[same]
    assert err.max().item() <= 1e-5 * a_ref.abs().max().item(), "input gradient"


def test_decoder_eval_sign_safeguard():
    """Inputs constructed so that many pre-activations are within 1e-7 of zero: the masks must be those of the
    fp32 FMA chain (mlp_l1_pairs), which the split product alone cannot guarantee."""
    from clid_slam_b200.ops.query import decoder_eval

    dec = _decoder(False, True, seed=3)
    W0 = dec.layers[0].weight.detach().double()
    b0 = dec.layers[0].bias.detach().double()
    gen = torch.Generator(device="cuda").manual_seed(5)
    n = 8192
    z = (torch.randn(n, 11, device="cuda", generator=gen) * 0.05).double()
    # move every sample onto the hyperplane of one unit: pre_j(z) = 0 up to fp32 rounding
    j = torch.arange(n, device="cuda") % 64
    w = W0[j]
    z = z - ((z * w).sum(1) + b0[j]).unsqueeze(1) * w / (w * w).sum(1, keepdim=True)
    z = z.float()
    out, a, mask = decoder_eval(dec, z, want_grad=True, want_mask=True)
    bits = torch.arange(32, device="cuda")
    on = torch.cat([(mask[:, 0:1] >> bits) & 1, (mask[:, 1:2] >> bits) & 1], dim=1).bool()
    pre64 = z.double() @ W0.T + b0
    sure = pre64.abs() > 2e-7
    assert torch.equal(on[sure], (pre64 > 0)[sure])
    # the logit is continuous across a flip: still accurate for every sample
    _, out_ref, _ = _reference(dec, z)
    assert (out.double() - out_ref).abs().max().item() <= 2e-6 * max(out_ref.abs().max().item(), 1e-3)
