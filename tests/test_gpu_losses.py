"""GPU: the fused loss heads (csrc/loss.cu) against the torch ops of the reference's step
(src/utils/loss.py l1_loss = |a - b|.mean(); SURVEY 8d weighted-sum loss), and the backward-pass
preparation on the side stream against the in-line path."""
import math

import pytest
import torch

import gsplat_b200 as gb
from importlib import import_module
from oracle import splat_oracle as so
from tests import util

pytestmark = pytest.mark.gpu
losses = import_module("mini-3d-gaussian-splatting_b200.losses")


@pytest.mark.parametrize("shape", [(3, 7, 5), (3, 96, 128), (3, 1080, 1920), (1, 1, 1), (3, 33, 17)])
def test_l1_loss_value_and_gradient_match_torch(shape):
    g = torch.Generator().manual_seed(5)
    x = torch.rand(shape, generator=g).cuda().requires_grad_(True)
    t = torch.rand(shape, generator=g).cuda()
    t.view(-1)[::7] = x.detach().view(-1)[::7]                   # exact ties: sign(0) = 0 like torch
    want = (x - t).abs().mean()
    (gw,) = torch.autograd.grad(want, x)
    got = losses.l1_loss(x, t)
    got.backward()
    assert abs(got.item() - float(want)) <= 2e-6 * max(1.0, abs(float(want)))
    assert torch.equal(x.grad, gw)
    # unaligned views take the scalar path
    if x.numel() > 9:
        xs = x.detach().view(-1)[1:-2].clone().requires_grad_(True)
        ts = t.view(-1)[1:-2]
        got = losses.l1_loss(xs, ts)
        got.backward()
        want = (xs.detach() - ts).abs().mean()
        assert abs(got.item() - float(want)) <= 2e-6
        assert torch.equal(xs.grad, torch.sign(xs.detach() - ts) / xs.numel())


def test_l1_loss_is_deterministic_and_leaves_its_workspace_clean():
    g = torch.Generator().manual_seed(6)
    x = torch.rand(3, 1080, 1920, generator=g).cuda()
    t = torch.rand(3, 1080, 1920, generator=g).cuda()
    vals = {losses.l1_loss(x, t).item() for _ in range(5)}
    assert len(vals) == 1
    small = torch.rand(3, 4, 4).cuda()
    assert abs(losses.l1_loss(small, torch.zeros_like(small)).item() - float(small.mean())) < 1e-6


def test_weighted_sum_loss_matches_dot_products_and_returns_the_weights_as_gradients():
    H, W = 96, 128
    w = [t.cuda() for t in so.loss_weights(H, W)]
    g = torch.Generator().manual_seed(7)
    outs = [torch.rand(3, H, W, generator=g).cuda().requires_grad_(True), torch.rand(1, H, W, generator=g).cuda().requires_grad_(True),
            (10 * torch.rand(1, H, W, generator=g)).cuda().requires_grad_(True)]
    coeffs = [1.0, 1.0, 0.1]
    want = sum(c * (wi.double() * o.detach().double()).sum() for c, wi, o in zip(coeffs, w, outs))
    got = losses.weighted_sum_loss(outs, w, coeffs)
    assert abs(got.item() - float(want)) <= 1e-5 * abs(float(want))
    got.backward()
    for c, wi, o in zip(coeffs, w, outs):
        assert torch.equal(o.grad, wi * c if c != 1.0 else wi)
    with pytest.raises(RuntimeError):
        losses.l1_loss(torch.zeros(3), torch.zeros(3))            # CPU tensors: no fallback


def test_fused_loss_through_the_renderer_equals_the_torch_loss():
    s = so.scene_aniso(1500, 3)
    s["scaling"] = s["scaling"] + math.log(3.0)
    cam = so.camera_orbit(1, 7, 128, 96)
    bg = torch.tensor([0.2, 0.1, 0.3])
    H, W = cam.height, cam.width
    tgt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(1)).cuda()
    grads = []
    for fused in (False, True):
        model = util.cuda_model_from_params(s)
        rd = gb.GaussianRenderer()
        out = rd.render(util.cuda_camera(cam), model, gb.RenderSettings(H, W, bg.cuda()))
        loss = losses.l1_loss(out["image"], tgt) if fused else (out["image"] - tgt).abs().mean()
        loss.backward()
        grads.append({k: getattr(model, "_" + k).grad.clone() for k in ("xyz", "scaling", "rotation", "opacity", "features_dc")})
    for k in grads[0]:
        assert util.rel_err(grads[1][k], grads[0][k]) < 1e-5, k


def test_backward_preparation_on_the_side_stream_changes_nothing():
    s = so.scene_aniso(3000, 11)
    s["scaling"] = s["scaling"] + math.log(3.0)
    cam = so.camera_orbit(3, 9, 160, 112)
    bg = torch.tensor([0.1, 0.2, 0.3])
    res = []
    for side in (True, False):
        rd = gb.GaussianRenderer()
        rd.side_stream_prep = side
        c_out, c_grads, _, _, _ = util.cuda_render_with_grads(cam, s, bg, renderer=rd)
        res.append((c_out, {k: v for k, v in c_grads.items() if v is not None}))
    for k in ("image", "alpha", "depth"):
        assert torch.equal(res[0][0][k], res[1][0][k])
    for k in res[0][1]:
        assert util.rel_err(res[0][1][k], res[1][1][k]) < 1e-5, k       # atomics: summation order differs run to run


@pytest.mark.parametrize("bg", [(0.0, 0.0, 0.0), (0.3, 0.1, 0.2)])
def test_image_only_loss_takes_the_lean_backward_and_matches_oracle_autograd(bg):
    """The reference's train step differentiates the image only (optimizer.py:137-139): alpha and depth then carry no
    gradient, gs_raster_bwd gets NULL for them and runs its instantiation without the depth terms.  Gradients against
    oracle autograd of the same L1 loss (a non-zero background keeps the dL/dA term alive)."""
    s = so.scene_aniso(1500, 29)
    s["scaling"] = s["scaling"] + math.log(3.0)
    s["opacity"] = s["opacity"] + 1.0
    cam = so.camera_orbit(4, 9, 144, 96)
    bgt = torch.tensor(bg)
    H, W = cam.height, cam.width
    tgt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(3))
    leaf = {k: s[k].clone().requires_grad_(True) for k in util.PARAM_KEYS}
    o = so.render_from_params(cam, leaf["xyz"], leaf["scaling"], leaf["rotation"], leaf["opacity"], leaf["features_dc"], bgt, H, W)
    (o["image"] - tgt).abs().mean().backward()
    model = util.cuda_model_from_params(s)
    rd = gb.GaussianRenderer()
    out = rd.render(util.cuda_camera(cam), model, gb.RenderSettings(H, W, bgt.cuda()))
    assert util.max_abs(out["image"], o["image"]) < util.IMG_TOL
    losses.l1_loss(out["image"], tgt.cuda()).backward()
    for k in util.PARAM_KEYS:
        assert util.rel_err(getattr(model, "_" + k).grad, leaf[k].grad) < 1e-3, k
