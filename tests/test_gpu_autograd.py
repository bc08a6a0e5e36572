"""GPU tests (-m gpu) of the training callers' gradients: the hand-written backward passes against
torch autograd through a float64 restatement of the reference's per-time-step recursion (hmm.py:92-101, :203-206;
mixture_gaussian.py:141-214, :312-336).  Tolerance 1e-3 relative (fp32 kernels vs float64 autograd)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hm():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pytorch_hmm_b200 as m
    return m


def _ref_loglik(log_b, log_P, log_p0):
    """log p(o) by the reference's forward recursion (hmm.py:92-101), float64 torch with autograd."""
    la = log_p0 + log_b[:, 0]
    for t in range(1, log_b.shape[1]):
        la = torch.logsumexp(la.unsqueeze(2) + log_P.unsqueeze(0), dim=1) + log_b[:, t]
    return la


@pytest.mark.parametrize("K,T,B", [(3, 12, 2), (5, 40, 3), (12, 130, 4), (17, 33, 2), (32, 20, 1), (40, 25, 3), (130, 18, 2), (600, 9, 2)])
def test_loglik_gradients_vs_float64_autograd(hm, K, T, B):
    from pytorch_hmm_b200.autograd import hmm_log_likelihood
    g = torch.Generator().manual_seed(K * 100 + T)
    log_b = (torch.randn(B, T, K, generator=g) * 2 - 5)
    log_P = torch.log_softmax(torch.randn(K, K, generator=g), -1)
    log_p0 = torch.log_softmax(torch.randn(K, generator=g), -1)
    w = torch.rand(B, generator=g) + 0.5
    # float64 autograd
    a, b, c = (t.double().clone().requires_grad_(True) for t in (log_b, log_P, log_p0))
    ll64 = torch.logsumexp(_ref_loglik(a, b, c), -1)
    (ll64 * w.double()).sum().backward()
    # kernels
    x, y, z = (t.cuda().clone().requires_grad_(True) for t in (log_b, log_P, log_p0))
    ll = hmm_log_likelihood(x, y, z, hm.ops.EMIS_LOG)
    (ll * w.cuda()).sum().backward()
    np.testing.assert_allclose(ll.detach().cpu().numpy(), ll64.detach().numpy(), rtol=1e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), a.grad.numpy(), rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(y.grad.cpu().numpy(), b.grad.numpy(), rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(z.grad.cpu().numpy(), c.grad.numpy(), rtol=1e-3, atol=1e-6)


def test_probability_emissions_gradient(hm):
    from pytorch_hmm_b200.autograd import hmm_log_likelihood
    g = torch.Generator().manual_seed(5)
    K, T, B = 6, 25, 3
    p = torch.rand(B, T, K, generator=g) * 0.9 + 0.05
    log_P = torch.log_softmax(torch.randn(K, K, generator=g), -1)
    log_p0 = torch.log_softmax(torch.randn(K, generator=g), -1)
    a = p.double().clone().requires_grad_(True)
    torch.logsumexp(_ref_loglik(torch.log(a + 1e-8), log_P.double(), log_p0.double()), -1).sum().backward()
    x = p.cuda().clone().requires_grad_(True)
    hmm_log_likelihood(x, log_P.cuda(), log_p0.cuda(), hm.ops.EMIS_PROB_FLOOR).sum().backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), a.grad.numpy(), rtol=1e-3, atol=1e-6)


def test_hmm_layer_training_step_matches_reference_formula(hm):
    """reference tests/test_hmm.py:189-208: one optimiser step changes the transitions; the gradient equals autograd through
    the reference's own loss -logsumexp_k log(exp(log alpha_{T-1,k}) + 1e-8) (hmm.py:203-206) for a short sequence."""
    torch.manual_seed(11)
    K, T, B = 5, 10, 2
    layer = hm.HMMLayer(K).cuda()
    obs = torch.rand(B, T, K).cuda()
    layer.train()
    loss = layer.compute_loss(obs)
    loss.backward()
    gt, gi = layer.log_transition_logits.grad.clone(), layer.log_initial_logits.grad.clone()
    assert torch.isfinite(gt).all() and torch.isfinite(gi).all() and gt.abs().sum() > 0
    # float64 restatement of the reference's loss
    lt = layer.log_transition_logits.detach().double().cpu().requires_grad_(True)
    li = layer.log_initial_logits.detach().double().cpu().requires_grad_(True)
    P = torch.softmax(lt, 1); P = P / P.sum(1, keepdim=True)
    p0 = torch.softmax(li, 0); p0 = p0 / p0.sum()
    la = _ref_loglik(torch.log(torch.sigmoid(obs.double().cpu()) + 1e-8), torch.log(P + 1e-8), torch.log(p0 + 1e-8))
    ref = -torch.logsumexp(torch.log(torch.exp(la) + 1e-8), -1).mean()
    ref.backward()
    np.testing.assert_allclose(loss.item(), ref.item(), rtol=1e-4)
    np.testing.assert_allclose(gt.cpu().numpy(), lt.grad.numpy(), rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(gi.cpu().numpy(), li.grad.numpy(), rtol=2e-3, atol=1e-6)
    before = layer.get_transition_matrix().detach().clone()
    opt = torch.optim.Adam(layer.parameters(), lr=0.01)
    opt.step()
    assert not torch.allclose(before, layer.get_transition_matrix().detach())


def test_mixture_layer_gradient_flow_matches_reference_autograd(hm):
    """reference tests/test_mixture_gaussian.py:159-176: loss = -scores.mean() back-propagates to every parameter.  The
    sub-gradient along the decoded path equals autograd through a torch.max Viterbi (mixture_gaussian.py:312-336)."""
    torch.manual_seed(3)
    K, C, D, B, T = 6, 3, 16, 2, 30
    m = hm.MixtureGaussianHMMLayer(K, D, num_components=C).cuda()
    with torch.no_grad():
        m.means.mul_(4.0)
    x = (m.means.detach()[torch.randint(0, K, (B, T)), torch.randint(0, C, (B, T))] + torch.randn(B, T, D).cuda())
    m.train()
    states, scores = m(x, return_log_probs=True)
    (-scores.mean()).backward()
    for n, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    # the kernel's own score is the same number
    m.eval()
    with torch.no_grad():
        st2, sc2 = m(x, return_log_probs=True)
    assert torch.equal(states, st2)
    np.testing.assert_allclose(scores.detach().cpu().numpy(), sc2.cpu().numpy(), rtol=1e-5)
    # float64 torch.max Viterbi with autograd
    P = {n: p.detach().double().cpu().requires_grad_(True) for n, p in m.named_parameters()}
    xd = x.double().cpu()
    logw = torch.log(torch.clamp(torch.softmax(P["mixture_weights_logits"], -1), min=1e-8))
    diff = xd[:, :, None, None, :] - P["means"][None, None]
    comp = -0.5 * ((diff ** 2 / torch.exp(P["log_vars"])).sum(-1) + P["log_vars"].sum(-1) + D * math.log(2 * math.pi)) + logw
    mx = comp.max(-1, keepdim=True)[0]
    logb = (mx + torch.log(torch.clamp(torch.exp(comp - mx).sum(-1, keepdim=True), min=1e-8))).squeeze(-1)
    ltr = torch.log(torch.clamp(torch.softmax(P["transition_logits"], -1), min=1e-8))
    delta = logb[:, 0] - math.log(K)
    for t in range(1, T):
        delta = (delta.unsqueeze(2) + ltr.unsqueeze(0)).max(1)[0] + logb[:, t]
    ref = delta.max(-1)[0]
    (-ref.mean()).backward()
    np.testing.assert_allclose(scores.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-5)
    for n, p in m.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), P[n].grad.numpy(), rtol=1e-3, atol=1e-6, err_msg=n)


# ---------------------------------------------------------------------------------------------------------
# posteriors (HMMLayer.forward in training mode, supervised cross-entropy) and emission parameters
# ---------------------------------------------------------------------------------------------------------
def _ref_posteriors(log_b, log_P, log_p0):
    """gamma by the reference's log-space forward-backward (hmm.py:92-126), float64 torch with autograd."""
    B, T, K = log_b.shape
    la = [log_p0 + log_b[:, 0]]
    for t in range(1, T):
        la.append(torch.logsumexp(la[-1].unsqueeze(2) + log_P.unsqueeze(0), dim=1) + log_b[:, t])
    lb = [torch.zeros(B, K, dtype=log_b.dtype)]
    for t in range(T - 2, -1, -1):
        lb.insert(0, torch.logsumexp(log_P.unsqueeze(0) + (log_b[:, t + 1] + lb[0]).unsqueeze(1), dim=2))
    lp = torch.stack(la, 1) + torch.stack(lb, 1)
    return torch.exp(lp - torch.logsumexp(lp, -1, keepdim=True))


@pytest.mark.parametrize("K,T,B", [(3, 9, 2), (5, 40, 3), (12, 200, 2), (17, 33, 2), (32, 15, 1), (4, 1, 2)])
@pytest.mark.parametrize("mode", ["log", "prob"])
def test_posterior_gradients_vs_float64_autograd(hm, K, T, B, mode):
    from pytorch_hmm_b200.autograd import hmm_posteriors
    g = torch.Generator().manual_seed(K * 10 + T)
    if mode == "log":
        e = torch.randn(B, T, K, generator=g) * 2 - 5
        emode, to_logb = hm.ops.EMIS_LOG, (lambda v: v)
    else:
        e = torch.rand(B, T, K, generator=g) * 0.9 + 0.05
        emode, to_logb = hm.ops.EMIS_PROB_FLOOR, (lambda v: torch.log(v + 1e-8))
    log_P = torch.log_softmax(torch.randn(K, K, generator=g), -1)
    log_p0 = torch.log_softmax(torch.randn(K, generator=g), -1)
    G = torch.randn(B, T, K, generator=g)
    a, b, c = (t.double().clone().requires_grad_(True) for t in (e, log_P, log_p0))
    gam64 = _ref_posteriors(to_logb(a), b, c)
    (gam64 * G.double()).sum().backward()
    x, y, z = (t.cuda().clone().requires_grad_(True) for t in (e, log_P, log_p0))
    gam, fwd, bwd = hmm_posteriors(x, y, z, emode)
    assert not fwd.requires_grad and not bwd.requires_grad
    (gam * G.cuda()).sum().backward()
    np.testing.assert_allclose(gam.detach().cpu().numpy(), gam64.detach().numpy(), rtol=1e-4, atol=1e-7)
    # 1e-3 relative on the gradients (fp32 kernels vs float64 autograd), absolute floor for entries that are numerically zero
    np.testing.assert_allclose(x.grad.cpu().numpy(), a.grad.numpy(), rtol=1e-3, atol=2e-5)
    if T > 1:                                                        # (a single frame never touches the transition matrix)
        np.testing.assert_allclose(y.grad.cpu().numpy(), b.grad.numpy(), rtol=1e-3, atol=5e-5)
    else:
        assert float(y.grad.abs().max()) == 0.0
    np.testing.assert_allclose(z.grad.cpu().numpy(), c.grad.numpy(), rtol=1e-3, atol=2e-5)


def test_hmm_layer_supervised_loss_has_gradients(hm):
    """HMMLayer.compute_loss(observations, target_alignment) (hmm_layer.py:161-167): cross-entropy on the training-mode posteriors
    back-propagates to the transition / initial logits and the observations; equals float64 autograd through the reference formula."""
    torch.manual_seed(21)
    K, T, B = 5, 12, 3
    layer = hm.HMMLayer(K).cuda().train()
    obs = torch.randn(B, T, K).cuda().requires_grad_(True)
    target = torch.randint(0, K, (B, T)).cuda()
    loss = layer.compute_loss(obs, target)
    loss.backward()
    gt, gi, go = layer.log_transition_logits.grad, layer.log_initial_logits.grad, obs.grad
    assert gt is not None and gi is not None and go is not None and gt.abs().sum() > 0
    lt = layer.log_transition_logits.detach().double().cpu().requires_grad_(True)
    li = layer.log_initial_logits.detach().double().cpu().requires_grad_(True)
    od = obs.detach().double().cpu().requires_grad_(True)
    P = torch.softmax(lt, 1); P = P / P.sum(1, keepdim=True)
    p0 = torch.softmax(li, 0); p0 = p0 / p0.sum()
    gam = _ref_posteriors(torch.log(torch.sigmoid(od) + 1e-8), torch.log(P + 1e-8), torch.log(p0 + 1e-8))
    ref = torch.nn.functional.cross_entropy(gam.reshape(-1, K), target.cpu().reshape(-1))
    ref.backward()
    np.testing.assert_allclose(loss.item(), ref.item(), rtol=1e-4)
    np.testing.assert_allclose(gt.cpu().numpy(), lt.grad.numpy(), rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(gi.cpu().numpy(), li.grad.numpy(), rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(go.cpu().numpy(), od.grad.numpy(), rtol=2e-3, atol=1e-6)


@pytest.mark.parametrize("K,C,D,N", [(12, 4, 80, 300), (6, 1, 16, 200), (5, 3, 33, 150), (10, 1, 80, 257)])
def test_emission_gradients_vs_float64_autograd(hm, K, C, D, N):
    from pytorch_hmm_b200.autograd import gmm_log_probs
    g = torch.Generator().manual_seed(K + C + D)
    means = torch.randn(K, C, D, generator=g)
    log_vars = 0.3 * torch.randn(K, C, D, generator=g)
    logits = torch.randn(K, C, generator=g)
    x = means[torch.randint(0, K, (N,), generator=g), torch.randint(0, C, (N,), generator=g)] + torch.randn(N, D, generator=g)
    G = torch.randn(N, K, generator=g)
    P = {k: v.double().clone().requires_grad_(True) for k, v in (("means", means), ("log_vars", log_vars), ("logits", logits), ("x", x))}
    logw = torch.log(torch.clamp(torch.softmax(P["logits"], -1), min=1e-8))
    diff = P["x"][:, None, None, :] - P["means"][None]
    comp = -0.5 * ((diff ** 2 / torch.exp(P["log_vars"])).sum(-1) + P["log_vars"].sum(-1) + D * math.log(2 * math.pi)) + logw
    ref = torch.logsumexp(comp, -1)
    (ref * G.double()).sum().backward()
    Q = {k: v.cuda().clone().requires_grad_(True) for k, v in (("means", means), ("log_vars", log_vars), ("logits", logits), ("x", x))}
    lw = torch.log(torch.clamp(torch.softmax(Q["logits"], -1), min=1e-8))
    if C == 1:
        out = gmm_log_probs(Q["x"], Q["means"].squeeze(1), Q["log_vars"].squeeze(1), None, 1.0)
    else:
        out = gmm_log_probs(Q["x"], Q["means"], Q["log_vars"], lw, 1.0)
    (out * G.cuda()).sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-4)
    scale = float(P["means"].grad.abs().max())
    np.testing.assert_allclose(Q["means"].grad.cpu().numpy(), P["means"].grad.numpy(), rtol=2e-3, atol=2e-4 * scale)
    np.testing.assert_allclose(Q["log_vars"].grad.cpu().numpy(), P["log_vars"].grad.numpy(), rtol=2e-3, atol=2e-4 * float(P["log_vars"].grad.abs().max()))
    np.testing.assert_allclose(Q["x"].grad.cpu().numpy(), P["x"].grad.numpy(), rtol=2e-3, atol=2e-4 * float(P["x"].grad.abs().max()))
    if C > 1:
        np.testing.assert_allclose(Q["logits"].grad.cpu().numpy(), P["logits"].grad.numpy(), rtol=2e-3, atol=1e-4 * float(P["logits"].grad.abs().max()))


def test_gaussian_hmm_layer_compute_loss_trains_every_parameter(hm):
    """GaussianHMMLayer.compute_loss (hmm_layer.py:342-359) must be differentiable w.r.t. means, log_scales and the transition /
    initial logits.  Short sequence, small D: nothing is floored, so the gradient equals float64 autograd
    through the reference's own formula -mean logsumexp_k log(exp(log alpha_{T-1,k}) + 1e-8)."""
    torch.manual_seed(31)
    K, D, B, T = 4, 3, 2, 6
    m = hm.GaussianHMMLayer(K, D).cuda().train()
    with torch.no_grad():
        m.means.mul_(0.5)
    x = torch.randn(B, T, D).cuda() * 0.5
    loss = m.compute_loss(x)
    loss.backward()
    grads = {n: p.grad for n, p in m.named_parameters()}
    for n, gr in grads.items():
        assert gr is not None and torch.isfinite(gr).all() and gr.abs().sum() > 0, n
    mu = m.means.detach().double().cpu().requires_grad_(True)
    ls = m.log_scales.detach().double().cpu().requires_grad_(True)
    lt = m.hmm_layer.log_transition_logits.detach().double().cpu().requires_grad_(True)
    li = m.hmm_layer.log_initial_logits.detach().double().cpu().requires_grad_(True)
    xd = x.double().cpu()
    diff = xd[:, :, None, :] - mu[None, None]
    logp = -0.5 * (D * math.log(2 * math.pi) + (2 * ls).sum(-1)) - 0.5 * (diff ** 2 / torch.exp(2 * ls)).sum(-1)
    P = torch.softmax(lt, 1); P = P / P.sum(1, keepdim=True)
    p0 = torch.softmax(li, 0); p0 = p0 / p0.sum()
    la = _ref_loglik(torch.log(torch.exp(logp) + 1e-8), torch.log(P + 1e-8), torch.log(p0 + 1e-8))
    ref = -torch.logsumexp(torch.log(torch.exp(la) + 1e-8), -1).mean()
    ref.backward()
    np.testing.assert_allclose(loss.item(), ref.item(), rtol=1e-4)
    np.testing.assert_allclose(grads["means"].cpu().numpy(), mu.grad.numpy(), rtol=5e-3, atol=1e-6)
    np.testing.assert_allclose(grads["log_scales"].cpu().numpy(), ls.grad.numpy(), rtol=5e-3, atol=1e-6)
    np.testing.assert_allclose(grads["hmm_layer.log_transition_logits"].cpu().numpy(), lt.grad.numpy(), rtol=5e-3, atol=1e-6)
    opt = torch.optim.SGD(m.parameters(), lr=0.05)
    before = m.means.detach().clone()
    opt.step()
    assert not torch.equal(before, m.means.detach())


def test_observation_log_probs_are_differentiable(hm):
    """reference tests/test_hsmm.py:286-289 and the mixture layer's analogue: obs_log_probs.sum().backward() reaches the parameters."""
    torch.manual_seed(5)
    h = hm.HSMMLayer(4, 8, max_duration=6).cuda()
    h.get_observation_log_probs(torch.randn(2, 10, 8).cuda()).sum().backward()
    assert h.observation_means.grad is not None and h.observation_means.grad.abs().sum() > 0
    assert h.observation_log_vars.grad is not None
    mm = hm.MixtureGaussianHMMLayer(5, 8, num_components=2).cuda()
    mm.get_observation_log_probs(torch.randn(2, 10, 8).cuda()).sum().backward()
    for n in ("means", "log_vars", "mixture_weights_logits"):
        assert getattr(mm, n).grad is not None and getattr(mm, n).grad.abs().sum() > 0, n
