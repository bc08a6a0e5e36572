"""GPU tests (-m gpu) for the Baum-Welch E-step kernels against the float64 oracle (oracle/hmm_oracle.c::orc_bw_stats_f64).
A9 has no reference implementation (formulas only, docs/01_hmm_theory.md:196-227): parity is against the restated
formulas plus the EM identities (monotone log-likelihood)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hm():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pytorch_hmm_b200 as m
    return m


def _problem(seed, B, T, K, C, D):
    rng = np.random.default_rng(seed)
    means = rng.standard_normal((K, C, D)) * 1.5
    var = np.exp(0.3 * rng.standard_normal((K, C, D)))
    w = rng.dirichlet(np.ones(C) * 3, size=K); P = rng.dirichlet(np.ones(K) * 2, size=K); p0 = rng.dirichlet(np.ones(K) * 2)
    s = rng.integers(0, K, (B, T)); c = rng.integers(0, C, (B, T))
    x = (means[s, c] + np.sqrt(var[s, c]) * rng.standard_normal((B, T, D))).astype(np.float32)
    return x, means, var, w, P, p0


# (the second row of shapes: K a multiple of 4 takes the bulk-staged xi kernel -- tasks of 32 frames, so T around its multiples,
# batches that give a warp several tasks, and every lanes-per-frame variant)
@pytest.mark.parametrize("B,T,K,C,D", [(3, 40, 4, 2, 8), (2, 257, 12, 4, 80), (5, 1, 3, 1, 5), (2, 65, 7, 3, 33),
                                       (1, 2, 8, 1, 4), (40, 34, 16, 1, 4), (3, 97, 20, 2, 8), (2, 33, 24, 1, 8), (2, 66, 28, 1, 4),
                                       (1, 129, 32, 1, 4), (300, 130, 12, 2, 8)])
def test_e_step_statistics_vs_float64_oracle(hm, B, T, K, C, D):
    from pytorch_hmm_b200 import baum_welch as bw
    x, means, var, w, P, p0 = _problem(B * 7 + K, B, T, K, C, D)
    f32 = lambda a: torch.from_numpy(np.asarray(a, np.float32))
    params = bw.GMMHMMParams(f32(P), f32(p0), f32(w), f32(means), f32(var))
    tr = bw.BaumWelch(params)
    ll = tr.e_step(torch.from_numpy(x))
    stats = tr.stats.cpu().numpy()
    # oracle on the same fp32-rounded parameters
    m32, v32, w32, P32, p32 = (np.asarray(a, np.float32).astype(np.float64) for a in (means, var, w, P, p0))
    comp = np.log(w32)[None, None] - 0.5 * (((x[:, :, None, None, :].astype(np.float64) - m32[None, None]) ** 2 / v32[None, None]).sum(-1)
                                            + np.log(v32).sum(-1)[None, None] + D * np.log(2 * np.pi))
    ref = c_oracle.bw_stats_f64(x, comp, np.log(P32), np.log(p32))
    sl = bw.stats_slices(K, C, D)
    np.testing.assert_allclose(ll.double().sum().item(), ref["loglik"], rtol=1e-5)
    np.testing.assert_allclose(stats[sl["extra"]][0], ref["loglik"], rtol=1e-5)
    assert (tr._n_frames, tr._n_seqs) == (B * T, B)              # written into the statistics vector by m_step()
    # 1e-4 relative on every statistic (atol for entries that are numerically zero)
    np.testing.assert_allclose(stats[sl["gamma1"]], ref["gamma1"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(stats[sl["xi"]].reshape(K, K), ref["xi"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(stats[sl["occ"]].reshape(K, C), ref["occ"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(stats[sl["sx"]].reshape(K, C, D), ref["sx"], rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(stats[sl["sxx"]].reshape(K, C, D), ref["sxx"], rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(stats[sl["occ"]].sum(), B * T, rtol=1e-5)
    if T > 1:
        np.testing.assert_allclose(stats[sl["xi"]].sum(), B * (T - 1), rtol=1e-5)


def test_e_step_long_utterances_use_the_sweep_workspace(hm):
    """T >= 8192 at small batch is where ops.forward_backward(method="auto") switches to the time-parallel scan, which does not
    leave the scaled alpha / beta in the caller's workspace: the E-step must pin the sweep.  xi and the
    occupancies against the float64 oracle."""
    from pytorch_hmm_b200 import baum_welch as bw
    B, T, K, C, D = 2, 8300, 5, 2, 8
    assert hm.ops.use_time_parallel_scan(B, T, K)
    x, means, var, w, P, p0 = _problem(41, B, T, K, C, D)
    f32 = lambda a: torch.from_numpy(np.asarray(a, np.float32))
    tr = bw.BaumWelch(bw.GMMHMMParams(f32(P), f32(p0), f32(w), f32(means), f32(var)))
    ll = tr.e_step(torch.from_numpy(x))
    stats = tr.stats.cpu().numpy()
    m32, v32, w32, P32, p32 = (np.asarray(a, np.float32).astype(np.float64) for a in (means, var, w, P, p0))
    comp = np.log(w32)[None, None] - 0.5 * (((x[:, :, None, None, :].astype(np.float64) - m32[None, None]) ** 2 / v32[None, None]).sum(-1)
                                            + np.log(v32).sum(-1)[None, None] + D * np.log(2 * np.pi))
    ref = c_oracle.bw_stats_f64(x, comp, np.log(P32), np.log(p32))
    sl = bw.stats_slices(K, C, D)
    assert np.isfinite(stats).all()
    np.testing.assert_allclose(ll.double().sum().item(), ref["loglik"], rtol=1e-5)
    np.testing.assert_allclose(stats[sl["xi"]].reshape(K, K), ref["xi"], rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(stats[sl["occ"]].reshape(K, C), ref["occ"], rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(stats[sl["xi"]].sum(), B * (T - 1), rtol=1e-5)


def test_em_increases_likelihood_and_recovers_structure(hm):
    from pytorch_hmm_b200 import baum_welch as bw
    x, means, var, w, P, p0 = _problem(3, 16, 200, 4, 2, 6)
    rng = np.random.default_rng(99)
    f32 = lambda a: torch.from_numpy(np.asarray(a, np.float32))
    start = bw.GMMHMMParams(f32(np.full((4, 4), 0.25)), f32(np.full(4, 0.25)), f32(np.full((4, 2), 0.5)),
                            f32(means + 0.7 * rng.standard_normal(means.shape)), f32(np.ones_like(var)))
    tr = bw.BaumWelch(start)
    hist = tr.fit([torch.from_numpy(x[:8]), torch.from_numpy(x[8:])], n_iter=8)
    assert all(b >= a - 1e-4 for a, b in zip(hist, hist[1:])), hist           # EM never decreases the likelihood
    assert hist[-1] > hist[0] + 0.05


def test_from_layer(hm):
    from pytorch_hmm_b200 import baum_welch as bw
    layer = hm.MixtureGaussianHMMLayer(5, 8, num_components=2).cuda()
    tr = bw.BaumWelch.from_layer(layer)
    ll = tr.e_step(torch.randn(3, 30, 8))
    assert ll.shape == (3,) and torch.isfinite(ll).all()
    assert np.isfinite(tr.m_step())
