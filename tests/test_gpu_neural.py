"""GPU parity tests (-m gpu) of the time-varying-transition recursions (csrc/recursion_tv.cu) against the golden
fixture of the real reference's NeuralHMM (tests/golden/neural.npz, oracle/make_golden.py::neural) and the C oracle.
Bars: Viterbi states / log_delta BIT-EXACT on identical fp32 inputs; posteriors within max(1e-4, the reference's own fp32 error) of the
float64 oracle; log-likelihood 1e-4 relative."""
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


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("tag", ["tv", "static", "tv12"])
def test_neural_recursion_vs_reference_golden(hm, golden, tag):
    g = golden("neural")
    K = g[f"{tag}_log_obs"].shape[-1]
    rec = hm.NeuralHMMRecursion(K)
    lo, lt, li = _t(g[f"{tag}_log_obs"]), _t(g[f"{tag}_log_trans"]), _t(g[f"{tag}_log_init"])
    states, delta = rec.viterbi_decode(lo, lt, li)
    assert states.dtype == torch.int64
    assert np.array_equal(states.cpu().numpy(), g[f"{tag}_states"])
    assert np.array_equal(delta.cpu().numpy(), g[f"{tag}_log_delta"])
    post, fwd, bwd = rec.forward(lo, lt, li)
    # the reference's own posteriors are fp32 log-space (error grows with T): gate against float64, report the reference's error
    if lt.dim() == 4:
        la, lb, gam, ll = c_oracle.tv_forward_backward_f64(g[f"{tag}_log_obs"], g[f"{tag}_log_trans"], g[f"{tag}_log_init"])
    else:
        la, lb, gam, ll = c_oracle.forward_backward_f64(g[f"{tag}_log_obs"].astype(np.float64), g[f"{tag}_log_trans"].astype(np.float64),
                                                        g[f"{tag}_log_init"].astype(np.float64))
    mask = gam > 1e-6
    eng = float(np.max(np.abs(post.cpu().numpy() - gam)[mask] / gam[mask]))
    ref = float(np.max(np.abs(g[f"{tag}_posterior"] - gam)[mask] / gam[mask]))
    print(f"{tag}: posterior rel err vs float64: engine {eng:.3e}, reference {ref:.3e}")
    assert eng <= max(1e-4, ref)
    np.testing.assert_allclose(post.cpu().numpy(), g[f"{tag}_posterior"], rtol=max(1e-4, 3 * ref), atol=1e-6)
    big = la > -80
    np.testing.assert_allclose(fwd.cpu().numpy()[big], g[f"{tag}_forward"][big], rtol=1e-3)
    bigb = lb > -80
    np.testing.assert_allclose(bwd.cpu().numpy()[bigb], g[f"{tag}_backward"][bigb], rtol=1e-3)
    np.testing.assert_allclose(rec.compute_likelihood(lo, lt, li).cpu().numpy(), g[f"{tag}_likelihood"], rtol=1e-4)
    np.testing.assert_allclose(rec.log_likelihood(lo, lt, li).cpu().numpy(), ll, rtol=1e-4)


@pytest.mark.parametrize("K,T,B", [(1, 4, 2), (2, 1, 3), (3, 2, 1), (4, 33, 5), (7, 65, 3), (12, 500, 4), (16, 129, 2), (17, 40, 2), (24, 30, 2),
                                   (32, 50, 2), (12, 9000, 2)])
def test_tv_kernels_vs_c_oracle(hm, K, T, B):
    rng = np.random.default_rng(6100 + K + T)
    logb = (rng.standard_normal((B, T, K)) * 3.0 - 10.0).astype(np.float32)
    Pt = rng.random((B, T, K, K)).astype(np.float32) ** 2 + 0.01
    Pt /= Pt.sum(-1, keepdims=True)
    logT = np.log(Pt + np.float32(1e-8)).astype(np.float32)
    if K > 2:
        logT[:, :, 0, 1] = logT[:, :, 2, 1]                         # exact ties between predecessors: the lowest index wins
    p0 = np.full(K, 1.0 / K, np.float32)
    logp0 = np.log(p0 + np.float32(1e-8)).astype(np.float32)
    st, dl, psi = c_oracle.tv_viterbi_f32(logb, logT, logp0)
    r = hm.ops.tv_viterbi(_t(logb), _t(logT), _t(logp0), want_psi=True)
    assert np.array_equal(r["delta"].cpu().numpy(), dl)
    assert np.array_equal(r["psi"].cpu().numpy().astype(np.int32), psi)
    assert np.array_equal(r["states"].cpu().numpy(), st)
    assert np.array_equal(r["score"].cpu().numpy(), dl[:, -1].max(-1))
    la, lb, gam, ll = c_oracle.tv_forward_backward_f64(logb, logT, logp0)
    f = hm.ops.tv_forward_backward(_t(logb), torch.exp(_t(logT)), torch.exp(_t(logp0)), want=("gamma", "log_alpha", "log_beta"))
    np.testing.assert_allclose(f["gamma"].cpu().numpy(), gam, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(f["loglik"].cpu().numpy(), ll, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(f["log_alpha"].cpu().numpy(), la, rtol=1e-4, atol=2e-3)
    np.testing.assert_allclose(f["log_beta"].cpu().numpy(), lb, rtol=1e-4, atol=2e-3)
    np.testing.assert_allclose(f["gamma"].sum(-1).cpu().numpy(), 1.0, atol=1e-5)
