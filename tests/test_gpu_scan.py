"""GPU parity tests (-m gpu) of the time-parallel forward-backward (csrc/scan_smallk.cu): same contract as the sequential
sweeps -- posteriors and log-likelihood within 1e-4 relative of the float64 oracle -- plus agreement with the sweeps."""
import numpy as np
import pytest
import torch

from oracle import c_oracle

pytestmark = pytest.mark.gpu
RTOL = 1e-4


@pytest.fixture(scope="module")
def hm():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pytorch_hmm_b200 as m
    return m


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", torch.float32)


def _case(rng, K, T, B, mode, hm):
    P = rng.random((K, K)) ** 3 + 0.02
    P /= P.sum(1, keepdims=True)
    p0 = rng.random(K) + 0.1
    p0 /= p0.sum()
    if mode == "prob":
        e = rng.random((B, T, K)).astype(np.float32)
        e[rng.random((B, T, K)) < 0.2] = 0.0
        logb = np.log(e + np.float32(1e-8)).astype(np.float64)
        emode = hm.ops.EMIS_PROB_FLOOR
    elif mode == "log":
        e = (rng.standard_normal((B, T, K)) * 4.0 - 50.0).astype(np.float32)
        logb = e.astype(np.float64)
        emode = hm.ops.EMIS_LOG
    else:
        e = (rng.standard_normal((B, T, K)) * 15.0 - 100.0).astype(np.float32)
        logb = np.log(np.exp(e - e.max(-1, keepdims=True)).astype(np.float32) + np.float32(1e-8)).astype(np.float64)
        emode = hm.ops.EMIS_LOG_NORM_FLOOR
    return e, logb, emode, (P + 1e-8).astype(np.float32), (p0 + 1e-8).astype(np.float32)


@pytest.mark.parametrize("K,T,B", [(1, 5, 2), (2, 1, 1), (3, 2, 2), (4, 33, 3), (5, 64, 1), (8, 65, 2), (12, 1000, 3), (12, 4097, 2),
                                   (13, 300, 1), (16, 2049, 2), (20, 500, 1), (32, 700, 2)])
@pytest.mark.parametrize("mode", ["prob", "log", "norm_floor"])
def test_scan_forward_backward_vs_float64(hm, K, T, B, mode):
    rng = np.random.default_rng(1900 + K + T)
    e, logb, emode, Pe, p0e = _case(rng, K, T, B, mode, hm)
    la, lb, gam, ll = c_oracle.forward_backward_f64(logb, np.log(Pe.astype(np.float64)), np.log(p0e.astype(np.float64)))
    r = hm.ops.forward_backward(_dev(e), emode, _dev(Pe), _dev(p0e), want=("gamma", "fwd", "bwd", "log_alpha", "log_beta"),
                                method="scan")
    np.testing.assert_allclose(r["gamma"].cpu().numpy(), gam, rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(r["loglik"].cpu().numpy(), ll, rtol=RTOL, atol=1e-4)
    np.testing.assert_allclose(r["log_alpha"].cpu().numpy(), la, rtol=RTOL, atol=1e-3)
    np.testing.assert_allclose(r["log_beta"].cpu().numpy(), lb, rtol=RTOL, atol=1e-3)
    big = la > -80
    np.testing.assert_allclose(r["fwd"].cpu().numpy()[big], np.exp(la)[big], rtol=1e-3)
    # log-likelihood only
    r1 = hm.ops.forward_backward(_dev(e), emode, _dev(Pe), _dev(p0e), want=(), method="scan")
    np.testing.assert_allclose(r1["loglik"].cpu().numpy(), ll, rtol=RTOL, atol=1e-4)


def test_scan_long_sequence_matches_sweeps_and_float64(hm):
    """T = 60 000 at B = 2 (the regime the scan is for): float64 oracle on sequence 0, sweeps on both; and the drop-in class
    picks the scan by itself for this shape."""
    K, T, B = 12, 60000, 2
    rng = np.random.default_rng(77)
    e, logb, emode, Pe, p0e = _case(rng, K, T, B, "norm_floor", hm)
    rs = hm.ops.forward_backward(_dev(e), emode, _dev(Pe), _dev(p0e), want=("gamma",), method="scan")
    rq = hm.ops.forward_backward(_dev(e), emode, _dev(Pe), _dev(p0e), want=("gamma",), method="sweep")
    _, _, gam, ll = c_oracle.forward_backward_f64(logb[:1], np.log(Pe.astype(np.float64)), np.log(p0e.astype(np.float64)))
    np.testing.assert_allclose(rs["gamma"][:1].cpu().numpy(), gam, rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(rs["loglik"][:1].cpu().numpy(), ll, rtol=1e-6)
    np.testing.assert_allclose(rs["gamma"].cpu().numpy(), rq["gamma"].cpu().numpy(), rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(rs["loglik"].cpu().numpy(), rq["loglik"].cpu().numpy(), rtol=1e-6)
    assert hm.ops.use_time_parallel_scan(B, T, K) and not hm.ops.use_time_parallel_scan(256, 2000, 12)
