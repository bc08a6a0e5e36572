"""GPU parity tests (-m gpu) of the large-K (32 < K <= 512) cluster kernels (csrc/recursion_largek.cu), through the C ABI.

Same bars as the small-K path: Viterbi states / delta / score BIT-EXACT given identical fp32 log-emissions;
posteriors and log-likelihoods within 1e-4 relative of the float64 oracle; and the golden fixture of the real reference
(tests/golden/largek.npz: K = 64 skip-left-to-right, K = 512 ergodic, BASELINE config 5 family)."""
import math

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


def _dev(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype)


def _exchange_ok(ws):
    """The sweep kernels raise a flag in the last 256 bytes of the workspace if a DSMEM exchange wait ever timed out."""
    return int(ws[-256:].view(torch.int32)[0].item()) == 0


@pytest.mark.parametrize("K,T,B", [(33, 7, 1), (40, 50, 3), (64, 64, 4), (65, 30, 5), (100, 45, 2), (128, 33, 9),
                                   (129, 20, 4), (200, 25, 3), (256, 40, 2), (257, 12, 1), (384, 18, 6), (500, 16, 2),
                                   (512, 30, 5), (512, 1, 2), (48, 2, 1), (512, 10, 130), (64, 20, 150),
                                   (513, 12, 3), (600, 25, 9), (1024, 9, 2), (2048, 4, 1)])   # > 512: one launch per frame (recursion_xlk.cu)
def test_largek_viterbi_bit_exact(hm, K, T, B):
    rng = np.random.default_rng(7000 + K + T)
    logb = (rng.standard_normal((B, T, K)) * 3.0).astype(np.float32)
    logb[rng.random((B, T, K)) < 0.15] = np.float32(math.log(1e-8))      # floor-induced exact ties
    P = rng.random((K, K)).astype(np.float32) + 0.01
    logP = np.log(P / P.sum(1, keepdims=True)).astype(np.float32)
    logP[0, 1] = logP[0, 2]
    logP[:, 5] = logP[:, 4]                                               # whole columns tie: lowest predecessor wins
    logp0 = np.log(np.full(K, 1.0 / K)).astype(np.float32)
    st, delta, psi, score = c_oracle.viterbi_f32(logb, logP, logp0)
    ws = hm.ops.viterbi_workspace(B, T, K, "cuda")
    r = hm.ops.viterbi(_dev(logb), hm.ops.EMIS_LOG, _dev(logP), _dev(logp0), want_delta=True, want_psi=True, workspace=ws)
    torch.cuda.synchronize()
    assert K > 512 or _exchange_ok(ws)
    assert np.array_equal(r["delta"].cpu().numpy(), delta)
    # packed backpointers: uint8 up to K = 256, uint16 above (int16 storage; K <= 512 so no sign issue)
    assert r["psi"].dtype == (torch.uint8 if K <= 256 else torch.int16)
    assert np.array_equal(r["psi"].cpu().numpy().astype(np.int32), psi)
    assert np.array_equal(r["states"].cpu().numpy(), st)
    assert np.array_equal(r["score"].cpu().numpy(), score)
    # without the delta output the trellis lives in the workspace; same path
    r2 = hm.ops.viterbi(_dev(logb), hm.ops.EMIS_LOG, _dev(logP), _dev(logp0), want_delta=False)
    assert np.array_equal(r2["states"].cpu().numpy(), st)


@pytest.mark.parametrize("K,T,B", [(33, 9, 2), (50, 60, 3), (64, 100, 4), (96, 40, 5), (130, 50, 2), (256, 64, 3),
                                   (300, 30, 1), (512, 80, 6), (512, 1, 1), (512, 8, 130), (520, 40, 3), (1000, 16, 10)])
@pytest.mark.parametrize("mode", ["prob", "log", "norm_floor"])
def test_largek_forward_backward_vs_float64(hm, K, T, B, mode):
    rng = np.random.default_rng(8000 + K + T)
    P = rng.random((K, K)) ** 4 + 1e-3                                    # uneven rows
    P /= P.sum(1, keepdims=True)
    p0 = rng.random(K) + 0.1
    p0 /= p0.sum()
    if mode == "prob":
        e = torch.softmax(torch.from_numpy(2.0 * rng.standard_normal((B, T, K))), -1).numpy().astype(np.float32)
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
    Pe, p0e = (P + 1e-8).astype(np.float32), (p0 + 1e-8).astype(np.float32)
    la, lb, gam, ll = c_oracle.forward_backward_f64(logb, np.log(Pe.astype(np.float64)), np.log(p0e.astype(np.float64)))
    ws = hm.ops.fb_workspace(B, T, K, "cuda")
    r = hm.ops.forward_backward(_dev(e), emode, _dev(Pe), _dev(p0e), want=("gamma", "fwd", "bwd", "log_alpha", "log_beta"),
                                workspace=ws)
    torch.cuda.synchronize()
    assert K > 512 or _exchange_ok(ws)
    # 1e-4 relative on posteriors (atol 1e-7 for numerically-zero entries) and on the log-likelihood
    np.testing.assert_allclose(r["gamma"].cpu().numpy(), gam, rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(r["loglik"].cpu().numpy(), ll, rtol=RTOL, atol=1e-4)
    np.testing.assert_allclose(r["log_alpha"].cpu().numpy(), la, rtol=RTOL, atol=1e-3)
    np.testing.assert_allclose(r["log_beta"].cpu().numpy(), lb, rtol=RTOL, atol=1e-3)
    np.testing.assert_allclose(r["gamma"].sum(-1).cpu().numpy(), 1.0, atol=1e-5)
    big = la > -80
    np.testing.assert_allclose(r["fwd"].cpu().numpy()[big], np.exp(la)[big], rtol=1e-3)
    # log-likelihood only (forward sweep alone)
    r1 = hm.ops.forward_backward(_dev(e), emode, _dev(Pe), _dev(p0e), want=())
    np.testing.assert_allclose(r1["loglik"].cpu().numpy(), ll, rtol=RTOL, atol=1e-4)


@pytest.mark.parametrize("tag", ["k64", "k512"])
def test_largek_drop_in_vs_reference_golden(hm, golden, tag):
    g = golden("largek")
    hmm = hm.HMMPyTorch(torch.from_numpy(g[f"{tag}_P"]), None, device="cuda")
    obs = torch.from_numpy(g[f"{tag}_obs"]).cuda()
    post, fwd, bwd = hmm.forward_backward(obs)
    np.testing.assert_allclose(post.cpu().numpy(), g[f"{tag}_posterior"], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(fwd.cpu().numpy(), g[f"{tag}_forward"], rtol=RTOL, atol=1e-37)
    np.testing.assert_allclose(bwd.cpu().numpy(), g[f"{tag}_backward"], rtol=RTOL, atol=1e-37)
    np.testing.assert_allclose(hmm.compute_likelihood(obs).cpu().numpy(), g[f"{tag}_likelihood"], rtol=RTOL)
    # bit-exact once the reference's own fp32 log-observations are fed (stored in the fixture: ATen's vectorised CPU log
    # is not guaranteed to round identically on every host CPU)
    log_obs = torch.from_numpy(g[f"{tag}_log_obs"])
    r = hm.ops.viterbi(log_obs.cuda(), hm.ops.EMIS_LOG, _dev(g[f"{tag}_log_P"]), _dev(g[f"{tag}_log_p0"]), want_delta=True)
    assert np.array_equal(r["states"].cpu().numpy(), g[f"{tag}_states"])
    np.testing.assert_array_equal(r["delta"].cpu().numpy(), g[f"{tag}_log_delta"])
    # through the class (GPU logf may differ from ATen's in the last bit): delta within 1e-5, path equal or a near-tie
    states, delta = hmm.viterbi_decode(obs)
    np.testing.assert_allclose(delta.cpu().numpy(), g[f"{tag}_log_delta"], rtol=1e-5, atol=1e-5)
    ours, ref = states.cpu().numpy(), g[f"{tag}_states"]
    d = g[f"{tag}_log_delta"]
    for b in range(ours.shape[0]):
        if not np.array_equal(ours[b], ref[b]):
            assert abs(d[b, -1, ours[b, -1]] - d[b, -1, ref[b, -1]]) < 1e-4 * max(1.0, abs(d[b, -1].max()))


def test_largek_config5_slice_properties(hm):
    """BASELINE config 5 shape family at a size the float64 oracle finishes in seconds (K = 512 ergodic, B = 8, T = 100)
    plus size-independent properties: posteriors sum to 1, Viterbi score = delta at the decoded last state,
    every decoded transition attains the max of its backpointer equation."""
    K, B, T = 512, 8, 100
    P = hm.create_transition_matrix(K, "ergodic")
    g = torch.Generator().manual_seed(5001)
    obs = torch.softmax(torch.randn(B, T, K, generator=g), dim=-1)
    hmm = hm.HMMPyTorch(P, None, device="cuda")
    post, _, _ = hmm.forward_backward(obs.cuda())
    log_obs = torch.log(obs + 1e-8)
    _, _, gam, ll = c_oracle.forward_backward_f64(log_obs.numpy().astype(np.float64), hmm.log_P.cpu().numpy().astype(np.float64),
                                                  hmm.log_p0.cpu().numpy().astype(np.float64))
    np.testing.assert_allclose(post.cpu().numpy(), gam, rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(hmm.log_likelihood(obs.cuda()).cpu().numpy(), ll, rtol=RTOL)
    np.testing.assert_allclose(post.sum(-1).cpu().numpy(), 1.0, atol=1e-5)
    r = hm.ops.viterbi(log_obs.cuda(), hm.ops.EMIS_LOG, hmm.log_P.cuda(), hmm.log_p0.cuda(), want_delta=True)
    st, dl, _, sc = c_oracle.viterbi_f32(log_obs.numpy(), hmm.log_P.cpu().numpy(), hmm.log_p0.cpu().numpy())
    assert np.array_equal(r["states"].cpu().numpy(), st)
    assert np.array_equal(r["delta"].cpu().numpy(), dl)
    assert np.array_equal(r["score"].cpu().numpy(), sc)


# (K, T, B): B <= 90 at K = 512 (8-CTA clusters, at most 15 co-resident) runs 3 sequences per group, larger batches 4
@pytest.mark.parametrize("K,T,B", [(512, 24, 13), (200, 30, 7), (64, 17, 1), (512, 12, 97), (512, 9, 200), (384, 11, 131)])
def test_largek_both_group_sizes(hm, K, T, B):
    """The sweeps run 3 or 4 sequences per group (chosen from the batch size: recursion_largek.cu lk_launch); ragged batches on
    either side of the switch: Viterbi bit-exact, posteriors / log-likelihood within 1e-4 of float64."""
    rng = np.random.default_rng(9100 + K + B)
    logb = (rng.standard_normal((B, T, K)) * 3.0 - 20.0).astype(np.float32)
    P = rng.random((K, K)).astype(np.float32) ** 3 + 0.01
    P /= P.sum(1, keepdims=True)
    logP = np.log(P).astype(np.float32)
    p0 = np.full(K, 1.0 / K, np.float32)
    st, delta, _, score = c_oracle.viterbi_f32(logb, logP, np.log(p0))
    ws = hm.ops.viterbi_workspace(B, T, K, "cuda")
    r = hm.ops.viterbi(_dev(logb), hm.ops.EMIS_LOG, _dev(logP), _dev(np.log(p0)), want_delta=True, workspace=ws)
    torch.cuda.synchronize()
    assert K > 512 or _exchange_ok(ws)
    assert np.array_equal(r["delta"].cpu().numpy(), delta)
    assert np.array_equal(r["states"].cpu().numpy(), st)
    assert np.array_equal(r["score"].cpu().numpy(), score)
    _, _, gam, ll = c_oracle.forward_backward_f64(logb.astype(np.float64), np.log(P.astype(np.float64)), np.log(p0.astype(np.float64)))
    ws = hm.ops.fb_workspace(B, T, K, "cuda")
    f = hm.ops.forward_backward(_dev(logb), hm.ops.EMIS_LOG, _dev(P), _dev(p0), want=("gamma",), workspace=ws)
    torch.cuda.synchronize()
    assert _exchange_ok(ws)
    np.testing.assert_allclose(f["gamma"].cpu().numpy(), gam, rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(f["loglik"].cpu().numpy(), ll, rtol=RTOL, atol=1e-4)


@pytest.mark.parametrize("K,T,B", [(512, 40, 5), (130, 60, 3), (96, 33, 2)])
@pytest.mark.parametrize("kind", ["integers", "flat", "zeros"])
def test_largek_viterbi_exact_ties(hm, K, T, B, kind):
    """Ties everywhere: torch.max / argmax return the FIRST maximum (hmm.py:167, :174), so must the sweep's maxima and the
    traceback's warp-wide arg-max (two integer-key reductions; +0 and -0 are equal)."""
    rng = np.random.default_rng(500 + K + len(kind))
    if kind == "integers":
        logb = -rng.integers(0, 3, (B, T, K)).astype(np.float32)
        logP = -rng.integers(1, 3, (K, K)).astype(np.float32)
        logp0 = -rng.integers(1, 3, K).astype(np.float32)
    elif kind == "flat":
        logb = np.full((B, T, K), -1.25, np.float32); logP = np.full((K, K), -3.5, np.float32); logp0 = np.full(K, -2.0, np.float32)
    else:
        logb = np.zeros((B, T, K), np.float32); logb[:, ::2] = -0.0
        logP = np.zeros((K, K), np.float32); logp0 = np.zeros(K, np.float32)
    st, delta, _, score = c_oracle.viterbi_f32(logb, logP, logp0)
    r = hm.ops.viterbi(_dev(logb), hm.ops.EMIS_LOG, _dev(logP), _dev(logp0), want_delta=True)
    torch.cuda.synchronize()
    assert np.array_equal(r["delta"].cpu().numpy(), delta)
    assert np.array_equal(r["states"].cpu().numpy(), st)
    assert np.array_equal(r["score"].cpu().numpy(), score)


# (K, T, B): the launch shape (groups per cluster x sequences per register pass x passes per group) is chosen from the batch size, the
# cluster size and the number of sweeps in the launch (recursion_largek.cu lk_choose_shape): batches on either side of its switches
@pytest.mark.parametrize("K,T,B", [(512, 10, 24), (512, 8, 64), (512, 6, 100), (512, 5, 200), (130, 20, 50), (300, 12, 31), (64, 9, 150)])
def test_largek_fused_launch_equals_the_separate_passes(hm, K, T, B):
    """forward + backward + Viterbi in ONE launch (two-pass groups) against the stand-alone calls (one-pass groups): bit for bit --
    a sequence's arithmetic must not depend on how the batch was cut into clusters, groups and passes."""
    rng = np.random.default_rng(9300 + K + B)
    logb = (rng.standard_normal((B, T, K)) * 3.0 - 20.0).astype(np.float32)
    P = rng.random((K, K)).astype(np.float32) ** 3 + 0.01
    P /= P.sum(1, keepdims=True)
    logP = np.log(P).astype(np.float32)
    p0 = np.full(K, 1.0 / K, np.float32)
    e = _dev(logb)
    sep_f = hm.ops.forward_backward(e, hm.ops.EMIS_LOG, _dev(P), _dev(p0), want=("gamma", "fwd", "bwd"))
    sep_v = hm.ops.viterbi(e, hm.ops.EMIS_LOG, _dev(logP), _dev(np.log(p0)), want_delta=True)
    fwd_only = hm.ops.forward_backward(e, hm.ops.EMIS_LOG, _dev(P), _dev(p0), want=())
    r = hm.ops.forward_backward_viterbi(e, hm.ops.EMIS_LOG, hm.ops.EMIS_LOG, _dev(P), _dev(p0), _dev(logP), _dev(np.log(p0)))
    torch.cuda.synchronize()
    for k in ("gamma", "fwd", "bwd", "loglik"):
        assert torch.equal(r[k], sep_f[k]), k
    assert torch.equal(fwd_only["loglik"], sep_f["loglik"])
    for k in ("states", "delta", "score"):
        assert torch.equal(r[k], sep_v[k]), k
    st, delta, _, score = c_oracle.viterbi_f32(logb, logP, np.log(p0))
    assert np.array_equal(r["states"].cpu().numpy(), st) and np.array_equal(r["delta"].cpu().numpy(), delta)


@pytest.mark.parametrize("K,T,B", [(5, 120_000, 1), (12, 100_000, 2), (3, 30_000, 3)])
def test_smallk_viterbi_has_no_length_limit(hm, K, T, B):
    """Past the length whose traceback tables fit one CTA's shared memory (~24 k frames at K <= 4, ~95 k at K <= 16)
    the small-K Viterbi call takes the cluster kernel's sweep and path-only traceback: bit-exact all the same."""
    rng = np.random.default_rng(K + T)
    logb = (rng.standard_normal((B, T, K)) * 2.0).astype(np.float32)
    P = rng.random((K, K)).astype(np.float32) + 0.05
    logP = np.log(P / P.sum(1, keepdims=True)).astype(np.float32)
    logp0 = np.log(np.full(K, 1.0 / K)).astype(np.float32)
    st, delta, psi, score = c_oracle.viterbi_f32(logb, logP, logp0)
    r = hm.ops.viterbi(_dev(logb), hm.ops.EMIS_LOG, _dev(logP), _dev(logp0), want_delta=True)
    assert np.array_equal(r["states"].cpu().numpy(), st)
    assert np.array_equal(r["delta"].cpu().numpy(), delta)
    assert np.array_equal(r["score"].cpu().numpy(), score)
