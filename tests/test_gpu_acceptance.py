"""Acceptance tests (-m gpu): the reference's own test STRATEGY (SURVEY section 4 -- property tests on random inputs: shapes,
rows sum to one, indices in range, monotone left-to-right paths, finiteness under extreme inputs, determinism, train/eval
switch) restated against the drop-in classes.  The reference's test files are cited for each property; nothing is read
from /root/reference at run time."""
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


# ---- HMMPyTorch (reference tests/test_hmm.py:60-120, :300-333) ---------------------------------------------------------
@pytest.mark.parametrize("K,B,T", [(5, 3, 40), (10, 2, 200), (50, 2, 30), (3, 1, 1)])
def test_core_properties(hm, K, B, T):
    torch.manual_seed(K + T)
    P = hm.create_left_to_right_matrix(K, 0.7)
    hmm = hm.HMMPyTorch(P, None, device="cuda")
    obs = torch.softmax(torch.randn(B, T, K), -1).cuda()
    post, fwd, bwd = hmm.forward_backward(obs)
    assert post.shape == fwd.shape == bwd.shape == (B, T, K) and post.dtype == torch.float32
    assert torch.allclose(post.sum(-1), torch.ones(B, T, device="cuda"), atol=1e-5)          # test_hmm.py:77
    assert torch.isfinite(post).all() and (post >= 0).all()
    states, delta = hmm.viterbi_decode(obs)
    assert states.shape == (B, T) and states.dtype == torch.int64 and delta.shape == (B, T, K)
    assert (states >= 0).all() and (states < K).all()                                         # test_hmm.py:91-92
    ll = hmm.compute_likelihood(obs)
    assert ll.shape == (B,) and torch.isfinite(ll).all()                                      # test_hmm.py:119
    # sequences are independent: a batch equals its rows processed alone
    p1, _, _ = hmm.forward_backward(obs[:1])
    s1, _ = hmm.viterbi_decode(obs[:1])
    assert torch.equal(p1, post[:1]) and torch.equal(s1, states[:1])
    # 2-D input: Viterbi squeezes the batch dimension, forward_backward does not (hmm.py:79-80, :180-182)
    s2, d2 = hmm.viterbi_decode(obs[0])
    p2, _, _ = hmm.forward_backward(obs[0])
    assert s2.shape == (T,) and d2.shape == (T, K) and p2.shape == (1, T, K)


def test_left_to_right_paths_are_monotone(hm):
    """test_hmm.py:95 -- with (almost) one-hot evidence along a monotone path the decoded path never goes back."""
    K, T = 6, 60
    P = hm.create_left_to_right_matrix(K, 0.8)
    hmm = hm.HMMPyTorch(P, None, device="cuda")
    path = torch.arange(T) * K // T
    obs = torch.full((1, T, K), 0.02)
    obs[0, torch.arange(T), path] = 0.9
    states, _ = hmm.viterbi_decode(obs.cuda())
    assert (states[0, 1:] >= states[0, :-1]).all()


def test_constructor_errors(hm):
    with pytest.raises(ValueError):
        hm.HMM(torch.ones(3, 4))                                                              # hmm.py:33-36
    with pytest.raises(ValueError):
        hm.HMM(torch.ones(3, 3), torch.ones(4))
    hmm = hm.HMMPyTorch(torch.eye(4) + 0.1, device="cuda")
    with pytest.raises(AssertionError):
        hmm.viterbi_decode(torch.rand(2, 5, 3).cuda())                                        # hmm.py:150


# ---- HMMLayer / GaussianHMMLayer (reference tests/test_hmm.py:125-208) ----------------------------------------------------
def test_hmm_layer_modes(hm):
    torch.manual_seed(2)
    K, B, T = 6, 3, 25
    layer = hm.HMMLayer(K).cuda()
    x = torch.randn(B, T, K).cuda()
    layer.train()
    post = layer(x)
    assert post.shape == (B, T, K) and torch.allclose(post.sum(-1), torch.ones(B, T, device="cuda"), atol=1e-5)
    assert not isinstance(layer(x, return_alignment=True), tuple)                            # tuple only in eval, hmm_layer.py:133-140
    layer.eval()
    post, align = layer(x, return_alignment=True)
    assert align.shape == (B, T) and ((post == 0) | (post == 1)).all() and (post.sum(-1) == 1).all()   # one-hot, :124-128
    assert torch.equal(post.argmax(-1), align)
    soft = hm.HMMLayer(K, viterbi_inference=False).cuda().eval()
    p2, a2 = soft(x, return_alignment=True)
    assert torch.allclose(p2.sum(-1), torch.ones(B, T, device="cuda"), atol=1e-5) and torch.equal(a2, p2.argmax(-1))
    for ly in (layer, soft):
        assert torch.isfinite(ly.compute_loss(x)) and ly.compute_loss(x).dim() == 0
        tgt = torch.randint(0, K, (B, T)).cuda()
        assert torch.isfinite(ly.compute_loss(x, tgt))
    T_mat = layer.get_transition_matrix()
    assert torch.allclose(T_mat.sum(1), torch.ones(K, device="cuda"), atol=1e-6)


def test_gaussian_hmm_layer(hm):
    torch.manual_seed(3)
    g = hm.GaussianHMMLayer(5, 12, normalize_emissions=True).cuda()
    x = torch.randn(2, 30, 12).cuda()
    g.train()
    post = g(x)
    assert post.shape == (2, 30, 5) and torch.allclose(post.sum(-1), torch.ones(2, 30, device="cuda"), atol=1e-5)
    assert torch.isfinite(g.compute_loss(x))
    for cov in ("diag", "spherical", "full"):
        lp = hm.GaussianHMMLayer(4, 6, covariance_type=cov).cuda()._compute_gaussian_log_probs(torch.randn(2, 7, 6).cuda()).detach()
        assert lp.shape == (2, 7, 4) and torch.isfinite(lp).all()


# ---- MixtureGaussianHMMLayer (reference tests/test_mixture_gaussian.py:80-300) -----------------------------------------
def test_mixture_layer_properties(hm):
    torch.manual_seed(4)
    m = hm.MixtureGaussianHMMLayer(5, 20, num_components=3).cuda().eval()
    x = torch.randn(4, 50, 20).cuda()
    states, none = m(x)
    assert none is None and states.shape == (4, 50) and (states >= 0).all() and (states < 5).all()
    lp = m.get_observation_log_probs(x).detach()
    assert lp.shape == (4, 50, 5) and torch.isfinite(lp).all() and (lp <= 0).all()            # :110-121
    A = m.get_transition_matrix()
    assert (A >= 0).all() and (A <= 1).all() and torch.allclose(A.sum(1), torch.ones(5, device="cuda"), atol=1e-6)
    with torch.no_grad():
        s1, p1 = m(x, return_log_probs=True)
        s2, p2 = m(x, return_log_probs=True)
    assert torch.equal(s1, s2) and torch.equal(p1, p2)                                        # determinism in eval, :281-296
    for T in (1, 10, 100, 500):                                                               # :178-197
        s, p = m(torch.randn(2, T, 20).cuda(), return_log_probs=True)
        assert s.shape == (2, T) and p.shape == (2,)
    info = m.get_model_info()
    assert info["num_states"] == 5 and info["total_parameters"] > 0


@pytest.mark.parametrize("cov", ["diag", "tied", "spherical"])
def test_mixture_covariance_types(hm, cov):
    m = hm.MixtureGaussianHMMLayer(4, 8, num_components=2, covariance_type=cov).cuda().eval()
    s, p = m(torch.randn(2, 30, 8).cuda(), return_log_probs=True)
    assert s.shape == (2, 30) and torch.isfinite(p).all()


def test_mixture_numerical_stability_with_extreme_inputs(hm):
    """test_mixture_gaussian.py:138-157: inputs of 1e-10, 1e10 and 1e5-scaled noise must decode to valid states."""
    m = hm.MixtureGaussianHMMLayer(3, 10, num_components=2).cuda().eval()
    for data in (torch.full((1, 20, 10), 1e-10), torch.full((1, 20, 10), 1e10), torch.randn(1, 20, 10) * 1e5):
        s, _ = m(data.cuda())
        assert (s >= 0).all() and (s < 3).all()
    fixed = hm.MixtureGaussianHMMLayer(4, 6, num_components=2, learnable_transitions=False).cuda().eval()
    s, _ = fixed(torch.randn(1, 40, 6).cuda())
    assert s.shape == (1, 40)


def test_mixture_large_model_memory(hm):
    """test_mixture_gaussian.py:231-260: K=50, C=5, D=80, B=8, T=2000 must stay under 2000 MB (the reference materialises
    [B,T,S,C,D] temporaries; here the emission is one kernel and K = 50 takes the cluster Viterbi)."""
    m = hm.MixtureGaussianHMMLayer(50, 80, num_components=5).cuda().eval()
    x = torch.randn(8, 2000, 80).cuda()
    torch.cuda.synchronize(); torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    s, p = m(x, return_log_probs=True)
    torch.cuda.synchronize()
    assert s.shape == (8, 2000) and torch.isfinite(p).all() and (s >= 0).all() and (s < 50).all()
    assert (torch.cuda.max_memory_allocated() - base) / 2 ** 20 < 2000


# ---- HSMMLayer (reference tests/test_hsmm.py:60-200, :340-360) ------------------------------------------------------------
@pytest.mark.parametrize("dist", ["gamma", "poisson", "weibull"])
def test_hsmm_layer_properties(hm, dist):
    torch.manual_seed(5)
    h = hm.HSMMLayer(5, 12, duration_distribution=dist, max_duration=10).cuda().eval()
    A = h.get_transition_matrix()
    assert torch.all(torch.diagonal(A) == 0)                                                  # test_hsmm.py:100-113
    assert torch.allclose(A.sum(1), torch.ones(5, device="cuda"), atol=1e-5)
    dp = h.get_duration_probabilities()
    assert dp.shape == (5, 10) and (dp >= 0).all() and torch.isfinite(dp).all()
    assert (h.get_expected_durations() > 0).all()
    x = torch.randn(3, 40, 12).cuda()
    states, scores = h(x)
    assert states.shape == (3, 40) and states.dtype == torch.int64 and scores.shape == (3,)
    assert (states >= 0).all() and (states < 5).all() and torch.isfinite(scores).all()
    post, ll = h.forward_backward(x)
    assert torch.allclose(post.sum(-1), torch.ones(3, 40, device="cuda"), atol=1e-4) and torch.isfinite(ll).all()
    assert (ll >= scores - 1e-3).all()                                                        # sum over segmentations >= best one


# ---- StreamingHMMProcessor (reference tests/test_streaming.py:60-200) ---------------------------------------------------
def test_streaming_processor_api(hm):
    torch.manual_seed(6)
    p = hm.StreamingHMMProcessor(6, 8, chunk_size=16, overlap_size=4, lookahead_frames=2, max_delay_frames=64,
                                 use_beam_search=False).cuda().eval()
    r = p.process_chunk(torch.randn(4, 8).cuda())
    assert isinstance(r, hm.StreamingResult) and r.status == "buffering"
    r = p.process_chunk(torch.randn(24, 8).cuda())
    assert r.status != "buffering" and (r.decoded_states >= 0).all() and (r.decoded_states < 6).all()
    assert 0.0 <= float(r.confidence) <= 1.0
    p.reset_streaming_state()
    assert p.process_chunk(torch.randn(4, 8).cuda()).status == "buffering"
    assert isinstance(p.get_performance_stats(), dict)


# ---- the throughput front end returns exactly what the plain calls return -------------------------------------------------
def test_engine_sharded_and_host_paths_match_single_pass_bit_for_bit(hm):
    """HMMInferenceEngine cuts the batch into shards on several streams (and, with host_io, pipelines the PCIe copies):
    utterances are independent, so every output must equal the un-sharded pass bit for bit -- including the posterior,
    whose arithmetic must not depend on where a frame lands in a launch."""
    from pytorch_hmm_b200.engine import HMMInferenceEngine
    torch.manual_seed(9)
    K, C, D, B, T = 12, 4, 80, 10, 333
    layer = hm.MixtureGaussianHMMLayer(K, D, num_components=C).cuda().eval()
    x = torch.randn(B, T, D).cuda() + layer.means.detach()[torch.randint(0, K, (B, T)), torch.randint(0, C, (B, T))]
    whole = HMMInferenceEngine(layer, B, T, shard=B, n_streams=1)
    ref = {k: v.clone() for k, v in whole.run_device(x).items()}
    torch.cuda.synchronize()
    sharded = HMMInferenceEngine(layer, B, T, shard=3, n_streams=3, host_io=True)
    out = sharded.run_device(x)
    torch.cuda.synchronize()
    for k in ("posterior", "forward", "backward", "log_delta", "states", "score", "loglik"):
        assert torch.equal(out[k], ref[k]), k
    xh = x.cpu().pin_memory()
    host = {k: torch.empty(ref[k].shape, dtype=ref[k].dtype).pin_memory() for k in ("posterior", "log_delta", "states")}
    sharded.run_host(xh, host)
    torch.cuda.synchronize()
    for k, v in host.items():
        assert torch.equal(v, ref[k].cpu()), k
    g = whole.capture_device(x)
    whole.out["posterior"].zero_()
    g.replay(); torch.cuda.synchronize()
    assert torch.equal(whole.out["posterior"], ref["posterior"])
    # and the plain drop-in calls agree with the engine
    st, sc = layer(x, return_log_probs=True)
    assert torch.equal(st, ref["states"]) and torch.equal(sc.detach(), ref["score"])
