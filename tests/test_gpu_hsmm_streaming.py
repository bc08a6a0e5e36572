"""GPU parity tests (-m gpu) for the explicit-duration (HSMM) recursions and the streaming kernels, against the golden
fixtures of the real reference and the C oracle.  Nothing here reads /root/reference."""
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


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _load_hsmm_layer(hm, g, tag, dist):
    K, D = g[f"{tag}_observation_means"].shape
    Dm = g[f"{tag}_dur_probs"].shape[1]
    m = hm.HSMMLayer(K, D, duration_distribution=dist, max_duration=Dm).cuda()
    sd = {k: torch.from_numpy(g[f"{tag}_{k}"]) for k in m.state_dict() if f"{tag}_{k}" in g.files}
    m.load_state_dict(sd, strict=False)
    return m


@pytest.mark.parametrize("tag,dist", [("gamma", "gamma"), ("poisson", "poisson"), ("weibull", "weibull")])
def test_hsmm_layer_vs_reference_golden(hm, golden, tag, dist):
    g = golden("hsmm")
    m = _load_hsmm_layer(hm, g, tag, dist)
    # host-side tables equal the reference's (same torch formulas; GPU libm may differ in the last bits)
    np.testing.assert_allclose(m.get_duration_probabilities().detach().cpu().numpy(), g[f"{tag}_dur_probs"], rtol=2e-5, atol=1e-30)
    np.testing.assert_allclose(m.get_transition_matrix().detach().cpu().numpy(), g[f"{tag}_trans"], rtol=1e-6)
    x = _dev(g[f"{tag}_x"])
    logb = m.get_observation_log_probs(x).detach()
    np.testing.assert_allclose(logb.cpu().numpy(), g[f"{tag}_logb"], rtol=1e-5, atol=1e-4)
    states, scores = m(x)
    assert states.dtype == torch.int64 and states.shape == g[f"{tag}_states"].shape
    np.testing.assert_allclose(scores.cpu().numpy(), g[f"{tag}_scores"], rtol=1e-5)
    assert np.array_equal(states.cpu().numpy(), g[f"{tag}_states"])


@pytest.mark.parametrize("tag", ["gamma", "poisson", "weibull"])
def test_hsmm_viterbi_kernel_bit_exact_on_reference_inputs(hm, golden, tag):
    """Identical fp32 tables and log-emissions in -> identical states AND scores out (hsmm.py:245-354 operation order)."""
    g = golden("hsmm")
    log_dur = torch.log(torch.from_numpy(g[f"{tag}_dur_probs"]) + 1e-8)
    log_trans = torch.log(torch.from_numpy(g[f"{tag}_trans"]) + 1e-8)
    states, scores = hm.ops.hsmm_viterbi(_dev(g[f"{tag}_logb"]), log_dur.cuda(), log_trans.cuda(), sum_order=0)
    assert np.array_equal(states.cpu().numpy(), g[f"{tag}_states"])
    assert np.array_equal(scores.cpu().numpy(), g[f"{tag}_scores"])


@pytest.mark.parametrize("K,Dm,T,B", [(2, 3, 7, 2), (5, 8, 30, 3), (6, 10, 200, 4), (10, 20, 300, 2), (3, 1, 9, 1), (4, 25, 12, 2)])
def test_hsmm_viterbi_vs_c_oracle(hm, K, Dm, T, B):
    rng = np.random.default_rng(K * 31 + Dm)
    logb = (rng.standard_normal((B, T, K)) * 4 - 20).astype(np.float32)
    logdur = np.log(rng.random((K, Dm)) + 1e-3).astype(np.float32)
    A = rng.random((K, K)).astype(np.float32) + 0.05
    np.fill_diagonal(A, 0.0)
    logA = np.log(A / A.sum(1, keepdims=True) + 1e-8).astype(np.float32)
    st, sc = c_oracle.hsmm_viterbi_f32(logb, logdur, logA)
    states, scores = hm.ops.hsmm_viterbi(_dev(logb), _dev(logdur), _dev(logA), sum_order=0)
    assert np.array_equal(scores.cpu().numpy(), sc)
    assert np.array_equal(states.cpu().numpy(), st)


@pytest.mark.parametrize("K,Dm", [(10, 20), (6, 10)])
@pytest.mark.parametrize("kind", ["integers", "merging", "flat"])
def test_hsmm_viterbi_exact_ties_and_roundings(hm, K, Dm, kind):
    """The winner of a cell is the FIRST (state, duration) whose own fp32 total equals the maximum (hsmm.py:283-300): inputs built so
    that many candidates tie exactly (small integers), or differ by less than an ulp of the total (so that distinct candidates round to
    the same total and an EARLIER, smaller one must win), or are all equal.  K = 10, Dmax = 20 is the specialised kernel (searches
    bounded by the row's first arg-max); the other shape runs the general one."""
    rng = np.random.default_rng(77 + K + len(kind))
    B, T = 3, 150
    if kind == "integers":
        logb = -rng.integers(1, 4, (B, T, K)).astype(np.float32)
        logdur = -rng.integers(1, 3, (K, Dm)).astype(np.float32)
        logA = np.full((K, K), -2.0, np.float32)
    elif kind == "merging":
        logb = (-rng.integers(1, 3, (B, T, K)) * 64.0).astype(np.float32)                      # totals around -1e4: ulp ~ 1e-3
        logdur = (-1.0 - rng.random((K, Dm)) * 1e-4).astype(np.float32)                        # differences far below that ulp
        logA = (-2.0 - rng.random((K, K)) * 1e-4).astype(np.float32)
    else:
        logb = np.full((B, T, K), -1.5, np.float32)
        logdur = np.full((K, Dm), -0.75, np.float32)
        logA = np.full((K, K), -2.25, np.float32)
    np.fill_diagonal(logA, np.float32(np.log(1e-8)))
    st, sc = c_oracle.hsmm_viterbi_f32(logb, logdur, logA)
    states, scores = hm.ops.hsmm_viterbi(_dev(logb), _dev(logdur), _dev(logA), sum_order=0)
    assert np.array_equal(scores.cpu().numpy(), sc)
    assert np.array_equal(states.cpu().numpy(), st)


def _semimarkov(hm, g):
    K, D = g["observation_means"].shape
    Dm = g["forward_variables"].shape[2]
    m = hm.SemiMarkovHMM(K, D, max_duration=Dm, duration_distribution="gamma").cuda()
    sd = {"transition_logits": g["transition_logits"], "initial_logits": g["initial_logits"],
          "observation_means": g["observation_means"], "observation_logvars": g["observation_logvars"],
          "duration_model.alpha_params": g["duration_model.alpha_params"],
          "duration_model.beta_params": g["duration_model.beta_params"]}
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return m


def test_semimarkov_forward_vs_reference_golden(hm, golden):
    g = golden("semimarkov")
    m = _semimarkov(hm, g)
    np.testing.assert_allclose(m.duration_model.log_table().detach().cpu().numpy(), g["log_dur"], rtol=1e-5, atol=1e-6)
    res = m(_dev(g["x"]))
    # 1e-4 relative on the log-probability and on every finite forward variable (fp32 reference)
    np.testing.assert_allclose(res["log_probability"].item(), float(g["log_probability"]), rtol=1e-4)
    ref = g["forward_variables"]; ours = res["forward_variables"].cpu().numpy()
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(ours), fin)
    np.testing.assert_allclose(ours[fin], ref[fin], rtol=1e-4)


def test_semimarkov_viterbi_vs_reference_golden(hm, golden):
    g = golden("semimarkov")
    m = _semimarkov(hm, g)
    st, du, lp = m.viterbi_decode(_dev(g["x"][0]))
    assert np.array_equal(st.cpu().numpy(), g["vit_states"])
    assert np.array_equal(du.cpu().numpy(), g["vit_durations"])
    np.testing.assert_allclose(lp.item(), float(g["vit_logprob"]), rtol=1e-5)


def _semimarkov_sup(hm, g, dist):
    K, D = g[f"{dist}_observation_means"].shape
    m = hm.SemiMarkovHMM(K, D, max_duration=8, duration_distribution=dist, min_duration=2 if dist == "gaussian" else 1).cuda()
    m.load_state_dict({k: torch.from_numpy(g[f"{dist}_{k}"]) for k in m.state_dict()})
    return m


@pytest.mark.parametrize("dist", ["gamma", "poisson", "gaussian"])
def test_semimarkov_supervised_forward_vs_reference_golden(hm, golden, dist):
    """SemiMarkovHMM.forward(x, state_sequence, duration_sequence) (semi_markov.py:280-305): segments that tile T, stop short of T,
    and overrun T; one duration below min_duration (-inf) in the gaussian case.  1e-5 relative: fp32 sums of a few dozen terms."""
    g = golden("semimarkov_sup")
    m = _semimarkov_sup(hm, g, dist)
    with torch.no_grad():
        res = m(_dev(g[f"{dist}_x"]), _dev(g[f"{dist}_states"]), _dev(g[f"{dist}_durs"]))
    for k in ("log_observation", "log_duration", "log_transition", "log_probability"):
        ref = g[f"{dist}_{k}"]; ours = res[k].cpu().numpy()
        assert not res[k].requires_grad
        assert np.array_equal(np.isfinite(ours), np.isfinite(ref)), k
        np.testing.assert_allclose(ours[np.isfinite(ref)], ref[np.isfinite(ref)], rtol=1e-5, atol=1e-5, err_msg=k)


def test_semimarkov_supervised_forward_gradients(hm, golden):
    """Supervised training through the given segmentation: gradients w.r.t. every parameter against float64 autograd of the
    reference's formulas restated with tensor operations (same three terms)."""
    g = golden("semimarkov_sup")
    m = _semimarkov_sup(hm, g, "gamma")
    x, st, du = _dev(g["gamma_x"]), _dev(g["gamma_states"]), _dev(g["gamma_durs"])
    res = m(x, st, du)
    np.testing.assert_allclose(res["log_probability"].detach().cpu().numpy(), g["gamma_log_probability"], rtol=1e-5)
    res["log_probability"].sum().backward()
    ours = {n: p.grad.detach().double().cpu() for n, p in m.named_parameters() if p.grad is not None}
    P = {n: p.detach().double().cpu().requires_grad_(True) for n, p in m.named_parameters()}
    xd, T = x.double().cpu(), x.shape[1]
    F = torch.nn.functional
    total = 0.0
    for b in range(xd.shape[0]):
        t = 0
        for j in range(st.shape[1]):
            s_, d_ = int(st[b, j]), int(du[b, j])
            a = F.softplus(P["duration_model.alpha_params"][s_]) + 1e-6
            be = F.softplus(P["duration_model.beta_params"][s_]) + 1e-6
            total = total + (a - 1) * np.log(d_ + 1e-8) - be * d_ - (torch.lgamma(a) - a * torch.log(be))
            if j > 0:
                total = total + torch.log(F.softmax(P["transition_logits"], 1) + 1e-8)[int(st[b, j - 1]), s_]
        for j in range(st.shape[1]):
            s_, d_ = int(st[b, j]), int(du[b, j])
            if t + d_ > T:
                break
            lv = P["observation_logvars"][s_]
            total = total - 0.5 * lv.sum() - 0.5 * xd.shape[2] * np.log(2 * np.pi) \
                - 0.5 * (((xd[b, t:t + d_] - P["observation_means"][s_]) ** 2) / torch.exp(lv)).sum()
            t += d_
    total.backward()
    for n in ("observation_means", "observation_logvars", "transition_logits", "duration_model.alpha_params",
              "duration_model.beta_params"):
        np.testing.assert_allclose(ours[n].numpy(), P[n].grad.numpy(), rtol=2e-4, atol=2e-4, err_msg=n)


def test_hsmm_forward_vs_float64_oracle(hm):
    rng = np.random.default_rng(12)
    K, Dm, T = 5, 7, 60
    f = (rng.standard_normal((1, T, K)) * 3 - 10).astype(np.float32)
    segc = (rng.standard_normal(K) - 5).astype(np.float32)
    logdur = np.log(rng.random((K, Dm)) + 1e-2).astype(np.float32)
    A = rng.random((K, K)) + 0.05
    np.fill_diagonal(A, 0.0)
    logA = np.log(A / A.sum(1, keepdims=True) + 1e-8).astype(np.float32)
    logpi = np.log(np.full(K, 1.0 / K)).astype(np.float32)
    seg = np.full((T, K, Dm), -np.inf)
    for t in range(T):
        for s in range(K):
            for d in range(1, min(Dm, t + 1) + 1):
                seg[t, s, d - 1] = float(segc[s]) + f[0, t - d + 1:t + 1, s].astype(np.float64).sum()
    alpha, beta, tot = c_oracle.hsmm_forward_f64(seg, logdur, logA, logpi)
    r = hm.ops.hsmm_forward(_dev(f), _dev(logdur), _dev(logA), seg_const=_dev(segc), log_init=_dev(logpi))
    np.testing.assert_allclose(r["total"].item(), tot, rtol=1e-5)
    fin = np.isfinite(alpha)
    np.testing.assert_allclose(r["alpha"][0].cpu().numpy()[fin], alpha[fin], rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("K,Dm,T,B", [(2, 3, 6, 2), (3, 4, 7, 1), (5, 7, 60, 3), (10, 20, 300, 2), (12, 5, 150, 4), (32, 6, 40, 2),
                                      (4, 30, 25, 1)])
def test_hsmm_forward_backward_vs_float64_oracle(hm, K, Dm, T, B):
    """Duration-augmented forward-backward (BASELINE config 4 family).  No reference implementation exists (SURVEY finding 5):
    the oracle is the float64 restatement, itself pinned by brute-force enumeration in tests/test_host_cpu.py.
    Tolerance: posteriors 1e-4 relative (atol 1e-6: they are differences of running sums), log-likelihood 1e-5 relative."""
    from oracle import hsmm_post
    rng = np.random.default_rng(300 + K + Dm + T)
    f = (rng.standard_normal((B, T, K)) * 3 - 10).astype(np.float32)
    segc = (rng.standard_normal(K) - 5).astype(np.float32)
    logdur = np.log(rng.random((K, Dm)) + 1e-2).astype(np.float32)
    A = rng.random((K, K)) + 0.05
    np.fill_diagonal(A, 0.0)
    logA = np.log(A / A.sum(1, keepdims=True) + 1e-8).astype(np.float32)
    logpi = np.log(rng.dirichlet(np.ones(K))).astype(np.float32)
    r = hm.ops.hsmm_forward_backward(_dev(f), _dev(logdur), _dev(logA), seg_const=_dev(segc), log_init=_dev(logpi), want_beta=True)
    gam = r["gamma"].cpu().numpy()
    for b in range(B):
        g64, tot = hsmm_post.posteriors_f64(f[b].astype(np.float64), segc.astype(np.float64), logdur.astype(np.float64),
                                            logA.astype(np.float64), logpi.astype(np.float64))
        np.testing.assert_allclose(r["total"][b].item(), tot, rtol=1e-5)
        np.testing.assert_allclose(gam[b], g64, rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(gam.sum(-1), 1.0, atol=1e-4)                    # every frame is covered by exactly one segment
    assert r["beta_end"][:, -1].abs().max().item() == 0.0


def test_hsmm_forward_backward_config4_shape_properties(hm):
    """BASELINE config 4 shape (K=10, Dmax=20, T=2000) at B=4: posteriors in [0,1] and summing to 1 on every frame, sequence 0
    against the float64 oracle at full length (1e-4 relative, atol 2e-6), and the fp32 log-space forward kernel's total
    (its own rounding is ~1e-7 |log p|, hence 1e-5 relative here)."""
    rng = np.random.default_rng(4001)
    K, Dm, T, B = 10, 20, 2000, 4
    f = (rng.standard_normal((B, T, K)) * 2 - 100).astype(np.float32)
    logdur = np.log(rng.random((K, Dm)) + 1e-3).astype(np.float32)
    A = rng.random((K, K)) + 0.05
    np.fill_diagonal(A, 0.0)
    logA = np.log(A / A.sum(1, keepdims=True) + 1e-8).astype(np.float32)
    r = hm.ops.hsmm_forward_backward(_dev(f), _dev(logdur), _dev(logA))
    g = r["gamma"].cpu().numpy()
    assert np.isfinite(g).all() and g.min() >= 0.0 and g.max() <= 1.0
    np.testing.assert_allclose(g.sum(-1), 1.0, atol=1e-4)
    from oracle import hsmm_post
    g64, tot = hsmm_post.posteriors_f64(f[0].astype(np.float64), np.zeros(K), logdur.astype(np.float64), logA.astype(np.float64),
                                        np.zeros(K))
    np.testing.assert_allclose(r["total"][0].item(), tot, rtol=1e-6)
    np.testing.assert_allclose(g[0], g64, rtol=1e-4, atol=2e-6)
    fw = hm.ops.hsmm_forward(_dev(f), _dev(logdur), _dev(logA), want_alpha=False)
    np.testing.assert_allclose(r["total"].cpu().numpy(), fw["total"].cpu().numpy(), rtol=1e-5)


def test_streaming_greedy_vs_reference_golden(hm, golden):
    g = golden("streaming")
    p = hm.StreamingHMMProcessor(6, 8, chunk_size=16, overlap_size=4, lookahead_frames=2, max_delay_frames=64,
                                 use_beam_search=False).cuda()
    sd = {k.replace("__", "."): torch.from_numpy(g[k]) for k in g.files if k.startswith(("transition_logits", "emission_net"))}
    p.load_state_dict(sd)
    p.eval()
    feats = torch.from_numpy(g["feats"]).cuda()
    r1 = p.process_chunk(feats[:24])
    r2 = p.process_chunk(feats[24:48])
    assert r1.status == "decoded" and r2.status == "decoded"
    assert r1.metadata["frames_processed"] == int(g["chunk1_frames"]) and r2.metadata["frames_processed"] == int(g["chunk2_frames"])
    assert np.array_equal(r1.decoded_states.cpu().numpy(), g["chunk1_states"])
    assert np.array_equal(r2.decoded_states.cpu().numpy(), g["chunk2_states"])
    np.testing.assert_allclose(r1.confidence, float(g["chunk1_conf"]), rtol=1e-4)
    np.testing.assert_allclose(r2.confidence, float(g["chunk2_conf"]), rtol=1e-4)
    assert p.process_chunk(feats[:2]).status in ("decoded", "waiting_for_lookahead", "buffering")
    p.reset_streaming_state()
    assert p.process_chunk(feats[:4]).status == "buffering"


def test_greedy_kernel_vs_c_oracle(hm):
    rng = np.random.default_rng(3)
    K, T = 9, 150
    logb = np.log(rng.dirichlet(np.ones(K), size=T)).astype(np.float32)
    logA = np.log(rng.dirichlet(np.ones(K), size=K) + 1e-8).astype(np.float32)
    s1, sc1 = c_oracle.greedy_decode_f32(logb[:70], logA, -1)
    s2, sc2 = c_oracle.greedy_decode_f32(logb[70:], logA, int(s1[-1]))
    state = torch.full((1,), -1, dtype=torch.int32, device="cuda")
    a, sa = hm.ops.greedy_decode(_dev(logb[None, :70]), _dev(logA), state)
    b, sb = hm.ops.greedy_decode(_dev(logb[None, 70:]), _dev(logA), state)
    assert np.array_equal(a[0].cpu().numpy(), s1) and np.array_equal(b[0].cpu().numpy(), s2)
    np.testing.assert_allclose(sa[0].cpu().numpy(), sc1, rtol=1e-6)
    assert int(state.item()) == int(s2[-1])


def test_forward_chunk_equals_unchunked_and_float64(hm):
    """New functionality (no reference): the carried-state forward over chunks must equal one pass, and the float64 oracle."""
    rng = np.random.default_rng(5)
    B, T, K = 3, 400, 12
    l = (rng.standard_normal((B, T, K)) * 3 - 40).astype(np.float32)
    P = rng.dirichlet(np.ones(K), size=K).astype(np.float32)
    p0 = rng.dirichlet(np.ones(K)).astype(np.float32)
    la, lb, gam, ll = c_oracle.forward_backward_f64(l.astype(np.float64), np.log(P.astype(np.float64)), np.log(p0.astype(np.float64)))
    full = hm.ops.new_forward_state(B, K, "cuda")
    f_full = hm.ops.forward_chunk(_dev(l), hm.ops.EMIS_LOG, _dev(P), _dev(p0), full)
    st = hm.ops.new_forward_state(B, K, "cuda")
    parts = [hm.ops.forward_chunk(_dev(l[:, a:b]), hm.ops.EMIS_LOG, _dev(P), _dev(p0), st) for a, b in ((0, 37), (37, 160), (160, 161), (161, 400))]
    f_chunks = torch.cat(parts, dim=1)
    assert torch.equal(f_chunks, f_full)                                   # bit-identical: same arithmetic, state carried
    assert torch.equal(st["loglik"], full["loglik"])
    np.testing.assert_allclose(full["loglik"].cpu().numpy(), ll, rtol=1e-5)
    filt64 = np.exp(la - np.logaddexp.reduce(la, axis=-1, keepdims=True))
    np.testing.assert_allclose(f_full.cpu().numpy(), filt64, rtol=1e-4, atol=1e-7)


def test_streaming_forward_chunk_api(hm):
    p = hm.StreamingHMMProcessor(6, 8, use_beam_search=False).cuda().eval()
    x = torch.randn(50, 8).cuda()
    a = p.forward_chunk(x[:20]); b = p.forward_chunk(x[20:])
    p.reset_streaming_state()
    c = p.forward_chunk(x)
    assert torch.allclose(torch.cat([a["filtered"], b["filtered"]]), c["filtered"])
    assert torch.allclose(b["log_likelihood"], c["log_likelihood"])
    np.testing.assert_allclose(c["filtered"].sum(-1).cpu().numpy(), 1.0, atol=1e-5)


def test_factories(hm):
    m = hm.create_speech_hmm(12, 80, "mixture_gaussian", num_mixtures=4).cuda()     # the README call that raises in the reference
    st, sc = m(torch.randn(2, 30, 80).cuda(), return_log_probs=True)
    assert st.shape == (2, 30) and sc.shape == (2,)
    h = hm.create_speech_hmm(5, 16, "hsmm", max_duration=6).cuda()
    st, sc = h(torch.randn(2, 25, 16).cuda())
    assert st.shape == (2, 25) and torch.isfinite(sc).all()
    assert isinstance(hm.ModelFactory.create_realtime_model(4, 8), hm.StreamingHMMProcessor)
    with pytest.raises(ValueError):
        hm.create_speech_hmm(3, 4, "nope")


def test_streaming_async_front_end_equals_synchronous_calls(hm):
    """start_async_processing / add_audio_chunk_async / get_result_async (streaming.py:123-181): the worker thread decodes the
    chunks in arrival order, so the results equal those of process_chunk called directly on a second processor with the same
    parameters."""
    import time
    torch.manual_seed(77)
    a = hm.StreamingHMMProcessor(5, 25, chunk_size=40, use_beam_search=False).cuda().eval()
    b = hm.StreamingHMMProcessor(5, 25, chunk_size=40, use_beam_search=False).cuda().eval()
    b.load_state_dict(a.state_dict())
    chunks = [torch.randn(40, 25) for _ in range(6)]
    with torch.no_grad():
        want = [b.process_chunk(c) for c in chunks]
    a.start_async_processing()
    try:
        assert all(a.add_audio_chunk_async(c) for c in chunks)
        got, deadline = [], time.time() + 20.0
        while len(got) < len(chunks) and time.time() < deadline:
            r = a.get_result_async()
            if r is None:
                time.sleep(0.005)
            else:
                got.append(r)
    finally:
        a.stop_async_processing()
    assert len(got) == len(chunks) and a.get_result_async() is None
    for g, w in zip(got, want):
        assert isinstance(g, hm.StreamingResult) and g.status == w.status and g.chunk_id == w.chunk_id
        if w.decoded_states is None:
            assert g.decoded_states is None
        else:
            assert torch.equal(g.decoded_states.cpu(), w.decoded_states.cpu())
