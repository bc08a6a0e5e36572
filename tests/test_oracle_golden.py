"""CPU tests: the oracle (oracle/ref_port.py torch port + oracle/hmm_oracle.c C restatement) against the
golden fixtures that oracle/make_golden.py produced by running the real reference.  This is what "pins" the
oracle (task brief, point 3); the -m gpu tests then compare the CUDA path with the oracle."""
import numpy as np
import torch

from oracle import c_oracle, ref_port

CORE_TAGS = ["a", "b", "c", "d", "e"]


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def test_prepare_hmm_matches_reference(golden):
    g = golden("core")
    for tag in CORE_TAGS:
        p0 = _t(g[f"{tag}_p0"]) if f"{tag}_p0" in g.files else None
        log_P, log_p0 = ref_port.prepare_hmm(_t(g[f"{tag}_P"]), p0)
        assert np.array_equal(log_P.numpy(), g[f"{tag}_log_P"])
        assert np.array_equal(log_p0.numpy(), g[f"{tag}_log_p0"])


def test_port_forward_backward_bit_identical(golden):
    g = golden("core")
    for tag in CORE_TAGS:
        obs = _t(g[f"{tag}_obs"])
        post, fwd, bwd = ref_port.forward_backward(obs, _t(g[f"{tag}_log_P"]), _t(g[f"{tag}_log_p0"]))
        assert np.array_equal(post.numpy(), g[f"{tag}_posterior"]), tag
        assert np.array_equal(fwd.numpy(), g[f"{tag}_forward"]), tag
        assert np.array_equal(bwd.numpy(), g[f"{tag}_backward"]), tag
        ll = ref_port.compute_likelihood(obs, _t(g[f"{tag}_log_P"]), _t(g[f"{tag}_log_p0"]))
        assert np.array_equal(ll.numpy(), g[f"{tag}_likelihood"]), tag


def test_port_viterbi_bit_identical(golden):
    g = golden("core")
    for tag in CORE_TAGS:
        states, delta = ref_port.viterbi_decode(_t(g[f"{tag}_obs"]), _t(g[f"{tag}_log_P"]), _t(g[f"{tag}_log_p0"]))
        assert np.array_equal(states.numpy(), g[f"{tag}_states"]), tag
        assert np.array_equal(delta.numpy(), g[f"{tag}_log_delta"]), tag
        assert states.dtype == torch.int64


def test_c_viterbi_bit_identical_to_reference(golden):
    """The C restatement reproduces states AND the full delta trellis bit for bit (fp32 add/max only)."""
    g = golden("core")
    for tag in CORE_TAGS:
        obs = _t(g[f"{tag}_obs"])
        if obs.dim() == 2:
            obs = obs[None]
        log_obs = torch.log(obs + 1e-8).numpy()          # same ATen log as the reference used
        states, delta, psi, score = c_oracle.viterbi_f32(log_obs, g[f"{tag}_log_P"], g[f"{tag}_log_p0"])
        ref_states = g[f"{tag}_states"].reshape(states.shape)
        ref_delta = g[f"{tag}_log_delta"].reshape(delta.shape)
        assert np.array_equal(states, ref_states), tag
        assert np.array_equal(delta, ref_delta), tag
        _, _, psi_port = ref_port.viterbi_log(_t(log_obs), _t(g[f"{tag}_log_P"]), _t(g[f"{tag}_log_p0"]))
        assert np.array_equal(psi, psi_port.numpy().astype(np.int32)), tag


def test_floor_induced_ties_take_lowest_index(golden):
    """SURVEY finding 8: fully floored frames make every state tie; torch.max picks index 0."""
    g = golden("core")
    obs = _t(g["c_obs"])
    _, _, psi = ref_port.viterbi_log(torch.log(obs + 1e-8), _t(g["c_log_P"]), _t(g["c_log_p0"]))
    assert (psi[:, 1:] >= 0).all()
    states, _, psi_c, _ = c_oracle.viterbi_f32(torch.log(obs + 1e-8).numpy(), g["c_log_P"], g["c_log_p0"])
    assert np.array_equal(psi_c, psi.numpy().astype(np.int32))
    assert np.array_equal(states, g["c_states"])


def test_c_forward_backward_f64_vs_reference(golden):
    g = golden("core")
    for tag in CORE_TAGS:
        obs = g[f"{tag}_obs"].astype(np.float64)
        if obs.ndim == 2:
            obs = obs[None]
        la, lb, gam, ll = c_oracle.forward_backward_f64(np.log(obs.astype(np.float32) + np.float32(1e-8)).astype(np.float64),
                                                        g[f"{tag}_log_P"], g[f"{tag}_log_p0"])
        ref_post = g[f"{tag}_posterior"].reshape(gam.shape)
        # the reference itself is fp32: agreement to ~1e-5 relative at these short T (tolerance stated)
        np.testing.assert_allclose(gam, ref_post, rtol=2e-4, atol=1e-7)
        ref_fwd = g[f"{tag}_forward"].reshape(la.shape)
        nz = ref_fwd > 1e-30
        np.testing.assert_allclose(np.exp(la)[nz], ref_fwd[nz], rtol=2e-4)


def test_gaussian_emission(golden):
    g = golden("gaussian")
    lp = ref_port.gaussian_log_probs(_t(g["x"]), _t(g["means"]), _t(g["log_scales"]))
    assert np.array_equal(lp.numpy(), g["log_probs"])
    c = c_oracle.gmm_emission_f64(g["x"], g["means"], g["log_scales"], 2.0, None)
    np.testing.assert_allclose(c, g["log_probs"], rtol=2e-6)


def test_mixture_emission_and_viterbi(golden):
    g = golden("mixture")
    for tag in ("soft", "sharp"):
        x = _t(g[f"{tag}_x"])
        logb = ref_port.gmm_log_probs(x, _t(g[f"{tag}_means"]), _t(g[f"{tag}_log_vars"]), _t(g[f"{tag}_mixture_weights_logits"]))
        assert np.array_equal(logb.numpy(), g[f"{tag}_logb"]), tag
        log_trans = ref_port.safe_log(torch.softmax(_t(g[f"{tag}_transition_logits"]), dim=-1))
        assert np.array_equal(log_trans.numpy(), g[f"{tag}_log_trans"])
        states, scores, delta, psi = ref_port.mixture_viterbi(logb, log_trans)
        assert np.array_equal(states.numpy(), g[f"{tag}_states"]), tag
        assert np.array_equal(scores.numpy(), g[f"{tag}_scores"]), tag
        # C oracle: emission in double, Viterbi in fp32 with the uniform prior -log K
        logw = ref_port.safe_log(torch.softmax(_t(g[f"{tag}_mixture_weights_logits"]), dim=-1)).numpy()
        c = c_oracle.gmm_emission_f64(g[f"{tag}_x"], g[f"{tag}_means"], g[f"{tag}_log_vars"], 1.0, logw)
        np.testing.assert_allclose(c, g[f"{tag}_logb"], rtol=3e-6)
        K = logb.shape[-1]
        prior = np.full((K,), -np.float32(np.log(np.float64(K))), np.float32)
        # delta_0 = logb_0 - log K: the reference subtracts a Python float -> same as adding fl32(-log K)
        st_c, delta_c, psi_c, score_c = c_oracle.viterbi_f32(g[f"{tag}_logb"], g[f"{tag}_log_trans"], prior)
        assert np.array_equal(st_c, g[f"{tag}_states"]), tag
        assert np.array_equal(delta_c, delta.numpy()), tag
        assert np.array_equal(score_c, g[f"{tag}_scores"]), tag


def test_hsmm_viterbi_bit_identical(golden):
    g = golden("hsmm")
    for tag in ("gamma", "poisson", "weibull"):
        logdur = torch.log(_t(g[f"{tag}_dur_probs"]) + 1e-8).numpy()       # hsmm.py:227
        logA = torch.log(_t(g[f"{tag}_trans"]) + 1e-8).numpy()             # hsmm.py:229
        states, score = c_oracle.hsmm_viterbi_f32(g[f"{tag}_logb"], logdur, logA)
        assert np.array_equal(states, g[f"{tag}_states"]), tag
        assert np.array_equal(score, g[f"{tag}_scores"]), tag


def test_segsum_order(golden):
    """torch.sum over a strided fp32 slice = 4 interleaved partial sums; the C oracle's seg_sum4 relies on it."""
    g = golden("segsum")
    x = g["x"]
    for t in range(40):
        for d in range(1, 21):
            v = x[t:t + d, 2]
            p = np.zeros(4, np.float32)
            q = d // 4
            for i in range(q):
                for k in range(4):
                    p[k] = np.float32(p[k] + v[4 * i + k])
            for i in range(4 * q, d):
                p[0] = np.float32(p[0] + v[i])
            for k in range(1, 4):
                p[0] = np.float32(p[0] + p[k])
            assert p[0] == g["sums"][t, d]


def _semimarkov_seg(g):
    """seg[t][s][d-1] as SemiMarkovHMM._compute_segment_observation_logprob defines it (semi_markov.py:411-425):
    the Gaussian constant is counted ONCE per segment."""
    x = g["x"][0].astype(np.float64)
    mu = g["observation_means"].astype(np.float64); lv = g["observation_logvars"].astype(np.float64)
    T, D = x.shape; K = mu.shape[0]; Dm = g["forward_variables"].shape[2]
    seg = np.full((T, K, Dm), -np.inf)
    for t in range(T):
        for s in range(K):
            for d in range(1, min(Dm, t + 1) + 1):
                fr = x[t - d + 1:t + 1]
                seg[t, s, d - 1] = (-0.5 * lv[s].sum() - 0.5 * D * np.log(2 * np.pi)
                                    - 0.5 * (((fr - mu[s]) ** 2) / np.exp(lv[s])).sum())
    return seg


def test_semimarkov_forward(golden):
    g = golden("semimarkov")
    seg = _semimarkov_seg(g)
    logA = np.log(torch.softmax(_t(g["transition_logits"]), dim=1).numpy().astype(np.float64) + 1e-8)
    logpi = np.log(torch.softmax(_t(g["initial_logits"]), dim=0).numpy().astype(np.float64) + 1e-8)
    alpha, beta, tot = c_oracle.hsmm_forward_f64(seg, g["log_dur"], logA, logpi)
    np.testing.assert_allclose(tot, float(g["log_probability"]), rtol=1e-5)
    ref = g["forward_variables"]
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(alpha), fin)
    np.testing.assert_allclose(alpha[fin], ref[fin], rtol=1e-5)
    # beta has no reference: the forward/backward identity must hold at every t where a segment can end
    K = alpha.shape[1]
    endv = np.logaddexp.reduce(alpha, axis=2)
    for t in range(alpha.shape[0]):
        v = np.logaddexp.reduce(endv[t] + beta[t])
        # sum over "a segment ends at t" is <= total; equality only at t = T-1
        assert v <= tot + 1e-9
    np.testing.assert_allclose(np.logaddexp.reduce(endv[-1] + beta[-1]), tot, rtol=1e-12)


def test_streaming_greedy(golden):
    g = golden("streaming")
    logA = torch.log(torch.softmax(_t(g["transition_logits"]), dim=-1) + 1e-8).numpy()
    n1 = int(g["chunk1_frames"]); n2 = int(g["chunk2_frames"])
    s1, sc1 = c_oracle.greedy_decode_f32(g["logb"][:n1], logA, -1)
    assert np.array_equal(s1, g["chunk1_states"])
    s2, sc2 = c_oracle.greedy_decode_f32(g["logb"][n1:n1 + n2], logA, int(s1[-1]))
    assert np.array_equal(s2, g["chunk2_states"])
    np.testing.assert_allclose(np.exp(sc1).mean(), float(g["chunk1_conf"]), rtol=1e-5)
    np.testing.assert_allclose(np.exp(sc2).mean(), float(g["chunk2_conf"]), rtol=1e-5)


def test_bw_stats_consistency():
    """A9 has no reference implementation (formulas only, docs/01_hmm_theory.md:196-227): check the identities
    sum_j xi(i,j) = sum_{t<T-1} gamma_t(i), occupancies sum to T, and agreement with a direct numpy evaluation."""
    rng = np.random.default_rng(7)
    B, T, K, C, D = 2, 15, 3, 2, 4
    x = rng.standard_normal((B, T, D)).astype(np.float32)
    comp = rng.standard_normal((B, T, K, C)) * 2.0
    P = rng.random((K, K)) + 0.1; P /= P.sum(1, keepdims=True)
    p0 = np.full((K,), 1.0 / K)
    st = c_oracle.bw_stats_f64(x, comp, np.log(P), np.log(p0))
    lb = np.logaddexp.reduce(comp, axis=3)
    la, lbe, gam, ll = c_oracle.forward_backward_f64(lb, np.log(P), np.log(p0))
    np.testing.assert_allclose(st["loglik"], ll.sum(), rtol=1e-12)
    np.testing.assert_allclose(st["gamma1"], gam[:, 0].sum(0), rtol=1e-10)
    np.testing.assert_allclose(st["xi"].sum(1), gam[:, :-1].sum((0, 1)), rtol=1e-9)
    np.testing.assert_allclose(st["occ"].sum(), B * T, rtol=1e-10)
    resp = gam[..., None] * np.exp(comp - lb[..., None])
    np.testing.assert_allclose(st["sx"], np.einsum("btkc,btd->kcd", resp, x.astype(np.float64)), rtol=1e-9, atol=1e-12)


def test_oracle_pinned_at_large_k(golden):
    """K = 64 and K = 512 (BASELINE config 5 family): the torch port and the C Viterbi reproduce the reference bit for
    bit; the float64 forward-backward agrees with the reference's fp32 posteriors."""
    g = golden("largek")
    for tag in ("k64", "k512"):
        obs = _t(g[f"{tag}_obs"])
        log_P, log_p0 = ref_port.prepare_hmm(_t(g[f"{tag}_P"]), None)
        assert np.array_equal(log_P.numpy(), g[f"{tag}_log_P"])
        post, fwd, bwd = ref_port.forward_backward(obs, log_P, log_p0)
        assert np.array_equal(post.numpy(), g[f"{tag}_posterior"]), tag
        states, delta = ref_port.viterbi_decode(obs, log_P, log_p0)
        assert np.array_equal(states.numpy(), g[f"{tag}_states"]), tag
        log_obs = torch.log(obs + 1e-8).numpy()
        st, dl, _, _ = c_oracle.viterbi_f32(log_obs, g[f"{tag}_log_P"], g[f"{tag}_log_p0"])
        assert np.array_equal(st, g[f"{tag}_states"]) and np.array_equal(dl, g[f"{tag}_log_delta"]), tag
        _, _, gam, _ = c_oracle.forward_backward_f64(log_obs.astype(np.float64), g[f"{tag}_log_P"].astype(np.float64),
                                                      g[f"{tag}_log_p0"].astype(np.float64))
        np.testing.assert_allclose(gam, g[f"{tag}_posterior"], rtol=1e-4, atol=1e-7)


def test_neural_oracle_vs_reference_golden(golden):
    """orc_tv_viterbi_f32 / orc_tv_forward_backward_f64 restate NeuralHMM's recursions (neural.py:403-511): bit-identical Viterbi,
    posteriors to the reference's fp32 rounding (1e-3 absolute at T = 300), on the real class's outputs."""
    from oracle import c_oracle
    g = golden("neural")
    for tag in ("tv", "tv12"):
        st, dl, psi = c_oracle.tv_viterbi_f32(g[f"{tag}_log_obs"], g[f"{tag}_log_trans"], g[f"{tag}_log_init"])
        assert np.array_equal(st, g[f"{tag}_states"]) and np.array_equal(dl, g[f"{tag}_log_delta"])
        la, lb, gam, ll = c_oracle.tv_forward_backward_f64(g[f"{tag}_log_obs"], g[f"{tag}_log_trans"], g[f"{tag}_log_init"])
        np.testing.assert_allclose(gam, g[f"{tag}_posterior"], atol=1e-3)
        np.testing.assert_allclose(np.exp(la), g[f"{tag}_forward"], rtol=1e-3, atol=1e-37)
        np.testing.assert_allclose(np.exp(lb), g[f"{tag}_backward"], rtol=1e-3, atol=1e-37)
        sat = np.log(np.sum(np.exp(la[:, -1]) + 1e-8, axis=-1))
        np.testing.assert_allclose(sat, g[f"{tag}_likelihood"], rtol=1e-4)
    # the static-transition case is the fixed-matrix recursion on log-emissions
    st, dl, psi, sc = c_oracle.viterbi_f32(g["static_log_obs"], g["static_log_trans"], g["static_log_init"])
    assert np.array_equal(st, g["static_states"]) and np.array_equal(dl, g["static_log_delta"])


# ------------------------------------------------------------------------------------------------
# alignment utilities: oracle/alignment_port.py against the real reference's outputs
# ------------------------------------------------------------------------------------------------
def _finite_close(a, b, tol):
    assert np.array_equal(np.isfinite(a), np.isfinite(b))
    m = np.isfinite(b)
    return np.all(np.abs(a[m] - b[m]) <= tol * np.maximum(1.0, np.abs(b[m])))


def test_alignment_port_ctc_matches_reference(golden):
    from oracle import alignment_port as ap
    g = golden("alignment")
    for tag in ("a", "b"):
        args = (g[f"ctc_{tag}_log_probs"], g[f"ctc_{tag}_targets"], g[f"ctc_{tag}_input_lengths"], g[f"ctc_{tag}_target_lengths"],
                int(g[f"ctc_{tag}_blank"]))
        _, ll = ap.ctc_forward(*args)
        assert _finite_close(ll, g[f"ctc_{tag}_loglik"], 1e-6), tag
        assert _finite_close(ap.ctc_backward(*args), g[f"ctc_{tag}_log_beta"], 1e-6), tag
    # the forward log-likelihood is minus torch's CTC loss (an implementation the reference does not share code with)
    _, ll = ap.ctc_forward(g["ctc_a_log_probs"], g["ctc_a_targets"], g["ctc_a_input_lengths"], g["ctc_a_target_lengths"], 0)
    assert np.allclose(-ll, g["ctc_a_torch_nll"], rtol=1e-6)


def test_alignment_port_dtw_bit_identical(golden):
    from oracle import alignment_port as ap
    g = golden("alignment")
    for tag in ("rand", "ties"):
        for pat in ("symmetric", "asymmetric", "rabiner_juang"):
            pi, pj, cost = ap.dtw(g[f"dtw_{tag}_dist"], pat)
            assert np.array_equal(cost, g[f"dtw_{tag}_{pat}_cost"]), (tag, pat)
            assert np.array_equal(pi, g[f"dtw_{tag}_{pat}_path_i"]) and np.array_equal(pj, g[f"dtw_{tag}_{pat}_path_j"]), (tag, pat)
