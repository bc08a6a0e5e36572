"""GPU parity tests (-m gpu) at the FULL BASELINE.json shapes, through the public classes / engine and the C ABI, against the
C oracle (oracle/hmm_oracle.c) on the same fp32 inputs.

  C1  K=10, D=80, B=32, T=1000    GaussianHMMLayer(normalize_emissions=True) emission + HMMPyTorch semantics
  C2  K=12, C=4, D=80, B=256, T=2000   HMMInferenceEngine (the bench.py step)
  C4  K=10, Dmax=20, B=128, T=2000  HSMMLayer explicit-duration Viterbi (+ forward-backward properties)
  C5  K=512 ergodic, B=64, T=4000   HMMPyTorch.viterbi_decode semantics (+ forward-backward on a slice)

Bars: Viterbi states / delta / score BIT-EXACT on identical fp32 log-emissions (the kernel's own emissions are fed to the
oracle); posteriors: engine error <= max(1e-4, error of an fp32 log-space recursion like the reference's) against float64
(SURVEY finding 9); log-likelihood 1e-4 relative.  Where the scalar C oracle would take minutes (C4, C5) the oracle checks a
fixed subset of the sequences of the full-shape GPU run (sequences are independent) and size-independent properties cover
the rest.  Also: a bounded repeat-run stress of the large-K cluster kernels (the race hunt of tools/lk_stress.py).
"""
import math

import numpy as np
import pytest
import torch

from oracle import c_oracle, ref_port

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hm():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pytorch_hmm_b200 as m
    return m


def _gamma_gate(ours, gam64, gam32):
    mask = gam64 > 1e-6
    engine_err = float(np.max(np.abs(ours - gam64)[mask] / gam64[mask]))
    ref_err = float(np.max(np.abs(gam32 - gam64)[mask] / gam64[mask]))
    return engine_err, ref_err


def test_config2_engine_full_shape(hm):
    """BASELINE configs[1] exactly as bench.py runs it: B=256, T=2000, K=12, C=4, D=80 through HMMInferenceEngine."""
    import bench
    from pytorch_hmm_b200.engine import HMMInferenceEngine
    K, C, D, B, T = bench.K_STATES, bench.N_MIX, bench.FEAT, bench.BATCH, bench.SEQ
    model = bench.make_model()
    x = bench.make_frames(model, B, T, 2001).cuda()
    layer = hm.MixtureGaussianHMMLayer(K, D, num_components=C).cuda().eval()
    layer.load_state_dict({k: v.cuda() for k, v in model.items()})
    eng = HMMInferenceEngine(layer, B, T, shard=B, n_streams=1)
    out = {k: v.clone() for k, v in eng.run_device(x).items()}
    logb = eng.slots[0].logb.clone()
    torch.cuda.synchronize()
    lb = logb.cpu()
    # emission vs float64 (1e-5 relative; |l| ~ 100-200)
    logw = ref_port.safe_log(torch.softmax(model["mixture_weights_logits"], -1)).numpy()
    sub = slice(0, 16)
    ref_logb = c_oracle.gmm_emission_f64(x[sub].cpu().numpy(), model["means"].numpy(), model["log_vars"].numpy(), 1.0, logw)
    np.testing.assert_allclose(lb[sub].numpy(), ref_logb, rtol=1e-5, atol=1e-4)
    # Viterbi (mixture_gaussian.py:290-338 semantics) on the engine's own log-emissions: bit-exact, every sequence
    P = layer.get_transition_matrix().detach()
    log_trans = layer._safe_log(P).cpu().numpy()
    prior = np.full((K,), -math.log(K), np.float32)
    st, dl, _, sc = c_oracle.viterbi_f32(lb.numpy(), log_trans, prior)
    assert np.array_equal(out["states"].cpu().numpy(), st)
    assert np.array_equal(out["log_delta"].cpu().numpy(), dl)
    assert np.array_equal(out["score"].cpu().numpy(), sc)
    # forward-backward (hmm.py:66-130 on per-frame max-normalised probabilities): every sequence against float64
    hmm = hm.HMMPyTorch(P, None, device="cuda")
    obs = torch.exp(lb - lb.max(-1, keepdim=True)[0])
    log_obs32 = torch.log(obs + 1e-8).numpy()
    logP32, logp032 = hmm.log_P.cpu().numpy(), hmm.log_p0.cpu().numpy()
    la, lbt, gam64, ll64 = c_oracle.forward_backward_f64(log_obs32.astype(np.float64), logP32.astype(np.float64),
                                                         logp032.astype(np.float64))
    _, _, gam32 = c_oracle.forward_backward_f32(log_obs32, logP32, logp032)
    engine_err, ref_err = _gamma_gate(out["posterior"].cpu().numpy(), gam64, gam32)
    print(f"C2 full shape: gamma rel err vs float64: engine {engine_err:.3e}, fp32 log-space (reference-like) {ref_err:.3e}")
    assert engine_err <= max(1e-4, ref_err)
    np.testing.assert_allclose(out["loglik"].cpu().numpy(), ll64, rtol=1e-4)
    # forward / backward are returned in probability space (hmm.py:127-128): compare where they have not underflowed
    big = la > -80
    np.testing.assert_allclose(out["forward"].cpu().numpy()[big], np.exp(la)[big], rtol=1e-3)
    bigb = lbt > -80
    np.testing.assert_allclose(out["backward"].cpu().numpy()[bigb], np.exp(lbt)[bigb], rtol=1e-3)


def test_config1_full_shape(hm):
    """BASELINE configs[0]: left-to-right K=10, D=80 diag-Gaussian, B=32, T=1000: layer emission -> HMMPyTorch recursions."""
    K, D, B, T = 10, 80, 32, 1000
    torch.manual_seed(1001)
    g = hm.GaussianHMMLayer(K, D, normalize_emissions=True).cuda().eval()
    P = hm.create_left_to_right_matrix(K, 0.7)
    path = (torch.arange(T) * K // T).expand(B, T)
    x = (g.means.detach().cpu()[path] + torch.randn(B, T, D)).cuda()
    logb = g._compute_gaussian_log_probs(x).detach()
    ref_logb = c_oracle.gmm_emission_f64(x.cpu().numpy(), g.means.detach().cpu().numpy(), g.log_scales.detach().cpu().numpy(),
                                         2.0, None)
    np.testing.assert_allclose(logb.cpu().numpy(), ref_logb, rtol=1e-5, atol=1e-4)
    hmm = hm.HMMPyTorch(P, None, device="cuda")
    lb = logb.cpu()
    obs = torch.exp(lb - lb.max(-1, keepdim=True)[0])            # what the reference is fed
    log_obs32 = torch.log(obs + 1e-8).numpy()
    logP32, logp032 = hmm.log_P.cpu().numpy(), hmm.log_p0.cpu().numpy()
    # Viterbi on the identical fp32 log-observations: bit-exact
    st, dl, _, _ = c_oracle.viterbi_f32(log_obs32, logP32, logp032)
    r = hm.ops.viterbi(torch.from_numpy(log_obs32).cuda(), hm.ops.EMIS_LOG, hmm.log_P.cuda(), hmm.log_p0.cuda(), want_delta=True)
    assert np.array_equal(r["states"].cpu().numpy(), st)
    assert np.array_equal(r["delta"].cpu().numpy(), dl)
    # forward-backward through the drop-in class on the probabilities
    post, fwd, bwd = hmm.forward_backward(obs.cuda())
    la, lbt, gam64, ll64 = c_oracle.forward_backward_f64(log_obs32.astype(np.float64), logP32.astype(np.float64),
                                                         logp032.astype(np.float64))
    _, _, gam32 = c_oracle.forward_backward_f32(log_obs32, logP32, logp032)
    engine_err, ref_err = _gamma_gate(post.cpu().numpy(), gam64, gam32)
    print(f"C1 full shape: gamma rel err vs float64: engine {engine_err:.3e}, fp32 log-space (reference-like) {ref_err:.3e}")
    assert engine_err <= max(1e-4, ref_err)
    np.testing.assert_allclose(hmm.log_likelihood(obs.cuda()).cpu().numpy(), ll64, rtol=1e-4)
    # the layer's own forward (train: posteriors; eval: one-hot Viterbi of the same model)
    g.hmm_layer.train()
    with torch.no_grad():
        # the layer's transition matrix is softmax(log(P + 1e-8)) = P up to the floor; posteriors agree to 1e-4 with float64 of the
        # layer's own effective matrix
        pl = g(x)
    assert pl.shape == (B, T, K)
    np.testing.assert_allclose(pl.sum(-1).cpu().numpy(), 1.0, atol=1e-5)


def test_config5_full_shape_viterbi_and_fb_slice(hm):
    """BASELINE configs[4]: K=512 ergodic, B=64, T=4000.  The GPU runs the full shape; the scalar C oracle checks 6 of the 64
    sequences bit for bit (each is 1 G add/compare pairs), the final score of each against its own delta row, and the
    forward-backward on a T=200 slice against float64."""
    K, B, T = 512, 64, 4000
    P = hm.create_transition_matrix(K, "ergodic")
    gen = torch.Generator().manual_seed(5001)
    obs = torch.softmax(torch.randn(B, T, K, generator=gen), -1)
    hmm = hm.HMMPyTorch(P, None, device="cuda")
    log_obs = torch.log(obs + 1e-8)                              # CPU ATen log -> the identical fp32 inputs for both sides
    ws = hm.ops.viterbi_workspace(B, T, K, "cuda")
    r = hm.ops.viterbi(log_obs.cuda(), hm.ops.EMIS_LOG, hmm.log_P.cuda(), hmm.log_p0.cuda(), want_delta=True, workspace=ws)
    torch.cuda.synchronize()
    assert int(ws[-256:].view(torch.int32)[0].item()) == 0      # no DSMEM exchange time-out
    states, delta, score = r["states"].cpu().numpy(), r["delta"].cpu().numpy(), r["score"].cpu().numpy()
    pick = [0, 1, 17, 31, 42, 63]
    st, dl, _, sc = c_oracle.viterbi_f32(log_obs[pick].numpy(), hmm.log_P.cpu().numpy(), hmm.log_p0.cpu().numpy())
    assert np.array_equal(delta[pick], dl)
    assert np.array_equal(states[pick], st)
    assert np.array_equal(score[pick], sc)
    # properties on every sequence: the score is the maximum of the last delta row, reached at the last state of the path
    assert np.array_equal(score, delta[:, -1].max(-1))
    assert np.array_equal(delta[np.arange(B), -1, states[:, -1]], score)
    assert states.min() >= 0 and states.max() < K
    # forward-backward on a slice against float64
    Ts = 200
    sl = log_obs[:8, :Ts]
    trans, init = hmm._effective_probs(torch.device("cuda", 0))
    f = hm.ops.forward_backward(obs[:8, :Ts].cuda(), hm.ops.EMIS_PROB_FLOOR, trans, init, want=("gamma",))
    _, _, gam64, ll64 = c_oracle.forward_backward_f64(sl.numpy().astype(np.float64), hmm.log_P.cpu().numpy().astype(np.float64),
                                                      hmm.log_p0.cpu().numpy().astype(np.float64))
    np.testing.assert_allclose(f["gamma"].cpu().numpy(), gam64, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(f["loglik"].cpu().numpy(), ll64, rtol=1e-4)


def test_config4_full_shape_hsmm(hm):
    """BASELINE configs[3]: HSMM K=10, Dmax=20, D=80, B=128, T=2000.  Explicit-duration Viterbi (hsmm.py:245-354) at the full
    shape; the scalar oracle (72 M candidate evaluations per sequence) checks 12 of the 128 sequences bit for bit; the
    duration-augmented forward-backward (new functionality, parity unpinned) is checked by its identities on all of them."""
    K, D, Dm, B, T = 10, 80, 20, 128, 2000
    torch.manual_seed(4001)
    m = hm.HSMMLayer(K, D, duration_distribution="gamma", max_duration=Dm).cuda().eval()
    # x sampled from the model's own Gaussians along a random segmentation
    seg = torch.randint(0, K, (B, T // 10 + 1)).repeat_interleave(10, 1)[:, :T]
    x = (m.observation_means.detach().cpu()[seg] + torch.randn(B, T, D)).cuda()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        logb = m.get_observation_log_probs(x).detach()
        states, scores = m._viterbi_from_log_probs(logb)
        gamma, ll = m.forward_backward(x)
    torch.cuda.synchronize()
    log_dur, log_trans = m._tables(torch.device("cuda", 0))
    pick = list(range(0, B, 11))
    st, sc = c_oracle.hsmm_viterbi_f32(logb[pick].cpu().numpy(), log_dur.cpu().numpy(), log_trans.cpu().numpy())
    assert np.array_equal(states[pick].cpu().numpy(), st)
    assert np.array_equal(scores[pick].cpu().numpy(), sc)
    s = states.cpu().numpy()
    assert s.min() >= 0 and s.max() < K
    # every segment of the decoded path is at most Dmax long (the recursion has no self-transition, hsmm.py:294)
    run = np.ones(B, np.int64); longest = np.ones(B, np.int64)
    for t in range(1, T):
        same = s[:, t] == s[:, t - 1]
        run = np.where(same, run + 1, 1)
        longest = np.maximum(longest, run)
    assert longest.max() <= 2 * Dm          # two consecutive segments of one state are impossible, so <= Dmax; slack for clarity
    assert longest.max() <= Dm
    g = gamma.cpu().numpy()
    assert np.isfinite(g).all() and np.isfinite(ll.cpu().numpy()).all()
    np.testing.assert_allclose(g.sum(-1), 1.0, atol=2e-4)       # state-occupancy posteriors sum to one per frame
    assert (ll.cpu().numpy() >= scores.cpu().numpy() - 1e-3 * np.abs(scores.cpu().numpy())).all()   # sum over paths >= best path


def test_largek_repeat_run_stress(hm, golden):
    """Bounded form of tools/lk_stress.py: the cluster kernels are deterministic, so every repeat of a fixed input must be
    bit-identical to the first (and, for the K=64 fixture, to the reference's own delta); the exchange time-out flag stays 0."""
    g = golden("largek")
    dev = "cuda"
    cases = []
    P64 = torch.from_numpy(g["k64_P"]).to(dev) + 1e-8
    cases.append(("k64", torch.from_numpy(g["k64_log_obs"]).to(dev), torch.from_numpy(g["k64_log_P"]).to(dev),
                  torch.from_numpy(g["k64_log_p0"]).to(dev), P64, torch.full((64,), 1.0 / 64, device=dev),
                  torch.from_numpy(g["k64_log_delta"]).to(dev)))
    torch.manual_seed(1)
    for K, B, T in ((512, 13, 64), (200, 7, 50)):
        logb = torch.randn(B, T, K, device=dev) * 3 - 20
        Pm = torch.rand(K, K, device=dev) ** 3 + 0.01
        Pm = Pm / Pm.sum(1, keepdim=True)
        p0 = torch.full((K,), 1.0 / K, device=dev)
        cases.append((f"K{K}", logb, torch.log(Pm), torch.log(p0), Pm, p0, None))
    for name, logb, logP, logp0, Pm, p0, gold in cases:
        B, T, K = logb.shape
        wsv = hm.ops.viterbi_workspace(B, T, K, dev)
        wsf = hm.ops.fb_workspace(B, T, K, dev)
        first_d = first_g = None
        for i in range(200):
            r = hm.ops.viterbi(logb, hm.ops.EMIS_LOG, logP, logp0, want_delta=True, workspace=wsv)
            f = hm.ops.forward_backward(logb, hm.ops.EMIS_LOG, Pm, p0, want=("gamma",), workspace=wsf)
            if first_d is None:
                first_d, first_g = r["delta"].clone(), f["gamma"].clone()
                if gold is not None:
                    assert torch.equal(first_d, gold.reshape(first_d.shape)), name
            else:
                assert torch.equal(r["delta"], first_d), f"{name}: run {i} delta differs from run 0"
                assert torch.equal(f["gamma"], first_g), f"{name}: run {i} gamma differs from run 0"
        torch.cuda.synchronize()
        assert int(wsv[-256:].view(torch.int32)[0].item()) == 0, name
