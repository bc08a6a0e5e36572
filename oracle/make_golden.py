"""oracle/make_golden.py -- writes tests/golden/*.npz by RUNNING THE REAL REFERENCE.  TEST INFRASTRUCTURE ONLY.

Run in the build container (the reference does not travel to the GPU box):

    python oracle/make_golden.py            # needs /root/reference

The reference's own tests hold no golden vectors, so parity is pinned by the outputs of
the reference itself on seeded inputs.  Every array the fixtures hold was produced by a reference class or
function, called through its public API; inputs are stored next to the outputs so the fixtures are
self-contained.  One shim is applied, and only for the ``semimarkov`` fixture: SemiMarkovHMM._unsupervised_forward
crashes at semi_markov.py:353 because it passes a Python float to torch.logaddexp (SURVEY.md finding 5); the
generator wraps torch.logaddexp so that float arguments are promoted to tensors -- the reference code itself is
executed unmodified.
"""
from __future__ import annotations

import io
import os
import sys
import contextlib

import numpy as np
import torch

REF = os.environ.get("HMM_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference not found at {REF}")
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import pytorch_hmm  # noqa: F401  (prints a banner on import)
    return pytorch_hmm


def _np(t):
    return t.detach().cpu().numpy()


def core(ref):
    """A1-A4: HMM.__init__, forward_backward, viterbi_decode, compute_likelihood (hmm.py)."""
    from pytorch_hmm.utils import create_left_to_right_matrix, create_transition_matrix
    out = {}
    g = torch.Generator().manual_seed(1101)

    def run(tag, P, p0, obs):
        hmm = ref.HMMPyTorch(P, p0)
        post, fwd, bwd = hmm.forward_backward(obs)
        states, delta = hmm.viterbi_decode(obs)
        ll = hmm.compute_likelihood(obs)
        out.update({f"{tag}_P": _np(P), f"{tag}_obs": _np(obs),
                    f"{tag}_log_P": _np(hmm.log_P), f"{tag}_log_p0": _np(hmm.log_p0),
                    f"{tag}_log_obs": _np(torch.log(obs + 1e-8)),   # hmm.py:152, on the generating machine's ATen
                    f"{tag}_posterior": _np(post), f"{tag}_forward": _np(fwd), f"{tag}_backward": _np(bwd),
                    f"{tag}_states": _np(states), f"{tag}_log_delta": _np(delta), f"{tag}_likelihood": _np(ll)})
        if p0 is not None:
            out[f"{tag}_p0"] = _np(p0)

    # a: small dense model with an explicit prior
    P = torch.rand(4, 4, generator=g) + 0.05
    p0 = torch.rand(4, generator=g) + 0.1
    run("a", P, p0, torch.softmax(torch.randn(2, 12, 4, generator=g), dim=-1))
    # b: BASELINE config 1 transition structure (left-to-right K=10, self loop 0.7), default uniform prior
    run("b", create_left_to_right_matrix(10, 0.7), None,
        torch.softmax(2.0 * torch.randn(3, 60, 10, generator=g), dim=-1))
    # c: floor-induced exact ties (SURVEY finding 8): most probabilities are exactly 0 -> log(1e-8)
    obs = torch.zeros(2, 30, 6)
    hot = torch.randint(0, 6, (2, 30), generator=g)
    obs.scatter_(2, hot.unsqueeze(-1), 1.0)
    obs[:, ::3] = 0.0                                    # whole frames floored: all states tie
    run("c", create_transition_matrix(6, "ergodic"), None, obs)
    # d: 2-D input (auto-batched; Viterbi squeezes, forward_backward does not) hmm.py:79-80,180-182
    run("d", create_transition_matrix(5, "left_to_right_skip"), None,
        torch.softmax(torch.randn(9, 5, generator=g), dim=-1))
    # e: K=12 longer sequence (headline K), T=200
    P = torch.softmax(0.1 * torch.randn(12, 12, generator=g), dim=-1)
    run("e", P, None, torch.softmax(3.0 * torch.randn(2, 200, 12, generator=g), dim=-1))
    np.savez_compressed(os.path.join(OUT, "core.npz"), **out)


def gaussian(ref):
    """A5/A6: GaussianHMMLayer emission (hmm_layer.py:270-323) and HMMLayer train/eval switch (:91-142)."""
    torch.manual_seed(1201)
    layer = ref.GaussianHMMLayer(10, 80, covariance_type="diag")
    with torch.no_grad():
        layer.log_scales.copy_(0.2 * torch.randn(10, 80))
    x = torch.randn(2, 16, 80) + layer.means.detach()[torch.randint(0, 10, (2, 16))]
    with torch.no_grad():
        lp = layer._compute_gaussian_log_probs(x)
    out = {"means": _np(layer.means), "log_scales": _np(layer.log_scales), "x": _np(x), "log_probs": _np(lp)}
    # HMMLayer: train -> forward_backward posterior, eval -> one-hot Viterbi (+ alignment)
    torch.manual_seed(1202)
    hl = ref.HMMLayer(7, learnable_transitions=True, transition_type="left_to_right", self_loop_prob=0.7)
    with torch.no_grad():
        hl.log_transition_logits.add_(0.3 * torch.randn(7, 7))
        hl.log_initial_logits.add_(0.3 * torch.randn(7))
    xin = torch.randn(2, 25, 7)
    hl.train()
    with torch.no_grad():
        post_train = hl(xin)
    hl.eval()
    with torch.no_grad():
        post_eval, align = hl(xin, return_alignment=True)
        st, sc = hl.align(xin)
        nll = hl.compute_loss(xin)
    out.update({"hl_log_transition_logits": _np(hl.log_transition_logits), "hl_log_initial_logits": _np(hl.log_initial_logits),
                "hl_x": _np(xin), "hl_post_train": _np(post_train), "hl_post_eval": _np(post_eval),
                "hl_alignment": _np(align), "hl_align_states": _np(st), "hl_align_scores": _np(sc),
                "hl_nll": _np(nll)})
    np.savez_compressed(os.path.join(OUT, "gaussian.npz"), **out)


def mixture(ref):
    """A7/A8: MixtureGaussianHMMLayer emission + private Viterbi (mixture_gaussian.py:157-365), two regimes."""
    out = {}
    for tag, seed, sharp in (("soft", 2001, False), ("sharp", 2002, True)):
        torch.manual_seed(seed)
        m = ref.MixtureGaussianHMMLayer(12, 80, num_components=4)
        with torch.no_grad():
            if sharp:                                     # means x6, log_vars ~ 0.3 N
                m.means.mul_(6.0)
                m.log_vars.copy_(0.3 * torch.randn_like(m.log_vars))
        # sample x from the model: random state/component per frame
        B, T = 2, 40
        s = torch.randint(0, 12, (B, T)); c = torch.randint(0, 4, (B, T))
        x = m.means.detach()[s, c] + torch.exp(0.5 * m.log_vars.detach()[s, c]) * torch.randn(B, T, 80)
        m.eval()
        with torch.no_grad():
            logb = m.get_observation_log_probs(x)
            states, scores = m(x, return_log_probs=True)
            log_trans = m._safe_log(m.get_transition_matrix())
        out.update({f"{tag}_means": _np(m.means), f"{tag}_log_vars": _np(m.log_vars),
                    f"{tag}_mixture_weights_logits": _np(m.mixture_weights_logits),
                    f"{tag}_transition_logits": _np(m.transition_logits), f"{tag}_x": _np(x),
                    f"{tag}_logb": _np(logb), f"{tag}_log_trans": _np(log_trans),
                    f"{tag}_states": _np(states), f"{tag}_scores": _np(scores)})
    np.savez_compressed(os.path.join(OUT, "mixture.npz"), **out)


def hsmm(ref):
    """A10-A12: HSMMLayer duration tables, emission and Viterbi (hsmm.py:115-354).  Tiny: the reference is O(T K^2 D^2) Python."""
    out = {}
    for tag, dist, seed in (("gamma", "gamma", 4001), ("poisson", "poisson", 4002), ("weibull", "weibull", 4003)):
        torch.manual_seed(seed)
        K, D, Dm, T, B = 4, 8, 6, 18, 2
        m = ref.HSMMLayer(K, D, duration_distribution=dist, max_duration=Dm)
        with torch.no_grad():
            m.observation_means.mul_(12.0)
            m.transition_logits.mul_(8.0)
            if dist == "gamma":
                m.duration_rate.fill_(0.7)
            elif dist == "poisson":
                m.duration_lambda.fill_(3.0)
            else:
                m.duration_scale.fill_(3.0)
        seq = torch.randint(0, K, (B, T))
        x = m.observation_means.detach()[seq] + torch.randn(B, T, D)
        with torch.no_grad():
            states, scores = m(x)
            out.update({f"{tag}_{n}": _np(p) for n, p in m.named_parameters()})
            out.update({f"{tag}_x": _np(x), f"{tag}_states": _np(states), f"{tag}_scores": _np(scores),
                        f"{tag}_dur_probs": _np(m.get_duration_probabilities()),
                        f"{tag}_trans": _np(m.get_transition_matrix()),
                        f"{tag}_logb": _np(m.get_observation_log_probs(x))})
    np.savez_compressed(os.path.join(OUT, "hsmm.npz"), **out)


def semimarkov(ref):
    """A13: SemiMarkovHMM._unsupervised_forward (semi_markov.py:308-383) and viterbi_decode (:455-570)."""
    real = torch.logaddexp

    def shim(a, b):                                       # the only shim: promote Python floats to tensors
        if not torch.is_tensor(a):
            a = torch.tensor(float(a))
        if not torch.is_tensor(b):
            b = torch.tensor(float(b))
        return real(a, b)

    torch.manual_seed(4101)
    K, D, Dm, T = 3, 4, 5, 12
    m = ref.SemiMarkovHMM(K, D, max_duration=Dm, duration_distribution="gamma")
    with torch.no_grad():
        m.duration_model.alpha_params.copy_(torch.tensor([1.0, 2.0, 3.0]))
        m.duration_model.beta_params.copy_(torch.tensor([0.5, 1.0, 1.5]))
        m.initial_logits.copy_(torch.tensor([0.3, -0.2, 0.1]))
    x = torch.randn(1, T, D) + m.observation_means.detach()[torch.randint(0, K, (T,))][None]
    torch.logaddexp = shim
    try:
        with torch.no_grad():
            res = m(x)
    finally:
        torch.logaddexp = real
    with torch.no_grad():
        st, du, lp = m.viterbi_decode(x[0])
        logdur = torch.stack([m.duration_model._compute_parametric_distribution(s, torch.arange(1, Dm + 1).float())
                              for s in range(K)])
    out = {n: _np(p) for n, p in m.named_parameters()}
    out.update({"x": _np(x), "log_probability": _np(res["log_probability"]),
                "forward_variables": _np(res["forward_variables"]), "log_dur": _np(logdur),
                "vit_states": _np(st), "vit_durations": _np(du), "vit_logprob": np.float32(float(lp))})
    np.savez_compressed(os.path.join(OUT, "semimarkov.npz"), **out)


def semimarkov_supervised(ref):
    """A13, alignment given: SemiMarkovHMM._supervised_forward (semi_markov.py:280-305) with its three terms
    (:385-453, :100-153).  Three sequences: segments that tile T exactly, segments that stop short of T, and segments that overrun T
    (the reference stops adding observation terms at the first segment that does not fit, :404-411)."""
    torch.manual_seed(4201)
    K, D, Dm, T = 4, 5, 8, 20
    out = {}
    for dist in ("gamma", "poisson", "gaussian"):
        m = ref.SemiMarkovHMM(K, D, max_duration=Dm, duration_distribution=dist, min_duration=2 if dist == "gaussian" else 1)
        x = torch.randn(3, T, D)
        states = torch.tensor([[0, 2, 1, 3, 0], [3, 1, 0, 2, 1], [1, 0, 3, 2, 0]])
        durs = torch.tensor([[4, 6, 3, 5, 2], [2, 3, 4, 1, 2], [7, 6, 5, 4, 3]])
        with torch.no_grad():
            res = m(x, states, durs)
        out.update({f"{dist}_{n}": _np(p) for n, p in m.named_parameters()})
        out.update({f"{dist}_x": _np(x), f"{dist}_states": states.numpy(), f"{dist}_durs": durs.numpy()})
        out.update({f"{dist}_{k}": _np(v) for k, v in res.items()})
    np.savez_compressed(os.path.join(OUT, "semimarkov_sup.npz"), **out)


def streaming(ref):
    """A14: StreamingHMMProcessor greedy path (streaming.py:183-320), two consecutive chunks."""
    torch.manual_seed(5101)
    p = ref.StreamingHMMProcessor(6, 8, chunk_size=16, overlap_size=4, lookahead_frames=2,
                                  max_delay_frames=64, use_beam_search=False)
    p.eval()
    feats = torch.randn(48, 8)
    with torch.no_grad():
        logb = p.emission_net(feats)
        r1 = p.process_chunk(feats[:24])
        r2 = p.process_chunk(feats[24:48])
        s1, c1 = r1.decoded_states, r1.confidence
        s2, c2 = r2.decoded_states, r2.confidence
    out = {n.replace(".", "__"): _np(v) for n, v in p.state_dict().items()}
    out.update({"feats": _np(feats), "logb": _np(logb), "chunk1_states": _np(s1), "chunk2_states": _np(s2),
                "chunk1_conf": np.float32(c1), "chunk2_conf": np.float32(c2),
                "chunk1_status": np.array(r1.status), "chunk2_status": np.array(r2.status),
                "chunk1_frames": np.int64(r1.metadata["frames_processed"]),
                "chunk2_frames": np.int64(r2.metadata["frames_processed"])})
    np.savez_compressed(os.path.join(OUT, "streaming.npz"), **out)


def segsum_probe():
    """Pins the summation order of torch.sum on a strided fp32 slice (used by HSMMLayer, hsmm.py:266,285)."""
    g = torch.Generator().manual_seed(4201)
    x = (50 * torch.randn(64, 5, generator=g) - 100).float()
    sums = np.zeros((40, 21), np.float32)
    for t in range(40):
        for d in range(1, 21):
            sums[t, d] = float(torch.sum(x[t:t + d, 2]))
    np.savez_compressed(os.path.join(OUT, "segsum.npz"), x=_np(x), sums=sums)


def largek(ref):
    """A2/A3 at K > 32 (BASELINE config 5 family): K = 64 skip-left-to-right and K = 512 ergodic on
    softmax(randn) observations (examples/benchmark.py:160-162)."""
    from pytorch_hmm.utils import create_transition_matrix
    out = {}
    g = torch.Generator().manual_seed(5001)
    for tag, K, kind, B, T in (("k64", 64, "left_to_right_skip", 3, 40), ("k512", 512, "ergodic", 2, 24)):
        P = create_transition_matrix(K, kind)
        obs = torch.softmax(torch.randn(B, T, K, generator=g), dim=-1)
        hmm = ref.HMMPyTorch(P, None)
        post, fwd, bwd = hmm.forward_backward(obs)
        states, delta = hmm.viterbi_decode(obs)
        out.update({f"{tag}_P": _np(P), f"{tag}_obs": _np(obs), f"{tag}_log_P": _np(hmm.log_P), f"{tag}_log_p0": _np(hmm.log_p0),
                    f"{tag}_log_obs": _np(torch.log(obs + 1e-8)),   # hmm.py:152, on the generating machine's ATen
                    f"{tag}_posterior": _np(post), f"{tag}_forward": _np(fwd), f"{tag}_backward": _np(bwd),
                    f"{tag}_states": _np(states), f"{tag}_log_delta": _np(delta),
                    f"{tag}_likelihood": _np(hmm.compute_likelihood(obs))})
    np.savez_compressed(os.path.join(OUT, "largek.npz"), **out)


def neural(ref):
    """NeuralHMM (neural.py:355-519).  The observation / transition networks are torch modules outside the
    path; the fixture stores THEIR outputs (log-emissions, per-frame transition probabilities) as inputs of the recursion and the
    class's own forward / viterbi_decode / compute_likelihood results."""
    from pytorch_hmm.neural import NeuralHMM
    import torch.nn.functional as F
    out = {}
    for tag, K, D, ctx, B, T, seed in (("tv", 6, 8, 4, 3, 40, 7101), ("static", 5, 8, 0, 2, 25, 7102), ("tv12", 12, 10, 6, 2, 300, 7103)):
        torch.manual_seed(seed)
        m = NeuralHMM(num_states=K, observation_dim=D, context_dim=ctx, hidden_dim=16)
        m.eval()
        x = torch.randn(B, T, D)
        c = torch.randn(B, T, ctx) if ctx > 0 else None
        with torch.no_grad():
            post, fwd, bwd = m(x, c)
            states, delta = m.viterbi_decode(x, c)
            ll = m.compute_likelihood(x, c)
            log_obs = m.observation_model(x)
            if ctx > 0:
                log_trans = torch.log(m.transition_model(c) + 1e-8)
            else:
                log_trans = torch.log(F.softmax(m.transition_matrix, dim=1) + 1e-8)
            log_init = torch.log(F.softmax(m.initial_logits, dim=0) + 1e-8)
        out.update({f"{tag}_log_obs": _np(log_obs), f"{tag}_log_trans": _np(log_trans), f"{tag}_log_init": _np(log_init),
                    f"{tag}_posterior": _np(post), f"{tag}_forward": _np(fwd), f"{tag}_backward": _np(bwd),
                    f"{tag}_states": _np(states), f"{tag}_log_delta": _np(delta), f"{tag}_likelihood": _np(ll)})
    np.savez_compressed(os.path.join(OUT, "neural.npz"), **out)


def alignment(ref):
    """CTC forward / backward trellises (alignment/ctc.py:32-199) and DTW (alignment/dtw.py:47-153)."""
    from pytorch_hmm.alignment.ctc import ctc_alignment_path, ctc_backward_algorithm, ctc_forward_algorithm
    from pytorch_hmm.alignment.dtw import compute_distance_matrix, compute_dtw_path
    out = {}
    g = torch.Generator().manual_seed(6101)
    # a: ragged batch (one full-length utterance, one short, one with an empty target, repeated labels)
    T, B, C, L = 24, 4, 7, 5
    lp = torch.log_softmax(2.0 * torch.randn(T, B, C, generator=g), dim=-1)
    targets = torch.tensor([[1, 2, 2, 3, 6], [4, 4, 4, 1, 0], [0, 0, 0, 0, 0], [5, 1, 5, 1, 5]])
    in_len = torch.tensor([24, 17, 9, 24])
    tg_len = torch.tensor([5, 4, 0, 5])
    # b: blank id in the middle of the alphabet, target that does not fit the utterance (log-likelihood -inf)
    T2, B2, C2, L2 = 10, 2, 5, 6
    lp2 = torch.log_softmax(torch.randn(T2, B2, C2, generator=g), dim=-1)
    targets2 = torch.tensor([[0, 1, 3, 3, 4, 0], [1, 1, 1, 1, 1, 1]])
    in_len2 = torch.tensor([10, 7])
    tg_len2 = torch.tensor([6, 6])
    for tag, args, blank in (("a", (lp, targets, in_len, tg_len), 0), ("b", (lp2, targets2, in_len2, tg_len2), 2)):
        ll = ctc_forward_algorithm(*args, blank_id=blank)
        lb = ctc_backward_algorithm(*args, blank_id=blank)
        al = ctc_alignment_path(*args, blank_id=blank)
        out.update({f"ctc_{tag}_log_probs": _np(args[0]), f"ctc_{tag}_targets": _np(args[1]), f"ctc_{tag}_input_lengths": _np(args[2]),
                    f"ctc_{tag}_target_lengths": _np(args[3]), f"ctc_{tag}_blank": np.int64(blank), f"ctc_{tag}_loglik": _np(ll),
                    f"ctc_{tag}_log_beta": _np(lb)})
        for b, a in enumerate(al):
            out[f"ctc_{tag}_align{b}"] = _np(a)
    # torch's own CTC loss on (a): an independent check of the forward log-likelihood's value
    flat = torch.cat([targets[b, : tg_len[b]] for b in range(B)])
    out["ctc_a_torch_nll"] = _np(torch.nn.functional.ctc_loss(lp, flat, in_len, tg_len, blank=0, reduction="none"))
    # DTW: euclidean distance matrices, the three step patterns; one matrix with exact ties (integers)
    x, y = torch.randn(23, 6, generator=g), torch.randn(31, 6, generator=g)
    d = compute_distance_matrix(x, y, "euclidean")
    ties = torch.randint(0, 3, (17, 12), generator=g).float()
    for tag, mat in (("rand", d), ("ties", ties)):
        out[f"dtw_{tag}_dist"] = _np(mat)
        for pat in ("symmetric", "asymmetric", "rabiner_juang"):
            pi, pj, cost = compute_dtw_path(mat, pat)
            out.update({f"dtw_{tag}_{pat}_path_i": _np(pi), f"dtw_{tag}_{pat}_path_j": _np(pj), f"dtw_{tag}_{pat}_cost": _np(cost)})
    out["dtw_x"], out["dtw_y"] = _np(x), _np(y)
    np.savez_compressed(os.path.join(OUT, "alignment.npz"), **out)


def covariance(ref):
    """the other covariance types of MixtureGaussianHMMLayer -- tied, spherical (mixture_gaussian.py:242-269) and
    full (Cholesky parameters, :216-240, :271-288): emission log-likelihoods and the layer's Viterbi on them."""
    out = {}
    for tag, seed in (("tied", 7001), ("spherical", 7002), ("full", 7003)):
        torch.manual_seed(seed)
        K, C, D = 6, 3, 16
        m = ref.MixtureGaussianHMMLayer(K, D, num_components=C, covariance_type=tag)
        with torch.no_grad():
            m.means.mul_(3.0)
            if tag == "full":
                m.cholesky_params.add_(0.15 * torch.randn_like(m.cholesky_params))
            else:
                m.log_vars.copy_(0.3 * torch.randn_like(m.log_vars))
            m.mixture_weights_logits.copy_(torch.randn_like(m.mixture_weights_logits))
        B, T = 2, 30
        s = torch.randint(0, K, (B, T)); c = torch.randint(0, C, (B, T))
        x = m.means.detach()[s, c] + 0.8 * torch.randn(B, T, D)
        m.eval()
        with torch.no_grad():
            logb = m.get_observation_log_probs(x)
            states, scores = m(x, return_log_probs=True)
            log_trans = m._safe_log(m.get_transition_matrix())
        for name, par in m.state_dict().items():
            out[f"{tag}_sd_{name}"] = _np(par)
        out.update({f"{tag}_x": _np(x), f"{tag}_logb": _np(logb), f"{tag}_log_trans": _np(log_trans),
                    f"{tag}_states": _np(states), f"{tag}_scores": _np(scores)})
    np.savez_compressed(os.path.join(OUT, "covariance.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = _import_reference()
    torch.set_num_threads(1)
    sections = {"core": core, "gaussian": gaussian, "mixture": mixture, "hsmm": hsmm, "semimarkov": semimarkov, "semimarkov_sup": semimarkov_supervised,
                "streaming": streaming, "largek": largek, "neural": neural, "alignment": alignment, "covariance": covariance}
    only = [a for a in sys.argv[1:] if a in sections or a == "segsum"]        # e.g. `make_golden.py largek`
    for name, fn in sections.items():
        if not only or name in only:
            fn(ref)
    if not only or "segsum" in only:
        segsum_probe()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
