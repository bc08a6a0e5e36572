"""CPU tests (-m "not gpu"): host-side logic of the drop-in classes, and that the C-ABI library loads and exports
every symbol include/hmm_b200.h declares.  No compute call is made (there is no GPU here and no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import pytorch_hmm_b200 as hm
from pytorch_hmm_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hmm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hmmb200_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    path = build.build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _declared_symbols()
    assert len(names) >= 9
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hmm_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"
    assert _lib.load().hmmb200_abi_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_fails_loudly():
    lib = _lib.load()
    assert lib.hmmb200_device_check(-1) == -5                       # HMMB200_ENODEVICE
    assert b"no CUDA device" in lib.hmmb200_last_error()
    hmm = hm.HMMPyTorch(torch.eye(3) + 0.1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hmm.forward_backward(torch.rand(2, 4, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hm.MixtureGaussianHMMLayer(3, 4)(torch.randn(1, 5, 4))


def test_size_queries_need_no_device():
    lib = _lib.load()
    # fp32 section (D * pairs * 4 + 2 * pairs) + tensor-core section (flag/centre/const + four fp16 W matrices)
    assert lib.hmmb200_gmm_packed_floats(12, 4, 80) == (80 * 24 * 4 + 48) + (4 + 80 + 48 + 2 * 48 * 80)
    assert lib.hmmb200_gmm_packed_floats(3, 5, 7) == ((7 * 8 * 4 + 16 + 3) // 4) * 4        # D % 4 != 0: fp32 section only
    assert lib.hmmb200_fb_workspace_bytes(256, 2000, 12) >= 256 * 2000 * (2 * 12 + 2) * 4
    assert lib.hmmb200_viterbi_workspace_bytes(256, 2000, 12) == 0          # backpointers fit in shared memory
    assert lib.hmmb200_viterbi_workspace_bytes(2, 9000, 16) == 2 * 9000 * 16


def test_hmm_init_matches_reference(golden):
    g = golden("core")
    for tag in "abcde":
        p0 = torch.from_numpy(g[f"{tag}_p0"]) if f"{tag}_p0" in g.files else None
        h = hm.HMMPyTorch(torch.from_numpy(g[f"{tag}_P"]), p0)
        assert np.array_equal(h.log_P.numpy(), g[f"{tag}_log_P"])
        assert np.array_equal(h.log_p0.numpy(), g[f"{tag}_log_p0"])
    with pytest.raises(ValueError):
        hm.HMM(torch.ones(2, 3))
    with pytest.raises(ValueError):
        hm.HMM(torch.ones(3, 3), torch.ones(4))
    h = hm.HMM(np.ones((3, 3), np.float32))                                  # numpy input accepted (hmm.py:24-25)
    assert h.K == 3


def test_transition_builders_match_reference(golden):
    g = golden("core")
    assert np.array_equal(hm.create_left_to_right_matrix(10, 0.7).numpy(), g["b_P"])
    assert np.array_equal(hm.create_transition_matrix(6, "ergodic").numpy(), g["c_P"])
    assert np.array_equal(hm.create_transition_matrix(5, "left_to_right_skip").numpy(), g["d_P"])
    P = hm.create_transition_matrix(4, "circular")
    assert torch.allclose(P.sum(1), torch.ones(4))
    with pytest.raises(ValueError):
        hm.create_transition_matrix(4, "nope")


def test_layer_parameter_names_and_state_dict_compat(golden):
    hl = hm.HMMLayer(7)
    assert set(hl.state_dict()) == {"log_transition_logits", "log_initial_logits"}
    assert set(hm.HMMLayer(5, learnable_transitions=False).state_dict()) == {"transition_matrix", "log_initial_logits"}
    gl = hm.GaussianHMMLayer(4, 6)
    assert set(gl.state_dict()) == {"means", "log_scales", "hmm_layer.log_transition_logits", "hmm_layer.log_initial_logits"}
    mg = hm.MixtureGaussianHMMLayer(12, 80, num_components=4)
    assert set(mg.state_dict()) == {"transition_logits", "mixture_weights_logits", "means", "log_vars"}
    assert mg.means.shape == (12, 4, 80) and mg.log_vars.shape == (12, 4, 80)
    g = golden("gaussian")
    hl.load_state_dict({"log_transition_logits": torch.from_numpy(g["hl_log_transition_logits"]),
                        "log_initial_logits": torch.from_numpy(g["hl_log_initial_logits"])})
    assert torch.allclose(hl.get_transition_matrix().sum(1), torch.ones(7), atol=1e-6)
    fixed = hm.MixtureGaussianHMMLayer(4, 3, learnable_transitions=False)
    assert fixed.get_transition_matrix()[0, 0].item() == pytest.approx(0.8)
    with pytest.raises(ValueError):
        hm.MixtureGaussianHMMLayer(4, 3, covariance_type="bogus")
    with pytest.raises(ValueError):
        hm.GaussianHMMLayer(4, 3, covariance_type="bogus")


def test_hsmm_posterior_oracle_vs_brute_force():
    """HSMM backward / posteriors have no reference implementation: pin the float64 oracle by enumerating every
    segmentation at tiny sizes."""
    import numpy as np
    from oracle import hsmm_post
    rng = np.random.default_rng(41)
    for K, Dm, T in ((2, 3, 6), (3, 2, 5), (3, 4, 7)):
        f = rng.standard_normal((T, K)) * 2 - 3
        segc = rng.standard_normal(K) - 1
        logdur = np.log(rng.random((K, Dm)) + 0.05)
        A = rng.random((K, K)) + 0.05
        np.fill_diagonal(A, 0.0)
        logA = np.log(A / A.sum(1, keepdims=True) + 1e-8)
        logpi = np.log(rng.dirichlet(np.ones(K)))
        g1, t1 = hsmm_post.posteriors_f64(f, segc, logdur, logA, logpi)
        g2, t2 = hsmm_post.brute_force(f, segc, logdur, logA, logpi)
        np.testing.assert_allclose(t1, t2, rtol=1e-12)
        np.testing.assert_allclose(g1, g2, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(g1.sum(-1), 1.0, rtol=1e-9)


def test_baum_welch_oracle_vs_brute_force():
    """Baum-Welch (A9) has formulas only in the reference (docs/01_hmm_theory.md:196-227): pin the float64 E-step oracle
    (oracle/hmm_oracle.c::orc_bw_stats_f64) by first principles -- enumerate every state path of a tiny GMM-HMM, weight it by its
    joint probability with the observations, and count: first-frame posteriors, transition counts, component occupancies and the
    occupancy-weighted first / second moments."""
    import itertools
    import numpy as np
    from oracle import c_oracle
    rng = np.random.default_rng(97)
    for B, T, K, Cn, D in ((2, 4, 3, 2, 2), (1, 5, 2, 3, 1), (3, 1, 3, 1, 2), (1, 6, 3, 1, 3)):
        x = rng.standard_normal((B, T, D)).astype(np.float32)
        means = rng.standard_normal((K, Cn, D)); var = np.exp(0.3 * rng.standard_normal((K, Cn, D)))
        w = rng.dirichlet(np.ones(Cn) * 3, size=K); P = rng.dirichlet(np.ones(K) * 2, size=K); p0 = rng.dirichlet(np.ones(K) * 2)
        xd = x.astype(np.float64)
        comp = np.log(w)[None, None] - 0.5 * (((xd[:, :, None, None, :] - means[None, None]) ** 2 / var[None, None]).sum(-1)
                                              + np.log(var).sum(-1)[None, None] + D * np.log(2 * np.pi))          # [B,T,K,C]
        got = c_oracle.bw_stats_f64(x, comp, np.log(P), np.log(p0))
        dens = np.exp(comp)                                      # joint density of (x_t, component c) given state k
        b = dens.sum(-1)                                         # [B,T,K]
        resp = dens / b[..., None]
        g1 = np.zeros(K); xi = np.zeros((K, K)); occ = np.zeros((K, Cn)); sx = np.zeros((K, Cn, D)); sxx = np.zeros((K, Cn, D)); ll = 0.0
        for u in range(B):
            paths = list(itertools.product(range(K), repeat=T))
            wts = np.array([p0[s[0]] * np.prod([P[s[t - 1], s[t]] for t in range(1, T)]) * np.prod([b[u, t, s[t]] for t in range(T)])
                            for s in paths])
            Z = wts.sum(); ll += np.log(Z)
            for s, wt in zip(paths, wts / Z):
                g1[s[0]] += wt
                for t in range(T):
                    if t:
                        xi[s[t - 1], s[t]] += wt
                    o = wt * resp[u, t, s[t]]                    # [C]
                    occ[s[t]] += o
                    sx[s[t]] += o[:, None] * xd[u, t][None]
                    sxx[s[t]] += o[:, None] * (xd[u, t] ** 2)[None]
        np.testing.assert_allclose(got["loglik"], ll, rtol=1e-12)
        for name, want in (("gamma1", g1), ("xi", xi), ("occ", occ), ("sx", sx), ("sxx", sxx)):
            np.testing.assert_allclose(got[name], want, rtol=1e-9, atol=1e-12, err_msg=f"{name} B={B} T={T} K={K} C={Cn}")


def test_duration_model_forward_matches_reference(golden):
    """DurationModel.forward (semi_markov.py:63-153), host side: log p_{s_i}(d_i) summed per sequence against the reference's
    'log_duration' of the supervised SemiMarkovHMM fixture (a duration below min_duration gives -inf), and the no-durations form
    against the rows of the table."""
    import numpy as np
    import pytorch_hmm_b200 as hm
    g = golden("semimarkov_sup")
    for dist in ("gamma", "poisson", "gaussian"):
        dm = hm.DurationModel(4, 8, dist, min_duration=2 if dist == "gaussian" else 1)
        dm.load_state_dict({k[len("duration_model."):]: torch.from_numpy(g[f"{dist}_{k}"])
                            for k in (f[len(dist) + 1:] for f in g.files if f.startswith(f"{dist}_duration_model."))})
        st, du = torch.from_numpy(g[f"{dist}_states"]), torch.from_numpy(g[f"{dist}_durs"])
        with torch.no_grad():
            got = dm(st.flatten(), du.flatten()).view(st.shape).sum(1).numpy()
            rows = dm(st[0])
        ref = g[f"{dist}_log_duration"]
        assert np.array_equal(np.isfinite(got), np.isfinite(ref))
        np.testing.assert_allclose(got[np.isfinite(ref)], ref[np.isfinite(ref)], rtol=1e-5)
        assert torch.equal(rows, dm.log_table()[st[0]])
