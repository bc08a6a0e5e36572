"""GPU parity tests (-m gpu) of the alignment kernels (csrc/alignment.cu) against the golden fixture of the real
reference (tests/golden/alignment.npz, oracle/make_golden.py::alignment) and the numpy oracle (oracle/alignment_port.py).
Bars: DTW cost matrices and paths BIT-EXACT (fp32 adds and minima only); CTC trellises within 1e-5 relative (fp32 log-sum-exp: the
device's expf/logf are not ATen's), identical -inf pattern."""
import numpy as np
import pytest
import torch

from oracle import alignment_port as ap

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def al():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pytorch_hmm_b200.alignment as m
    return m


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _close(a, b, tol):
    a, b = np.asarray(a), np.asarray(b)
    assert np.array_equal(np.isfinite(a), np.isfinite(b)), "the -inf pattern differs"
    m = np.isfinite(b)
    err = np.abs(a[m] - b[m]) / np.maximum(1.0, np.abs(b[m]))
    assert err.size == 0 or err.max() <= tol, err.max()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_ctc_trellises_vs_reference_golden(al, golden, tag):
    g = golden("alignment")
    args = (_t(g[f"ctc_{tag}_log_probs"]), _t(g[f"ctc_{tag}_targets"]), _t(g[f"ctc_{tag}_input_lengths"]), _t(g[f"ctc_{tag}_target_lengths"]))
    blank = int(g[f"ctc_{tag}_blank"])
    ll = al.ctc_forward_algorithm(*args, blank_id=blank)
    _close(ll.cpu().numpy(), g[f"ctc_{tag}_loglik"], 1e-5)
    lb = al.ctc_backward_algorithm(*args, blank_id=blank)
    assert lb.shape == g[f"ctc_{tag}_log_beta"].shape
    _close(lb.cpu().numpy(), g[f"ctc_{tag}_log_beta"], 1e-5)
    paths = al.ctc_alignment_path(*args, blank_id=blank)
    for b, p in enumerate(paths):
        assert np.array_equal(p.cpu().numpy(), g[f"ctc_{tag}_align{b}"])
    if tag == "a":
        _close(-ll.cpu().numpy(), g["ctc_a_torch_nll"], 1e-5)


def test_ctc_larger_random_batch_vs_oracle(al):
    rng = np.random.default_rng(7)
    T, B, C, L = 150, 6, 40, 30
    lp = torch.log_softmax(torch.from_numpy(rng.standard_normal((T, B, C)).astype(np.float32)) * 2, dim=-1).numpy()
    tg = rng.integers(1, C, size=(B, L))
    tg[2, 3:9] = tg[2, 3]                                     # repeated labels: the skip transition is forbidden
    il = np.array([150, 150, 120, 77, 61, 150])
    tl = np.array([30, 12, 25, 30, 0, 1])
    la_o, ll_o = ap.ctc_forward(lp, tg, il, tl, 0)
    lb_o = ap.ctc_backward(lp, tg, il, tl, 0)
    from pytorch_hmm_b200.alignment.ctc import ctc_forward_trellis
    la, ll = ctc_forward_trellis(_t(lp), _t(tg), _t(il), _t(tl), 0)
    _close(la.cpu().numpy(), la_o, 2e-5)
    _close(ll.cpu().numpy(), ll_o, 2e-5)
    lb = al.ctc_backward_algorithm(_t(lp), _t(tg), _t(il), _t(tl), 0)
    _close(lb.cpu().numpy(), lb_o, 2e-5)
    # alpha_t(s) + beta_t(s) summed over s is the log-likelihood at every live frame (a size-independent property of the pair)
    tot = torch.logsumexp(la + lb, dim=-1).cpu().numpy()
    for b in range(B):
        if np.isfinite(ll_o[b]):
            assert np.allclose(tot[b, : il[b]], ll_o[b], rtol=2e-5), b
    # and it equals minus torch's own CTC loss
    flat = torch.cat([torch.from_numpy(tg[b, : tl[b]]) for b in range(B)])
    nll = torch.nn.functional.ctc_loss(torch.from_numpy(lp), flat, torch.from_numpy(il), torch.from_numpy(tl), blank=0, reduction="none")
    _close(-ll.cpu().numpy(), nll.numpy(), 2e-5)
    # the posterior alignment visits the expanded target monotonically
    paths = al.ctc_posterior_alignment(_t(lp), _t(tg), _t(il), _t(tl), 0)
    assert len(paths) == B and all(len(p) == il[b] for b, p in enumerate(paths))


def test_ctc_long_target_more_positions_than_threads(al):
    rng = np.random.default_rng(8)
    T, B, C, L = 40, 2, 9, 600                                # 2L+1 = 1201 expanded positions > 1024 threads
    lp = torch.log_softmax(torch.from_numpy(rng.standard_normal((T, B, C)).astype(np.float32)), dim=-1).numpy()
    tg = rng.integers(1, C, size=(B, L))
    il, tl = np.array([40, 33]), np.array([15, 600])
    lb = al.ctc_backward_algorithm(_t(lp), _t(tg), _t(il), _t(tl), 0)
    _close(lb.cpu().numpy(), ap.ctc_backward(lp, tg, il, tl, 0), 2e-5)
    ll = al.ctc_forward_algorithm(_t(lp), _t(tg), _t(il), _t(tl), 0)
    _close(ll.cpu().numpy(), ap.ctc_forward(lp, tg, il, tl, 0)[1], 2e-5)


@pytest.mark.parametrize("tag", ["rand", "ties"])
@pytest.mark.parametrize("pattern", ["symmetric", "asymmetric", "rabiner_juang"])
def test_dtw_bit_exact_vs_reference_golden(al, golden, tag, pattern):
    g = golden("alignment")
    pi, pj, cost = al.compute_dtw_path(_t(g[f"dtw_{tag}_dist"]), pattern)
    assert np.array_equal(cost.cpu().numpy(), g[f"dtw_{tag}_{pattern}_cost"])
    assert np.array_equal(pi.cpu().numpy(), g[f"dtw_{tag}_{pattern}_path_i"])
    assert np.array_equal(pj.cpu().numpy(), g[f"dtw_{tag}_{pattern}_path_j"])


def test_dtw_larger_and_batched(al):
    rng = np.random.default_rng(9)
    x = torch.from_numpy(rng.standard_normal((3, 300, 16)).astype(np.float32)).cuda()
    y = torch.from_numpy(rng.standard_normal((3, 421, 16)).astype(np.float32)).cuda()
    aligner = al.DTWAligner()
    pis, pjs, costs = aligner(x, y)
    for b in range(3):
        d = al.compute_distance_matrix(x[b], y[b]).cpu().numpy()
        pi_o, pj_o, cost_o = ap.dtw(d, "symmetric")
        assert np.array_equal(pis[b].cpu().numpy(), pi_o) and np.array_equal(pjs[b].cpu().numpy(), pj_o)
        assert float(costs[b]) == float(cost_o[-1, -1])
    # quantised distances: many exact ties, the tie order decides the path
    d = torch.from_numpy(rng.integers(0, 4, size=(257, 129)).astype(np.float32)).cuda()
    for pattern in ("symmetric", "rabiner_juang"):
        pi, pj, cost = al.compute_dtw_path(d, pattern)
        pi_o, pj_o, cost_o = ap.dtw(d.cpu().numpy(), pattern)
        assert np.array_equal(cost.cpu().numpy(), cost_o)
        assert np.array_equal(pi.cpu().numpy(), pi_o) and np.array_equal(pj.cpu().numpy(), pj_o)
    # path properties at a size the Python oracle does not reach: monotone unit steps from (0,0) to (N-1,M-1); dtw(x,x) = 0 on the diagonal
    big = torch.rand(1500, 2000, device="cuda")
    pi, pj, cost = al.compute_dtw_path(big)
    assert int(pi[0]) == 0 and int(pj[0]) == 0 and int(pi[-1]) == 1499 and int(pj[-1]) == 1999
    di, dj = (pi[1:] - pi[:-1]), (pj[1:] - pj[:-1])
    assert bool(((di >= 0) & (di <= 1) & (dj >= 0) & (dj <= 1) & (di + dj >= 1)).all())
    assert float(cost[-1, -1]) == pytest.approx(float(big[pi, pj].sum()), rel=1e-5)
    z = torch.randn(200, 8, device="cuda")
    pi, pj, c = al.dtw_alignment(z, z)
    assert float(c) == 0.0 and torch.equal(pi, pj)
