"""CPU checks of the mathematical reformulations the CUDA kernels rely on (numpy emulations of the kernels'
algorithms against the oracle's direct restatement of the reference).  They document and pin:
  * V1: dual-softmax mutual-NN from log-sum-exp statistics, no L x S confidence matrix;
  * the tile-partial merge of (reference, sum) pairs incl. the shared-reference fast path and its underflow
    fallback rule used by EpiScoreLse;
  * V3: linear attention without the /S ... *S rescale;
  * V4: window gather == F.unfold + rearrange;
  * the analytic dustbin handling of the Sinkhorn sweeps;
  * the fp16 hi/lo split with power-of-two weight pre-scaling.
"""
import numpy as np
import pytest
import torch

from cases import build_cfg
from oracle import loftr_oracle as O

F32 = np.float32


def _lse_partials_emulation(sim, tile_m=128, tile_n=256, blk=32, shared_ref=True):
    """numpy emulation of EpiScoreLse: per (32-row x 32-col) block a (reference, sum) partial for rows and columns,
    merged exactly like lse_merge_kernel.  Returns (rowLSE, colLSE, number of blocks that took the fallback)."""
    L, S = sim.shape
    row_m = np.full(L, -1e30, F32); row_l = np.zeros(L, F32)
    col_m = np.full(S, -1e30, F32); col_l = np.zeros(S, F32)
    fallbacks = 0
    for i0 in range(0, L, blk):
        for j0 in range(0, S, blk):
            z = sim[i0:i0 + blk, j0:j0 + blk].astype(F32)
            use_shared = shared_ref and z.shape == (blk, blk)
            if use_shared:
                g = z.max()
                e = np.exp((z - g).astype(F32)).astype(F32)
                e[e < 1.18e-38] = 0.0                       # ex2.approx.ftz flushes denormals
                rs, cs = e.sum(1), e.sum(0)
                if (rs >= 1e-30).all() and (cs >= 1e-30).all():
                    rref, cref = np.full(z.shape[0], g, F32), np.full(z.shape[1], g, F32)
                else:
                    use_shared = False
                    fallbacks += 1
            if not use_shared:
                rref, cref = z.max(1), z.max(0)
                rs = np.exp(z - rref[:, None]).sum(1).astype(F32)
                cs = np.exp(z - cref[None, :]).sum(0).astype(F32)
            for (m, l, ref, s, sl) in ((row_m, row_l, rref, rs, slice(i0, i0 + blk)), (col_m, col_l, cref, cs, slice(j0, j0 + blk))):
                mn = np.maximum(m[sl], ref)
                l[sl] = l[sl] * np.exp(m[sl] - mn) + s * np.exp(ref - mn)
                m[sl] = mn
    return row_m + np.log(row_l), col_m + np.log(col_l), fallbacks


def _matches_from_lse(sim, row_lse, col_lse, thr, border, hw0, hw1):
    """EpiScoreArgmax + match_flag_kernel + ordered compaction, in numpy."""
    j_star = (2 * sim - col_lse[None, :]).argmax(1)
    i_star = (2 * sim - row_lse[:, None]).argmax(0)
    out = []
    (h0, w0), (h1, w1) = hw0, hw1
    for i, j in enumerate(j_star):
        if i_star[j] != i:
            continue
        y0, x0, y1, x1 = i // w0, i % w0, j // w1, j % w1
        if border > 0 and not (border <= y0 < h0 - border and border <= x0 < w0 - border and
                               border <= y1 < h1 - border and border <= x1 < w1 - border):
            continue
        conf = np.exp(2 * sim[i, j] - row_lse[i] - col_lse[j])
        if conf > thr:
            out.append((i, j, conf))
    return out


@pytest.mark.parametrize("spread", ["benign", "huge"])
def test_two_pass_dual_softmax_equals_reference_formulation(spread):
    rs = np.random.RandomState(3)
    h0, w0, h1, w1, c = 12, 16, 10, 16, 64
    L, S = h0 * w0, h1 * w1
    f0 = rs.standard_normal((1, L, c)).astype(F32) * (1.5 if spread == "benign" else 6.0)
    f1 = rs.standard_normal((1, S, c)).astype(F32) * (1.5 if spread == "benign" else 6.0)
    k = 100
    src, dst = rs.permutation(L)[:k], rs.permutation(S)[:k]
    f1[0, dst] = f0[0, src] * rs.uniform(0.3, 1.2, (k, 1)).astype(F32)
    if spread == "huge":
        f0[0, ::5] *= 0.01
    cfg = dict(build_cfg({"thr": 0.05})["match_coarse"])
    ref = O.coarse_matching(f0, f1, cfg, (h0 * 8, w0 * 8), (h0, w0), (h1, w1))
    sim = ((f0[0] / np.sqrt(c)) @ (f1[0] / np.sqrt(c)).T / F32(0.1)).astype(F32)
    row_lse, col_lse, fallbacks = _lse_partials_emulation(sim)
    if spread == "huge":
        assert sim.max() - sim.min() > 150 and fallbacks > 0, "the case must exercise the underflow fallback"
    else:
        assert fallbacks == 0
    np.testing.assert_allclose(row_lse, O._logsumexp(sim, 1), rtol=0, atol=2e-4)
    np.testing.assert_allclose(col_lse, O._logsumexp(sim, 0), rtol=0, atol=2e-4)
    got = _matches_from_lse(sim, row_lse, col_lse, cfg["thr"], cfg["border_rm"], (h0, w0), (h1, w1))
    assert [(i, j) for i, j, _ in got] == list(zip(ref["i_ids"].tolist(), ref["j_ids"].tolist()))
    np.testing.assert_allclose([c_ for _, _, c_ in got], ref["mconf"], rtol=1e-3)


def test_shared_reference_without_fallback_would_be_wrong():
    """Why the fallback exists: with a single block reference, a weak row inside a strong block underflows."""
    sim = np.full((32, 32), -200.0, F32)
    sim[0, :] = 0.0                               # one dominant row; rows 1.. are ~200 nats below
    e = np.exp(sim - sim.max())
    assert (e[1:].sum(1) == 0).all()              # their sums vanish relative to the block maximum
    row_lse, _, fallbacks = _lse_partials_emulation(sim, shared_ref=True)
    assert fallbacks == 1
    np.testing.assert_allclose(row_lse[1:], -200.0 + np.log(32.0), atol=1e-4)


def test_linear_attention_rescale_is_a_noop_in_fp32():
    rs = np.random.RandomState(0)
    n, l, s, h, d = 2, 50, 70, 8, 32
    q, k, v = (rs.standard_normal((n, t, h, d)).astype(F32) for t in (l, s, s))
    qm, km = rs.uniform(size=(n, l)) > 0.2, rs.uniform(size=(n, s)) > 0.3
    ref = O.linear_attention(q, k, v, qm, km)
    Q, K = O._elu_feature_map(q) * qm[..., None, None], O._elu_feature_map(k) * km[..., None, None]
    V = v * km[..., None, None]
    KV = np.einsum("nshd,nshv->nhdv", K, V)
    out = np.einsum("nlhd,nhdv->nlhv", Q, KV) / (np.einsum("nlhd,nhd->nlh", Q, K.sum(1)) + 1e-6)[..., None]
    np.testing.assert_allclose(out, ref, rtol=2e-5, atol=2e-6)
    assert np.abs(ref[~qm]).max() == 0           # masked queries produce exactly zero messages


def test_window_gather_equals_unfold():
    rs = np.random.RandomState(1)
    n, c, hc, wc, W, stride = 2, 8, 6, 7, 5, 4
    feat = rs.standard_normal((n, c, hc * stride, wc * stride)).astype(F32)
    unf = torch.nn.functional.unfold(torch.from_numpy(feat), kernel_size=(W, W), stride=stride, padding=W // 2)
    unf = unf.reshape(n, c, W * W, hc * wc).permute(0, 3, 2, 1).numpy()          # n l ww c
    b = np.array([0, 1, 1, 0]); idx = np.array([0, hc * wc - 1, 17, 9])
    got = O.gather_windows(feat, b, idx, wc, W, stride)
    np.testing.assert_array_equal(got, unf[b, idx])


def test_sinkhorn_dustbins_handled_analytically():
    """One Sinkhorn half-iteration over the real block + a scalar dustbin term == the same on the materialised
    (L+1) x (S+1) couplings matrix (how lse_merge_kernel / bin_lse_kernel treat the dustbins)."""
    rs = np.random.RandomState(2)
    L, S, alpha = 30, 41, F32(0.7)
    z = rs.standard_normal((1, L, S)).astype(F32) * 3
    full = O.log_optimal_transport(z, alpha, 3)[0]
    norm = F32(-np.log(L + S))
    log_mu_bin, log_nu_bin = np.log(S) + norm, np.log(L) + norm
    u, v, bu, bv = np.zeros(L, F32), np.zeros(S, F32), F32(0), F32(0)
    for _ in range(3):
        u = norm - np.logaddexp(O._logsumexp(z[0] + v[None, :], 1), alpha + bv)
        bu = log_mu_bin - np.logaddexp(O._logsumexp((alpha + v)[None, :], 1)[0], alpha + bv)
        v = norm - np.logaddexp(O._logsumexp(z[0] + u[:, None], 0), alpha + bu)
        bv = log_nu_bin - np.logaddexp(O._logsumexp((alpha + u)[None, :], 1)[0], alpha + bu)
    np.testing.assert_allclose(z[0] + u[:, None] + v[None, :] - norm, full[:L, :S], rtol=0, atol=2e-5)


def test_split_with_power_of_two_prescale_keeps_22_bits_for_small_weights():
    rs = np.random.RandomState(4)
    w = (rs.standard_normal(100000) * 0.03).astype(F32)            # typical conv / linear weight magnitude

    def split(x):
        hi = x.astype(np.float16)
        lo = (x - hi.astype(F32)).astype(np.float16)
        return hi.astype(np.float64) + lo.astype(np.float64)

    plain = np.abs(split(w) - w.astype(np.float64)).max() / np.abs(w).max()
    e = int(np.floor(np.log2(4096.0 / np.abs(w).max())))
    scaled = np.abs(split(w * F32(2.0 ** e)) * 2.0 ** (-e) - w.astype(np.float64)).max() / np.abs(w).max()
    assert scaled < 3e-7                      # ~2^-22 of the tensor's scale
    assert plain > scaled                     # without pre-scaling the lo halves sit in fp16's subnormal range
