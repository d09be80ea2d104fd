"""The merge rule of the split-key attention launch (csrc/attn_tc.cu: nsplit, k_attn_combine), pinned on the CPU in float64.

Each CTA of a split launch returns, per query row, the un-normalised output O_s = sum_k P_sk V_k and the row sum l_s = sum_k P_sk
with P_sk = 2^(c S_k - m_s), where m_s is the kernel's *lazy* reference maximum of that key range: any value within 2^8 of the true
maximum, not necessarily the maximum itself.  k_attn_combine computes (w_0 O_0 + w_1 O_1) / (w_0 l_0 + w_1 l_1) with
w_s = 2^(m_s - max(m_0, m_1)).  That must equal softmax(S / sqrt(d)) V (AttentionBlock.forward, HYB:292-305) for ANY finite m_s."""
import numpy as np


def _partial(S, V, c, m):
    P = np.exp2(c * S - m[:, None])
    return P @ V, P.sum(axis=1)


def test_merge_of_two_key_ranges_equals_the_full_softmax_for_any_reference_maxima():
    rng = np.random.default_rng(5)
    nq, nk, d = 32, 256, 96
    c = d ** -0.5 * np.log2(np.e)
    for peaked in (False, True):
        S = rng.standard_normal((nq, nk)) * (8.0 if peaked else 1.0)
        if peaked:
            S *= np.linspace(0.2, 4.0, nk)[None, :]            # the second key range scores far higher or lower
        V = rng.standard_normal((nk, d))
        A = np.exp(S / np.sqrt(d) - (S / np.sqrt(d)).max(axis=1, keepdims=True))
        ref = (A / A.sum(axis=1, keepdims=True)) @ V
        h = nk // 2
        for slack in (0.0, 3.0, 7.9):                          # lazy maxima: up to 2^8 below the true one (in exp2 units)
            m0 = c * S[:, :h].max(axis=1) - slack * rng.random(nq)
            m1 = c * S[:, h:].max(axis=1) - slack * rng.random(nq)
            O0, l0 = _partial(S[:, :h], V[:h], c, m0)
            O1, l1 = _partial(S[:, h:], V[h:], c, m1)
            M = np.maximum(m0, m1)
            w0, w1 = np.exp2(m0 - M), np.exp2(m1 - M)
            out = (w0[:, None] * O0 + w1[:, None] * O1) / (w0 * l0 + w1 * l1)[:, None]
            assert np.abs(out - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())
