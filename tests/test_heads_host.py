"""Host-side logic of the engine transfer heads (rgcn_b200/heads.py) that needs no GPU: operand layouts, the padded
K | V projection weights, the dropout-mask shape.  The kernels themselves are covered by the `-m gpu` tests."""
import torch

from rgcn_b200 import heads as H


def test_rows16_and_padded_views():
    a = torch.arange(5 * 63, dtype=torch.float32).reshape(5, 63)
    r = H.rows16(a)
    assert r.shape == (5, 63) and r.stride() == (64, 1) and torch.equal(r, a)
    base = r._base if r._base is not None else r
    assert base.shape == (5, 64) and float(base[:, 63].abs().max()) == 0.0           # zero padding
    assert H.rows16(r) is r                                                          # already addressable: no copy
    p = H.padded_like(7, 137, 'cpu')
    assert p.shape == (7, 137) and p.stride() == (140, 1)
    e = torch.randn(3, 4, 63)
    st = H.stack_rows16(e)
    assert st.shape == (3, 4, 63) and st.stride() == (4 * 64, 64, 1) and torch.equal(st, e)
    flat = H._flat_rows16(st)
    assert flat.shape == (12, 63) and flat.stride() == (64, 1) and flat.data_ptr() == st.data_ptr()   # a view, no copy
    assert torch.equal(flat, e.reshape(12, 63))
    flat2 = H._flat_rows16(e)                                                         # packed input: padded copy
    assert flat2.stride() == (64, 1) and torch.equal(flat2, e.reshape(12, 63))


def test_kv_projection_weights_put_both_halves_on_16_byte_boundaries():
    emb = 63
    att = torch.nn.MultiheadAttention(emb, 3)
    with torch.no_grad():
        att.in_proj_bias.uniform_(-1, 1)
    w, b, ep = H._kv_weights(att.in_proj_weight, att.in_proj_bias, emb)
    assert ep == 64 and w.shape == (128, emb) and b.shape == (128,)
    x = torch.randn(10, emb)
    kv = x @ w.t() + b
    k_ref = x @ att.in_proj_weight[emb:2 * emb].t() + att.in_proj_bias[emb:2 * emb]
    v_ref = x @ att.in_proj_weight[2 * emb:].t() + att.in_proj_bias[2 * emb:]
    assert torch.allclose(kv[:, :emb], k_ref, atol=1e-6) and torch.allclose(kv[:, ep:ep + emb], v_ref, atol=1e-6)
    assert float(kv[:, emb:ep].abs().max()) == 0.0 and float(kv[:, ep + emb:].abs().max()) == 0.0   # pad columns
    w2, b2, ep2 = H._kv_weights(torch.randn(3 * 32, 32), None, 32)                    # no padding needed, no bias
    assert ep2 == 32 and w2.shape == (64, 32) and b2 is None


def test_keep_mask_is_row_zero_of_the_attention_weight_dropout():
    torch.manual_seed(0)
    keep = H.attention_keep_mask(50, 3, 3, 0.2, 'cpu')
    assert keep.shape == (150, 3) and keep.is_contiguous()
    vals = set(keep.unique().tolist())
    assert vals <= {0.0, 1.25} and len(vals) == 2                                     # kept entries scaled by 1 / (1 - p)
    torch.manual_seed(0)
    full = torch.nn.functional.dropout(torch.ones(150, 3, 3), p=0.2, training=True)
    assert torch.equal(keep, full[:, 0, :])


def test_engine_head_is_the_first_attention_row_of_the_module():
    """The algebra the head kernels implement (query of summary 0 only, softmax over the S summaries per head, output
    projection) restated in torch and checked against nn.MultiheadAttention itself on the CPU."""
    torch.manual_seed(1)
    S, n, emb = 3, 40, 63
    att = torch.nn.MultiheadAttention(emb, S, dropout=0.2).eval()
    e = torch.randn(S, n, emb)
    want = att(e, e, e)[0][0]
    d = emb // S
    w, b = att.in_proj_weight, att.in_proj_bias
    q = (e[0] @ w[:emb].t() + b[:emb]).view(n, S, d)
    k = (e @ w[emb:2 * emb].t() + b[emb:2 * emb]).view(S, n, S, d)
    v = (e @ w[2 * emb:].t() + b[2 * emb:]).view(S, n, S, d)
    score = torch.einsum('nhd,snhd->nhs', q, k) / d ** 0.5
    o = torch.einsum('nhs,snhd->nhd', torch.softmax(score, -1), v).reshape(n, emb)
    got = o @ att.out_proj.weight.t() + att.out_proj.bias
    assert torch.allclose(got, want, atol=1e-5)
