"""The CUDA path (drop-in layer classes over the C-ABI) against the golden vectors produced by
the reference's own source (tools/make_golden.py).  fp32 vs the float64 goldens: 1e-5 relative."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def G(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def T(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).cuda()


def close(got, want, rtol=1e-5, atol=5e-6):
    np.testing.assert_allclose(got.detach().cpu().numpy().astype(np.float64), want, rtol=rtol, atol=atol)


def setp(param, value):
    with torch.no_grad():
        param.copy_(T(value).reshape(param.shape))


def test_embedding_concat_bit_exact(rtf):
    g = G("dlrm_embedding_concat")
    tabs = [T(g[f"t{i}"]) for i in range(5)]
    out = rtf.embed_fwd(tabs, T(g["sparse"], torch.int32))
    assert np.array_equal(out.cpu().numpy(), g["out"].astype(np.float32))
    # per-field Keras-style Embedding layers + concat, as the model files write it
    embs = [rtf.layers.Embedding(t.shape[0], t.shape[1]) for t in tabs]
    for e, t in zip(embs, tabs):
        setp(e.embeddings, t.cpu().numpy())
    sp = T(g["sparse"], torch.int32)
    out2 = torch.cat([embs[i](sp[:, i]) for i in range(5)], -1)
    assert torch.equal(out2, out)
    # float-typed ids are cast like Keras does (A1)
    assert torch.equal(embs[0](sp[:, 0].float()), embs[0](sp[:, 0]))


def test_fm_layer(rtf):
    g = G("ctr_fm_layer")
    layer = rtf.layers.FM(221)
    out2 = layer([T(g["first"]), T(g["second2"])])
    setp(layer.w, g["w"])
    close(layer([T(g["first"]), T(g["second2"])]), g["out2"], atol=1e-5 * np.abs(g["out2"]).max())
    out3 = layer([T(g["first"]), T(g["second3"])])
    assert out3.shape == (64, 1)
    close(out3, g["out3"], atol=1e-5 * np.abs(g["out3"]).max())


def test_fm_model(rtf):
    g = G("ctr_fm_model")
    fn = [int(n) for n in g["feat_nums"]]
    fc = [[{"feat": f"I{i}"} for i in range(13)], [{"feat": f"C{i}", "feat_num": n, "embed_dim": 8} for i, n in enumerate(fn)]]
    m = rtf.FMModel(fc, k=8)
    m.load_reference_weights(g["w0"], g["w"], g["V"])
    close(m([T(g["dense"]), T(g["sparse"], torch.int32)]), g["out"])


def test_din_attention_layer(rtf):
    g = G("ctr_attention_layer")
    layer = rtf.layers.AttentionLayer(1, activation="sigmoid")
    q, k, v = T(g["q"]), T(g["k"]), T(g["v"])
    layer([q, k, v, T(g["mask"])])
    setp(layer.att_dense_kernel, g["W"]); setp(layer.att_dense_bias, g["b"])
    close(layer([q, k, v, T(g["mask"])]), g["out_mask"])
    close(layer([q, k, v, None]), g["out_nomask"])


def test_ctr_multihead_attention(rtf):
    g = G("ctr_multihead_attention")
    x = T(g["x"])
    layer = rtf.layers.ctr.MultiHeadAttention(head_size=16, head_num=2, use_res=True)
    layer(x)
    for d, n in ((layer.q_dense, "Wq"), (layer.k_dense, "Wk"), (layer.v_dense, "Wv"), (layer.res_dense, "W0")):
        setp(d.kernel, g[n])
    close(layer(x), g["out"])
    l1 = rtf.layers.ctr.MultiHeadAttention(head_size=8, head_num=1)
    l1([x, x, x])
    for d, n in ((l1.q_dense, "Wq1"), (l1.k_dense, "Wk1"), (l1.v_dense, "Wv1")):
        setp(d.kernel, g[n])
    close(l1([x, x, x]), g["out1"])


def _load_encoder(enc, g, prefix=""):
    m = enc.mha
    for d, n in ((m.wq, "q"), (m.wk, "k"), (m.wv, "v")):
        setp(d.kernel, g[prefix + "w" + n]); setp(d.bias, g[prefix + "b" + n])
    setp(enc.layernorm1.gamma, g[prefix + "ln1_g"]); setp(enc.layernorm1.beta, g[prefix + "ln1_b"])
    setp(enc.layernorm2.gamma, g[prefix + "ln2_g"]); setp(enc.layernorm2.beta, g[prefix + "ln2_b"])
    setp(enc.ffn.conv1.kernel, g[prefix + "w1"]); setp(enc.ffn.conv1.bias, g[prefix + "b1"])
    setp(enc.ffn.conv2.kernel, g[prefix + "w2"]); setp(enc.ffn.conv2.bias, g[prefix + "b2"])


def test_match_mha_and_transformer_encoder(rtf):
    g = G("match_transformer_encoder")
    enc = rtf.layers.TransformerEncoder(64, 1, 128)
    x, mask = T(g["x"]), T(g["mask"])
    enc([x, mask])
    _load_encoder(enc, g)
    close(enc.mha(x, x, x, mask), g["out_mha"])
    close(enc([x, mask]), g["out_enc"], rtol=2e-5, atol=2e-5)
    enc4 = rtf.layers.TransformerEncoder(32, 4, 48)
    enc4([T(g["x4"]), T(g["mask4"])])
    _load_encoder(enc4, g, "h4_")
    close(enc4([T(g["x4"]), T(g["mask4"])]), g["out_enc4"], rtol=2e-5, atol=2e-5)


def test_pooling_sampled_softmax_dice(rtf):
    g = G("match_pooling_layer")
    ts = [T(g[k]) for k in ("t0", "t1", "t2")]
    for mode in ("mean", "sum", "max"):
        close(rtf.layers.PoolingLayer(mode)(ts), g[mode])
    g = G("match_sampled_softmax_layer")
    layer = rtf.layers.SampledSoftmaxLayer(num_sampled=5)
    sv = (T(g["sampled"], torch.int64), T(g["true_exp"]), T(g["samp_exp"]))
    loss = layer([T(g["item"]), T(g["user"]), T(g["labels"], torch.int64)], sampled_values=sv)
    close(loss, g["loss"], atol=1e-5)
    g = G("ctr_dice")
    dice = rtf.layers.Dice().eval()
    x = T(g["x"])
    dice(x)
    setp(dice.alpha, g["alpha"])
    close(dice(x), g["out"])


def test_sasrec_model(rtf):
    from recommend_tf2_b200.models import SASRec
    g = G("match_sasrec")
    m = SASRec(item_num=100, embed_dim=64, blocks=2, seq_len=10, neg_len=100)
    seq, pos, neg = (T(g[k], torch.int32) for k in ("seq", "pos", "neg"))
    m([seq, pos, neg])
    for w, n in zip(m.tables.weights, ("seq_table", "pos_table", "neg_table")):
        setp(w, g[n])
    for bi, enc in enumerate(m.encoder_layer):
        _load_encoder(enc, g, f"b{bi}_")
    logits, loss = m([seq, pos, neg])
    close(logits, g["logits"], rtol=5e-5, atol=5e-5)
    close(loss, g["loss"], rtol=1e-5)
