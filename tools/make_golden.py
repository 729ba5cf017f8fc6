#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the REFERENCE'S OWN layer source over the numpy
TensorFlow stand-in (tools/tf_shim).  Run in the build container (needs /root/reference):

    python tools/make_golden.py

Each file holds the inputs, the weights the reference layer created, and the reference layer's
outputs (float64).  tests/test_golden_cpu.py checks the oracle against them, the GPU tests check
the CUDA path against them.  Shapes follow the reference's shape fixtures.
TF's kernels themselves are restated by the shim, not executed — see oracle/__init__.py.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("RTF_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "tools", "tf_shim"))
sys.path.insert(0, os.path.join(REF, "src"))

import tensorflow as tf  # noqa: E402  (the shim)

OUT = os.environ.get("RTF_GOLDEN_OUT", os.path.join(ROOT, "tests", "golden"))


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: np.asarray(v) for k, v in arrays.items()})
    print(f"  {name}: " + ", ".join(f"{k}{tuple(np.asarray(v).shape)}" for k, v in arrays.items()))


def created_since(n):
    return tf.CREATED_LAYERS[n:]


def main():
    rng = np.random.default_rng(7)
    tf._seed(11)
    from ctr.layers import modules as C       # reference: src/ctr/layers/modules.py
    from ctr.layers import util as CU         # reference: src/ctr/layers/util.py
    from match.layers import modules as M     # reference: src/match/layers/modules.py

    # ---- a3: ctr FM layer, DeepFM shapes (2-D second input) and the 3-D docstring shape
    B = 8
    first = rng.normal(0, 0.3, (B, 13 + 26 * 8))
    second2 = rng.normal(0, 0.3, (B, 26 * 8))
    fm = C.FM(first.shape[1])
    out2 = fm([first, second2])
    second3 = rng.normal(0, 0.3, (B, 26, 8))
    out3 = fm([first, second3])
    save("ctr_fm_layer", first=first, second2=second2, second3=second3, w=fm.w, out2=out2, out3=out3)

    # ---- a6: DIN AttentionLayer (hidden_unit=1, 'sigmoid' as src/ctr/din/train.py:31), maxlen 10
    B, L, d = 6, 10, 16
    q, k, v = rng.normal(0, 0.5, (B, d)), rng.normal(0, 0.5, (B, L, d)), rng.normal(0, 0.5, (B, L, d))
    lens = rng.integers(1, L + 1, B)
    mask = (np.arange(L)[None, :] < lens[:, None]).astype(np.float64)
    mask[0] = 0                                           # one fully padded sample
    att = C.AttentionLayer(1, activation="sigmoid")
    out_mask = att([q, k, v, mask])
    out_nomask = att([q, k, v, None])                     # "mask is not a tf.Tensor" branch
    try:
        C.AttentionLayer(1)                               # default activation='prelu'
        prelu_ok = 1
    except ValueError:
        prelu_ok = 0
    save("ctr_attention_layer", q=q, k=k, v=v, mask=mask, W=att.att_dense.kernel, b=att.att_dense.bias,
         out_mask=out_mask, out_nomask=out_nomask, prelu_accepted=prelu_ok)

    # ---- a7: ctr MultiHeadAttention (AutoInt interacting layer), 39 fields x 16, 2 heads x 16, res
    B, F, dm = 4, 39, 16
    x = rng.normal(0, 0.3, (B, F, dm))
    n0 = len(tf.CREATED_LAYERS)
    mha = C.MultiHeadAttention(head_size=16, head_num=2, use_res=True)
    out = mha(x)
    dens = [l for l in created_since(n0) if isinstance(l, tf.keras.layers.Dense)]
    assert len(dens) == 4
    mha1 = C.MultiHeadAttention(head_size=8, head_num=1, use_res=False, activation="relu")
    n1 = len(tf.CREATED_LAYERS)
    out1 = mha1([x, x, x])
    dens1 = [l for l in created_since(n1) if isinstance(l, tf.keras.layers.Dense)]
    save("ctr_multihead_attention", x=x, Wq=dens[0].kernel, Wk=dens[1].kernel, Wv=dens[2].kernel,
         W0=dens[3].kernel, out=out, Wq1=dens1[0].kernel, Wk1=dens1[1].kernel, Wv1=dens1[2].kernel,
         out1=out1)

    # ---- a8: ctr scaled_dot_product_attention (dead code in the reference) incl. mask=None quirk
    qh, kh, vh = (rng.normal(0, 1, (2, 2, 5, 4)) for _ in range(3))
    m4 = (rng.random((2, 2, 5, 1)) < 0.6).astype(np.float64)
    save("ctr_sdpa_util", q=qh, k=kh, v=vh, mask=m4, out_mask=CU.scaled_dot_product_attention(qh, kh, vh, m4),
         out_nomask=CU.scaled_dot_product_attention(qh, kh, vh, None))

    # ---- a9: match MultiHeadAttention + TransformerEncoder, SASRec fixture (len 10, d 64)
    B, L, d = 3, 10, 64
    x = rng.normal(0, 1, (B, L, d))
    lens = rng.integers(1, L + 1, B)
    mask = (np.arange(L)[None, :] >= (L - lens)[:, None]).astype(np.float64)[:, :, None]  # pre-padding
    enc = M.TransformerEncoder(d, num_heads=1, ffn_hidden_unit=128)
    out_enc = enc([x, mask])
    mh = enc.mha
    out_mha = mh(x, x, x, mask)
    enc4 = M.TransformerEncoder(32, num_heads=4, ffn_hidden_unit=48)
    x4 = rng.normal(0, 1, (2, 7, 32))
    mask4 = np.ones((2, 7, 1))
    mask4[0, :3] = 0
    out_enc4 = enc4([x4, mask4])
    w = lambda e, p="": {  # noqa: E731
        p + "wq": e.mha.wq.kernel, p + "bq": e.mha.wq.bias, p + "wk": e.mha.wk.kernel, p + "bk": e.mha.wk.bias,
        p + "wv": e.mha.wv.kernel, p + "bv": e.mha.wv.bias, p + "ln1_g": e.layernorm1.gamma,
        p + "ln1_b": e.layernorm1.beta, p + "ln2_g": e.layernorm2.gamma, p + "ln2_b": e.layernorm2.beta,
        p + "w1": e.ffn.conv1.kernel, p + "b1": e.ffn.conv1.bias, p + "w2": e.ffn.conv2.kernel,
        p + "b2": e.ffn.conv2.bias}
    save("match_transformer_encoder", x=x, mask=mask, out_mha=out_mha, out_enc=out_enc, x4=x4, mask4=mask4,
         out_enc4=out_enc4, **w(enc), **w(enc4, "h4_"))

    # ---- a2: PoolingLayer
    ts = [rng.normal(0, 1, (4, 3, 5)) for _ in range(3)]
    save("match_pooling_layer", t0=ts[0], t1=ts[1], t2=ts[2],
         mean=M.PoolingLayer("mean")(ts), sum=M.PoolingLayer("sum")(ts), max=M.PoolingLayer("max")(ts),
         single=M.PoolingLayer("mean")(ts[0]))

    # ---- a11: SampledSoftmaxLayer as YoutubeDNN wires it (weights = item tower output, classes = 32)
    B, n = 48, 32
    item, user = rng.normal(0, 1, (B, 1, n)), rng.normal(0, 1, (B, 1, n))
    labels = rng.integers(0, 2, (B, 1))
    ssl = M.SampledSoftmaxLayer(num_sampled=5)
    loss = ssl([item, user, labels])
    s, te, se = tf.LAST_SAMPLED_VALUES
    save("match_sampled_softmax_layer", item=item, user=user, labels=labels, sampled=s, true_exp=te,
         samp_exp=se, loss=loss)

    # ---- a12: Dice (BatchNormalization inference mode: moving mean 0 / var 1)
    x = rng.normal(0, 1, (16, 8))
    dice = C.Dice()
    save("ctr_dice", x=x, alpha=dice.alpha, out=dice(x))

    # ---- a4: ctr.fm.model.FM (one-hot form), 13 dense + 26 sparse, k = 8
    from ctr.fm.model import FM as FMModel
    feat_nums = [int(n_) for n_ in rng.integers(2, 40, 26)]
    fc = [[{"feat": f"I{i}"} for i in range(13)],
          [{"feat": f"C{i}", "feat_num": n_, "embed_dim": 8} for i, n_ in enumerate(feat_nums)]]
    fmm = FMModel(fc, k=8)
    fmm.w0 = fmm.w0 + 0.1                                  # non-trivial bias
    B = 32
    dense = rng.random((B, 13))
    sparse = np.stack([rng.integers(0, n_, B) for n_ in feat_nums], 1).astype(np.int32)
    save("ctr_fm_model", dense=dense, sparse=sparse, feat_nums=np.asarray(feat_nums), w0=fmm.w0, w=fmm.w,
         V=fmm.V, out=fmm.call([dense, sparse]))

    # ---- a10: match SASRec forward with the commented-out fixture at sasrec/model.py:121-127
    from match.sasrec.model import SASRec
    user_features = [{"feat": "user_id", "feat_num": 100, "feat_len": 1, "embed_dim": 8},
                     {"feat": "seq_item", "feat_num": 100, "feat_len": 10, "embed_dim": 64},
                     {"feat": "pos_item", "feat_num": 100, "feat_len": 1, "embed_dim": 64},
                     {"feat": "neg_item", "feat_num": 100, "feat_len": 100, "embed_dim": 64}]
    item_features = [{"feat": "item_id", "feat_num": 100, "feat_len": 1, "embed_dim": 32}]
    sas = SASRec(user_features, item_features, att_hidden_unit=64, blocks=2)
    B = 5
    seq = rng.integers(1, 100, (B, 10)).astype(np.int32)
    for i in range(B):
        seq[i, : rng.integers(0, 9)] = 0                   # pad_sequences default 'pre'
    pos = rng.integers(1, 100, (B, 1)).astype(np.int32)
    neg = rng.integers(1, 100, (B, 100)).astype(np.int32)
    logits = sas.call([seq, pos, neg])
    el = sas.user_embed_layers
    blocks = {}
    for bi, e in enumerate(sas.encoder_layer):
        blocks.update(w(e, f"b{bi}_"))
    save("match_sasrec", seq=seq, pos=pos, neg=neg, logits=logits, loss=sas._losses[-1],
         seq_table=el["embed_seq_item"].embeddings, pos_table=el["embed_pos_item"].embeddings,
         neg_table=el["embed_neg_item"].embeddings, **blocks)

    # ---- a1: Embedding lookup + concat exactly as src/ctr/dlrm/model.py:30-37,45-46
    from tensorflow.keras.layers import Embedding
    rows = [1460, 583, 3, 24, 305]
    embs = [Embedding(input_dim=r, input_length=1, output_dim=16, embeddings_initializer="random_uniform")
            for r in rows]
    sparse = np.stack([rng.integers(0, r, 12) for r in rows], 1).astype(np.int32)
    out = tf.concat([embs[i](sparse[:, i]) for i in range(sparse.shape[1])], axis=-1)
    save("dlrm_embedding_concat", sparse=sparse, out=out, **{f"t{i}": e.embeddings for i, e in enumerate(embs)})


if __name__ == "__main__":
    print(f"reference: {REF}")
    main()
