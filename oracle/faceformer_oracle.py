"""Oracle: FaceFormer-disentangle decoder path (TEST INFRASTRUCTURE; see oracle/__init__.py).

fp32 torch-on-CPU restatement of models/faceformer_disentangle.py
  init_biased_mask :56-77, enc_dec_mask :80-88, PeriodicPositionalEncoding :92-107,
  Faceformer.forward_ff :435-482 (teacher-forced and autoregressive branches), predict :767-812
(models/faceformer_vert.py:759-840 is the same arithmetic with hidden_states = audio only, :434)
and of the torch.nn.TransformerDecoderLayer it instantiates at :195-196
(post-LN, ReLU, batch_first, eps 1e-5; eval mode => dropout inactive).
Pinned by tests/golden/ff_*.npz (outputs of the reference's own predict/forward_ff via oracle/make_golden.py).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

N_HEAD = 4
MAX_SEQ = 600


def get_slopes(n: int) -> list[float]:
    # faceformer_disentangle.py:57-67
    def pow2(n):
        start = 2 ** (-2 ** -(math.log2(n) - 3))
        return [start * start ** i for i in range(n)]
    if math.log2(n).is_integer():
        return pow2(n)
    c = 2 ** math.floor(math.log2(n))
    return pow2(c) + get_slopes(2 * c)[0::2][: n - c]


def init_biased_mask(n_head: int = N_HEAD, max_seq_len: int = MAX_SEQ, period: int = 30) -> torch.Tensor:
    """:56-77, literal."""
    slopes = torch.Tensor(get_slopes(n_head))
    bias = torch.arange(0, max_seq_len, period).unsqueeze(1).repeat(1, period).view(-1) // period
    bias = -torch.flip(bias, dims=[0])
    alibi = torch.zeros(max_seq_len, max_seq_len)
    for i in range(max_seq_len):
        alibi[i, : i + 1] = bias[-(i + 1):]
    alibi = slopes[:, None, None] * alibi[None]
    mask = (torch.triu(torch.ones(max_seq_len, max_seq_len)) == 1).transpose(0, 1)
    mask = mask.float().masked_fill(mask == 0, float("-inf")).masked_fill(mask == 1, 0.0)
    return mask[None] + alibi


def enc_dec_mask(dataset: str, T: int, S: int) -> torch.Tensor:
    """:80-88 ; True = masked."""
    mask = torch.ones(T, S)
    if dataset == "BIWI":
        for i in range(T):
            mask[i, i * 2: i * 2 + 2] = 0
    elif dataset == "vocaset":
        for i in range(T):
            mask[i, i] = 0
    return mask == 1


def ppe_table(d_model: int, period: int, max_seq_len: int = MAX_SEQ) -> torch.Tensor:
    """:92-104 -> [1, L, d_model]."""
    pe = torch.zeros(period, d_model)
    position = torch.arange(0, period, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).repeat(1, max_seq_len // period + 1, 1)


def _mha(sd, prefix, q_in, kv_in, mask, pmask=None):
    """torch.nn.MultiheadAttention forward (batch_first, 3-D float or 2-D bool mask), 4 heads; pmask [B,4,Tq,Tk]: the (pre-scaled)
    dropout mask on the attention probabilities in train mode."""
    fd = q_in.shape[-1]
    W, b = sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"]
    q = F.linear(q_in, W[:fd], b[:fd])
    k = F.linear(kv_in, W[fd:2 * fd], b[fd:2 * fd])
    v = F.linear(kv_in, W[2 * fd:], b[2 * fd:])
    B, Tq, _ = q.shape
    Tk = k.shape[1]
    hd = fd // N_HEAD
    q = q.view(B, Tq, N_HEAD, hd).transpose(1, 2)
    k = k.view(B, Tk, N_HEAD, hd).transpose(1, 2)
    v = v.view(B, Tk, N_HEAD, hd).transpose(1, 2)
    s = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(hd)
    if mask.dtype == torch.bool:
        s = s.masked_fill(mask, float("-inf"))
    else:
        s = s + mask
    a = torch.softmax(s, dim=-1)
    if pmask is not None:
        a = a * pmask
    o = torch.matmul(a, v).transpose(1, 2).reshape(B, Tq, fd)
    return F.linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])


def decoder_layer(sd, x, mem, tgt_mask, memory_mask, drop=None):
    """nn.TransformerDecoderLayer (post-norm, relu). drop (train mode): {site: pre-scaled mask of THIS clip} for the attention-probability
    dropouts of the two MultiheadAttentions (sa, ca) and dropout1 / dropout2 / the feed-forward dropout / dropout3 (d1, d2, act, d3)."""
    p = "transformer_decoder.layers.0."
    fd = x.shape[-1]
    d = (lambda name, t: t) if drop is None else (lambda name, t: t * drop[name].reshape(t.shape))

    def ln(i, t):
        return F.layer_norm(t, (fd,), sd[p + f"norm{i}.weight"], sd[p + f"norm{i}.bias"], 1e-5)

    x = ln(1, x + d("d1", _mha(sd, p + "self_attn.", x, x, tgt_mask, None if drop is None else drop["sa"])))
    x = ln(2, x + d("d2", _mha(sd, p + "multihead_attn.", x, mem, memory_mask, None if drop is None else drop["ca"])))
    ff = F.linear(d("act", F.relu(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"]))),
                  sd[p + "linear2.weight"], sd[p + "linear2.bias"])
    return ln(3, x + d("d3", ff))


@torch.no_grad()
def forward_ff(sd, template, hidden_states, obj_embedding, frame_num, teacher_forcing, gt_verts=None,
               period=30, dataset="vocaset", merge=True, reg=None):
    """Faceformer.forward_ff :435-482, literal (the AR branch re-runs the whole prefix each step).

    template [1,1,V3]; hidden_states [B,T,36+fd] (or [B,T,fd] with merge=False); obj_embedding [B,fd].
    """
    fd = obj_embedding.shape[-1]
    biased_mask = init_biased_mask(N_HEAD, MAX_SEQ, period)
    pe = ppe_table(fd, period)
    mix = F.linear(hidden_states, sd["v_merge2hidden.weight"], sd["v_merge2hidden.bias"]) if merge else hidden_states
    outs = []
    for j in range(len(mix)):
        hs = mix[j:j + 1]
        style = obj_embedding.unsqueeze(1)[j:j + 1]
        if teacher_forcing:
            vin = torch.cat([template, gt_verts[j:j + 1][:, :-1]], 1) - template
            vin = F.linear(vin, sd["vertice_map.weight"], sd["vertice_map.bias"]) + style
            vin = vin + pe[:, : vin.shape[1]]
            n = vin.shape[1]
            drop = None
            if reg is not None:           # train mode (teacher-forced branch only): PPE dropout + the decoder layer's, clip j's slices
                B = len(mix)
                mk = reg["masks"]
                vin = vin * mk["ppe"].view(B, n, -1)[j:j + 1]
                drop = {"sa": mk["dec.sa"][j:j + 1], "ca": mk["dec.ca"][j:j + 1], "d1": mk["dec.d1"].view(B, n, -1)[j:j + 1],
                        "d2": mk["dec.d2"].view(B, n, -1)[j:j + 1], "act": mk["dec.act"].view(B, n, -1)[j:j + 1],
                        "d3": mk["dec.d3"].view(B, n, -1)[j:j + 1]}
            out = decoder_layer(sd, vin, hs, biased_mask[:, :n, :n], enc_dec_mask(dataset, n, hs.shape[1]), drop)
            out = F.linear(out, sd["vertice_map_r.weight"], sd["vertice_map_r.bias"])
        else:
            emb = style
            for i in range(frame_num):
                vin = emb + pe[:, : emb.shape[1]]
                n = vin.shape[1]
                out = decoder_layer(sd, vin, hs, biased_mask[:, :n, :n], enc_dec_mask(dataset, n, hs.shape[1]))
                out = F.linear(out, sd["vertice_map_r.weight"], sd["vertice_map_r.bias"])
                new = F.linear(out[:, -1, :], sd["vertice_map.weight"], sd["vertice_map.bias"]).unsqueeze(1) + style
                emb = torch.cat((emb, new), 1)
        outs.append(out)
    return torch.cat(outs) + template


@torch.no_grad()
def forward_ff_cached(sd, template, hidden_states, obj_embedding, frame_num, period=30, merge=True,
                      return_hidden=False):
    """O(T) equivalent of the AR branch: causal => row k of step i equals row k of step k
    (checked against the literal loop in tests/test_oracle_golden.py). Used where the
    literal O(T^2) loop is too slow (10 s clips, CPU baseline timing)."""
    p = "transformer_decoder.layers.0."
    fd = obj_embedding.shape[-1]
    hd = fd // N_HEAD
    slopes = torch.tensor(get_slopes(N_HEAD))
    pe = ppe_table(fd, period)[0]
    mix = F.linear(hidden_states, sd["v_merge2hidden.weight"], sd["v_merge2hidden.bias"]) if merge else hidden_states
    W, b = sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"]
    Wc, bc = sd[p + "multihead_attn.in_proj_weight"], sd[p + "multihead_attn.in_proj_bias"]

    def ln(i, t):
        return F.layer_norm(t, (fd,), sd[p + f"norm{i}.weight"], sd[p + f"norm{i}.bias"], 1e-5)

    outs, hids = [], []
    for j in range(len(mix)):
        mem = mix[j]
        ca_all = F.linear(F.linear(mem, Wc[2 * fd:], bc[2 * fd:]), sd[p + "multihead_attn.out_proj.weight"],
                          sd[p + "multihead_attn.out_proj.bias"])            # degenerate cross-attn
        style = obj_embedding[j]
        emb = style
        Ks, Vs, ys = [], [], []
        for i in range(frame_num):
            x = emb + pe[i]
            q = F.linear(x, W[:fd], b[:fd]).view(N_HEAD, hd)
            Ks.append(F.linear(x, W[fd:2 * fd], b[fd:2 * fd]).view(N_HEAD, hd))
            Vs.append(F.linear(x, W[2 * fd:], b[2 * fd:]).view(N_HEAD, hd))
            K = torch.stack(Ks, 1)                                            # [H, i+1, hd]
            V = torch.stack(Vs, 1)
            s = torch.einsum("hd,hjd->hj", q, K) / math.sqrt(hd)
            jj = torch.arange(i + 1)
            s = s - slopes[:, None] * ((i - jj) // period)[None].float()
            o = torch.einsum("hj,hjd->hd", torch.softmax(s, -1), V).reshape(fd)
            x = ln(1, x + F.linear(o, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"]))
            x = ln(2, x + ca_all[i])
            ff = F.linear(F.relu(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                          sd[p + "linear2.weight"], sd[p + "linear2.bias"])
            y = ln(3, x + ff)
            ys.append(y)
            v = F.linear(y, sd["vertice_map_r.weight"], sd["vertice_map_r.bias"])
            emb = F.linear(v, sd["vertice_map.weight"], sd["vertice_map.bias"]) + style
        Y = torch.stack(ys)
        hids.append(Y)
        outs.append(F.linear(Y, sd["vertice_map_r.weight"], sd["vertice_map_r.bias"]))
    out = torch.stack(outs) + template
    return (out, torch.stack(hids)) if return_hidden else out


@torch.no_grad()
def predict(sd_ff, sd_w2v, template, audio, emo_embed, n_subjects=8, period=30, cached=False, w2v_layers=12):
    """Faceformer.predict :767-812 with the FanEncoder branch replaced by a supplied emo_embed [B,T,30]
    (image CNN is out of scope); returns vertices [B,T,V3]."""
    from .wav2vec2_oracle import wav2vec2_forward
    B = audio.shape[0]
    one_hot = torch.zeros(B, n_subjects)
    one_hot[:, 0] = 1
    obj = F.linear(one_hot, sd_ff["obj_vector.weight"])                                   # :771-773
    ha = wav2vec2_forward(sd_w2v, audio, layers=w2v_layers)                               # :775
    ha = F.linear(ha, sd_ff["audio_feature_map.weight"], sd_ff["audio_feature_map.bias"])  # :776
    T = ha.shape[1]
    hs = torch.cat([sd_ff["learnable_eye_embed"].expand(B, T, -1), emo_embed[:, :T], ha], dim=-1)  # :808
    if cached:
        return forward_ff_cached(sd_ff, template, hs, obj, T, period)
    return forward_ff(sd_ff, template, hs, obj, T, teacher_forcing=False, period=period)      # :810
