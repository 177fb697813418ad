"""CPU oracle of the CLIP-FSAR few-shot inference path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file,
and only as the checker / reported baseline. The product (clip_fsar_b200/) never imports it and has no CPU path.

It is a functional fp32 restatement (torch CPU tensor ops, no nn.Module, no nn.MultiheadAttention, no conv2d)
of the reference algorithm in /root/reference/models/base/few_shot.py; every function cites the lines it follows.

Parity pin: the reference ships no tests, golden vectors or fixtures for this path, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: oracle/gen_golden.py imports the unmodified
reference from /root/reference (CPU, stubs for ipdb/ftfy, CLIP checkpoint download replaced by seeded weights),
runs CNN_OTAM_CLIPFSAR.forward and stores inputs seeds + outputs under tests/golden/; tests/test_oracle_golden.py
checks this file against those fixtures (max |delta| ~1e-6 relative, fp32 summation-order noise only).
"""
import math

import numpy as np
import torch

LN_EPS = 1e-5  # nn.LayerNorm default, few_shot.py:605-611 / 974


def _t(a):
    return a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))


def layer_norm(x, w, b):
    """LayerNorm.forward, few_shot.py:608-611: statistics in fp32, biased variance, eps 1e-5."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * w + b


def quick_gelu(x):
    """QuickGELU.forward, few_shot.py:614-616."""
    return x * torch.sigmoid(1.702 * x)


def quick_gelu_16(x, dt):
    """QuickGELU as the CUDA epilogue evaluates it when operands are 16-bit: x sigmoid(1.702 x) = h + h tanh(0.851 x),
    h = x / 2, in packed 16-bit arithmetic (every intermediate rounded to `dt`, the final fma rounded once)."""
    r = lambda v: v.to(dt).float()
    x16 = r(x)
    t = r(torch.tanh(r(r(torch.tensor(0.851)) * x16)))
    h = r(0.5 * x16)
    return r(h * t + h)


def gelu_erf(x):
    """nn.GELU() (exact erf form) used by FeedForward, few_shot.py:1648."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def patchify(frames, P):
    """conv1 with kernel = stride = P and no bias (few_shot.py:659, 672-674) is a matmul over flattened patches:
    [n,3,S,S] -> [n, G*G, 3*P*P] with k = c*P*P + ky*P + kx, patches in row-major (py, px) order."""
    n, C, S, _ = frames.shape
    G = S // P
    x = frames.reshape(n, C, G, P, G, P).permute(0, 2, 4, 1, 3, 5)
    return x.reshape(n, G * G, C * P * P)


def _rnd(x, dt):
    """Round to the tensor-core operand type and back (emulates the 16-bit operand storage of the CUDA path)."""
    return x if dt is None else x.to(dt).float()


def residual_block(x, sd, p, heads, dt=None, mask=None):
    """ResidualAttentionBlock.forward, few_shot.py:633-640 with nn.MultiheadAttention(d, heads) (623, 635):
    no mask, no dropout in eval, head_dim ** -0.5 scaling of q. x: [n, L, D].
    dt != None additionally rounds every GEMM operand to that 16-bit type exactly where the CUDA path stores one
    (LN outputs, weights, QKV, softmax numerators, attention output, GELU output); accumulation, residual stream,
    LN / softmax statistics stay fp32 in both."""
    n, L, D = x.shape
    dh = D // heads
    y = _rnd(layer_norm(x, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"]), dt)
    qkv = _rnd(y @ _rnd(sd[p + "attn.in_proj_weight"], dt).T + sd[p + "attn.in_proj_bias"], dt)
    q, k, v = qkv.split(D, dim=-1)
    q = q.reshape(n, L, heads, dh).transpose(1, 2)
    k = k.reshape(n, L, heads, dh).transpose(1, 2)
    v = v.reshape(n, L, heads, dh).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * dh ** -0.5
    if mask is not None:   # additive attn_mask of the text transformer (few_shot.py:635, 777-783)
        s = s + mask
    if dt is None:
        o = torch.softmax(s, dim=-1) @ v
    else:  # un-normalised numerators are the 16-bit P operand; the fp32 row sum divides afterwards
        e = torch.exp(s - s.max(-1, keepdim=True).values)
        o = (_rnd(e, dt) @ v) / e.sum(-1, keepdim=True)
    o = _rnd(o.transpose(1, 2).reshape(n, L, D), dt)
    x = x + o @ _rnd(sd[p + "attn.out_proj.weight"], dt).T + sd[p + "attn.out_proj.bias"]
    y = _rnd(layer_norm(x, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"]), dt)
    pre = y @ _rnd(sd[p + "mlp.c_fc.weight"], dt).T + sd[p + "mlp.c_fc.bias"]
    hdn = quick_gelu(pre) if dt is None else quick_gelu_16(pre, dt)
    return x + hdn @ _rnd(sd[p + "mlp.c_proj.weight"], dt).T + sd[p + "mlp.c_proj.bias"]


def vit_forward(sd, g, frames, taps=None, chunk=16, operand_dtype=None):
    """VisionTransformer.forward, few_shot.py:671-688: frames [n,3,S,S] -> [n, embed_dim].
    operand_dtype=torch.float16 emulates the CUDA path's operand rounding (see residual_block)."""
    sd = {k: _t(v) for k, v in sd.items()}
    frames = _t(frames).float()
    dt = operand_dtype
    D, P = g["width"], g["patch_size"]
    outs = []
    for s in range(0, frames.shape[0], chunk):
        f = frames[s:s + chunk]
        x = _rnd(patchify(f, P), dt) @ _rnd(sd["backbone.conv1.weight"].reshape(D, -1), dt).T    # 672-674
        cls = sd["backbone.class_embedding"].expand(x.shape[0], 1, D)                         # 675
        x = torch.cat([cls, x], dim=1) + sd["backbone.positional_embedding"]                  # 675-676
        x = layer_norm(x, sd["backbone.ln_pre.weight"], sd["backbone.ln_pre.bias"])           # 677
        if taps is not None and s == 0:
            taps["ln_pre"] = x.clone()
        for i in range(g["layers"]):                                                           # 679-681
            x = residual_block(x, sd, "backbone.transformer.resblocks.%d." % i, g["heads"], dt)
            if taps is not None and s == 0:
                taps["block%d" % i] = x.clone()
        x = layer_norm(x[:, 0, :], sd["backbone.ln_post.weight"], sd["backbone.ln_post.bias"])  # 683
        outs.append(x @ sd["backbone.proj"])                                                   # 686
    return torch.cat(outs, dim=0)


def text_encode(sd, tg, tokens, operand_dtype=None):
    """CLIP.encode_text, few_shot.py:793-806: token + positional embedding, the ResidualAttentionBlocks of
    Transformer (643-651) under build_attention_mask (777-783: -inf above the diagonal), ln_final, the row of the
    largest token id (end of text) @ text_projection. sd: CLIP state_dict keys of the text side. tokens: int [n, C]."""
    sd = {k: _t(v) for k, v in sd.items()}
    tokens = _t(tokens).long()
    n, C = tokens.shape
    x = sd["token_embedding.weight"][tokens] + sd["positional_embedding"]                   # 794-796
    mask = torch.full((C, C), float("-inf")).triu_(1)                                        # 777-783
    for i in range(tg["layers"]):                                                            # 797-799
        x = residual_block(x, sd, "transformer.resblocks.%d." % i, tg["heads"], operand_dtype, mask=mask)
    x = layer_norm(x, sd["ln_final.weight"], sd["ln_final.bias"])                            # 800
    return x[torch.arange(n), tokens.argmax(dim=-1)] @ sd["text_projection"]                 # 804


def modulator(sd, g, x):
    """Transformer_v1.forward(x, x, x), few_shot.py:990-999, with PreNormattention_qkv (971-977: ONE shared
    LayerNorm, residual adds the un-normalised q), Attention_qkv (1055-1073: bias-free q/k/v projections,
    scale dim_head ** -0.5, softmax, to_out Linear + bias; dropout inactive in eval) and FeedForward
    (1643-1654: Linear, exact GELU, Linear). x: [n_seq, n_tok, E]."""
    sd = {k: _t(v) for k, v in sd.items()}
    x = _t(x).float()
    H, dh = g["mod_heads"], g["mod_dim_head"]
    for l in range(g["mod_depth"]):
        p = "context2.layers.%d." % l
        b, n, _ = x.shape
        y = layer_norm(x, sd[p + "0.norm.weight"], sd[p + "0.norm.bias"])
        q = (y @ sd[p + "0.fn.to_q.weight"].T).reshape(b, n, H, dh).transpose(1, 2)
        k = (y @ sd[p + "0.fn.to_k.weight"].T).reshape(b, n, H, dh).transpose(1, 2)
        v = (y @ sd[p + "0.fn.to_v.weight"].T).reshape(b, n, H, dh).transpose(1, 2)
        att = torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, dim=-1)
        o = (att @ v).transpose(1, 2).reshape(b, n, H * dh)
        x = o @ sd[p + "0.fn.to_out.0.weight"].T + sd[p + "0.fn.to_out.0.bias"] + x
        hdn = gelu_erf(x @ sd[p + "1.net.0.weight"].T + sd[p + "1.net.0.bias"])
        x = hdn @ sd[p + "1.net.3.weight"].T + sd[p + "1.net.3.bias"] + x
    return x


def cos_sim(x, y, epsilon=0.01):
    """cos_sim, few_shot.py:1115-1124: x y^T / (|x| |y|^T + 0.01) — epsilon is added to the PRODUCT of the norms."""
    x, y = _t(x), _t(y)
    num = x @ y.transpose(-1, -2)
    den = x.norm(dim=-1).unsqueeze(-1) @ y.norm(dim=-1).unsqueeze(-1).transpose(-1, -2) + epsilon
    return num / den


def otam_cum_dist(dists, lbda=0.5):
    """OTAM_cum_dist_v2, few_shot.py:2657-2687. dists [Q, way, T, T'] -> [Q, way]."""
    d = _t(dists).float()
    Qn, Wn, L, M = d.shape
    d = torch.cat([torch.zeros(Qn, Wn, L, 1), d, torch.zeros(Qn, Wn, L, 1)], dim=3)      # F.pad (1,1), 2663
    M2 = M + 2
    c = torch.zeros_like(d)
    for m in range(1, M2):                                                              # top row, 2668-2671
        c[:, :, 0, m] = d[:, :, 0, m] + c[:, :, 0, m - 1]
    for l in range(1, L):
        c[:, :, l, 1] = d[:, :, l, 1] - lbda * torch.log(                               # 2677
            torch.exp(-c[:, :, l - 1, 0] / lbda) + torch.exp(-c[:, :, l - 1, 1] / lbda) + torch.exp(-c[:, :, l, 0] / lbda))
        for m in range(2, M2 - 1):                                                      # 2680-2681
            c[:, :, l, m] = d[:, :, l, m] - lbda * torch.log(
                torch.exp(-c[:, :, l - 1, m - 1] / lbda) + torch.exp(-c[:, :, l, m - 1] / lbda))
        c[:, :, l, -1] = d[:, :, l, -1] - lbda * torch.log(                             # 2685
            torch.exp(-c[:, :, l - 1, -2] / lbda) + torch.exp(-c[:, :, l - 1, -1] / lbda) + torch.exp(-c[:, :, l, -2] / lbda))
    return c[:, :, -1, -1]


def otam_scalar(d, lbda=0.5):
    """Same recurrence for ONE [T, T'] matrix in plain Python floats (independent cross-check of otam_cum_dist)."""
    L, M = len(d), len(d[0])
    pad = [[0.0] + [float(v) for v in row] + [0.0] for row in d]
    c = [[0.0] * (M + 2) for _ in range(L)]
    for m in range(1, M + 2):
        c[0][m] = pad[0][m] + c[0][m - 1]
    for l in range(1, L):
        for m in range(1, M + 2):
            if m == 1 or m == M + 1:
                s = math.exp(-c[l - 1][m - 1] / lbda) + math.exp(-c[l - 1][m] / lbda) + math.exp(-c[l][m - 1] / lbda)
            else:
                s = math.exp(-c[l - 1][m - 1] / lbda) + math.exp(-c[l][m - 1] / lbda)
            c[l][m] = pad[l][m] - lbda * math.log(s)
    return c[L - 1][M + 1]


def preprocess_u8(frames_u8, crop, resize=(256, 256), mean=(0.48145466, 0.4578275, 0.40821073),
                  std=(0.26862954, 0.26130258, 0.27577711)):
    """Test-time loader transform of the reference, restated: ToTensorVideo (uint8 [T,H,W,C] -> float / 255),
    KineticsResizedCropFewshot._get_controlled_crop with short_side_range = [TEST_SCALE, TEST_SCALE], one spatial crop,
    idx = TEST_CENTER_CROP (bilinear resize, align_corners = False, then the centre crop; transformations.py:676-716)
    and NormalizeVideo (ssv2_few_shot.py:633-642). Returns fp32 [T, 3, crop, crop] (the task-dict frame layout)."""
    x = _t(frames_u8).float().permute(0, 3, 1, 2) / 255.0                              # [T, C, H, W]
    x = torch.nn.functional.interpolate(x, size=(int(resize[0]), int(resize[1])), mode="bilinear", align_corners=False)
    y0, x0 = (int(resize[0]) - crop) // 2, (int(resize[1]) - crop) // 2
    x = x[:, :, y0:y0 + crop, x0:x0 + crop]
    m = torch.tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
    sd = torch.tensor(std, dtype=torch.float32).view(1, 3, 1, 1)
    return ((x - m) / sd).contiguous()


def class_index(labels):
    """Rank of each label among the sorted distinct labels: torch.unique(support_labels) is sorted
    (few_shot.py:2950/2960/2965) and extract_class_indices (1127-1136) selects by equality."""
    labels = _t(labels).long()
    uniq = torch.unique(labels)
    return torch.stack([(uniq == v).nonzero()[0, 0] for v in labels]), uniq.numel()


def text_probabilities(sd, text_test, target_feats, real_support_labels, cls, way):
    """p_text of the EVAL_TEXT / COMBINE branches, few_shot.py:2836-2849 (= 2857-2870): per-class mean of the support
    videos' text features, unit-normalised (no epsilon here), softmax over classes of scale * cosine."""
    txt = _t(text_test).float()[_t(real_support_labels).long()]
    txt = torch.stack([txt[cls == c].mean(0) for c in range(way)])
    img = _t(target_feats).float().mean(1)
    img = img / img.norm(dim=1, keepdim=True)
    txt = txt / txt.norm(dim=1, keepdim=True)
    return torch.softmax(sd["scale"] * img @ txt.t(), dim=1)


def head_forward(sd, g, text_train, text_test, support_feats, target_feats, support_labels, real_support_labels,
                 merge_before=False, single_direct=False, lbda=0.5, text_mode=0, text_coff=0.9):
    """CNN_OTAM_CLIPFSAR.forward eval else-branch after get_feats, few_shot.py:2936-2990.
    support_feats [S,T,E], target_feats [Q,T,E]. Returns dict with logits, class_logits and intermediates."""
    sd = {k: _t(v) for k, v in sd.items()}
    sup, tgt = _t(support_feats).float(), _t(target_feats).float()
    text_train, text_test = _t(text_train).float(), _t(text_test).float()
    T = sup.shape[1]
    cls, way = class_index(support_labels)
    if text_mode == 1:                                                                   # TRAIN.EVAL_TEXT, 2835-2852
        p_text = text_probabilities(sd, text_test, tgt, real_support_labels, cls, way)
        return {"logits": p_text, "class_logits": None, "class_index": cls}
    # 2936-2939 (classification_layer is an empty nn.Sequential)
    class_logits = cos_sim(torch.cat([sup, tgt], 0).mean(1), text_train) * sd["scale"]
    context = text_test[_t(real_support_labels).long()].unsqueeze(1)                     # 2946
    tgt_mod = modulator(sd, g, tgt)                                                       # 2948
    if merge_before:                                                                      # 2949-2954
        sup = torch.stack([sup[cls == c].mean(0) for c in range(way)])
        context = torch.stack([context[cls == c].mean(0) for c in range(way)])
    sup_mod = modulator(sd, g, torch.cat([sup, context], dim=1))[:, :T, :]                # 2955-2956
    if not merge_before:                                                                  # 2959-2962
        sup_mod = torch.stack([sup_mod[cls == c].mean(0) for c in range(way)])
    Q = tgt_mod.shape[0]
    E = tgt_mod.shape[2]
    sim = cos_sim(tgt_mod.reshape(Q * T, E), sup_mod.reshape(way * T, E))                 # 2970-2973
    dists = (1 - sim).reshape(Q, T, way, T).permute(0, 2, 1, 3)                           # 2974-2976
    if single_direct:                                                                     # 2979-2982
        cum = otam_cum_dist(dists, lbda)
    else:
        cum = otam_cum_dist(dists, lbda) + otam_cum_dist(dists.transpose(2, 3), lbda)
    if text_mode == 2:                                                                   # TRAIN.COMBINE, 2921-2930
        p_text = text_probabilities(sd, text_test, tgt, real_support_labels, cls, way)
        p_vis = torch.softmax((8 - cum) / 8.0, dim=1)
        return {"logits": p_text.pow(text_coff) * p_vis.pow(1.0 - text_coff), "class_logits": None, "target_mod": tgt_mod,
                "protos": sup_mod, "dists": dists.contiguous(), "cum_dists": cum, "class_index": cls}
    # 2986-2989: prototypes are already one per sorted class, so the class reduction is the identity
    return {"logits": -cum, "class_logits": class_logits, "target_mod": tgt_mod, "protos": sup_mod,
            "dists": dists.contiguous(), "cum_dists": cum, "class_index": cls}


def episode_forward(sd, g, text_train, text_test, task, n_frames, merge_before=False, single_direct=False,
                    lbda=0.5, operand_dtype=None, text_mode=0, text_coff=0.9):
    """CNN_OTAM_CLIPFSAR.forward (eval), few_shot.py:2772-2990, on a task dict of numpy arrays / tensors."""
    sup = vit_forward(sd, g, task["support_set"], operand_dtype=operand_dtype)            # get_feats 2760-2765
    tgt = vit_forward(sd, g, task["target_set"], operand_dtype=operand_dtype)
    E = sup.shape[-1]
    out = head_forward(sd, g, text_train, text_test, sup.reshape(-1, n_frames, E), tgt.reshape(-1, n_frames, E),
                       task["support_labels"], task["real_support_labels"], merge_before, single_direct, lbda,
                       text_mode, text_coff)
    out["support_feats"] = sup.reshape(-1, n_frames, E)
    out["target_feats"] = tgt.reshape(-1, n_frames, E)
    return out
