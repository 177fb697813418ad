"""Seeded synthetic weights and episodes (numpy PCG64: bit-stable across machines and library versions).

There is no network for CLIP checkpoints or video datasets, so benchmarks, tests and golden fixtures all use
random-init weights of the real geometry and synthetic 224x224 frames. Key names and shapes follow the
reference head's state_dict (few_shot.py:655-669, 619-631, 979-988, 1035-1053, 1643-1652);
episode dicts follow the reference dataset's task dict (datasets/base/ssv2_few_shot.py:267-285).
"""
import numpy as np

GEOMETRIES = {
    # name: image, patch, width, layers, heads, embed
    "ViT-B/16": dict(image_size=224, patch_size=16, width=768, layers=12, heads=12, embed_dim=512),
    "ViT-L/14": dict(image_size=224, patch_size=14, width=1024, layers=24, heads=16, embed_dim=768),
    "ViT-B/32": dict(image_size=224, patch_size=32, width=768, layers=12, heads=12, embed_dim=512),
    # small towers with the same structure, for tests that must finish in seconds on CPU
    "tiny": dict(image_size=32, patch_size=16, width=128, layers=2, heads=2, embed_dim=128),
    "small": dict(image_size=64, patch_size=16, width=256, layers=3, heads=4, embed_dim=256),
    # ViT-L/14 token / width geometry (257 tokens, width 1024, patch K = 588) with 2 layers, for parity tests
    "l14-2layer": dict(image_size=224, patch_size=14, width=1024, layers=2, heads=16, embed_dim=768),
    # ViT-B/32 token geometry (50 tokens, patch K = 3072) with 2 layers
    "b32-2layer": dict(image_size=224, patch_size=32, width=768, layers=2, heads=12, embed_dim=512),
}


# CLIP text towers paired with the visual geometries above (few_shot.py:849-886 build_model reads them off the
# checkpoint: transformer_width, heads = width // 64, layers, context 77, vocab 49408).
TEXT_GEOMETRIES = {
    "ViT-B/16": dict(width=512, layers=12, heads=8, context_length=77, vocab_size=49408),
    "ViT-L/14": dict(width=768, layers=12, heads=12, context_length=77, vocab_size=49408),
    "ViT-B/32": dict(width=512, layers=12, heads=8, context_length=77, vocab_size=49408),
    "tiny": dict(width=128, layers=2, heads=2, context_length=77, vocab_size=49408),
    "small": dict(width=256, layers=3, heads=4, context_length=77, vocab_size=49408),
}


def text_state_dict_shapes(tg, embed_dim):
    """CLIP state_dict keys of the text side (few_shot.py:735-744) -> shapes."""
    W = tg["width"]
    s = {"token_embedding.weight": (tg["vocab_size"], W), "positional_embedding": (tg["context_length"], W),
         "ln_final.weight": (W,), "ln_final.bias": (W,), "text_projection": (W, embed_dim)}
    for i in range(tg["layers"]):
        p = "transformer.resblocks.%d." % i
        s[p + "attn.in_proj_weight"] = (3 * W, W)
        s[p + "attn.in_proj_bias"] = (3 * W,)
        s[p + "attn.out_proj.weight"] = (W, W)
        s[p + "attn.out_proj.bias"] = (W,)
        for ln in ("ln_1", "ln_2"):
            s[p + ln + ".weight"] = (W,)
            s[p + ln + ".bias"] = (W,)
        s[p + "mlp.c_fc.weight"] = (4 * W, W)
        s[p + "mlp.c_fc.bias"] = (4 * W,)
        s[p + "mlp.c_proj.weight"] = (W, 4 * W)
        s[p + "mlp.c_proj.bias"] = (W,)
    return s


def synth_text_state_dict(tg, embed_dim, seed=3):
    """Seeded text-tower weights with spread-out statistics (perturbed LayerNorm affine terms, sharper attention) so the
    class embeddings differ visibly between prompts."""
    rng = np.random.default_rng(seed)
    out = {}
    W = tg["width"]
    for name, shape in text_state_dict_shapes(tg, embed_dim).items():
        if name == "token_embedding.weight":
            v = 0.5 * rng.standard_normal(shape, dtype=np.float32)
        elif name == "positional_embedding":
            v = 0.2 * rng.standard_normal(shape)
        elif name.endswith(".weight") and (".ln_" in name or name.startswith("ln_final")):
            v = 1.0 + 0.25 * rng.standard_normal(shape)
        elif name.endswith(".bias") and (".ln_" in name or name.startswith("ln_final")):
            v = 0.15 * rng.standard_normal(shape)
        elif name.endswith("bias"):
            v = 0.1 * rng.standard_normal(shape)
        elif name == "text_projection":
            v = W ** -0.5 * rng.standard_normal(shape)
        else:
            gain = 2.0 if "in_proj_weight" in name else 1.0
            v = gain * shape[1] ** -0.5 * rng.standard_normal(shape)
        out[name] = np.ascontiguousarray(v, dtype=np.float32)
    return out


def full_geometry(name, mod_depth=1):
    g = dict(GEOMETRIES[name])
    g.update(mod_heads=8, mod_dim_head=g["embed_dim"] // 8, mod_mlp_dim=2048, mod_depth=mod_depth)
    return g


def state_dict_shapes(g):
    """name -> shape of every parameter of the head (reference names without the 'head.' prefix)."""
    D, E, P = g["width"], g["embed_dim"], g["patch_size"]
    tokens = (g["image_size"] // P) ** 2 + 1
    inner, F = g["mod_heads"] * g["mod_dim_head"], g["mod_mlp_dim"]
    s = {"scale": (1,)}
    b = "backbone."
    s[b + "class_embedding"] = (D,)
    s[b + "positional_embedding"] = (tokens, D)
    s[b + "proj"] = (D, E)
    s[b + "conv1.weight"] = (D, 3, P, P)
    for ln in ("ln_pre", "ln_post"):
        s[b + ln + ".weight"] = (D,)
        s[b + ln + ".bias"] = (D,)
    for i in range(g["layers"]):
        p = b + "transformer.resblocks.%d." % i
        s[p + "attn.in_proj_weight"] = (3 * D, D)
        s[p + "attn.in_proj_bias"] = (3 * D,)
        s[p + "attn.out_proj.weight"] = (D, D)
        s[p + "attn.out_proj.bias"] = (D,)
        for ln in ("ln_1", "ln_2"):
            s[p + ln + ".weight"] = (D,)
            s[p + ln + ".bias"] = (D,)
        s[p + "mlp.c_fc.weight"] = (4 * D, D)
        s[p + "mlp.c_fc.bias"] = (4 * D,)
        s[p + "mlp.c_proj.weight"] = (D, 4 * D)
        s[p + "mlp.c_proj.bias"] = (D,)
    for l in range(g["mod_depth"]):
        p = "context2.layers.%d." % l
        s[p + "0.norm.weight"] = (E,)
        s[p + "0.norm.bias"] = (E,)
        for n in ("to_q", "to_k", "to_v"):
            s[p + "0.fn.%s.weight" % n] = (inner, E)
        s[p + "0.fn.to_out.0.weight"] = (E, inner)
        s[p + "0.fn.to_out.0.bias"] = (E,)
        s[p + "1.net.0.weight"] = (F, E)
        s[p + "1.net.0.bias"] = (F,)
        s[p + "1.net.3.weight"] = (E, F)
        s[p + "1.net.3.bias"] = (E,)
    return s


def synth_state_dict(g, seed=0, spread=True):
    """Random weights. `spread=True` perturbs LayerNorm affine terms and biases so that frame features, frame
    distances and logits are not degenerate (default init gives logits ~8.13 +- 0.01)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in state_dict_shapes(g).items():
        if name == "scale":
            v = np.full(shape, 1.0 if not spread else 1.5, dtype=np.float32)
        elif name.endswith("norm.weight") or (".ln_" in name and name.endswith(".weight")):
            v = 1.0 + (0.25 if spread else 0.0) * rng.standard_normal(shape)
        elif name.endswith("norm.bias") or (".ln_" in name and name.endswith(".bias")):
            v = (0.15 if spread else 0.0) * rng.standard_normal(shape)
        elif name.endswith(".bias") or name.endswith("in_proj_bias"):
            v = (0.1 if spread else 0.02) * rng.standard_normal(shape)
        elif name.endswith("class_embedding") or name.endswith("positional_embedding"):
            v = g["width"] ** -0.5 * rng.standard_normal(shape) * (4.0 if spread else 1.0)
        elif name.endswith("backbone.proj"):
            v = g["width"] ** -0.5 * rng.standard_normal(shape)
        else:  # dense weights [out, in, ...]
            fan_in = int(np.prod(shape[1:]))
            gain = 1.0
            if spread and ("in_proj_weight" in name or "to_q" in name or "to_k" in name):
                gain = 2.0  # sharper attention maps
            v = gain * fan_in ** -0.5 * rng.standard_normal(shape)
        out[name] = np.ascontiguousarray(v, dtype=np.float32)
    return out


def synth_text_features(n_classes, embed_dim, seed=7):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n_classes, embed_dim)).astype(np.float32)


def _smooth_field(rng, n, size, cells):
    """n random low-frequency fields [n, 3, size, size]: a coarse grid upsampled by pixel replication."""
    coarse = rng.standard_normal((n, 3, cells, cells)).astype(np.float32)
    rep = -(-size // cells)
    return np.kron(coarse, np.ones((1, 1, rep, rep), dtype=np.float32))[:, :, :size, :size]


def synth_episode(way=5, shot=1, queries_per_class=1, n_frames=8, image_size=224, n_test_classes=24, seed=1000,
                  structured=True):
    """One few-shot episode as the reference dataset produces it (ssv2_few_shot.py:190-285): support/target videos
    shuffled, labels as fp32. `structured` gives every class its own low-frequency pattern (plus per-video and
    per-frame variation) so that prototypes of different classes are actually different."""
    rng = np.random.default_rng(seed)
    S, Q, T = way * shot, way * queries_per_class, n_frames
    sup_cls = np.repeat(np.arange(way), shot)
    tgt_cls = np.repeat(np.arange(way), queries_per_class)
    sup_cls = sup_cls[rng.permutation(S)]
    tgt_cls = tgt_cls[rng.permutation(Q)]
    real_ids = rng.choice(n_test_classes, size=way, replace=False)

    def videos(cls_of_video, class_fields):
        n = len(cls_of_video)
        if not structured:
            return rng.standard_normal((n * T, 3, image_size, image_size)).astype(np.float32)
        vid = _smooth_field(rng, n, image_size, 4)
        drift = _smooth_field(rng, n, image_size, 2)
        out = np.empty((n, T, 3, image_size, image_size), dtype=np.float32)
        for t in range(T):
            phase = np.float32((t - (T - 1) / 2.0) / max(T, 1))
            out[:, t] = 0.9 * class_fields[cls_of_video] + 0.5 * vid + 1.2 * phase * drift
        out += 0.35 * rng.standard_normal(out.shape).astype(np.float32)
        return out.reshape(n * T, 3, image_size, image_size)

    class_fields = _smooth_field(rng, way, image_size, 7) if structured else None
    return {
        "support_set": videos(sup_cls, class_fields),
        "support_labels": sup_cls.astype(np.float32),
        "target_set": videos(tgt_cls, class_fields),
        "target_labels": tgt_cls.astype(np.float32),
        "real_support_labels": real_ids[sup_cls].astype(np.float32),
        "real_target_labels": real_ids[tgt_cls].astype(np.float32),
        "batch_class_list": real_ids.astype(np.float32),
    }


def ragged_support(task, n_frames, keep_counts):
    """Unequal shots per class (the reference averages whatever members a class has, few_shot.py:2949-2962): keep only
    the first keep_counts[c] support videos of class c, in their (shuffled) order of appearance."""
    labels = task["support_labels"].astype(np.int64)
    seen = {}
    keep = []
    for i, c in enumerate(labels):
        seen[c] = seen.get(c, 0) + 1
        if seen[c] <= keep_counts[c]:
            keep.append(i)
    keep = np.array(keep)
    out = dict(task)
    frames = task["support_set"].reshape(len(labels), n_frames, *task["support_set"].shape[1:])
    out["support_set"] = np.ascontiguousarray(frames[keep].reshape(-1, *task["support_set"].shape[1:]))
    out["support_labels"] = np.ascontiguousarray(task["support_labels"][keep])
    out["real_support_labels"] = np.ascontiguousarray(task["real_support_labels"][keep])
    return out


def synth_raw_frames(n_frames, height, width, seed=77):
    """uint8 [n, H, W, 3] "camera" frames (8x8 blocks plus noise, so interpolation errors are visible) for the
    pre-processing path; regenerated from the seed by tests instead of being stored in the fixtures."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(n_frames, height // 8 + 1, width // 8 + 1, 3))
    frames = np.kron(base, np.ones((1, 8, 8, 1), dtype=np.int64))[:, :height, :width, :]
    return np.clip(frames + rng.integers(-20, 21, size=frames.shape), 0, 255).astype(np.uint8)


# Algorithmic FLOPs of the frame encoder: 2 * MAC of patch-embed, QKV, QK^T, PV, out-proj, fc1, fc2
# and the final projection; no padding, no element-wise work.
def vit_flops_per_frame(g):
    D, E, P = g["width"], g["embed_dim"], g["patch_size"]
    G2 = (g["image_size"] // P) ** 2
    L = G2 + 1
    per_layer = L * D * 3 * D + 2 * g["heads"] * L * L * 64 + L * D * D + 2 * L * D * 4 * D
    return 2.0 * (G2 * 3 * P * P * D + g["layers"] * per_layer + D * E)
