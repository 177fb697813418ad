"""CNN_OTAM_CLIPFSAR_SM100 — drop-in head for the reference's HEAD_REGISTRY (models/base/base_blocks.py:21).

Same constructor `(cfg)`, same `forward(dict) -> {'logits', 'class_logits'}` and the SAME parameter names as the
reference head CNN_OTAM_CLIPFSAR (models/base/few_shot.py:2690-2993), so `utils/checkpoint.py:329`
`load_state_dict(strict=False)` fills it from a reference `.pyth` checkpoint. All compute goes through the C ABI
of libfsar_sm100.so (clip_fsar_b200/lib.py); PyTorch only owns the fp32 master parameters, device buffers and the
stream. There is no torch / CPU fallback: forward on a machine without an sm_100 GPU raises.

Scope: the eval forward (few_shot.py:2834-2990) — the visual else-branch honouring
TRAIN.MERGE_BEFORE, TRAIN.SINGLE_DIRECT, TRAIN.TRANSFORMER_DEPTH and DATA.NUM_INPUT_FRAMES, and the text branches
TRAIN.EVAL_TEXT / TRAIN.COMBINE (+ TEXT_COFF). Training mode raises NotImplementedError.

text_features_{train,test} (few_shot.py:2714-2728): given a CLIP checkpoint (VIDEO.HEAD.CLIP_CHECKPOINT) the class
prompts are tokenised on the host (the reference's own BPE tokenizer) and encoded by the library's text tower
(fsar_text_encode) on first use; otherwise they are passed in / loaded from a .pt file.
"""
import os
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import lib as _lib
from . import synth as _synth


class _Holder(nn.Module):
    """Parameter container: exists only so that state_dict() reproduces the reference key names."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the compute lives in libfsar_sm100.so")


def _linear_params(out_f, in_f, bias=True):
    h = _Holder()
    h.weight = nn.Parameter(torch.empty(out_f, in_f))
    if bias:
        h.bias = nn.Parameter(torch.empty(out_f))
    return h


def _ln_params(dim):
    h = _Holder()
    h.weight = nn.Parameter(torch.ones(dim))
    h.bias = nn.Parameter(torch.zeros(dim))
    return h


def _build_backbone(g):
    """Parameter tree of VisionTransformer (few_shot.py:655-669) / ResidualAttentionBlock (620-631)."""
    D, E, P = g["width"], g["embed_dim"], g["patch_size"]
    tokens = (g["image_size"] // P) ** 2 + 1
    bb = _Holder()
    bb.conv1 = _Holder()
    bb.conv1.weight = nn.Parameter(torch.empty(D, 3, P, P))
    bb.class_embedding = nn.Parameter(torch.empty(D))
    bb.positional_embedding = nn.Parameter(torch.empty(tokens, D))
    bb.ln_pre = _ln_params(D)
    bb.transformer = _Holder()
    blocks = []
    for _ in range(g["layers"]):
        blk = _Holder()
        blk.attn = _Holder()
        blk.attn.in_proj_weight = nn.Parameter(torch.empty(3 * D, D))
        blk.attn.in_proj_bias = nn.Parameter(torch.empty(3 * D))
        blk.attn.out_proj = _linear_params(D, D)
        blk.ln_1 = _ln_params(D)
        blk.mlp = _Holder()
        blk.mlp.c_fc = _linear_params(4 * D, D)
        blk.mlp.c_proj = _linear_params(D, 4 * D)
        blk.ln_2 = _ln_params(D)
        blocks.append(blk)
    bb.transformer.resblocks = nn.ModuleList(blocks)
    bb.ln_post = _ln_params(D)
    bb.proj = nn.Parameter(torch.empty(D, E))
    return bb


def _build_context2(g):
    """Parameter tree of Transformer_v1 (few_shot.py:979-988): layers.{l}.0 = PreNormattention_qkv(norm, fn =
    Attention_qkv(to_q, to_k, to_v, to_out.0)), layers.{l}.1 = FeedForward(net.0, net.3)."""
    E, inner, Fh = g["embed_dim"], g["mod_heads"] * g["mod_dim_head"], g["mod_mlp_dim"]
    ctx = _Holder()
    layers = []
    for _ in range(g["mod_depth"]):
        att = _Holder()
        att.norm = _ln_params(E)
        att.fn = _Holder()
        att.fn.to_q = _linear_params(inner, E, bias=False)
        att.fn.to_k = _linear_params(inner, E, bias=False)
        att.fn.to_v = _linear_params(inner, E, bias=False)
        att.fn.to_out = nn.ModuleDict({"0": _linear_params(E, inner)})
        ff = _Holder()
        ff.net = nn.ModuleDict({"0": _linear_params(Fh, E), "3": _linear_params(E, Fh)})
        layers.append(nn.ModuleList([att, ff]))
    ctx.layers = nn.ModuleList(layers)
    return ctx


def _cfg_get(node, name, default=None):
    return getattr(node, name, default) if node is not None and hasattr(node, name) else default


class CNN_OTAM_CLIPFSAR_SM100(nn.Module):
    """B200 (sm_100a) implementation of CNN_OTAM_CLIPFSAR's inference forward. Register with
    `clip_fsar_b200.register.register()` and select with `VIDEO.HEAD.NAME: CNN_OTAM_CLIPFSAR_SM100`."""

    def __init__(self, cfg, text_features_train=None, text_features_test=None, tokenizer=None):
        super().__init__()
        self.args = cfg
        name = cfg.VIDEO.HEAD.BACKBONE_NAME
        if name not in _synth.GEOMETRIES:
            raise ValueError("CNN_OTAM_CLIPFSAR_SM100 supports ViT backbones %s, got %r (RN50 is out of scope)"
                             % (sorted(_synth.GEOMETRIES), name))
        depth = int(_cfg_get(cfg.TRAIN, "TRANSFORMER_DEPTH", 0) or 1)                 # few_shot.py:2736-2739
        self.geometry = _synth.full_geometry(name, depth)
        self.mid_dim = self.geometry["embed_dim"]
        self.num_frames = int(cfg.DATA.NUM_INPUT_FRAMES)
        self.merge_before = bool(_cfg_get(cfg.TRAIN, "MERGE_BEFORE", False))          # 2949
        self.single_direct = bool(_cfg_get(cfg.TRAIN, "SINGLE_DIRECT", False))        # 2979
        # text branches of the eval forward: EVAL_TEXT wins over COMBINE (few_shot.py:2835 / 2855, if / elif)
        self.text_mode = 1 if _cfg_get(cfg.TRAIN, "EVAL_TEXT", False) else (2 if _cfg_get(cfg.TRAIN, "COMBINE", False) else 0)
        self.text_coff = float(_cfg_get(cfg.TRAIN, "TEXT_COFF", 0) or 0.9)            # 2925-2928
        self.class_real_train = list(_cfg_get(cfg.TRAIN, "CLASS_NAME", []) or [])
        self.class_real_test = list(_cfg_get(cfg.TEST, "CLASS_NAME", []) or [])

        self.backbone = _build_backbone(self.geometry)
        self.context2 = _build_context2(self.geometry)
        self.scale = nn.Parameter(torch.ones(1))                                      # 2733-2734
        self.mid_layer = nn.Sequential()
        self.classification_layer = nn.Sequential()
        self._init_parameters()

        head_cfg = cfg.VIDEO.HEAD
        self._text_state = None          # text side of the CLIP checkpoint, encoded on the device at first use
        self._text_tokens = None
        ckpt = _cfg_get(head_cfg, "CLIP_CHECKPOINT", None)
        if ckpt:
            self._text_state = self.load_clip_visual(ckpt)
        # text_features_{train,test}: plain attributes, not buffers (few_shot.py:2720, 2728)
        tf_path = _cfg_get(head_cfg, "TEXT_FEATURES", None)
        if text_features_train is None and tf_path:
            blob = torch.load(tf_path, map_location="cpu")
            text_features_train, text_features_test = blob["train"], blob["test"]
        if text_features_train is None and self._text_state:
            # few_shot.py:2714-2728: tokenize(prompt.format(class)) -> encode_text, with the library's text tower
            self.set_clip_text(self._text_state, tokenizer, _cfg_get(cfg.TEST, "PROMPT", None))
            text_features_train = torch.zeros(len(self.class_real_train), self.mid_dim)     # placeholders until the
            text_features_test = torch.zeros(len(self.class_real_test), self.mid_dim)       # first forward on a GPU
        if text_features_train is None:
            if not _cfg_get(head_cfg, "SYNTHETIC_TEXT", False):
                raise ValueError("text features are required: pass text_features_train/test, or set VIDEO.HEAD."
                                 "TEXT_FEATURES to a {'train','test'} .pt file, or VIDEO.HEAD.SYNTHETIC_TEXT: true")
            text_features_train = torch.from_numpy(
                _synth.synth_text_features(max(len(self.class_real_train), 1), self.mid_dim, 7))
            text_features_test = torch.from_numpy(
                _synth.synth_text_features(max(len(self.class_real_test), 1), self.mid_dim, 8))
        self.text_features_train = torch.as_tensor(text_features_train, dtype=torch.float32)
        self.text_features_test = torch.as_tensor(text_features_test, dtype=torch.float32)

        self._engine = None
        self._pushed_versions = None
        self._max_videos = int(_cfg_get(head_cfg, "MAX_VIDEOS", 0) or 0)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._mark_dirty())

    # ------------------------------------------------------------------ parameters
    def _init_parameters(self):
        """Random init in the spirit of the reference constructors (few_shot.py:661-669; nn.Linear defaults).
        Real runs overwrite it from a checkpoint."""
        g = self.geometry
        with torch.no_grad():
            for name, p in self.named_parameters():
                if name == "scale" or ".ln_" in name or name.endswith("norm.weight") or name.endswith("norm.bias"):
                    continue
                if name.endswith("bias"):
                    p.zero_()
                elif name in ("backbone.class_embedding", "backbone.positional_embedding", "backbone.proj"):
                    p.normal_(0.0, g["width"] ** -0.5)
                else:
                    fan_in = p[0].numel()
                    p.normal_(0.0, fan_in ** -0.5)

    def load_clip_visual(self, path):
        """Fill `backbone.*` from an OpenAI CLIP checkpoint (state_dict or TorchScript archive): keys 'visual.*'.
        Returns the text side of the checkpoint (CLIP.encode_text's parameters) or None if it has none."""
        try:
            sd = torch.jit.load(path, map_location="cpu").state_dict()
        except RuntimeError:
            sd = torch.load(path, map_location="cpu")
            sd = sd.get("state_dict", sd)
        vis = {k[len("visual."):]: v.float() for k, v in sd.items() if k.startswith("visual.")}
        missing = self.backbone.load_state_dict(vis, strict=False)
        if missing.missing_keys:
            raise ValueError("CLIP checkpoint %s lacks visual keys: %s" % (path, missing.missing_keys[:5]))
        text = {k: v.float() for k, v in sd.items() if torch.is_tensor(v) and v.dim() > 0 and not k.startswith("visual.")}
        return text if "token_embedding.weight" in text else None

    def set_clip_text(self, clip_text_state, tokenizer=None, prompt=None):
        """Have text_features_{train,test} computed by the library's CLIP text tower from `clip_text_state` (the
        non-visual tensors of a CLIP state_dict) at the next forward. Prompts as in few_shot.py:2714-2727:
        TEST.PROMPT.format(name) or "a photo of {name}". `tokenizer(list[str]) -> int [n, 77]` defaults to the
        reference's tokenize (few_shot.py:393-429; importable once clip_fsar_b200.register.register() ran)."""
        if tokenizer is None:
            try:
                from models.base.few_shot import tokenize as tokenizer
            except ImportError as e:
                raise ImportError("no tokenizer: pass tokenizer=, or make the reference tree importable "
                                  "(clip_fsar_b200.register.register())") from e
        fmt = prompt if prompt else "a photo of {}"
        self._text_tokens = tuple(torch.as_tensor(tokenizer([fmt.format(n) for n in names])).to(torch.int32)
                                  for names in (self.class_real_train, self.class_real_test))
        W = clip_text_state["ln_final.weight"].shape[0]
        layers = 1 + max(int(k.split(".")[2]) for k in clip_text_state if k.startswith("transformer.resblocks."))
        self._text_geometry = dict(width=W, layers=layers, heads=W // 64,
                                   context_length=clip_text_state["positional_embedding"].shape[0],
                                   vocab_size=clip_text_state["token_embedding.weight"].shape[0])
        self._text_state = {k: v for k, v in clip_text_state.items() if k != "logit_scale"}
        self._text_pending = True
        self._mark_dirty()

    def _mark_dirty(self):
        self._pushed_versions = None

    def set_text_features(self, train, test):
        """Replace text_features_{train,test} ([n_cls, embed_dim] fp32; few_shot.py:2720, 2728)."""
        self.text_features_train = torch.as_tensor(train, dtype=torch.float32)
        self.text_features_test = torch.as_tensor(test, dtype=torch.float32)
        self._text_tokens, self._text_pending = None, False        # explicit features replace the text tower's
        if self._engine is not None and max(self.text_features_train.shape[0], self.text_features_test.shape[0]) \
                > self._engine.cfg.max_classes:
            self._engine.close()
            self._engine = None
        self._mark_dirty()

    def _ensure_engine(self, device, n_videos):
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        if self._engine is not None and (n_videos > self._engine.cfg.max_videos or self._engine.cfg.device != dev_index):
            self._engine.close()          # grown episode, or the module moved to another GPU (module.to / DDP device_ids)
            self._engine = None
        if self._engine is None:
            cap = max(n_videos, self._max_videos, 10)
            g = dict(self.geometry)
            n_cls = max(self.text_features_train.shape[0], self.text_features_test.shape[0], 1)
            # one pass when the episode fits a wave of 256-row blocks (80 frames for 5-way 1-shot), else whole-wave passes
            g.update(max_frames=min(cap * self.num_frames, _lib.best_pass_frames(g["image_size"], g["patch_size"])),
                     max_videos=cap, max_tokens=self.num_frames,
                     max_classes=n_cls, otam_lambda=0.5, device=dev_index)
            self._engine = _lib.Engine(**g)
            self._pushed_versions = None
            if self._text_tokens is not None:
                self._text_pending = True
        if getattr(self, "_text_pending", False):
            # CLIP.encode_text on the device (fsar_text_encode), once per engine
            if getattr(self._engine, "text_cfg", None) is None:
                self._engine.text_configure(**self._text_geometry)
                self._engine.load_clip_text_state_dict(self._text_state)
            self.text_features_train = self._engine.text_encode(self._text_tokens[0]).cpu()
            self.text_features_test = self._engine.text_encode(self._text_tokens[1]).cpu()
            self._text_pending = False
            self._pushed_versions = None
        versions = tuple(p._version for p in self.parameters())
        if versions != self._pushed_versions:
            for name, p in self.named_parameters():
                self._engine.set_weight(name, p.data)
            self._engine.set_weight("text_features_train", self.text_features_train)
            self._engine.set_weight("text_features_test", self.text_features_test)
            missing = self._engine.missing_weights()
            if missing:
                raise RuntimeError("libfsar_sm100: weights never set: %s" % missing[:5])
            self._pushed_versions = versions
        return self._engine

    # ------------------------------------------------------------------ the reference interface
    def forward(self, inputs):
        """inputs: the task dict of runs/test_net_few_shot.py:59-62 (few_shot.py:2773). Returns
        {'logits': [Q, way], 'class_logits': [S + Q, n_train_classes]} on the inputs' device."""
        if self.training:
            raise NotImplementedError("CNN_OTAM_CLIPFSAR_SM100 is the inference path: call model.eval() first "
                                      "(training/backward stays with the reference head)")
        support, target = inputs["support_set"], inputs["target_set"]
        if not support.is_cuda:
            raise _lib.FsarError(-2, "inputs are on %s: libfsar_sm100 has no CPU path (needs an sm_100 GPU)" % support.device)
        S_img = self.geometry["image_size"]
        for name, t in (("support_set", support), ("target_set", target)):
            if t.dim() != 4 or tuple(t.shape[1:]) != (3, S_img, S_img):
                # the reference fails here with a positional_embedding shape mismatch (few_shot.py:676)
                raise ValueError("%s has shape %s but %s expects frames of [3, %d, %d]: set DATA.TEST_CROP_SIZE / "
                                 "DATA.TRAIN_CROP_SIZE to %d" % (name, tuple(t.shape), self.args.VIDEO.HEAD.BACKBONE_NAME,
                                                                 S_img, S_img, S_img))
        support_labels = inputs["support_labels"]
        real = inputs["real_support_labels"]
        if self.text_features_test.shape[0] < 1:
            raise ValueError("text_features_test is empty: TEST.CLASS_NAME must list the test classes")
        T = self.num_frames
        if support.shape[0] % T or target.shape[0] % T:
            raise ValueError("frame counts %d / %d are not multiples of DATA.NUM_INPUT_FRAMES = %d" %
                             (support.shape[0], target.shape[0], T))
        S, Q = support.shape[0] // T, target.shape[0] // T
        if "batch_class_list" in inputs and inputs["batch_class_list"].numel() > 0:
            way = int(inputs["batch_class_list"].numel())          # shape-derived: no host sync
        else:
            way = int(torch.unique(support_labels).numel())        # what the reference does (few_shot.py:2965)
        eng = self._ensure_engine(support.device, S + Q)

        def f32(t):
            return t.detach().to(dtype=torch.float32).contiguous()

        logits, class_logits = eng.episode_forward(
            f32(support), f32(target), f32(support_labels).reshape(-1), f32(real).reshape(-1), T, way,
            self.merge_before, self.single_direct, n_train_classes=self.text_features_train.shape[0],
            text_mode=self.text_mode, text_coff=self.text_coff)
        return {"logits": logits, "class_logits": class_logits}

    def loss(self, task_dict, model_dict):
        """few_shot.py:2992-2993."""
        return F.cross_entropy(model_dict["logits"], task_dict["target_labels"].long())

    def engine(self):
        return self._engine


def warn_if_reference_missing():
    if not os.path.isdir("/root/reference"):
        warnings.warn("reference tree not present: HEAD_REGISTRY registration skipped")
