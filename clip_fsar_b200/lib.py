"""ctypes binding of libfsar_sm100.so (C ABI in include/fsar.h).

The reference (CLIP-FSAR) is pure Python and has no FFI of its own; this is the stub a maintainer adds to reach
the sm_100a kernels from `models/base/few_shot.py`-style code. PyTorch is used only for
device memory and streams: every call takes `tensor.data_ptr()` and `torch.cuda.current_stream().cuda_stream`.

There is no fallback: a missing library or a missing sm_100 device raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

LIB_NAME = "libfsar_sm100.so"
# FSAR_LIB_PATH selects another build of the same sources (e.g. the -DFSAR_BF16 operand-type variant)
LIB_PATH = os.environ.get("FSAR_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

FSAR_PROF_CLASSES = 13
EPI_STORE16, EPI_QGELU16, EPI_RESID32, EPI_PATCH32, EPI_STORE32 = 0, 1, 2, 3, 4

# every symbol include/fsar.h declares (tests check the library exports exactly these)
SYMBOLS = [
    "fsar_version", "fsar_class_name", "fsar_create", "fsar_destroy", "fsar_last_error", "fsar_set_weight",
    "fsar_missing_weights", "fsar_missing_weight", "fsar_vit_forward", "fsar_modulate", "fsar_otam_logits",
    "fsar_episode_forward", "fsar_episode_forward_host", "fsar_episode_submit_host", "fsar_episode_collect_host",
    "fsar_episodes_forward", "fsar_episodes_submit_host", "fsar_episodes_collect_host",
    "fsar_preprocess_u8", "fsar_vit_forward_u8", "fsar_episodes_submit_host_u8", "fsar_text_configure", "fsar_text_encode", "fsar_metrics_update", "fsar_peek", "fsar_operand_dtype", "fsar_op_layernorm", "fsar_op_gemm", "fsar_op_attention", "fsar_op_f32_to_16",
    "fsar_launch_count", "fsar_profile_begin", "fsar_profile_end",
]


class FsarError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libfsar_sm100 error %d: %s" % (code, message))
        self.code = code


class FsarConfig(Structure):
    _fields_ = [
        ("image_size", c_int32), ("patch_size", c_int32), ("width", c_int32), ("layers", c_int32),
        ("heads", c_int32), ("embed_dim", c_int32), ("mod_heads", c_int32), ("mod_dim_head", c_int32),
        ("mod_mlp_dim", c_int32), ("mod_depth", c_int32), ("max_frames", c_int32), ("max_videos", c_int32),
        ("max_tokens", c_int32), ("max_classes", c_int32), ("max_batch", c_int32), ("otam_lambda", c_float),
        ("device", c_int32),
    ]


class FsarEpisode(Structure):
    _fields_ = [
        ("support_frames", c_void_p), ("target_frames", c_void_p), ("support_labels", c_void_p),
        ("real_support_labels", c_void_p), ("n_support", c_int32), ("n_target", c_int32), ("n_frames", c_int32),
        ("way", c_int32), ("merge_before", c_int32), ("single_direct", c_int32), ("text_mode", c_int32),
        ("text_coff", c_float),
    ]


class FsarTextConfig(Structure):
    _fields_ = [("width", c_int32), ("layers", c_int32), ("heads", c_int32), ("context_length", c_int32),
                ("vocab_size", c_int32)]


class FsarProfile(Structure):
    _fields_ = [
        ("ms", c_double * FSAR_PROF_CLASSES), ("launches", c_int64 * FSAR_PROF_CLASSES),
        ("flops", c_double * FSAR_PROF_CLASSES), ("bytes", c_double * FSAR_PROF_CLASSES),
    ]


_lib = None


def load_library(path=None):
    """dlopen the in-tree shared library and declare the prototypes. Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("FSAR_LIB_PATH") or LIB_PATH     # the environment is read at load time, not import time
    if not os.path.exists(p):
        raise FileNotFoundError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(clip_fsar_b200/build.py); there is no CPU fallback" % p)
    lib = ctypes.CDLL(p)
    H = c_void_p
    lib.fsar_version.restype = c_int
    lib.fsar_class_name.restype = c_char_p
    lib.fsar_class_name.argtypes = [c_int]
    lib.fsar_create.argtypes = [POINTER(FsarConfig), POINTER(H)]
    lib.fsar_destroy.argtypes = [H]
    lib.fsar_destroy.restype = None
    lib.fsar_last_error.argtypes = [H]
    lib.fsar_last_error.restype = c_char_p
    lib.fsar_set_weight.argtypes = [H, c_char_p, c_void_p, c_int64, c_int]
    lib.fsar_missing_weights.argtypes = [H]
    lib.fsar_missing_weight.argtypes = [H, c_int]
    lib.fsar_missing_weight.restype = c_char_p
    lib.fsar_vit_forward.argtypes = [H, c_void_p, c_int, c_void_p, c_void_p]
    lib.fsar_modulate.argtypes = [H, c_void_p, c_int, c_int, c_void_p, c_void_p]
    lib.fsar_otam_logits.argtypes = [H, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                     c_void_p]
    lib.fsar_episode_forward.argtypes = [H, POINTER(FsarEpisode), c_void_p, c_void_p, c_void_p]
    lib.fsar_episode_forward_host.argtypes = [H, POINTER(FsarEpisode), c_void_p, c_void_p]
    lib.fsar_episode_submit_host.argtypes = [H, c_int, POINTER(FsarEpisode)]
    lib.fsar_episode_collect_host.argtypes = [H, c_int, c_void_p, c_void_p]
    lib.fsar_episodes_forward.argtypes = [H, POINTER(FsarEpisode), c_int, c_void_p, c_void_p, c_void_p]
    lib.fsar_episodes_submit_host.argtypes = [H, c_int, POINTER(FsarEpisode), c_int]
    lib.fsar_episodes_collect_host.argtypes = [H, c_int, c_void_p, c_void_p]
    F3 = c_float * 3
    lib.fsar_preprocess_u8.argtypes = [H, c_void_p, c_int, c_int, c_int, c_int, c_int, F3, F3, c_void_p, c_void_p]
    lib.fsar_vit_forward_u8.argtypes = [H, c_void_p, c_int, c_int, c_int, c_int, c_int, F3, F3, c_void_p, c_void_p]
    lib.fsar_episodes_submit_host_u8.argtypes = [H, c_int, POINTER(FsarEpisode), c_int, c_int, c_int, c_int, c_int, F3, F3]
    lib.fsar_text_configure.argtypes = [H, POINTER(FsarTextConfig)]
    lib.fsar_text_encode.argtypes = [H, c_void_p, c_int, c_void_p, c_void_p]
    lib.fsar_metrics_update.argtypes = [H, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.fsar_peek.argtypes = [H, c_char_p, c_void_p, c_int64, c_void_p]
    lib.fsar_peek.restype = c_int64
    lib.fsar_operand_dtype.restype = c_int
    lib.fsar_op_layernorm.argtypes = [H, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.fsar_op_gemm.argtypes = [H, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.fsar_op_attention.argtypes = [H, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.fsar_op_f32_to_16.argtypes = [H, c_void_p, c_void_p, c_int64, c_void_p]
    lib.fsar_launch_count.argtypes = [H]
    lib.fsar_launch_count.restype = c_int64
    lib.fsar_profile_begin.argtypes = [H]
    lib.fsar_profile_end.argtypes = [H, POINTER(FsarProfile)]
    if path is None:
        _lib = lib
    return lib


def best_pass_frames(image_size, patch_size, n_sm=148):
    """Frames per ViT pass that make every GEMM a whole wave of 256-row blocks on the CTA pairs of one B200
    (74 pairs x 256 rows / tokens per frame): 96 for ViT-B/16 (197 tokens), 73 for ViT-L/14 (257 tokens). Larger passes
    only grow the activations past the 126 MB L2 (profiles/README.md, pass size sweep)."""
    tokens = (image_size // patch_size) ** 2 + 1
    return max(1, (n_sm // 2) * 256 // tokens)


def geometry(backbone_name, num_frames=8, max_frames=None, max_videos=10, max_classes=128, mod_depth=1, device=0,
             max_batch=1):
    """fsar_config for the CLIP visual towers CNN_OTAM_CLIPFSAR can be built on (few_shot.py:2705-2713).
    ViT-L/14 is an extension: the reference head has no branch for it (SURVEY.md headline finding 2)."""
    table = {
        "ViT-B/16": dict(image_size=224, patch_size=16, width=768, layers=12, heads=12, embed_dim=512),
        "ViT-B/32": dict(image_size=224, patch_size=32, width=768, layers=12, heads=12, embed_dim=512),
        "ViT-L/14": dict(image_size=224, patch_size=14, width=1024, layers=24, heads=16, embed_dim=768),
    }
    if backbone_name not in table:
        raise ValueError("unsupported VIDEO.HEAD.BACKBONE_NAME %r (supported: %s)" % (backbone_name, sorted(table)))
    g = dict(table[backbone_name])
    g.update(mod_heads=8, mod_dim_head=g["embed_dim"] // 8, mod_mlp_dim=2048, mod_depth=int(mod_depth),
             max_frames=int(max_frames or min(max_videos * num_frames, best_pass_frames(g["image_size"], g["patch_size"]))),
             max_videos=int(max_videos),
             max_tokens=int(num_frames), max_classes=int(max_classes), max_batch=int(max_batch), otam_lambda=0.5,
             device=int(device))
    return g


class Engine:
    """One fsar_handle: owns packed weights + workspace on one device. Not thread-safe (one process per GPU)."""

    def __init__(self, **cfg):
        import torch  # device memory + streams only

        self._torch = torch
        self.lib = load_library()
        cfg = dict(cfg)
        cfg.setdefault("max_batch", 1)
        self.cfg = FsarConfig(**cfg)
        self._h = c_void_p()
        rc = self.lib.fsar_create(byref(self.cfg), byref(self._h))
        if rc != 0:
            raise FsarError(rc, (self.lib.fsar_last_error(None) or b"").decode())
        self.device = torch.device("cuda", self.cfg.device)
        self._slot_refs = {}       # host tensors of the batches in flight, per slot (released at collect)
        self.operand_dtype = torch.bfloat16 if self.lib.fsar_operand_dtype() == 1 else torch.float16

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.fsar_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise FsarError(rc, (self.lib.fsar_last_error(self._h) or b"").decode())

    def _stream(self):
        return c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    def _f32(self, t, name):
        torch = self._torch
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError("%s must be a contiguous fp32 CUDA tensor (got %s %s)" % (name, t.dtype, t.device))
        return c_void_p(t.data_ptr())

    # ------------------------------------------------------------------ weights
    def set_weight(self, name, tensor):
        torch = self._torch
        t = tensor.detach()
        if t.dtype != torch.float32:
            t = t.float()
        t = t.contiguous()
        self._check(self.lib.fsar_set_weight(self._h, name.encode(), c_void_p(t.data_ptr()), t.numel(),
                                             1 if t.is_cuda else 0))

    def load_state_dict(self, state, prefix=""):
        """Push every tensor whose key (minus `prefix`) the library knows; returns the list of ignored keys."""
        ignored = []
        for k, v in state.items():
            if not k.startswith(prefix):
                ignored.append(k)
                continue
            try:
                self.set_weight(k[len(prefix):], v)
            except FsarError as e:
                if e.code != -4:
                    raise
                ignored.append(k)
        return ignored

    def missing_weights(self):
        n = self.lib.fsar_missing_weights(self._h)
        return [self.lib.fsar_missing_weight(self._h, i).decode() for i in range(n)]

    # ------------------------------------------------------------------ the path
    def vit_forward(self, frames):
        torch = self._torch
        S = self.cfg.image_size
        if frames.dim() != 4 or tuple(frames.shape[1:]) != (3, S, S):
            raise ValueError("frames must be [n, 3, %d, %d] for this engine, got %s" % (S, S, tuple(frames.shape)))
        n = frames.shape[0]
        out = torch.empty((n, self.cfg.embed_dim), dtype=torch.float32, device=self.device)
        self._check(self.lib.fsar_vit_forward(self._h, self._f32(frames, "frames"), n, c_void_p(out.data_ptr()),
                                              self._stream()))
        return out

    def modulate(self, x):
        torch = self._torch
        out = torch.empty_like(x)
        self._check(self.lib.fsar_modulate(self._h, self._f32(x, "x"), x.shape[0], x.shape[1], c_void_p(out.data_ptr()),
                                           self._stream()))
        return out

    def otam_logits(self, q, protos, single_direct=False, return_intermediates=False):
        torch = self._torch
        Q, T, _ = q.shape
        way = protos.shape[0]
        logits = torch.empty((Q, way), dtype=torch.float32, device=self.device)
        dists = torch.empty((Q, way, T, T), dtype=torch.float32, device=self.device) if return_intermediates else None
        cum = torch.empty((Q, way), dtype=torch.float32, device=self.device) if return_intermediates else None
        self._check(self.lib.fsar_otam_logits(
            self._h, self._f32(q, "q"), self._f32(protos, "protos"), Q, way, T, int(bool(single_direct)),
            c_void_p(logits.data_ptr()), c_void_p(dists.data_ptr()) if dists is not None else None,
            c_void_p(cum.data_ptr()) if cum is not None else None, self._stream()))
        return (logits, dists, cum) if return_intermediates else logits

    def _episode(self, support, target, support_labels, real_support_labels, n_frames, way, merge_before,
                 single_direct, text_mode=0, text_coff=0.9, frame_shape=None):
        if frame_shape is None:
            frame_shape = (3, self.cfg.image_size, self.cfg.image_size)
        for name, t in (("support_set", support), ("target_set", target)):
            if t.dim() != 4 or tuple(t.shape[1:]) != tuple(frame_shape):
                raise ValueError("%s must be [frames, %d, %d, %d] for this engine (DATA.TEST_CROP_SIZE / image_size %d), "
                                 "got %s" % ((name,) + tuple(frame_shape) + (self.cfg.image_size, tuple(t.shape))))
        S = support.shape[0] // n_frames
        Q = target.shape[0] // n_frames
        if S * n_frames != support.shape[0] or Q * n_frames != target.shape[0]:
            raise ValueError("frame counts %d / %d are not multiples of NUM_INPUT_FRAMES=%d" %
                             (support.shape[0], target.shape[0], n_frames))
        if support_labels.numel() != S or real_support_labels.numel() != S:
            raise ValueError("expected %d support labels, got %d / %d" %
                             (S, support_labels.numel(), real_support_labels.numel()))
        for name, t in (("support_labels", support_labels), ("real_support_labels", real_support_labels)):
            if t.dtype != self._torch.float32 or not t.is_contiguous():
                raise ValueError("%s must be contiguous fp32 (labels are fp32 tensors holding integers, "
                                 "ssv2_few_shot.py:278-283), got %s" % (name, t.dtype))
        ep = FsarEpisode(support.data_ptr(), target.data_ptr(), support_labels.data_ptr(),
                         real_support_labels.data_ptr(), S, Q, n_frames, way, int(bool(merge_before)),
                         int(bool(single_direct)), int(text_mode), float(text_coff))
        return ep, S, Q

    def episode_forward(self, support, target, support_labels, real_support_labels, n_frames, way,
                        merge_before=False, single_direct=False, n_train_classes=None, want_class_logits=True,
                        text_mode=0, text_coff=0.9):
        """Device-resident episode: returns (logits [Q, way], class_logits [S + Q, n_train] or None).
        text_mode 1 / 2 = TRAIN.EVAL_TEXT / TRAIN.COMBINE (class_logits is None there, as in the reference)."""
        if text_mode:
            want_class_logits = False
        torch = self._torch
        for name, t in (("support_set", support), ("target_set", target), ("support_labels", support_labels),
                        ("real_support_labels", real_support_labels)):
            self._f32(t, name)
        ep, S, Q = self._episode(support, target, support_labels, real_support_labels, n_frames, way, merge_before,
                                 single_direct, text_mode, text_coff)
        logits = torch.empty((Q, way), dtype=torch.float32, device=self.device)
        cl = None
        if want_class_logits:
            cl = torch.empty((S + Q, int(n_train_classes)), dtype=torch.float32, device=self.device)
        self._check(self.lib.fsar_episode_forward(self._h, byref(ep), c_void_p(logits.data_ptr()),
                                                  c_void_p(cl.data_ptr()) if cl is not None else None,
                                                  self._stream()))
        return logits, cl

    def episodes_forward(self, episodes, n_frames, way, merge_before=False, single_direct=False, n_train_classes=None,
                         want_class_logits=True):
        """Throughput form: `episodes` is a list of (support, target, support_labels, real_support_labels) DEVICE
        tensors with the same way / shot / n_frames. One ViT sweep over all frames (passes of cfg.max_frames frames),
        head per episode. Returns (logits [n, Q, way], class_logits [n, S + Q, n_train] or None)."""
        torch = self._torch
        n = len(episodes)
        arr = (FsarEpisode * n)()
        for i, (sup, tgt, sl, rl) in enumerate(episodes):
            for name, t in (("support_set", sup), ("target_set", tgt), ("support_labels", sl), ("real_support_labels", rl)):
                self._f32(t, name)
            arr[i], S, Q = self._episode(sup, tgt, sl, rl, n_frames, way, merge_before, single_direct)
        logits = torch.empty((n, Q, way), dtype=torch.float32, device=self.device)
        cl = torch.empty((n, S + Q, int(n_train_classes)), dtype=torch.float32, device=self.device) if want_class_logits else None
        self._check(self.lib.fsar_episodes_forward(self._h, arr, n, c_void_p(logits.data_ptr()),
                                                   c_void_p(cl.data_ptr()) if cl is not None else None, self._stream()))
        return logits, cl

    def episodes_submit_host(self, slot, episodes, n_frames, way, merge_before=False, single_direct=False):
        """`episodes`: list of (support, target, support_labels, real_support_labels) HOST (ideally pinned) tensors."""
        n = len(episodes)
        arr = (FsarEpisode * n)()
        for i, (sup, tgt, sl, rl) in enumerate(episodes):
            arr[i], S, Q = self._host_episode(sup, tgt, sl, rl, n_frames, way, merge_before, single_direct)
        self._check(self.lib.fsar_episodes_submit_host(self._h, slot, arr, n))
        self._slot_refs[slot] = list(episodes)      # fsar.h lifetime rule: the buffers live until collect returns
        return n, S, Q

    def episodes_collect_host(self, slot, logits_out, class_logits_out=None):
        try:
            self._check(self.lib.fsar_episodes_collect_host(
                self._h, slot, c_void_p(logits_out.data_ptr()),
                c_void_p(class_logits_out.data_ptr()) if class_logits_out is not None else None))
        finally:
            self._slot_refs.pop(slot, None)

    def _host_episode(self, support, target, support_labels, real_support_labels, n_frames, way, merge_before,
                      single_direct):
        torch = self._torch
        for name, t in (("support_set", support), ("target_set", target), ("support_labels", support_labels),
                        ("real_support_labels", real_support_labels)):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("%s must be a contiguous fp32 HOST tensor" % name)
        return self._episode(support, target, support_labels, real_support_labels, n_frames, way, merge_before,
                             single_direct)

    def episode_submit_host(self, slot, support, target, support_labels, real_support_labels, n_frames, way,
                            merge_before=False, single_direct=False):
        ep, S, Q = self._host_episode(support, target, support_labels, real_support_labels, n_frames, way,
                                      merge_before, single_direct)
        self._check(self.lib.fsar_episode_submit_host(self._h, slot, byref(ep)))
        self._slot_refs[slot] = (support, target, support_labels, real_support_labels)
        return S, Q

    def episode_collect_host(self, slot, logits_out, class_logits_out=None):
        try:
            self._check(self.lib.fsar_episode_collect_host(
                self._h, slot, c_void_p(logits_out.data_ptr()),
                c_void_p(class_logits_out.data_ptr()) if class_logits_out is not None else None))
        finally:
            self._slot_refs.pop(slot, None)

    def episode_forward_host(self, support, target, support_labels, real_support_labels, n_frames, way,
                             merge_before=False, single_direct=False, n_train_classes=0):
        """HOST buffers in, HOST results out; H2D/D2H copies happen inside the call (runs/test_net_few_shot.py:61-62
        + the .item() reads at 174-178)."""
        torch = self._torch
        S, Q = self.episode_submit_host(0, support, target, support_labels, real_support_labels, n_frames, way,
                                        merge_before, single_direct)
        logits = torch.empty((Q, way), dtype=torch.float32)
        cl = torch.empty((S + Q, int(n_train_classes)), dtype=torch.float32) if n_train_classes else None
        self.episode_collect_host(0, logits, cl)
        return logits, cl

    # CLIP normalisation constants of the shipped configs (DATA.MEAN / DATA.STD, CLIPFSAR_K100_1shot_v1.yaml)
    CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
    CLIP_STD = (0.26862954, 0.26130258, 0.27577711)

    def preprocess_u8(self, frames_u8, resize=(256, 256), mean=None, std=None):
        """uint8 [n, H, W, 3] CUDA -> fp32 [n, 3, S, S]: the reference loader's test-time transform, on the device."""
        torch = self._torch
        if not (frames_u8.is_cuda and frames_u8.dtype == torch.uint8 and frames_u8.is_contiguous() and frames_u8.dim() == 4
                and frames_u8.shape[3] == 3):
            raise ValueError("frames_u8 must be a contiguous uint8 CUDA tensor [n, H, W, 3]")
        n, Hh, Ww, _ = frames_u8.shape
        S = self.cfg.image_size
        out = torch.empty((n, 3, S, S), dtype=torch.float32, device=self.device)
        F3 = c_float * 3
        self._check(self.lib.fsar_preprocess_u8(self._h, c_void_p(frames_u8.data_ptr()), n, Hh, Ww, int(resize[0]),
                                                int(resize[1]), F3(*(mean or self.CLIP_MEAN)), F3(*(std or self.CLIP_STD)),
                                                c_void_p(out.data_ptr()), self._stream()))
        return out

    def vit_forward_u8(self, frames_u8, resize=(256, 256), mean=None, std=None):
        """uint8 [n, H, W, 3] CUDA -> frame features [n, embed_dim]; resize / crop / normalise happen inside the patch
        gather (no fp32 crop in HBM). Same numbers as vit_forward(preprocess_u8(frames_u8))."""
        torch = self._torch
        if not (frames_u8.is_cuda and frames_u8.dtype == torch.uint8 and frames_u8.is_contiguous() and frames_u8.dim() == 4
                and frames_u8.shape[3] == 3):
            raise ValueError("frames_u8 must be a contiguous uint8 CUDA tensor [n, H, W, 3]")
        n, Hh, Ww, _ = frames_u8.shape
        out = torch.empty((n, self.cfg.embed_dim), dtype=torch.float32, device=self.device)
        F3 = c_float * 3
        self._check(self.lib.fsar_vit_forward_u8(self._h, c_void_p(frames_u8.data_ptr()), n, Hh, Ww, int(resize[0]),
                                                 int(resize[1]), F3(*(mean or self.CLIP_MEAN)), F3(*(std or self.CLIP_STD)),
                                                 c_void_p(out.data_ptr()), self._stream()))
        return out

    def episodes_submit_host_u8(self, slot, episodes, n_frames, way, resize=(256, 256), mean=None, std=None,
                                merge_before=False, single_direct=False):
        """`episodes`: list of (support_u8 [S*T, H, W, 3], target_u8 [Q*T, H, W, 3], support_labels, real_support_labels)
        HOST tensors (uint8 frames, fp32 labels). Raw bytes cross PCIe; pre-processing runs on the device."""
        torch = self._torch
        n = len(episodes)
        arr = (FsarEpisode * n)()
        Hh, Ww = episodes[0][0].shape[1:3]
        for i, (sup, tgt, sl, rl) in enumerate(episodes):
            for name, t in (("support_u8", sup), ("target_u8", tgt)):
                if t.is_cuda or t.dtype != torch.uint8 or not t.is_contiguous() or tuple(t.shape[1:]) != (Hh, Ww, 3):
                    raise ValueError("%s must be a contiguous uint8 HOST tensor [frames, %d, %d, 3]" % (name, Hh, Ww))
            S, Q = sup.shape[0] // n_frames, tgt.shape[0] // n_frames
            if S * n_frames != sup.shape[0] or Q * n_frames != tgt.shape[0]:
                raise ValueError("frame counts %d / %d are not multiples of NUM_INPUT_FRAMES=%d" % (sup.shape[0], tgt.shape[0], n_frames))
            for name, t in (("support_labels", sl), ("real_support_labels", rl)):
                if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != S:
                    raise ValueError("%s must be a contiguous fp32 HOST tensor with %d elements, got %s %s" %
                                     (name, S, t.dtype, tuple(t.shape)))
            arr[i] = FsarEpisode(sup.data_ptr(), tgt.data_ptr(), sl.data_ptr(), rl.data_ptr(), S, Q, n_frames, way,
                                 int(bool(merge_before)), int(bool(single_direct)), 0, 0.9)
        F3 = c_float * 3
        self._check(self.lib.fsar_episodes_submit_host_u8(self._h, slot, arr, n, Hh, Ww, int(resize[0]), int(resize[1]),
                                                          F3(*(mean or self.CLIP_MEAN)), F3(*(std or self.CLIP_STD))))
        self._slot_refs[slot] = list(episodes)
        return n, S, Q

    def metrics_update(self, logits, target_labels, counters, per_class=None):
        """counters (int64[3], device) += {top-1 hits, queries, round(sum cross-entropy * 1e6)}; no host sync."""
        torch = self._torch
        Q, way = logits.reshape(-1, logits.shape[-1]).shape
        if counters.dtype != torch.int64 or not counters.is_cuda or counters.numel() < 3:
            raise ValueError("counters must be a CUDA int64 tensor with 3 elements")
        self._check(self.lib.fsar_metrics_update(
            self._h, self._f32(logits, "logits"), self._f32(target_labels, "target_labels"), Q, way,
            c_void_p(counters.data_ptr()), c_void_p(per_class.data_ptr()) if per_class is not None else None,
            self._stream()))
        return counters

    # ------------------------------------------------------------------ CLIP text tower (init-time text features)
    def text_configure(self, width, layers, heads, context_length=77, vocab_size=49408):
        """Register the text tower's weights ('clip.*' = CLIP state_dict keys); then push them with set_weight /
        load_state_dict(prefix) and call text_encode."""
        tc = FsarTextConfig(width=int(width), layers=int(layers), heads=int(heads), context_length=int(context_length),
                            vocab_size=int(vocab_size))
        self._check(self.lib.fsar_text_configure(self._h, byref(tc)))
        self.text_cfg = tc

    def load_clip_text_state_dict(self, clip_state):
        """Push the text-side tensors of an OpenAI CLIP state_dict (CLIP.encode_text, few_shot.py:793-806): everything
        except 'visual.*', 'logit_scale' and the non-parameter entries of TorchScript archives."""
        skip = ("visual.", "logit_scale", "input_resolution", "context_length", "vocab_size")
        n = 0
        for k, v in clip_state.items():
            if k.startswith(skip):
                continue
            self.set_weight("clip." + k, v)
            n += 1
        return n

    def text_encode(self, tokens):
        """tokens: int [n, context_length] (tokenize(), few_shot.py:393-429) -> fp32 [n, embed_dim] on the device."""
        torch = self._torch
        t = tokens.to(device=self.device, dtype=torch.int32).contiguous()
        tc = getattr(self, "text_cfg", None)
        if tc is None:
            raise FsarError(-5, "text_encode: call text_configure first")
        if t.dim() != 2 or t.shape[1] != tc.context_length:
            raise ValueError("tokens must be [n, %d], got %s" % (tc.context_length, tuple(t.shape)))
        out = torch.empty(t.shape[0], self.cfg.embed_dim, device=self.device, dtype=torch.float32)
        self._check(self.lib.fsar_text_encode(self._h, c_void_p(t.data_ptr()), t.shape[0], c_void_p(out.data_ptr()),
                                              self._stream()))
        return out

    def peek(self, name, shape, dtype=None):
        torch = self._torch
        out = torch.empty(shape, dtype=dtype or torch.float32)
        n = self.lib.fsar_peek(self._h, name.encode(), c_void_p(out.data_ptr()), out.numel(), self._stream())
        if n < 0:
            self._check(int(n))
        if n != out.numel():
            raise FsarError(-1, "tap %s holds %d elements, asked for %d" % (name, n, out.numel()))
        return out

    # ------------------------------------------------------------------ single operators (parity tests)
    def op_f32_to_16(self, x):
        torch = self._torch
        out = torch.empty(x.shape, dtype=self.operand_dtype, device=self.device)
        self._check(self.lib.fsar_op_f32_to_16(self._h, self._f32(x, "x"), c_void_p(out.data_ptr()), x.numel(),
                                               self._stream()))
        return out

    def op_layernorm(self, x, gamma, beta, out16=False):
        torch = self._torch
        rows, dim = x.shape
        out = torch.empty((rows, dim), dtype=self.operand_dtype if out16 else torch.float32, device=self.device)
        self._check(self.lib.fsar_op_layernorm(self._h, self._f32(x, "x"), self._f32(gamma, "gamma"),
                                               self._f32(beta, "beta"), rows, dim, int(out16),
                                               c_void_p(out.data_ptr()), self._stream()))
        return out

    def op_gemm(self, a16, w16, bias=None, epi=EPI_STORE32, out=None):
        torch = self._torch
        M, K = a16.shape
        N, K2 = w16.shape
        if K != K2 or a16.dtype != self.operand_dtype or w16.dtype != self.operand_dtype:
            raise ValueError("op_gemm: A [M,K] and W [N,K] must share K and be %s" % self.operand_dtype)
        if out is None:
            dt = self.operand_dtype if epi in (EPI_STORE16, EPI_QGELU16) else torch.float32
            out = torch.zeros((M, N), dtype=dt, device=self.device)
        self._check(self.lib.fsar_op_gemm(self._h, c_void_p(a16.data_ptr()), c_void_p(w16.data_ptr()),
                                          self._f32(bias, "bias") if bias is not None else None, M, N, K, epi,
                                          c_void_p(out.data_ptr()), self._stream()))
        return out

    def op_attention(self, qkv16, n_frames, L, heads):
        torch = self._torch
        out = torch.empty((n_frames * L, heads * 64), dtype=self.operand_dtype, device=self.device)
        self._check(self.lib.fsar_op_attention(self._h, c_void_p(qkv16.data_ptr()), n_frames, L, heads,
                                               c_void_p(out.data_ptr()), self._stream()))
        return out

    # ------------------------------------------------------------------ instrumentation
    def launch_count(self):
        return int(self.lib.fsar_launch_count(self._h))

    def profile_begin(self):
        self._check(self.lib.fsar_profile_begin(self._h))

    def profile_end(self):
        prof = FsarProfile()
        self._check(self.lib.fsar_profile_end(self._h, byref(prof)))
        out = {}
        for k in range(FSAR_PROF_CLASSES):
            if prof.launches[k]:
                out[self.lib.fsar_class_name(k).decode()] = dict(ms=prof.ms[k], launches=int(prof.launches[k]),
                                                                 flops=prof.flops[k], bytes=prof.bytes[k])
        return out
