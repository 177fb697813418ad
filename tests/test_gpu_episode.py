"""GPU: the whole path through the C ABI against (a) the committed fixtures produced by the reference's forward,
(b) the CPU oracle with 16-bit operand emulation, and (c) size-independent properties at BASELINE.json's full size.

Tolerances. The CUDA path stores GEMM operands in fp16 (fp32 accumulation, fp32 residual stream / LN / softmax
statistics, fp32 head); the reference is fp32 end to end. north_star's bound is 1e-3 relative on the logits:
  * default-init weights (the configuration BASELINE.json names) -- EVERY flag variant has a `*_di` / `*_default_init`
    fixture produced by the reference: max|dlogit| / max|logit| <= 1e-3, AND the same bound on the row-centred logits
    (l - rowmean(l)), which removes the T-proportional soft-min offset that inflates max|logit|;
  * "spread" weights (sharpened attention, perturbed LN affine terms, see synth.py) are a deliberately harder
    stress case where fp16 operand rounding ALONE (oracle with operand_dtype=fp16, no GPU involved) already gives
    ~1.4e-3, so the bound there is 3e-3 on logits / 5e-3 rel-L2 on features, plus a TIGHT bound against the
    16-bit-emulating oracle which isolates kernel bugs from rounding.
"""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, regenerate

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel_l2(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def rel_max(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def rel_centred(a, b):
    """max |(a - rowmean(a)) - (b - rowmean(b))| / max|b|: the additive OTAM offset (proportional to T, doubled by the
    bidirectional sum) carries no class information; this is the error on what the argmax actually sees."""
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    ac, bc = a - a.mean(dim=1, keepdim=True), b - b.mean(dim=1, keepdim=True)
    return float((ac - bc).abs().max() / (b.abs().max() + 1e-30))


def make_engine(lib, meta, g, sd, tt, te, max_frames=None):
    n_vid = meta["way"] * (meta["shot"] + meta.get("qpc", 1))
    e = lib.Engine(**dict(g, max_frames=max_frames or n_vid * meta["T"], max_videos=n_vid, max_tokens=meta["T"],
                          max_classes=128, otam_lambda=0.5, device=0))
    ignored = e.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    assert ignored == []
    e.set_weight("text_features_train", torch.from_numpy(tt))
    e.set_weight("text_features_test", torch.from_numpy(te))
    assert e.missing_weights() == []
    return e


def run(e, meta, task, host=False):
    t = {k: torch.from_numpy(v) for k, v in task.items()}
    args = ("support_set", "target_set", "support_labels", "real_support_labels")
    if host:
        return e.episode_forward_host(*[t[k].pin_memory() for k in args], meta["T"], meta["way"], meta["merge_before"],
                                      meta["single_direct"], n_train_classes=meta["n_train"])
    return e.episode_forward(*[t[k].to(DEV) for k in args], meta["T"], meta["way"], meta["merge_before"],
                             meta["single_direct"], n_train_classes=meta["n_train"], text_mode=meta.get("text_mode", 0),
                             text_coff=meta.get("text_coff", 0.9))


@pytest.mark.parametrize("name", golden_names())
def test_episode_matches_reference_fixture(lib, name):
    meta, ref = load_golden(name)
    g, sd, tt, te, task = regenerate(meta)
    e = make_engine(lib, meta, g, sd, tt, te)
    logits, class_logits = run(e, meta, task)
    S, Q, T, E, way = len(task["support_labels"]), len(task["target_labels"]), meta["T"], g["embed_dim"], meta["way"]
    tol_logits = 3e-3 if meta["spread"] else 1e-3
    mode = meta.get("text_mode", 0)
    slim = ref["support_feats"].size == 0
    if not slim:
        assert rel_l2(e.peek("support_feats", (S, T, E)), ref["support_feats"]) < 5e-3
        assert rel_l2(e.peek("target_feats", (Q, T, E)), ref["target_feats"]) < 5e-3
    if mode != 1 and not slim:
        assert float(np.abs(e.peek("dists", (Q, way, T, T)).numpy() - ref["dists"]).max()) < 3e-3
    assert rel_max(logits, ref["logits"]) < tol_logits
    assert rel_centred(logits, ref["logits"]) < tol_logits
    assert logits.shape == ref["logits"].shape
    if mode == 0:
        assert rel_max(class_logits, ref["class_logits"]) < 3e-3 and class_logits.shape == ref["class_logits"].shape
        # argmax wherever the reference's own top-2 margin is above the stated tolerance (random-init rows can be tied
        # to 1e-4 of the logit scale; a tie is not a parity failure)
        top2 = np.sort(ref["logits"], axis=1)[:, -2:]
        sure = (top2[:, 1] - top2[:, 0]) > 2 * tol_logits * np.abs(ref["logits"]).max()
        assert (logits.cpu().numpy().argmax(1)[sure] == ref["logits"].argmax(1)[sure]).all()
    else:
        # text branches return probabilities (rows sum to ... <= 1) and class_logits = None (few_shot.py:2852, 2930);
        # the argmax is only compared where the reference's top-2 margin exceeds the 16-bit noise
        assert class_logits is None
        top2 = np.sort(ref["logits"], axis=1)[:, -2:]
        sure = (top2[:, 1] - top2[:, 0]) > 2e-3
        assert (logits.cpu().numpy().argmax(1)[sure] == ref["logits"].argmax(1)[sure]).all()
    e.close()


@pytest.mark.parametrize("name", ["tiny_5w1s", "tiny_5w5s_merge", "small_5w1s"])
def test_episode_matches_16bit_emulating_oracle_tightly(lib, name):
    from oracle import fsar_oracle as O
    meta, _ = load_golden(name)
    g, sd, tt, te, task = regenerate(meta)
    e = make_engine(lib, meta, g, sd, tt, te)
    logits, class_logits = run(e, meta, task)
    out = O.episode_forward(sd, g, tt, te, task, meta["T"], meta["merge_before"], meta["single_direct"],
                            operand_dtype=e.operand_dtype)
    S, T, E = meta["way"] * meta["shot"], meta["T"], g["embed_dim"]
    assert rel_l2(e.peek("support_feats", (S, T, E)), out["support_feats"]) < 1.5e-3      # tanh.approx vs torch.tanh
    assert rel_max(logits, out["logits"]) < 1e-3
    assert torch.equal(e.peek("class_index", (S,), torch.int32).long(), out["class_index"])
    e.close()


def test_vitl14_full_depth_16_frame_episode(lib):
    """BASELINE.json configs[3]: ViT-L/14 (24 layers, 257 tokens, width 1024), 5-way 1-shot, 16 frames = 160 frames, the
    WHOLE episode (modulator 8 x 96, OTAM 16 x 16) against (a) the fixture the reference's own classes produced and
    (b) the fp16-operand-emulating oracle's outputs stored in the same fixture (twice the depth of ViT-B/16: rounding
    accumulates over 24 blocks)."""
    meta, ref = load_golden("vitl14_5w1s_T16_default_init")
    g, sd, tt, te, task = regenerate(meta)
    assert g["layers"] == 24 and g["width"] == 1024
    e = make_engine(lib, meta, g, sd, tt, te, max_frames=80)      # two 80-frame passes
    logits, class_logits = run(e, meta, task)
    S, Q, T, E, way = 5, 5, 16, g["embed_dim"], 5
    sf, tf = e.peek("support_feats", (S, T, E)), e.peek("target_feats", (Q, T, E))
    assert rel_l2(sf, ref["support_feats"]) < 5e-3 and rel_l2(tf, ref["target_feats"]) < 5e-3
    assert rel_l2(sf, ref["emu16_support_feats"]) < 2e-3 and rel_l2(tf, ref["emu16_target_feats"]) < 2e-3
    assert float(np.abs(e.peek("dists", (Q, way, T, T)).numpy() - ref["dists"]).max()) < 3e-3
    assert rel_max(logits, ref["logits"]) < 1e-3 and rel_centred(logits, ref["logits"]) < 1e-3
    assert rel_max(logits, ref["emu16_logits"]) < 1e-3
    assert rel_max(class_logits, ref["class_logits"]) < 3e-3
    top2 = np.sort(ref["logits"], axis=1)[:, -2:]
    sure = (top2[:, 1] - top2[:, 0]) > 2e-3 * np.abs(ref["logits"]).max()
    assert (logits.cpu().numpy().argmax(1)[sure] == ref["logits"].argmax(1)[sure]).all()
    e.close()


@pytest.mark.parametrize("geom", ["l14-2layer", "b32-2layer"])
def test_other_clip_geometries_against_oracle(lib, geom):
    """BASELINE.json configs[3] names ViT-L/14 (257 tokens, width 1024, 14x14 patches -> K = 588 padded to 640; the
    tcgen05 attention instance with the scalar 257th token); ViT-B/32 has 50 tokens and a 3072-deep patch GEMM. The
    reference head has no branch for either, so the oracle restatement is the checker."""
    from clip_fsar_b200 import synth
    from oracle import fsar_oracle as O
    g = synth.full_geometry(geom)
    sd = synth.synth_state_dict(g, 3)
    e = lib.Engine(**dict(g, max_frames=6, max_videos=4, max_tokens=4, max_classes=8, otam_lambda=0.5, device=0))
    e.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    frames = torch.from_numpy(synth.synth_episode(2, 1, 1, 4, 224, 8, 11)["support_set"])      # 8 frames, 2 passes
    ref32 = O.vit_forward(sd, g, frames)
    ref16 = O.vit_forward(sd, g, frames, operand_dtype=e.operand_dtype)
    out = e.vit_forward(frames.to(DEV))
    assert rel_l2(out, ref32) < 5e-3 and rel_l2(out, ref16) < 1e-3
    x = torch.randn(3, 5, g["embed_dim"])
    assert rel_l2(e.modulate(x.to(DEV)), O.modulator(sd, g, x)) < 5e-6
    e.close()


def test_host_buffer_entry_point_equals_device_entry_point(lib):
    meta, _ = load_golden("tiny_5w5s_nomerge")
    g, sd, tt, te, task = regenerate(meta)
    e = make_engine(lib, meta, g, sd, tt, te)
    a, ca = run(e, meta, task)
    b, cb = run(e, meta, task, host=True)
    assert torch.equal(a.cpu(), b) and torch.equal(ca.cpu(), cb)               # same kernels, same order: bit equal
    e.close()


def test_pipelined_submit_collect(lib):
    meta, _ = load_golden("tiny_5w1s")
    g, sd, tt, te, task = regenerate(meta)
    e = make_engine(lib, meta, g, sd, tt, te)
    ref, _ = run(e, meta, task)
    t = {k: torch.from_numpy(v).pin_memory() for k, v in task.items()}
    args = (t["support_set"], t["target_set"], t["support_labels"], t["real_support_labels"], meta["T"], meta["way"])
    outs = [torch.empty(5, 5) for _ in range(6)]
    e.episode_submit_host(0, *args)
    for i in range(1, 6):
        e.episode_submit_host(i & 1, *args)
        e.episode_collect_host((i - 1) & 1, outs[i - 1])
    e.episode_collect_host(1, outs[5])
    for o in outs:
        assert torch.equal(o, ref.cpu())
    with pytest.raises(lib.FsarError):
        e.episode_collect_host(0, outs[0])                                     # nothing submitted
    e.close()


def test_batched_episodes_equal_single_episode_calls(lib):
    """fsar_episodes_forward regroups the frames of several episodes into ViT passes that ignore episode boundaries;
    per episode the result must be what fsar_episode_forward gives (same kernels on the same rows: bit equal for the
    head, and the ViT rows only move between GEMM tiles)."""
    from clip_fsar_b200 import synth
    meta, _ = load_golden("tiny_5w1s")
    g, sd, tt, te, _ = regenerate(meta)
    n_vid = meta["way"] * (meta["shot"] + 1)
    e = lib.Engine(**dict(g, max_frames=96, max_videos=n_vid, max_tokens=meta["T"], max_classes=128, max_batch=6,
                          otam_lambda=0.5, device=0))
    e.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    e.set_weight("text_features_train", torch.from_numpy(tt))
    e.set_weight("text_features_test", torch.from_numpy(te))
    keys = ("support_set", "target_set", "support_labels", "real_support_labels")
    eps_np = [synth.synth_episode(5, 1, 1, 8, g["image_size"], 24, 2000 + i) for i in range(6)]
    eps = [[torch.from_numpy(ep[k]).to(DEV) for k in keys] for ep in eps_np]
    singles = [e.episode_forward(*ep, 8, 5, n_train_classes=64) for ep in eps]
    logits, cl = e.episodes_forward(eps, 8, 5, n_train_classes=64)
    for i in range(6):
        assert rel_max(logits[i], singles[i][0]) < 2e-4 and rel_max(cl[i], singles[i][1]) < 2e-4
    # host entry points, two slots
    pinned = [[torch.from_numpy(ep[k]).pin_memory() for k in keys] for ep in eps_np]
    e.episodes_submit_host(0, pinned[:3], 8, 5)
    e.episodes_submit_host(1, pinned[3:], 8, 5)
    a, b = torch.empty(3, 5, 5), torch.empty(3, 5, 5)
    e.episodes_collect_host(0, a)
    e.episodes_collect_host(1, b)
    assert rel_max(torch.cat([a, b]), logits.cpu()) < 2e-4
    with pytest.raises(lib.FsarError):
        e.episodes_forward(eps + eps, 8, 5, n_train_classes=64)             # 12 > max_batch
    e.close()


def test_device_metrics_kernel_equals_caller_math(lib):
    """fsar_metrics_update vs the reference caller's math (F.cross_entropy, topks_correct; test_net_few_shot.py:111, 147)."""
    from clip_fsar_b200 import runner, synth
    g = synth.full_geometry("tiny")
    e = lib.Engine(**dict(g, max_frames=8, max_videos=10, max_tokens=8, max_classes=8, otam_lambda=0.5, device=0))
    gen = torch.Generator().manual_seed(5)
    dev_c, ref_c = runner.new_counters(DEV), runner.new_counters("cpu")
    per_class = torch.zeros(2 * 7, dtype=torch.int64, device=DEV)
    for _ in range(9):
        logits = torch.randn(13, 7, generator=gen) * 3
        logits[3, 2] = logits[3, 5] = logits[3].max() + 1.0              # tie: the first maximum must win
        tgt = torch.randint(0, 7, (13,), generator=gen).float()
        runner.update_counters_device(e, dev_c, logits.to(DEV), tgt.to(DEV), per_class)
        runner.update_counters(ref_c, logits, tgt)
    d, r = dev_c.cpu(), ref_c
    assert d[0] == r[0] and d[1] == r[1] == 9 * 13
    assert abs(int(d[2]) - int(r[2])) <= 9 * 13                         # each query's CE is rounded to 1e-6 separately
    assert int(per_class[7:].sum()) == 9 * 13 and int(per_class[:7].sum()) == int(d[0])
    e.close()


def test_sharded_evaluate_on_gpu_matches_oracle_accuracy(lib):
    from clip_fsar_b200 import runner
    from oracle import fsar_oracle as O
    meta, _ = load_golden("tiny_5w1s")
    g, sd, tt, te, _ = regenerate(meta)
    e = make_engine(lib, meta, g, sd, tt, te)

    def gpu_forward(task):
        return e.episode_forward(task["support_set"], task["target_set"], task["support_labels"], task["real_support_labels"],
                                 8, 5, n_train_classes=64)[0]

    def cpu_forward(task):
        return O.episode_forward(sd, g, tt, te, {k: v.numpy() for k, v in task.items()}, 8)["logits"]

    kw = dict(n_frames=8, image_size=g["image_size"])
    got = runner.evaluate(gpu_forward, 6, device=DEV, engine=e, **kw)
    want = runner.evaluate(cpu_forward, 6, device="cpu", **kw)
    assert got["n_total"] == want["n_total"] == 30 and got["n_correct"] == want["n_correct"]
    assert abs(got["loss"] - want["loss"]) < 5e-3
    e.close()


def test_vit_chunking_is_invisible(lib):
    meta, _ = load_golden("tiny_5w1s")
    g, sd, tt, te, task = regenerate(meta)
    frames = torch.from_numpy(task["support_set"]).to(DEV)
    big = make_engine(lib, meta, g, sd, tt, te, max_frames=64)
    small = make_engine(lib, meta, g, sd, tt, te, max_frames=7)               # 40 frames -> 6 passes
    assert torch.equal(big.vit_forward(frames), small.vit_forward(frames))
    big.close()
    small.close()


def test_capacity_and_argument_errors(lib):
    meta, _ = load_golden("tiny_5w1s")
    g, sd, tt, te, task = regenerate(meta)
    e = make_engine(lib, meta, g, sd, tt, te)
    t = {k: torch.from_numpy(v).to(DEV) for k, v in task.items()}
    with pytest.raises(lib.FsarError) as err:                                   # 16 frames per video > max_tokens 8
        e.episode_forward(t["support_set"][:32], t["target_set"][:32], t["support_labels"][:2],
                          t["real_support_labels"][:2], 16, 2, n_train_classes=64)
    assert err.value.code == -5
    with pytest.raises(ValueError):                                             # label count mismatch
        e.episode_forward(t["support_set"], t["target_set"], t["support_labels"][:3], t["real_support_labels"], 8, 5,
                          n_train_classes=64)
    with pytest.raises(ValueError):                                             # host tensor on the device entry point
        e.episode_forward(t["support_set"].cpu(), t["target_set"], t["support_labels"], t["real_support_labels"], 8, 5,
                          n_train_classes=64)
    e.close()


def test_out_of_range_real_labels_are_reported_not_silently_wrong(lib):
    """The reference raises IndexError at text_features_test[support_real_class.long()] (few_shot.py:2946). Here the labels
    live on the device and no entry point synchronises: the kernels clamp the index (nothing reads out of bounds) and raise
    a flag in mapped host memory; the host path reports it at collect (FSAR_E_INVALID), the stream-ordered path at the
    next call on the handle. A good episode afterwards works again."""
    meta, _ = load_golden("tiny_5w1s")
    g, sd, tt, te, task = regenerate(meta)
    e = make_engine(lib, meta, g, sd, tt, te)
    good = {k: torch.from_numpy(v) for k, v in task.items()}
    for bad_value in (float(meta["n_test"]), -1.0, float("nan"), 1.0e9):
        bad = dict(good)
        bad["real_support_labels"] = good["real_support_labels"].clone()
        bad["real_support_labels"][2] = bad_value
        args = lambda t, dev: [t[k].to(dev) if dev else t[k].pin_memory() for k in
                               ("support_set", "target_set", "support_labels", "real_support_labels")]
        # host path: reported by the collect of that very batch
        with pytest.raises(lib.FsarError) as err:
            e.episode_forward_host(*args(bad, None), meta["T"], meta["way"], n_train_classes=meta["n_train"])
        assert err.value.code == -1 and "real_support_labels" in str(err.value)
        # device path: the call itself is stream-ordered and returns; the NEXT call reports
        logits, _ = e.episode_forward(*args(bad, DEV), meta["T"], meta["way"], n_train_classes=meta["n_train"])
        torch.cuda.synchronize()
        assert torch.isfinite(logits).all()                        # clamped, not out of bounds
        with pytest.raises(lib.FsarError) as err:
            e.episode_forward(*args(good, DEV), meta["T"], meta["way"], n_train_classes=meta["n_train"])
        assert err.value.code == -1
        ok, _ = e.episode_forward(*args(good, DEV), meta["T"], meta["way"], n_train_classes=meta["n_train"])
        torch.cuda.synchronize()
    ref, _ = run(e, meta, task)
    assert torch.equal(ok, ref)
    # fewer classes announced than the labels hold (way = 4 for a 5-class episode): flagged the same way
    with pytest.raises(lib.FsarError) as err:
        e.episode_forward_host(*args(good, None), meta["T"], 4, n_train_classes=meta["n_train"])
    assert err.value.code == -1 and "way" in str(err.value)
    e.close()


def test_wrong_frame_geometry_is_rejected(lib):
    """A task dict built with another DATA.TEST_CROP_SIZE must not be read with the engine's stride (silent garbage): the
    binding raises before anything is enqueued; the reference fails with a positional_embedding shape mismatch."""
    meta, _ = load_golden("tiny_5w1s")
    g, sd, tt, te, task = regenerate(meta)
    e = make_engine(lib, meta, g, sd, tt, te)
    t = {k: torch.from_numpy(v).to(DEV) for k, v in task.items()}
    S = g["image_size"]
    wrong = torch.zeros(40, 3, S + 16, S + 16, device=DEV)
    with pytest.raises(ValueError, match="image_size"):
        e.episode_forward(wrong, t["target_set"], t["support_labels"], t["real_support_labels"], 8, 5, n_train_classes=64)
    with pytest.raises(ValueError):
        e.vit_forward(wrong)
    with pytest.raises(ValueError):                                             # labels must be fp32 (ssv2_few_shot.py:278-283)
        e.episode_forward(t["support_set"], t["target_set"], t["support_labels"].long(), t["real_support_labels"], 8, 5,
                          n_train_classes=64)
    e.close()


# ----------------------------------------------------------------------------------------------------------------
# BASELINE.json's headline size (ViT-B/16, 5-way 1-shot, 8 x 224^2): properties that need no CPU reference
@pytest.fixture(scope="module")
def full(lib):
    meta, ref = load_golden("vitb16_5w1s")
    g, sd, tt, te, task = regenerate(meta)
    e = make_engine(lib, meta, g, sd, tt, te)
    yield e, meta, task, ref
    e.close()


def test_full_size_is_deterministic_and_finite(full):
    e, meta, task, _ = full
    a, ca = run(e, meta, task)
    b, cb = run(e, meta, task)
    assert torch.equal(a, b) and torch.equal(ca, cb)
    assert torch.isfinite(a).all() and torch.isfinite(ca).all()


def test_full_size_support_order_invariance(full):
    """Prototypes are per sorted class (torch.unique, few_shot.py:2965): shuffling the support VIDEOS (with their
    labels) must not change the logits; shuffling the QUERIES permutes the rows."""
    e, meta, task, _ = full
    base, _ = run(e, meta, task)
    T = meta["T"]
    perm = np.array([3, 0, 4, 1, 2])
    t2 = dict(task)
    t2["support_set"] = task["support_set"].reshape(5, T, 3, 224, 224)[perm].reshape(5 * T, 3, 224, 224).copy()
    t2["support_labels"] = task["support_labels"][perm].copy()
    t2["real_support_labels"] = task["real_support_labels"][perm].copy()
    shuffled, _ = run(e, meta, t2)
    assert rel_max(shuffled, base) < 2e-4         # frames land in different GEMM tiles / chunk positions only
    t3 = dict(task)
    t3["target_set"] = task["target_set"].reshape(5, T, 3, 224, 224)[perm].reshape(5 * T, 3, 224, 224).copy()
    qperm, _ = run(e, meta, t3)
    assert rel_max(qperm, base[torch.from_numpy(perm).to(base.device)]) < 2e-4


def test_full_size_identical_query_and_support_is_the_nearest(full):
    """A query that IS a support video has zero frame distance to its own prototype diagonal: its own class must win."""
    e, meta, task, _ = full
    t2 = dict(task)
    t2["target_set"] = task["support_set"].copy()
    logits, _ = run(e, meta, t2)
    cls = torch.from_numpy(np.argsort(np.argsort(task["support_labels"]))).to(logits.device)
    # note: the support prototype also attends to the text token, so the distance is small, not exactly zero
    assert torch.equal(logits.argmax(1), cls)


def test_full_size_matches_fixture_and_counts_launches(full):
    e, meta, task, ref = full
    n0 = e.launch_count()
    logits, _ = run(e, meta, task)
    assert e.launch_count() - n0 > 50                                           # our kernels ran, not a fallback
    assert rel_max(logits, ref["logits"]) < 3e-3


@pytest.mark.parametrize("name", ["tiny_5w1s", "small_5w1s"])
def test_cls_only_last_block_equals_full_last_block(lib, name, monkeypatch):
    """The last transformer block evaluates Q / out_proj / ln_2 / MLP for the CLS row only (the only row
    VisionTransformer.forward keeps, few_shot.py:683). FSAR_FULL_LAST_BLOCK=1 computes every token row as the reference
    does: the frame features must agree to operand-rounding noise (the CLS attention row is SIMT fp32 instead of
    tcgen05, everything else is the same kernels on the same rows)."""
    meta, ref = load_golden(name)
    g, sd, tt, te, task = regenerate(meta)
    frames = torch.from_numpy(task["support_set"]).to(DEV)
    feats = {}
    for full_block in ("0", "1"):
        monkeypatch.setenv("FSAR_FULL_LAST_BLOCK", full_block)
        e = make_engine(lib, meta, g, sd, tt, te)
        n0 = e.launch_count()
        feats[full_block] = e.vit_forward(frames).cpu()
        launches = e.launch_count() - n0
        e.close()
        # patch gather + patch GEMM + ln_pre (which also emits block 0's ln_1) + 7 per block - that ln_1 + final projection;
        # the CLS-only block has 8 launches
        assert launches == 3 + 7 * g["layers"] + (1 if full_block == "0" else 0)
    assert rel_max(feats["0"], feats["1"]) < 5e-4
    assert rel_l2(feats["0"], ref["support_feats"].reshape(-1, g["embed_dim"])) < 5e-3
