"""GPU bring-up self test: runs each operator of libfsar_sm100 against torch / the CPU oracle and prints one JSON
line per check (never aborts on the first failure). Usage on the GPU box:
    for g in basic gemm attention vit head episode full; do timeout 300 python tools/selftest_gpu.py $g; done
Development tool (not part of the product path, not a pytest).
"""
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clip_fsar_b200 import lib as L  # noqa: E402
from clip_fsar_b200 import synth  # noqa: E402
from oracle import fsar_oracle as O  # noqa: E402

DEV = "cuda:0"


def report(check_name, **kw):
    print(json.dumps(dict(check=check_name, **kw)), flush=True)


def err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    d = (a - b).abs()
    return dict(maxabs=float(d.max()), ref_absmax=float(b.abs().max()),
                rel_l2=float((a - b).norm() / (b.norm() + 1e-30)), nan=bool(torch.isnan(a).any()))


def make_engine(geom, T=8, max_videos=10, max_frames=None, mod_depth=1):
    g = synth.full_geometry(geom, mod_depth)
    cfg = dict(g)
    cfg.update(max_frames=max_frames or max_videos * T, max_videos=max_videos, max_tokens=T, max_classes=128,
               otam_lambda=0.5, device=0)
    return L.Engine(**cfg), g


def guarded(fn):
    def run(*a, **k):
        try:
            fn(*a, **k)
        except Exception as e:  # noqa: BLE001
            report(fn.__name__, error=repr(e), tb=traceback.format_exc()[-800:])
    return run


@guarded
def basic():
    eng, g = make_engine("tiny")
    x = torch.randn(1000, 256, device=DEV)
    y = eng.op_f32_to_16(x)
    report("f32_to_16", **err(y, x.to(eng.operand_dtype)))
    for D in (128, 512, 768, 1024):
        x = torch.randn(777, D, device=DEV) * 3 + 1
        gm, bt = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
        ref = torch.nn.functional.layer_norm(x, (D,), gm, bt, 1e-5)
        report("layernorm32_D%d" % D, **err(eng.op_layernorm(x, gm, bt, False), ref))
        report("layernorm16_D%d" % D, **err(eng.op_layernorm(x, gm, bt, True), ref))
    report("launch_count", n=eng.launch_count())


@guarded
def gemm():
    eng, g = make_engine("tiny")
    dt = eng.operand_dtype
    shapes = [(128, 256, 64), (128, 256, 128), (256, 512, 256), (300, 256, 192), (1000, 768, 768), (1000, 2304, 768),
              (777, 768, 3072), (333, 128, 64), (100, 64, 64), (1970, 3072, 768), (15760, 768, 768)]
    for (M, N, K) in shapes:
        a = (torch.randn(M, K, device=DEV) * 0.5).to(dt)
        w = (torch.randn(N, K, device=DEV) * 0.5).to(dt)
        bias = torch.randn(N, device=DEV)
        ref = a.float() @ w.float().T + bias
        for epi, nm in ((L.EPI_STORE32, "store32"), (L.EPI_STORE16, "store16"), (L.EPI_QGELU16, "qgelu16"),
                        (L.EPI_RESID32, "resid32")):
            try:
                if epi == L.EPI_RESID32:
                    x0 = torch.randn(M, N, device=DEV)
                    out = eng.op_gemm(a, w, bias, epi, out=x0.clone())
                    r = ref + x0
                elif epi == L.EPI_QGELU16:
                    out = eng.op_gemm(a, w, bias, epi)
                    r = ref * torch.sigmoid(1.702 * ref)
                else:
                    out = eng.op_gemm(a, w, bias, epi)
                    r = ref
                torch.cuda.synchronize()
                report("gemm_%s_%dx%dx%d" % (nm, M, N, K), **err(out, r))
            except Exception as e:  # noqa: BLE001
                report("gemm_%s_%dx%dx%d" % (nm, M, N, K), error=repr(e))
    # timing of the big shapes
    for (M, N, K) in ((15760, 2304, 768), (15760, 768, 768), (15760, 3072, 768), (15760, 768, 3072)):
        a = (torch.randn(M, K, device=DEV) * 0.5).to(dt)
        w = (torch.randn(N, K, device=DEV) * 0.5).to(dt)
        out = torch.zeros(M, N, device=DEV, dtype=dt)
        for _ in range(3):
            eng.op_gemm(a, w, None, L.EPI_STORE16, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            eng.op_gemm(a, w, None, L.EPI_STORE16, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        report("gemm_time_%dx%dx%d" % (M, N, K), ms=ms, tflops=2.0 * M * N * K / ms / 1e9)


@guarded
def gemm_peak():
    """Steady-state rate of the GEMM main loop: shapes with whole waves and a long K so that ramp/tail/epilogue vanish."""
    eng, g = make_engine("tiny")
    dt = eng.operand_dtype
    for (M, N, K) in ((18944, 256, 16384), (18944, 512, 8192), (18944, 2304, 768), (18944, 768, 768), (18944, 3072, 768),
                      (18944, 768, 3072), (15760, 2304, 768), (15760, 768, 3072)):
        a = (torch.randn(M, K, device=DEV) * 0.1).to(dt)
        w = (torch.randn(N, K, device=DEV) * 0.1).to(dt)
        out = torch.zeros(M, N, device=DEV, dtype=dt)
        for _ in range(3):
            eng.op_gemm(a, w, None, L.EPI_STORE16, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        n = 20
        e0.record()
        for _ in range(n):
            eng.op_gemm(a, w, None, L.EPI_STORE16, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        report("gemm_peak_%dx%dx%d" % (M, N, K), ms=ms, tflops=2.0 * M * N * K / ms / 1e9,
               single=os.environ.get("FSAR_GEMM_SINGLE", "0"))


@guarded
def attention():
    eng, g = make_engine("tiny")
    dt = eng.operand_dtype
    for (n, Lt, H) in ((2, 5, 2), (3, 197, 2), (2, 197, 12), (2, 257, 4), (1, 64, 1), (1, 65, 1)):
        D = H * 64
        qkv = (torch.randn(n * Lt, 3 * D, device=DEV)).to(dt)
        out = eng.op_attention(qkv, n, Lt, H)
        q, k, v = qkv.float().reshape(n, Lt, 3, H, 64).permute(2, 0, 3, 1, 4)
        att = torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1)
        ref = (att @ v).transpose(1, 2).reshape(n * Lt, D)
        report("attention_n%d_L%d_H%d" % (n, Lt, H), **err(out, ref))
    n, Lt, H = 80, 197, 12
    qkv = torch.randn(n * Lt, 3 * H * 64, device=DEV).to(dt)
    for _ in range(3):
        eng.op_attention(qkv, n, Lt, H)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        eng.op_attention(qkv, n, Lt, H)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    report("attention_time_80x197x12", ms=ms, tflops=4.0 * n * H * Lt * Lt * 64 / ms / 1e9)


def load_weights(eng, g, seed=0, spread=True, n_train=64, n_test=24):
    sd = synth.synth_state_dict(g, seed, spread)
    eng.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    tt = synth.synth_text_features(n_train, g["embed_dim"], 7)
    te = synth.synth_text_features(n_test, g["embed_dim"], 8)
    eng.set_weight("text_features_train", torch.from_numpy(tt))
    eng.set_weight("text_features_test", torch.from_numpy(te))
    return sd, tt, te


@guarded
def vit():
    for geom in ("tiny", "small"):
        eng, g = make_engine(geom, max_frames=7)  # 7 < 16 frames: exercises chunking
        sd, _, _ = load_weights(eng, g)
        report("missing_%s" % geom, missing=eng.missing_weights())
        frames = torch.from_numpy(synth.synth_episode(2, 1, 1, 8, g["image_size"], 24, 5)["support_set"])
        ref = O.vit_forward(sd, g, frames)
        out = eng.vit_forward(frames.to(DEV))
        report("vit_%s" % geom, **err(out, ref))


@guarded
def head():
    eng, g = make_engine("tiny", T=32, max_videos=60)
    sd, tt, te = load_weights(eng, g)
    E = g["embed_dim"]
    for (n, t) in ((5, 8), (5, 9), (10, 17), (3, 33)):
        x = torch.randn(n, t, E)
        report("modulate_%dx%d" % (n, t), **err(eng.modulate(x.to(DEV)), O.modulator(sd, g, x)))
    for (Q, way, T) in ((5, 5, 8), (10, 10, 16), (20, 20, 32), (1, 3, 1), (2, 2, 2)):
        for single in (False, True):
            q, p = torch.randn(Q, T, E), torch.randn(way, T, E)
            sim = O.cos_sim(q.reshape(Q * T, E), p.reshape(way * T, E))
            d = (1 - sim).reshape(Q, T, way, T).permute(0, 2, 1, 3)
            cum = O.otam_cum_dist(d) if single else O.otam_cum_dist(d) + O.otam_cum_dist(d.transpose(2, 3))
            lg, dd, cc = eng.otam_logits(q.to(DEV), p.to(DEV), single, True)
            report("otam_Q%d_w%d_T%d_s%d" % (Q, way, T, single), dists=err(dd, d), logits=err(lg, -cum))


def run_golden(name, eng_cache={}):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    m = json.loads(str(z["meta"]))
    eng, g = make_engine(m["geom"], T=m["T"], max_videos=m["way"] * (m["shot"] + 1), mod_depth=m["mod_depth"])
    load_weights(eng, g, m["wseed"], m["spread"], m["n_train"], m["n_test"])
    task = synth.synth_episode(m["way"], m["shot"], 1, m["T"], g["image_size"], m["n_test"], m["eseed"], m["structured"])
    dev = {k: torch.from_numpy(v).to(DEV) for k, v in task.items()}
    logits, cl = eng.episode_forward(dev["support_set"], dev["target_set"], dev["support_labels"],
                                     dev["real_support_labels"], m["T"], m["way"], m["merge_before"], m["single_direct"],
                                     n_train_classes=m["n_train"])
    S, Q, T, E = m["way"] * m["shot"], m["way"], m["T"], g["embed_dim"]
    res = dict(logits=err(logits, torch.from_numpy(z["logits"])), class_logits=err(cl, torch.from_numpy(z["class_logits"])),
               support_feats=err(eng.peek("support_feats", (S, T, E)), torch.from_numpy(z["support_feats"])),
               target_feats=err(eng.peek("target_feats", (Q, T, E)), torch.from_numpy(z["target_feats"])),
               dists=err(eng.peek("dists", (Q, m["way"], T, T)), torch.from_numpy(z["dists"])),
               argmax_agree=float((logits.cpu().argmax(1) == torch.from_numpy(z["logits"]).argmax(1)).float().mean()))
    report("golden_" + name, **res)
    return eng, g, m, task, dev


@guarded
def episode():
    for name in ("tiny_5w1s", "tiny_5w5s_merge", "tiny_5w5s_nomerge", "tiny_10w1s_T16", "tiny_3w2s_T32_single",
                 "tiny_5w1s_depth2", "tiny_5w1s_default_init", "small_5w1s"):
        try:
            run_golden(name)
        except Exception as e:  # noqa: BLE001
            report("golden_" + name, error=repr(e), tb=traceback.format_exc()[-600:])


@guarded
def full():
    eng, g, m, task, dev = run_golden("vitb16_5w1s")
    # host path
    pin = {k: torch.from_numpy(v).pin_memory() for k, v in task.items()}
    lg, cl = eng.episode_forward_host(pin["support_set"], pin["target_set"], pin["support_labels"],
                                      pin["real_support_labels"], m["T"], m["way"], n_train_classes=m["n_train"])
    z = np.load(os.path.join(ROOT, "tests", "golden", "vitb16_5w1s.npz"))
    report("full_host_path", **err(lg, torch.from_numpy(z["logits"])))
    # timing, device-resident
    args = (dev["support_set"], dev["target_set"], dev["support_labels"], dev["real_support_labels"], m["T"], m["way"])
    for _ in range(3):
        eng.episode_forward(*args, n_train_classes=m["n_train"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    n = 20
    e0.record()
    for _ in range(n):
        eng.episode_forward(*args, n_train_classes=m["n_train"])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    report("full_episode_time", ms=ms, eps=1000.0 / ms, tflops=80 * synth.vit_flops_per_frame(g) / ms / 1e9)
    eng.profile_begin()
    eng.episode_forward(*args, n_train_classes=m["n_train"])
    prof = eng.profile_end()
    for k, v in prof.items():
        v["tflops"] = v["flops"] / v["ms"] / 1e9 if v["ms"] else 0
        v["gbs"] = v["bytes"] / v["ms"] / 1e6 if v["ms"] else 0
    report("full_profile", **prof)
    # pipelined host path
    t0 = time.time()
    n = 20
    eng.episode_submit_host(0, pin["support_set"], pin["target_set"], pin["support_labels"], pin["real_support_labels"], m["T"], m["way"])
    out = torch.empty(5, 5)
    for i in range(1, n):
        eng.episode_submit_host(i & 1, pin["support_set"], pin["target_set"], pin["support_labels"], pin["real_support_labels"], m["T"], m["way"])
        eng.episode_collect_host((i - 1) & 1, out)
    eng.episode_collect_host((n - 1) & 1, out)
    dt = time.time() - t0
    report("full_host_pipelined", ms=dt / n * 1e3, eps=n / dt, **err(out, torch.from_numpy(z["logits"])))


if __name__ == "__main__":
    groups = dict(basic=basic, gemm=gemm, gemm_peak=gemm_peak, attention=attention, vit=vit, head=head, episode=episode, full=full)
    for name in sys.argv[1:] or list(groups):
        report("group", name=name, gpu=torch.cuda.get_device_name(0))
        groups[name]()
