"""GPU: every operator of libfsar_sm100.so, called through the C ABI, against a plain torch fp32 computation
of the same op on the same (16-bit rounded) operands. Tolerances are written next to each assert."""
import pytest
import torch

from conftest import regenerate, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def eng(lib):
    from clip_fsar_b200 import synth
    g = synth.full_geometry("tiny")
    e = lib.Engine(**dict(g, max_frames=16, max_videos=60, max_tokens=32, max_classes=128, otam_lambda=0.5, device=0))
    yield e
    e.close()


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_f32_to_16_is_round_to_nearest(eng):
    for n in (1, 3, 4, 1000 * 256 + 1):
        x = torch.randn(n, device=DEV) * 50
        assert torch.equal(eng.op_f32_to_16(x), x.to(eng.operand_dtype))       # bit exact


@pytest.mark.parametrize("D", [128, 512, 768, 1024])
def test_layernorm(eng, D):
    x = torch.randn(777, D, device=DEV) * 3 + 1
    g, b = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
    ref = torch.nn.functional.layer_norm(x, (D,), g, b, 1e-5)
    assert rel_l2(eng.op_layernorm(x, g, b, False), ref) < 1e-6                # fp32 in / fp32 out
    assert rel_l2(eng.op_layernorm(x, g, b, True), ref) < 4e-4                 # 16-bit output rounding only


@pytest.mark.parametrize("rows,D", [(2048, 768), (5003, 768), (18912, 768), (4111, 1024), (3000, 512), (2500, 128)])
def test_layernorm_full_size(eng, rows, D):
    """Full-size launches (thousands of rows, every supported width, a row with a large dynamic range)."""
    x = torch.randn(rows, D, device=DEV) * 3 + 1
    x[rows // 2] *= 50.0                                                       # one row with a large dynamic range
    g, b = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
    ref = torch.nn.functional.layer_norm(x, (D,), g, b, 1e-5)
    out = eng.op_layernorm(x, g, b, True)
    assert rel_l2(out, ref) < 4e-4
    assert float((out.float() - ref).abs().max()) < 2e-2 * float(ref.abs().max())   # no row may be off
    assert rel_l2(eng.op_layernorm(x, g, b, False), ref) < 1e-6


def test_layernorm_rejects_unsupported_dim(eng, lib):
    x = torch.randn(4, 100, device=DEV)
    with pytest.raises(lib.FsarError):
        eng.op_layernorm(x, torch.ones(100, device=DEV), torch.zeros(100, device=DEV))


GEMM_SHAPES = [(128, 256, 64), (1, 64, 64), (127, 72, 136), (300, 256, 192), (1000, 768, 768), (1000, 2304, 768),
               (777, 768, 3072), (333, 128, 64), (1970, 3072, 768), (15760, 768, 768),
               # exactly two tile columns, ragged last column blocks (1280 = 5 x 256; 1096 is not a multiple of 64), and the
               # full 96-frame QKV shape
               (600, 512, 128), (300, 1280, 192), (515, 1096, 72), (18912, 2304, 768)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_tcgen05_gemm_all_epilogues(eng, lib, M, N, K):
    dt = eng.operand_dtype
    a = (torch.randn(M, K, device=DEV) * 0.5).to(dt)
    w = (torch.randn(N, K, device=DEV) * 0.5).to(dt)
    bias = torch.randn(N, device=DEV)
    ref = a.float() @ w.float().T + bias                                       # fp32 accumulate of the same operands
    assert rel_l2(eng.op_gemm(a, w, bias, lib.EPI_STORE32), ref) < 1e-5        # accumulation-order noise only
    assert rel_l2(eng.op_gemm(a, w, None, lib.EPI_STORE32), ref - bias) < 1e-5
    assert rel_l2(eng.op_gemm(a, w, bias, lib.EPI_STORE16), ref) < 4e-4        # + one 16-bit rounding
    # QuickGELU is evaluated as h + h tanh(0.851 x) in packed 16-bit arithmetic (one MUFU per two elements):
    # three 16-bit roundings + tanh.approx instead of one rounding
    assert rel_l2(eng.op_gemm(a, w, bias, lib.EPI_QGELU16), ref * torch.sigmoid(1.702 * ref)) < 1.5e-3
    x0 = torch.randn(M, N, device=DEV)
    assert rel_l2(eng.op_gemm(a, w, bias, lib.EPI_RESID32, out=x0.clone()), ref + x0) < 1e-5


def test_gemm_rejects_bad_shapes(eng, lib):
    dt = eng.operand_dtype
    with pytest.raises(lib.FsarError):
        eng.op_gemm(torch.zeros(8, 60, device=DEV, dtype=dt), torch.zeros(64, 60, device=DEV, dtype=dt))   # K % 8
    with pytest.raises(lib.FsarError):
        eng.op_gemm(torch.zeros(8, 64, device=DEV, dtype=dt), torch.zeros(12, 64, device=DEV, dtype=dt))   # N % 8


# 197 = ViT-B/16, 257 = ViT-L/14 (256 tokens on the tensor cores + one scalar token), 208 / 209 / 256 = the instance
# boundaries (MAXK = 208 with TMA-staged output, MAXK = 256 with direct output); 258+ tokens are rejected (the round-1
# mma.sync kernel that served them only exists in the -DFSAR_PROBES build)
@pytest.mark.parametrize("n,L,H", [(2, 5, 2), (3, 197, 2), (2, 197, 12), (2, 257, 4), (5, 257, 16), (1, 64, 1), (1, 65, 1), (1, 1, 1),
                                   (2, 208, 2), (2, 209, 2), (3, 256, 3), (2, 129, 1)])
def test_attention_core(eng, n, L, H):
    D = H * 64
    qkv = torch.randn(n * L, 3 * D, device=DEV).to(eng.operand_dtype)
    q, k, v = qkv.float().reshape(n, L, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1) @ v).transpose(1, 2).reshape(n * L, D)
    out = eng.op_attention(qkv, n, L, H)
    assert rel_l2(out, ref) < 6e-4                                             # 16-bit P and output rounding
    # no row may be off (the scalar token of L = 257 is one row in 257: a global norm would hide it)
    assert float((out.float().cpu() - ref.cpu()).abs().max()) < 2e-2 * float(ref.abs().max())


def test_attention_core_sharp_rows(eng):
    """Peaked softmax rows (large logits), L = 257: the scalar key must enter the row maximum, and rows whose mass sits
    on the scalar key / whose query is the scalar row must come out right."""
    n, L, H = 2, 257, 2
    D = H * 64
    qkv = torch.randn(n * L, 3 * D, device=DEV)
    qkv[:, :2 * D] *= 3.0                                                       # logits ~ N(0, 9^2 * 64 / 64)
    qkv = qkv.to(eng.operand_dtype)
    q, k, v = qkv.float().reshape(n, L, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1) @ v).transpose(1, 2).reshape(n * L, D)
    out = eng.op_attention(qkv, n, L, H).float().cpu()
    assert torch.isfinite(out).all()
    assert rel_l2(out, ref) < 2e-3
    rows = ref.reshape(n, L, D)
    assert rel_l2(out.reshape(n, L, D)[:, 256], rows[:, 256]) < 2e-3            # the scalar query row itself


def test_attention_rejects_too_many_tokens(eng, lib):
    for L in (258, 272, 300):
        with pytest.raises(lib.FsarError) as err:
            eng.op_attention(torch.zeros(L, 192, device=DEV, dtype=eng.operand_dtype), 1, L, 1)
        assert err.value.code == -1


@pytest.mark.parametrize("n,t", [(5, 8), (5, 9), (10, 17), (3, 33), (1, 1)])
def test_modulator_fp32(eng, n, t):
    from oracle import fsar_oracle as O
    meta, _ = load_golden("tiny_5w1s")
    g, sd, tt, te, _ = regenerate(meta)
    eng.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    x = torch.randn(n, t, g["embed_dim"])
    assert rel_l2(eng.modulate(x.to(DEV)), O.modulator(sd, g, x)) < 5e-6       # fp32 end to end


@pytest.mark.parametrize("Q,way,T", [(5, 5, 8), (10, 10, 16), (20, 20, 32), (1, 3, 1), (2, 2, 2), (7, 3, 5)])
@pytest.mark.parametrize("single", [False, True])
def test_cos_otam_fp32(eng, Q, way, T, single):
    from oracle import fsar_oracle as O
    E = 128
    q, p = torch.randn(Q, T, E), torch.randn(way, T, E)
    d = (1 - O.cos_sim(q.reshape(Q * T, E), p.reshape(way * T, E))).reshape(Q, T, way, T).permute(0, 2, 1, 3)
    cum = O.otam_cum_dist(d) if single else O.otam_cum_dist(d) + O.otam_cum_dist(d.transpose(2, 3))
    logits, dists, cum_gpu = eng.otam_logits(q.to(DEV), p.to(DEV), single, True)
    assert (dists.cpu() - d).abs().max() < 1e-6
    assert (logits.cpu() + cum).abs().max() < 1e-5 * max(1.0, float(cum.abs().max()))
    assert torch.equal(logits, -cum_gpu)


def test_otam_zero_vectors_do_not_nan(eng):
    # cos_sim's +0.01 on the product of norms (few_shot.py:1121) keeps all-zero frames finite
    q = torch.zeros(2, 4, 128, device=DEV)
    p = torch.randn(3, 4, 128, device=DEV)
    logits, dists, _ = eng.otam_logits(q, p, False, True)
    assert torch.isfinite(logits).all() and torch.allclose(dists, torch.ones_like(dists))


def test_otam_rejects_long_sequences(eng, lib):
    with pytest.raises(lib.FsarError):
        eng.otam_logits(torch.zeros(1, 33, 128, device=DEV), torch.zeros(1, 33, 128, device=DEV))


def test_weights_must_be_set_before_forward(lib):
    from clip_fsar_b200 import synth
    g = synth.full_geometry("tiny")
    e = lib.Engine(**dict(g, max_frames=8, max_videos=10, max_tokens=8, max_classes=64, otam_lambda=0.5, device=0))
    assert len(e.missing_weights()) == len(synth.state_dict_shapes(g)) + 2
    with pytest.raises(lib.FsarError) as err:
        e.vit_forward(torch.zeros(1, 3, 32, 32, device=DEV))
    assert err.value.code == -5
    with pytest.raises(lib.FsarError) as err:
        e.set_weight("backbone.nope", torch.zeros(3))
    assert err.value.code == -4
    with pytest.raises(lib.FsarError) as err:
        e.set_weight("scale", torch.zeros(3))
    assert err.value.code == -1
    e.close()
