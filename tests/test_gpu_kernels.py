"""GPU parity tests (B200): every C-ABI kernel against the oracle."""
import numpy as np
import pytest
import torch

from oracle import oracle_port as orc
from tests._golden import known_answers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda):
    from rl_sandbox_b200 import ops as _ops
    return _ops


# ---------------------------------------------------------------- K2 ---------------------------
@pytest.mark.parametrize("case", known_answers(), ids=lambda c: c["name"])
def test_k2_known_answers(ops, cuda, case):
    """reference test/dreamer/test_critic.py:14-62 vectors through the CUDA kernel."""
    vs = torch.tensor(case["vs"], device=cuda).view(-1, 1)
    rs = torch.tensor(case["rs"] + [0.0], device=cuda).view(-1, 1)
    ds = torch.tensor(case["ds"], device=cuda).view(-1, 1)
    out, _, _ = ops.lambda_return(rs, vs, ds, case["lam"])
    assert out.view(-1).tolist() == case["expected"]


@pytest.mark.parametrize("T,N", [(16, 800), (16, 801), (2, 5), (3, 4), (16, 65536), (31, 1000)])
def test_k2_bit_exact_vs_oracle(ops, cuda, T, N):
    g = torch.Generator().manual_seed(T * 1000 + N)
    r, v = torch.randn(T, N, generator=g), torch.randn(T, N, generator=g)
    d = (torch.rand(T, N, generator=g) > 0.1).float()
    vs_o, w_o, adv_o = orc.lambda_return_c(r.numpy(), v.numpy(), d.numpy(), 0.95)
    vs, w, adv = ops.lambda_return(r.to(cuda), v.to(cuda), d.to(cuda), 0.95)
    assert np.array_equal(vs.cpu().numpy(), vs_o)
    assert np.array_equal(w.cpu().numpy(), w_o)
    assert np.array_equal(adv.cpu().numpy(), adv_o)
    # and against the reference's own loop (ac.py:52-62) evaluated by torch on the CPU
    assert torch.equal(vs.cpu(), orc.lambda_return_loop(v, r[:-1], d, 0.95))


def test_k2_reference_shape_with_trailing_axis_and_nan_discount(ops, cuda):
    """(T, N, 1) tensors as DreamerV2.train passes them; Bernoulli.mode NaNs propagate like torch."""
    g = torch.Generator().manual_seed(5)
    r, v = torch.randn(16, 64, 1, generator=g), torch.randn(16, 64, 1, generator=g)
    d = torch.ones(16, 64, 1)
    d[3, 7, 0] = float("nan")
    vs, w, adv = ops.lambda_return(r.to(cuda), v.to(cuda), d.to(cuda), 0.95)
    ref = orc.lambda_return_loop(v, r[:-1], d, 0.95)
    assert vs.shape == (15, 64, 1) and w.shape == (16, 64, 1) and adv.shape == (14, 64, 1)
    assert torch.equal(torch.isnan(vs.cpu()), torch.isnan(ref))
    assert torch.equal(torch.nan_to_num(vs.cpu()), torch.nan_to_num(ref))


def test_k2_batch_major_warp_shuffle_variant(ops, cuda):
    g = torch.Generator().manual_seed(6)
    for T, N in [(16, 1000), (11, 333), (32, 64)]:
        r, v = torch.randn(T, N, generator=g), torch.randn(T, N, generator=g)
        d = (torch.rand(T, N, generator=g) > 0.1).float()
        vs_o, w_o, adv_o = orc.lambda_return_c(r.numpy(), v.numpy(), d.numpy(), 0.9)
        vs, w, adv = ops.lambda_return(r.t().contiguous().to(cuda), v.t().contiguous().to(cuda),
                                       d.t().contiguous().to(cuda), 0.9, batch_major=True)
        # fp32 scan tolerance of the north star: 1e-5
        torch.testing.assert_close(vs.t().cpu(), torch.from_numpy(vs_o), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(w.t().cpu(), torch.from_numpy(w_o), rtol=0, atol=0)
        torch.testing.assert_close(adv.t().cpu(), torch.from_numpy(adv_o), rtol=1e-5, atol=1e-5)


def test_k2_backward_matches_autograd_of_reference_loop(ops, cuda):
    g = torch.Generator().manual_seed(7)
    T, N = 16, 500
    r = torch.randn(T, N, generator=g).requires_grad_()
    v = torch.randn(T, N, generator=g).requires_grad_()
    d = torch.rand(T, N, generator=g).requires_grad_()
    gvs = torch.randn(T - 1, N, generator=g)
    (orc.lambda_return_loop(v, r[:-1], d, 0.95) * gvs).sum().backward()
    rc, vc, dc = (x.detach().to(cuda).requires_grad_() for x in (r, v, d))
    out = ops.LambdaReturnFn.apply(rc, vc, dc, 0.95)
    (out * gvs.to(cuda)).sum().backward()
    torch.testing.assert_close(rc.grad.cpu()[:-1], r.grad[:-1], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(vc.grad.cpu(), v.grad, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(dc.grad.cpu()[:-1], d.grad[:-1], rtol=1e-5, atol=1e-5)


def test_k2_linearity_at_full_size(ops, cuda):
    """size-independent property at the sweep's largest N: vs is linear in (r, v) for fixed d."""
    T, N = 16, 262144
    g = torch.Generator(device="cuda").manual_seed(8)
    r1, v1, r2, v2 = (torch.randn(T, N, device=cuda, generator=g) for _ in range(4))
    d = (torch.rand(T, N, device=cuda, generator=g) > 0.05).float()
    a = ops.lambda_return(r1, v1, d, 0.95)[0]
    b = ops.lambda_return(r2, v2, d, 0.95)[0]
    c = ops.lambda_return(r1 + r2, v1 + v2, d, 0.95)[0]
    torch.testing.assert_close(c, a + b, rtol=1e-5, atol=2e-5)
    # d == 0 everywhere => vs == r (test_critic.py:14-23 at scale)
    z = ops.lambda_return(r1, v1, torch.zeros_like(d), 0.95)[0]
    assert torch.equal(z, r1[:-1])


# ---------------------------------------------------------------- sampler / RNG ----------------
def test_sampler_bit_exact(ops, cuda):
    g = torch.Generator().manual_seed(2)
    for classes, rows in [(32, 100000), (17, 5000), (1, 10), (64, 1000), (3, 7)]:
        logits = torch.randn(rows, classes, generator=g) * 3
        un = torch.rand(rows, classes, generator=g)
        idx = ops.sample_categorical(logits.to(cuda), un.to(cuda)).cpu().long()
        assert torch.equal(idx, orc.sample_categorical(logits, un))


def test_sampler_edge_uniforms_and_ties(ops, cuda):
    un = torch.tensor([[0.0, 1.0, 0.5, 1e-30], [0.99999994, 0.99999994, 0.1, 0.2]])
    lg = torch.zeros(2, 4)
    idx = ops.sample_categorical(lg.to(cuda), un.to(cuda)).cpu().long()
    assert torch.equal(idx, orc.sample_categorical(lg, un))
    same = ops.sample_categorical(torch.zeros(5, 8, device=cuda), torch.full((5, 8), 0.25, device=cuda))
    assert same.tolist() == [0] * 5  # ties -> lowest index, like torch.argmax


def test_latent_sampler_bit_exact(ops, cuda):
    """The rollout's own latent draw (fast screening pass + reference-order redraw of unsure groups) gives the oracle's
    indices, bit for bit: random logits, explicit uniforms and Philox noise, ragged row counts, groups not a power of 2."""
    g = torch.Generator().manual_seed(12)
    for rows, groups, scale in [(20000, 32, 3.0), (4097, 32, 0.05), (333, 5, 10.0), (1, 32, 1.0), (129, 64, 1.0)]:
        logits = torch.randn(rows, groups, 32, generator=g) * scale
        un = torch.rand(rows, groups, 32, generator=g)
        idx, onehot = ops.sample_latent(logits.to(cuda), un.to(cuda), want_onehot=True)
        want = orc.sample_categorical(logits, un)
        assert torch.equal(idx.cpu().long(), want)
        assert torch.equal(onehot.cpu().view(rows, groups, 32).argmax(-1), want)
        assert torch.equal(onehot.cpu().sum(), torch.tensor(float(rows * groups)))
    rows, groups = 5000, 32
    logits = torch.randn(rows, groups, 32, generator=g)
    idx = ops.sample_latent(logits.to(cuda), None, seed=2 ** 41 + 77, row_offset=123456, step=9)
    un = torch.from_numpy(orc.philox_uniform(2 ** 41 + 77, 123456, 9, 0, groups * 32, rows)).view(rows, groups, 32)
    assert torch.equal(idx.cpu().long(), orc.sample_categorical(logits, un))


def test_latent_sampler_near_ties_take_the_exact_path(ops, cuda):
    """Scores engineered so that the runner-up sits within a few ulp of the winner (below the screening pass's error):
    only the bit-reproducible redraw can order them; plus exact ties, edge uniforms and non-finite logits."""
    g = torch.Generator().manual_seed(13)
    rows, groups = 3000, 32
    logits = torch.randn(rows, groups, 32, generator=g) * 2
    un = torch.rand(rows, groups, 32, generator=g)
    gum = orc.gumbel(un).view(rows, groups, 32)
    score = logits + gum
    best = score.max(-1, keepdim=True).values
    j = torch.randint(0, 32, (rows, groups, 1), generator=g)
    ulps = torch.randint(-3, 4, (rows, groups, 1), generator=g).float()
    target = best + ulps * torch.finfo(torch.float32).eps * best.abs()
    logits.scatter_(-1, j, target - gum.gather(-1, j))       # class j now scores within 3 ulp of the winner
    want = orc.sample_categorical(logits, un)
    got = ops.sample_latent(logits.to(cuda), un.to(cuda)).cpu().long()
    assert torch.equal(got, want)
    frac = (want == j.squeeze(-1)).float().mean().item()
    assert 0.2 < frac < 0.9   # the planted class wins some and loses some: the order really is decided at ulp level
    # exact ties -> lowest index; edge uniforms; -inf / +inf / NaN logits follow the reference-order scan
    lg = torch.zeros(4, 32, 32)
    u2 = torch.full((4, 32, 32), 0.25)
    u2[1, :, 5] = 0.0
    u2[1, :, 9] = 1.0
    u2[2, :, 3] = 0.99999994
    u2[2, :, 30] = 0.99999994
    u2[3] = torch.rand(32, 32, generator=g)
    lg[3, :, 4] = float("-inf")
    lg[3, 1, 7] = float("inf")
    lg[3, 2, 0] = float("nan")
    lg[3, 3, 11] = float("nan")
    got = ops.sample_latent(lg.to(cuda), u2.to(cuda)).cpu().long()
    assert torch.equal(got, orc.sample_categorical(lg, u2))
    assert got[0].tolist() == [0] * 32


def test_philox_bit_exact(ops, cuda):
    for seed, n0, t, stream, per_row, rows in [(0, 0, 0, 0, 1024, 64), (2 ** 40 + 12345, 1000000, 14, 1, 17, 300)]:
        u = ops.philox_uniform(seed, n0, t, stream, per_row, rows).cpu().numpy()
        assert np.array_equal(u, orc.philox_uniform(seed, n0, t, stream, per_row, rows))


# ---------------------------------------------------------------- packing / GEMM ---------------
def test_pack_roundtrip(ops, cuda):
    for rows, cols in [(300, 200), (128, 64), (1, 1), (1000, 1041)]:
        x = torch.randn(rows, cols, device=cuda)
        assert torch.equal(ops.unpack_rows(ops.pack_rows(x), rows, cols), x.bfloat16().float())


@pytest.fixture(params=[1, 2, 4], ids=lambda c: f"cluster{c}")
def cluster(request):
    """CTAs per cluster sharing a weight block through TMA multicast."""
    from rl_sandbox_b200 import _lib
    lib = _lib.load()
    assert lib.rlsb_set_cluster_size(request.param) == request.param
    yield request.param
    lib.rlsb_set_cluster_size(2)


@pytest.mark.parametrize("M,K,N", [(128, 64, 32), (300, 128, 64), (300, 448, 400), (1000, 2048, 3072),
                                   (5000, 1088, 1024), (777, 256, 600), (1, 64, 17), (129, 1280, 400),
                                   (40000, 448, 400)])
def test_tcgen05_gemm_vs_fp32_reference(ops, cuda, cluster, M, K, N):
    """acc in fp32 on bf16-rounded operands: compare with torch fp32 matmul on the same rounded operands."""
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g).to(cuda)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    rb, nb = ops.plan_blocks(N)
    kp = ops.round_up(K, 64)
    out, st = ops.gemm_bias(ops.pack_rows(x), kp, ops.pack_rows(w, row_block=rb, rows_pad=rb * nb, k_pad=kp), rb, nb,
                            b, M, N, want_stats=True)
    ref = (x.bfloat16().double() @ w.bfloat16().double().t() + b.double()).float()
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=2e-5)
    # per-(row, n-block) partials are (sum, sum of squares) over the block's valid columns
    mean = st[:, :M, 0].sum(0) / N
    var = st[:, :M, 1].sum(0) / N - mean ** 2
    torch.testing.assert_close(mean, ref.mean(-1), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(var, ref.var(-1, unbiased=False), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("M,K,N,use_ln,act", [(300, 448, 400, True, 1), (1000, 2048, 400, True, 1), (300, 256, 200, False, 1),
                                              (260, 448, 17, False, 0), (5000, 448, 400, True, 1)])
def test_staged_epilogue_output_is_bit_identical(ops, cuda, M, K, N, use_ln, act):
    """The full-row epilogue writes its 128 x 64 output tiles through shared memory + one bulk copy per tile (default) or with
    16-byte stores per thread: only the path to HBM differs, every bit of the packed image (padding included) is the same."""
    from rl_sandbox_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g).to(cuda)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    gam = (1 + 0.1 * torch.randn(N, generator=g)).to(cuda)
    bet = (0.1 * torch.randn(N, generator=g)).to(cuda)
    rb, nb = ops.plan_blocks(N)
    kp, okp = ops.round_up(K, 64), ops.round_up(N, 64)
    xp, wp = ops.pack_rows(x), ops.pack_rows(w, row_block=rb, rows_pad=rb, k_pad=kp)
    outs = []
    try:
        for staged in (1, 0):
            assert lib.rlsb_set_staged_output(staged) == staged
            outs.append(ops.gemm_ln_act(xp, kp, wp, rb, b, M, N, gam if use_ln else None, bet if use_ln else None, 1e-5, act,
                                        okp).clone())
    finally:
        lib.rlsb_set_staged_output(1)
    rows = ops.round_up(M, 128)
    a0, a1 = (ops.unpack_rows(o, rows, okp, k_pad=okp) for o in outs)
    assert torch.equal(a0[:M], a1[:M])


@pytest.mark.parametrize("M,K,N,use_ln,act", [(300, 448, 400, True, 1), (1000, 2048, 400, True, 1),
                                              (300, 256, 200, False, 1), (260, 448, 17, False, 0),
                                              (130, 384, 384, True, 2), (33000, 448, 400, True, 1)])
def test_tcgen05_gemm_layernorm_act_epilogue(ops, cuda, cluster, M, K, N, use_ln, act):
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g).to(cuda)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    gam = (1 + 0.1 * torch.randn(N, generator=g)).to(cuda)
    bet = (0.1 * torch.randn(N, generator=g)).to(cuda)
    rb, nb = ops.plan_blocks(N)
    assert nb == 1
    kp, okp = ops.round_up(K, 64), ops.round_up(N, 64)
    outp = ops.gemm_ln_act(ops.pack_rows(x), kp, ops.pack_rows(w, row_block=rb, rows_pad=rb, k_pad=kp), rb, b, M, N,
                           gam if use_ln else None, bet if use_ln else None, 1e-5, act, okp)
    y = ops.unpack_rows(outp, M, okp, k_pad=okp)
    pre = x.bfloat16().float() @ w.bfloat16().float().t() + b
    if use_ln:
        pre = torch.nn.functional.layer_norm(pre, (N,), gam, bet, 1e-5)
    ref = {0: lambda t: t, 1: torch.nn.functional.elu, 2: torch.relu}[act](pre)
    # output is rounded to bf16 (one ulp = 2^-8 relative)
    torch.testing.assert_close(y[:, :N], ref, rtol=2 ** -7, atol=1e-5)
    assert torch.equal(y[:, N:], torch.zeros_like(y[:, N:]))
