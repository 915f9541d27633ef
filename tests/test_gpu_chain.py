"""GPU parity of chains (paro_chain_forward): several fused linears + their element-wise neighbours in one launch.

  * a chain of independent steps, and a chain that feeds y forward, equal the same steps launched as one-step chains
    BIT FOR BIT (same kernel, fixed summation order whoever arrives last);
  * every stage of the decoder tail (o -> +residual/RMSNorm -> gate_up -> SiLU*up -> down -> +residual/RMSNorm -> qkv)
    vs the oracle applied to the GPU's own previous stage (teacher forcing): normwise relative error <= 1e-3
    (BASELINE.json's tolerance) in the activation dtype;
  * the only state a launch leaves behind is the bumped epoch: repeated launches and CUDA-graph replays (PDL edges between
    launches included) reproduce the eager result exactly.
"""
import numpy as np
import pytest
import torch

from paroquant_b200.checkpoint import make_synthetic_activations, make_synthetic_layer

pytestmark = pytest.mark.gpu
TOL = 1e-3
_TD = {"float16": torch.float16, "bfloat16": torch.bfloat16}


@pytest.fixture(scope="module")
def mods():
    from paroquant_b200 import chain
    from paroquant_b200.linear import ParoLinearKernel
    return chain, ParoLinearKernel


def _np(t):
    return t.float().cpu().numpy()


def _tail(mods, hidden, inter, qkv_parts, M, dt, seed=7):
    chain, PK = mods
    T = _TD[dt]
    Ls = {"o": make_synthetic_layer(hidden, [hidden], seed=seed), "gate_up": make_synthetic_layer(hidden, [inter, inter], seed=seed + 1),
          "down": make_synthetic_layer(inter, [hidden], seed=seed + 2), "qkv": make_synthetic_layer(hidden, qkv_parts, seed=seed + 3)}
    ks = {n: PK.from_buffers(L.to("cuda"), T, check_pairs=False) for n, L in Ls.items()}
    g = torch.Generator().manual_seed(seed)
    attn = make_synthetic_activations(M, hidden, seed=seed + 10, dtype=T).cuda()
    resid = make_synthetic_activations(M, hidden, seed=seed + 11, dtype=T).cuda()
    w1 = (1.0 + 0.1 * torch.randn(hidden, generator=g)).to(T).cuda()
    w2 = (1.0 + 0.1 * torch.randn(hidden, generator=g)).to(T).cuda()
    ch, bufs = chain.decoder_tail(ks["o"], ks["gate_up"], ks["down"], ks["qkv"], attn_out=attn, residual=resid,
                                  post_attn_norm=w1, next_input_norm=w2, eps=1e-5)
    return Ls, ks, ch, bufs, attn, resid, w1, w2


@pytest.mark.parametrize("dt", ["bfloat16", "float16"])
@pytest.mark.parametrize("M", [1, 3, 8, 16])
def test_decoder_tail_vs_oracle_stagewise(mods, oracle, dt, M):
    Ls, ks, ch, bufs, attn, resid, w1, w2 = _tail(mods, 512, 1024, [512, 128, 128], M, dt)
    ch()
    torch.cuda.synchronize()
    O = oracle
    d = {n: L.numpy_dict() for n, L in Ls.items()}
    h1 = O.add_residual(O.linear(_np(attn), d["o"], dt), _np(resid), dt)
    assert O.rel_err(_np(bufs["residual_mid"]), h1) < TOL
    act = O.linear(O.rms_norm(_np(bufs["residual_mid"]), _np(w1), 1e-5, dt), d["gate_up"], dt)
    assert O.rel_err(_np(bufs["mlp_act"]), act) < TOL
    h2 = O.add_residual(O.linear(O.silu_and_mul(_np(bufs["mlp_act"]), dt), d["down"], dt), _np(bufs["residual_mid"]), dt)
    assert O.rel_err(_np(bufs["residual_out"]), h2) < TOL
    qkv = O.linear(O.rms_norm(_np(bufs["residual_out"]), _np(w2), 1e-5, dt), d["qkv"], dt)
    assert O.rel_err(_np(bufs["qkv"]), qkv) < TOL


def _torch_rms(h, w, eps):
    hf = h.float()
    rstd = torch.rsqrt(hf.pow(2).mean(-1, keepdim=True) + eps)
    return ((hf * rstd).to(h.dtype) * w)


def _torch_silu_mul(y):
    k = y.shape[-1] // 2
    g = y[..., :k].float()
    return (g / (1.0 + torch.exp(-g))).to(y.dtype) * y[..., k:]


@pytest.mark.parametrize("M", [1, 4, 16])
def test_decoder_tail_llama_shapes_vs_unfused_gpu(mods, M):
    """Llama-3-8B shapes: each stage vs the single-linear kernel + plain torch element-wise ops on the same inputs."""
    Ls, ks, ch, bufs, attn, resid, w1, w2 = _tail(mods, 4096, 14336, [4096, 1024, 1024], M, "bfloat16", seed=21)
    ch()
    torch.cuda.synchronize()

    def err(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm())

    h1 = (ks["o"](attn).float() + resid.float()).to(attn.dtype)
    assert err(bufs["residual_mid"], h1) < TOL
    act = ks["gate_up"](_torch_rms(bufs["residual_mid"], w1, 1e-5))
    assert err(bufs["mlp_act"], act) < TOL
    h2 = (ks["down"](_torch_silu_mul(bufs["mlp_act"])).float() + bufs["residual_mid"].float()).to(attn.dtype)
    assert err(bufs["residual_out"], h2) < TOL
    qkv = ks["qkv"](_torch_rms(bufs["residual_out"], w2, 1e-5))
    assert err(bufs["qkv"], qkv) < TOL


@pytest.mark.parametrize("M", [1, 5, 16])
def test_chain_of_plain_linears_matches_single_launches(mods, M):
    """A chain that feeds y forward == the same steps launched as one-step chains, BIT FOR BIT (same kernel, fixed summation
    order whoever arrives last), and == the single-linear entry point (the cluster kernel: another summation order) within
    the tolerance; independent steps (every x given) wait for nothing and give the same bits."""
    chain, PK = mods
    T = torch.bfloat16
    shapes = [(1024, [256, 128]), (384, [1024]), (1024, [640]), (640, [48, 16, 32])]
    ks = [PK.from_buffers(make_synthetic_layer(K, p, seed=60 + i).to("cuda"), T, check_pairs=False) for i, (K, p) in enumerate(shapes)]
    x0 = make_synthetic_activations(M, 1024, seed=3, dtype=T).cuda()
    # dependent chain: 1024 -> 384 -> 1024 -> 640 -> 96
    ys = [torch.empty(M, k.shape.out_features, dtype=T, device="cuda") for k in ks]
    steps = [chain.ChainStep(ks[0], x=x0, y=ys[0])] + [chain.ChainStep(ks[i], y=ys[i]) for i in range(1, 4)]
    chain.ParoChain(steps, M)()
    ref = x0
    for k, y in zip(ks, ys):
        one = torch.empty_like(y)
        chain.ParoChain([chain.ChainStep(k, x=ref, y=one)], M)()
        assert torch.equal(y, one)
        lin = k(ref)
        assert float((lin.double() - y.double()).norm() / y.double().norm()) < 5e-4
        ref = one
    # independent steps (every x given): nothing waits, results unchanged
    xs = [x0] + [make_synthetic_activations(M, k.shape.in_features, seed=9 + i, dtype=T).cuda() for i, k in enumerate(ks[1:])]
    ys2 = [torch.empty_like(y) for y in ys]
    chain.ParoChain([chain.ChainStep(k, x=x, y=y) for k, x, y in zip(ks, xs, ys2)], M)()
    for k, x, y in zip(ks, xs, ys2):
        one = torch.empty_like(y)
        chain.ParoChain([chain.ChainStep(k, x=x, y=one)], M)()
        assert torch.equal(y, one)


def test_stream_kernel_serves_single_linears(mods, oracle, monkeypatch):
    """The chain kernel as the single-linear kernel (where the cluster kernel has no plan): a one-step chain vs the oracle on
    partial blocks, merged projections and every row count."""
    chain, PK = mods
    for K, parts, Ms in [(512, [64], (1, 2, 3, 8, 9, 16)), (640, [48, 16, 32], (1, 5, 16)), (256, [272, 16], (3, 12)), (1024, [256, 128], (1, 4, 7))]:
        L = make_synthetic_layer(K, parts, seed=43, bias=(K == 1024))
        k = PK.from_buffers(L.to("cuda"), torch.bfloat16, check_pairs=False)
        bias = None if L.bias is None else L.bias.to("cuda", torch.bfloat16)
        cache = {}
        for M in Ms:
            x = make_synthetic_activations(M, K, seed=50 + M, dtype=torch.bfloat16)
            y = torch.empty(M, sum(parts), dtype=torch.bfloat16, device="cuda")
            chain.ParoChain([chain.ChainStep(k, x=x.cuda(), y=y, bias=bias)], M)()
            ref = oracle.linear(x.float().numpy(), L.numpy_dict(), "bfloat16", W_cache=cache)
            assert oracle.rel_err(_np(y), ref) < TOL, (K, parts, M)


def test_repeated_launches_and_graph_replay_reproduce_eager(mods):
    """Tags are per launch: 3 eager launches, the warm-up, the capture and 3 graph replays (the chain between two single-linear
    launches, PDL edges included) give identical bits."""
    chain, PK = mods
    Ls, ks, ch, bufs, attn, resid, w1, w2 = _tail(mods, 1024, 2048, [1024, 256, 256], 4, "bfloat16", seed=33)
    pre = PK.from_buffers(make_synthetic_layer(1024, [1024], seed=90).to("cuda"), torch.bfloat16, check_pairs=False)
    x = make_synthetic_activations(4, 1024, seed=91, dtype=torch.bfloat16).cuda()

    def step():
        attn.copy_(pre(x))
        ch()
        return pre(bufs["residual_out"])

    outs = []
    for _ in range(3):
        y = step()
        torch.cuda.synchronize()
        outs.append((y.clone(), bufs["qkv"].clone(), bufs["residual_out"].clone()))
    for o in outs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(o, outs[0]))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        yg = step()
    for _ in range(3):
        bufs["qkv"].zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(yg, outs[0][0]) and torch.equal(bufs["qkv"], outs[0][1]) and torch.equal(bufs["residual_out"], outs[0][2])
    sync = ch.workspace[:8].view(torch.int32).tolist()
    assert sync == [7, 0], f"sync words [epoch, CTAs done] = {sync}: every launch bumps the epoch once and clears the counter"


def test_chain_argument_checks(mods):
    chain, PK = mods
    k = PK.from_buffers(make_synthetic_layer(256, [128], seed=1).to("cuda"), torch.bfloat16, check_pairs=False)
    x = torch.zeros(2, 256, dtype=torch.bfloat16, device="cuda")
    y = torch.zeros(2, 128, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError, match="needs an input"):
        chain.ParoChain([chain.ChainStep(k, y=y)], 2)()
    with pytest.raises(RuntimeError, match="RMSNORM"):
        chain.ParoChain([chain.ChainStep(k, x=x, y=y, x_op="rmsnorm", norm_weight=x[0].contiguous())], 2)()
    with pytest.raises(RuntimeError, match="1..16"):
        chain.ParoChain([chain.ChainStep(k, x=torch.zeros(17, 256, dtype=torch.bfloat16, device="cuda"))], 17)
    with pytest.raises(RuntimeError, match="alias"):
        r = torch.zeros(2, 128, dtype=torch.bfloat16, device="cuda")
        chain.ParoChain([chain.ChainStep(k, x=x, epilogue="add_residual", residual_in=r, residual_out=r)], 2)


@pytest.mark.parametrize("M", [1, 4])
def test_chain_mixes_group64_and_group128_steps(mods, oracle, M):
    """A chain whose steps have different record sizes (group_size 64: two scale / zero sets per record): the ring stage is
    sized for the largest, every step copies its own record size."""
    chain, PK = mods
    dt, T = "bfloat16", torch.bfloat16
    La = make_synthetic_layer(512, [768], group_size=64, seed=301)
    Lb = make_synthetic_layer(768, [256, 128], group_size=128, seed=302)
    Lc = make_synthetic_layer(384, [512], group_size=64, seed=303)
    ka, kb, kc = (PK.from_buffers(L.to("cuda"), T, check_pairs=False) for L in (La, Lb, Lc))
    x = make_synthetic_activations(M, 512, seed=310, dtype=T).cuda()
    ya = torch.empty(M, 768, dtype=T, device="cuda")
    yb = torch.empty(M, 384, dtype=T, device="cuda")
    yc = torch.empty(M, 512, dtype=T, device="cuda")
    ch = chain.ParoChain([chain.ChainStep(ka, x=x, y=ya), chain.ChainStep(kb, y=yb), chain.ChainStep(kc, y=yc)], M)
    ch()
    torch.cuda.synchronize()
    O = oracle
    assert O.rel_err(_np(ya), O.linear(_np(x), La.numpy_dict(), dt)) < TOL
    assert O.rel_err(_np(yb), O.linear(_np(ya), Lb.numpy_dict(), dt)) < TOL
    assert O.rel_err(_np(yc), O.linear(_np(yb), Lc.numpy_dict(), dt)) < TOL
