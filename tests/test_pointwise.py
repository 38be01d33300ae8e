"""GPU parity of the 1x1-convolution layer (sed_pointwise_forward) against torch.nn.functional.conv1d in FP32:
every kernel the dispatcher can pick (weight-stationary tcgen05, per-tile tcgen05, few-output-channel FFMA), ragged point
counts, the fused input affine + activation, the statistics / max-min epilogue, and a repeated-launch run of the
MMA-bound shape (raw input, 256 -> 128) whose producer warps outrun the TMA ring -- the shape that exposed a phase-parity
hazard between the two producer groups of pw_tc2_kernel.  Reference layers: src/SEDNet.py:292-342."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 2e-5      # FP16 hi/lo split, 22 significant bits per operand, FP32 accumulation: relative to the output scale


@pytest.fixture(scope="module")
def lib():
    assert torch.cuda.is_available()
    from sednet_b200.src import _lib
    _lib.load()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return _lib


def _case(B, N, Cin, Cout, act, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = torch.device("cuda")
    x = torch.randn(B, Cin, N, device=dev, generator=g)
    W = torch.randn(Cout, Cin, device=dev, generator=g) / Cin ** 0.5
    bias = torch.randn(Cout, device=dev, generator=g)
    if act is None:
        return x, W, bias, None, None, F.conv1d(x.double(), W.double()[:, :, None], bias.double())
    a = torch.rand(B, Cin, device=dev, generator=g) + 0.5
    s = torch.randn(B, Cin, device=dev, generator=g)
    z = (a[:, :, None] * x + s[:, :, None]).double()
    z = (z, F.relu(z), F.leaky_relu(z, 0.2))[act]
    return x, W, bias, a, s, F.conv1d(z, W.double()[:, :, None], bias.double())


def _run(lib, x, W, bias, a, s, act, want_stats=False):
    B, Cin, N = x.shape
    Cout = W.shape[0]
    P = (N + 127) // 128
    y = torch.full((B, Cout, N), float("nan"), device=x.device)
    stats = torch.zeros(B, P, (Cout + 31) // 32, 2, device=x.device, dtype=torch.float64) if want_stats else None
    mm = torch.zeros(B, P, Cout, 2, device=x.device) if want_stats else None
    lib.call("sed_pointwise_forward", lib.ptr(x), Cin * N, lib.ptr(W), Cin, lib.ptr(bias), lib.ptr(a), lib.ptr(s),
             act or 0, lib.ptr(y), Cout * N, lib.ptr(stats), lib.ptr(mm), B, Cin, Cout, N, lib.stream())
    torch.cuda.synchronize()
    return y, stats, mm


@pytest.mark.parametrize("B,N,Cin,Cout,act", [
    (2, 1000, 256, 128, None),      # segmentation head (raw input), ragged last tile
    (2, 1024, 256, 256, 1),         # conv + GroupNorm + ReLU chain
    (3, 777, 128, 256, 2),          # EdgeConv per-point GEMM shape, LeakyReLU input, odd N (no 16-byte rows: per-tile kernel)
    (2, 1500, 256, 1024, 0),        # mlp1: eight 128-channel groups
    (2, 1300, 256, 6, 1),           # primitive-type head: few output channels
    (2, 1300, 128, 2, 0),           # edge head
    (1, 600, 6, 128, None),         # first EdgeConv: Cin below the tensor-core kernels
    (2, 900, 192, 64, 1),           # Cin not a multiple of 64
    (1, 5, 64, 70, 1),              # fewer points than a lane group, Cout not a multiple of 32
    (2, 132, 33, 96, 2),            # odd Cin (zero-padded K), ragged second tile
    (1, 4096, 48, 200, None),       # two 128-channel groups, the second one partial
    (3, 260, 256, 8, 2),            # widest few-output-channel layer
])
def test_pointwise_matches_conv1d(lib, B, N, Cin, Cout, act):
    x, W, bias, a, s, ref = _case(B, N, Cin, Cout, act, seed=Cin + Cout)
    y, _, _ = _run(lib, x, W, bias, a, s, act)
    scale = float(ref.abs().max())
    assert float((y.double() - ref).abs().max()) <= TOL * scale


@pytest.mark.parametrize("N,Cin,Cout", [(1100, 256, 256), (1100, 128, 200), (333, 64, 96)])
def test_pointwise_epilogue_statistics(lib, N, Cin, Cout):
    """stats = per (128-point tile, 32-channel block) sum and sum of squares; mm = per (tile, channel) max and min."""
    B = 2
    x, W, bias, a, s, ref = _case(B, N, Cin, Cout, 1, seed=5)
    y, stats, mm = _run(lib, x, W, bias, a, s, 1, want_stats=True)
    P = (N + 127) // 128
    pad = P * 128 - N
    cpad = (Cout + 31) // 32 * 32 - Cout
    yp = F.pad(y.double(), (0, pad, 0, cpad)).view(B, (Cout + cpad) // 32, 32, P, 128)
    s1 = yp.sum((2, 4)).permute(0, 2, 1)
    s2 = (yp * yp).sum((2, 4)).permute(0, 2, 1)
    assert torch.allclose(stats[..., 0], s1, rtol=1e-5, atol=1e-3)
    assert torch.allclose(stats[..., 1], s2, rtol=1e-5, atol=1e-3)
    ymax = F.pad(y, (0, pad), value=float("-inf")).view(B, Cout, P, 128).amax(3).permute(0, 2, 1)
    ymin = F.pad(y, (0, pad), value=float("inf")).view(B, Cout, P, 128).amin(3).permute(0, 2, 1)
    assert torch.equal(mm[..., 0], ymax) and torch.equal(mm[..., 1], ymin)


@pytest.mark.parametrize("Cin,act", [(256, None), (256, 1), (128, None)])
def test_pointwise_repeated_full_size(lib, Cin, act):
    """16 clouds x 10 000 points, 30 launches: every launch must reproduce the first one bit for bit and match conv1d."""
    B, N, Cout = 16, 10000, 128
    x, W, bias, a, s, ref = _case(B, N, Cin, Cout, act, seed=9)
    first, _, _ = _run(lib, x, W, bias, a, s, act)
    assert float((first.double() - ref).abs().max()) <= TOL * float(ref.abs().max())
    for _ in range(29):
        y, _, _ = _run(lib, x, W, bias, a, s, act)
        assert torch.equal(y, first)
