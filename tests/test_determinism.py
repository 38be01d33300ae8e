"""Run-to-run determinism of the tcgen05 kernels at BASELINE.json's full sizes: every kernel of the hot path is launched
repeatedly on the same inputs and must return the same bits each time.  None of the kernels uses floating-point atomics or
an order-dependent reduction, so any difference is a synchronisation bug -- this is the test that would have caught the
barrier-phase hazard of pw_tc2_kernel (wrong tiles in two launches out of five, only at 16-32 clouds x 10 000 points,
invisible to compute-sanitizer because it was timing-dependent)."""
import numpy as np
import pytest
import torch

from sednet_b200 import synth
from util import t

pytestmark = pytest.mark.gpu
N = 10000
REPS = 8


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    from sednet_b200.src import _lib
    _lib.load()
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def clouds(dev):
    B = 16
    pts, nrm, lab, _ = synth.make_batch(B, N, seed0=4242)
    x6 = t(np.concatenate([pts, nrm], 2).transpose(0, 2, 1).copy()).to(dev)
    emb = torch.stack([t(synth.make_embedding(lab[b], 128, 0.02, 70 + b)) for b in range(B)]).to(dev)
    return x6, emb


def _same(outs):
    first = outs[0]
    for o in outs[1:]:
        for a, b in zip(first, o):
            if not torch.equal(a, b):
                return False
    return True


def test_knn_is_deterministic(dev, clouds):
    from sednet_b200.src import PointNet
    x6, _ = clouds
    feats = torch.randn(16, 64, N, device=dev, generator=torch.Generator(device="cuda").manual_seed(3))
    for k in (20, 64):
        assert _same([(PointNet.knn_points_normals(x6, k, k, 1.0),) for _ in range(REPS)])
        assert _same([(PointNet.knn(feats, k, k),) for _ in range(REPS)])


def test_forward_is_deterministic(dev, clouds):
    """Both networks' forward at 16 clouds (the batch size at which the pointwise hazard showed) -- every output tensor."""
    from sednet_b200.src import SEDNet
    x6, _ = clouds
    sd = synth.make_state_dict(1, randomize_gn=True)
    m = SEDNet.SEDNet(embedding=True, emb_size=128, primitives=True, num_primitives=6, mode=5, num_channels=6,
                      combine_label_prim=True, edge_module=True, late_fusion=True, nn_nb=64)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m = m.to(dev).eval()
    outs = []
    for _ in range(REPS):
        o = m(x6)
        outs.append((o[0].clone(), o[1].clone(), o[3].clone()))
    assert all(torch.isfinite(a).all() for a in outs[0])
    assert _same(outs)
    enc = [tuple(a.clone() for a in m.encode(x6)) for _ in range(4)]
    assert _same(enc)


@pytest.mark.parametrize("mode", [4, 1, 3])
def test_meanshift_is_deterministic(dev, clouds, mode):
    """Bandwidth, 50 iterations and nms on 8 clouds of planted embeddings, four times."""
    from sednet_b200.src.mean_shift import MeanShift
    _, emb = clouds
    ms = MeanShift(prec_mode=mode)
    outs = []
    for _ in range(4):
        res = []
        for b in range(8):
            np.random.seed(5)       # compute_bandwidth shuffles with NumPy's global RNG, as the reference does
            newX, center, bw, labels = ms.mean_shift(emb[b], 10000, 0.015, 50)
            res += [newX, center, torch.as_tensor(bw).reshape(1).to(dev).float(), labels]
        outs.append(tuple(res))
    assert _same(outs)


def test_pipeline_step_is_deterministic(dev, clouds):
    """The whole batched step (2 forwards + clustering + vote + fits) through the pipeline handle, default mode (4)."""
    from sednet_b200.pipeline import Pipeline
    x6, _ = clouds
    B = 8
    pts = x6[:B, :3].permute(0, 2, 1).contiguous()
    nrm = x6[:B, 3:].permute(0, 2, 1).contiguous()
    sd1, sd2 = synth.make_state_dict(1, randomize_gn=True), synth.make_state_dict(2, randomize_gn=True)
    pipe = Pipeline(B, N, k=64)
    pipe.set_weights(sd1, sd2)
    outs = []
    for _ in range(4):
        pipe.run_device(pts, nrm)
        torch.cuda.synchronize()
        outs.append(tuple(pipe.device_tensor(name)[:B].clone() for name in ("X", "labels", "pred_type", "params", "residual", "bw")))
    assert _same(outs)


def test_pipeline_weights_can_be_replaced(dev, clouds):
    """The handle keeps FP16 hi / lo images of its 1x1-convolution weights across steps (split once per set_weights): a second
    set_weights must drop them -- the outputs then equal those of a fresh handle built with the new weights, bit for bit."""
    from sednet_b200.pipeline import Pipeline
    x6, _ = clouds
    B = 4
    pts = x6[:B, :3].permute(0, 2, 1).contiguous()
    nrm = x6[:B, 3:].permute(0, 2, 1).contiguous()
    sd = [synth.make_state_dict(s, randomize_gn=True) for s in (1, 2, 3, 4)]

    def outputs(pipe):
        pipe.run_forward(pts, nrm)
        torch.cuda.synchronize()
        return tuple(pipe.device_tensor(name)[:B].clone() for name in ("X", "log_prob", "type_log_prob"))

    pipe = Pipeline(B, N, k=64)
    pipe.set_weights(sd[0], sd[1])
    first = outputs(pipe)
    assert _same([first, outputs(pipe)])                 # second step: prepared weights reused
    pipe.set_weights(sd[2], sd[3])
    second = outputs(pipe)
    assert not torch.equal(first[0], second[0])
    fresh = Pipeline(B, N, k=64)
    fresh.set_weights(sd[2], sd[3])
    assert _same([second, outputs(fresh)])
    pipe.close(); fresh.close()
