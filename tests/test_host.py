"""CPU: host-side logic -- state_dict compatibility, synthetic data determinism, sharding, error behaviour."""
import numpy as np
import pytest
import torch

from sednet_b200 import shard, synth
from sednet_b200.src import SEDNet, _lib


def _model(k=64):
    return SEDNet.SEDNet(embedding=True, emb_size=128, primitives=True, num_primitives=6, mode=5, num_channels=6,
                         combine_label_prim=True, edge_module=True, late_fusion=True, nn_nb=k)


def test_state_dict_layout_matches_reference():
    sd = _model().state_dict()
    assert set(sd) == set(synth.STATE_SHAPES) and len(sd) == 56
    for k, v in sd.items():
        assert tuple(v.shape) == synth.STATE_SHAPES[k], k
    assert sum(p.numel() for p in _model().parameters()) == 1351432


def test_load_state_dict_and_param_table_order():
    m = _model()
    sd = {k: torch.from_numpy(v) for k, v in synth.make_state_dict(3, randomize_gn=True).items()}
    m.load_state_dict(sd)  # strict
    assert len(_lib.PARAM_KEYS) == 45
    named = dict(m.named_parameters())
    for k in _lib.PARAM_KEYS:
        assert torch.equal(named[k], sd[k]), k


def test_unsupported_configuration_raises():
    with pytest.raises(NotImplementedError):
        SEDNet.SEDNet(mode=0)


def test_forward_without_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        _model()(torch.zeros(1, 6, 128))


def test_synth_is_deterministic_and_normalised():
    a, b = synth.make_cloud(7, 3000), synth.make_cloud(7, 3000)
    for x, y in zip(a[:4], b[:4]):
        assert np.array_equal(x, y)
    p, n = a[0], a[1]
    assert p.shape == (3000, 3) and abs(np.linalg.norm(n, axis=1) - 1).max() < 1e-5
    ext = np.max(p.max(0) - p.min(0))   # unit extent before the PCA rotation, so <= sqrt(3) after it
    assert 0.5 < ext < 1.75 and abs(p.mean(0)).max() < 0.05
    assert min(np.bincount(a[2])) >= 200


def test_shard_range_partitions():
    for total, world in ((512, 8), (10, 4), (3, 8), (64, 1)):
        spans = [shard.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_records_single_process():
    B, N, S = 3, 50, 4
    labels = torch.randint(0, S, (B, N))
    status = torch.tensor([[0, 1, 0, 2], [1, 1, 1, 1], [0, 0, 0, 0]], dtype=torch.int32)
    rec = shard.make_records(torch.arange(B), torch.tensor([2, 1, 4]), status, torch.ones(B, S), torch.ones(B), labels)
    assert rec.shape == (B, len(shard.RECORD_FIELDS))
    assert rec[:, 2].tolist() == [3, 0, 4]
    out = shard.gather_records(rec, 5)
    assert out.shape == (B, len(shard.RECORD_FIELDS)) and out[:, 0].tolist() == [0, 1, 2]
