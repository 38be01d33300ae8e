"""CPU: the drop-in boundary proven on the reference driver's OWN source lines.  With the alias of INTEGRATION.md section 1
installed (sys.modules["src.<name>"] -> sednet_b200.src.<name>), the import block, the guard function, the model
construction and the checkpoint-loading lines of /root/reference/generate_predictions_aug.py are read from the file and
executed unchanged (line ranges below); a checkpoint written by the UNMODIFIED reference SEDNet (with DataParallel's
"module." prefix) must load strictly into the replacement.  Needs /root/reference: skipped on the GPU box."""
import importlib
import os
import sys
import types

import pytest
import torch

import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference is not present on this machine")

DRIVER = os.path.join(ref_shim.REFERENCE_ROOT, "generate_predictions_aug.py")
ALIASED = ("SEDNet", "PointNet", "mean_shift", "primitive_forward", "fitting_utils", "primitives", "segment_utils",
           "smooth_normal_matrix", "fitting_optimization", "utils", "guard")


def _lines(lo, hi):
    with open(DRIVER) as f:
        src = f.readlines()
    return "".join(src[lo - 1:hi])


@pytest.fixture()
def aliased():
    """Reference modules imported first (to write the checkpoint), then the alias in front of them."""
    ref = ref_shim.load()
    ref_sednet = ref.SEDNet
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    import sednet_b200.src  # noqa: F401
    for name in ALIASED:
        sys.modules[f"src.{name}"] = importlib.import_module(f"sednet_b200.src.{name}")
    yield ref_sednet
    for name in ALIASED:
        sys.modules.pop(f"src.{name}", None)
    sys.modules.update(saved)


def test_driver_lines_run_on_the_replacement(aliased, tmp_path):
    ref_sednet = aliased
    # a checkpoint as the reference's training script leaves it: DataParallel prefix, the reference's own module
    torch.manual_seed(3)
    ref_model = ref_sednet.SEDNet(embedding=True, emb_size=128, primitives=True, num_primitives=6, mode=5, num_channels=6,
                                  combine_label_prim=True, edge_module=True, late_fusion=True, nn_nb=64)
    ckpt = tmp_path / "ckpt.pth"
    torch.save({"module." + k: v for k, v in ref_model.state_dict().items()}, ckpt)

    ns = {"__name__": "driver_under_test", "__file__": DRIVER, "torch": torch,
          "config": types.SimpleNamespace(pretrain_model_path=str(ckpt), pretrain_model_type_path=str(ckpt), num_train=0,
                                          num_val=0, num_test=1),
          "my_knn": 64}
    exec(_lines(1, 5), ns)          # import sys, logging, json, os
    exec(_lines(12, 21), ns)        # numpy, gen_test_vis, src.dataset_segments, src.smooth_normal_matrix.hpnet_process
    exec(_lines(25, 35), ns)        # def guard_mean_shift(ms, embedding, quantile, iterations, kernel_type)
    exec(_lines(40, 55), ns)        # sys.path, torch, src.SEDNet, src.segment_loss, src.segment_utils, src.mean_shift
    import sednet_b200.src.SEDNet as fast_sednet
    import sednet_b200.src.mean_shift as fast_ms
    import sednet_b200.src.segment_utils as fast_su
    import sednet_b200.src.smooth_normal_matrix as fast_snm
    assert ns["SEDNet"] is fast_sednet.SEDNet and ns["MeanShift"] is fast_ms.MeanShift
    assert ns["to_one_hot"] is fast_su.to_one_hot and ns["SIOU_matched_segments"] is fast_su.SIOU_matched_segments
    assert ns["SIOU_matched_segments_usecd"] is fast_su.SIOU_matched_segments_usecd
    assert ns["compute_type_miou_abc"] is fast_su.compute_type_miou_abc
    assert ns["hpnet_process"] is fast_snm.hpnet_process
    assert ns["EmbeddingLoss"].__module__ == "src.segment_loss"         # the reference's own module stays the reference's
    exec(_lines(139, 140), ns)      # userspace, Loss = EmbeddingLoss(margin=1.0)
    exec(_lines(142, 173), ns)      # model = SEDNet(...), model_inst = SEDNet(...), .cuda(), split_dict, ms = MeanShift()
    assert isinstance(ns["model"], fast_sednet.SEDNet) and isinstance(ns["model_inst"], fast_sednet.SEDNet)
    assert isinstance(ns["ms"], fast_ms.MeanShift)
    ns["model"].eval(); ns["model_inst"].eval()
    exec(_lines(190, 198), ns)      # state_dict = torch.load(...); strip "module."; load_state_dict (strict), twice
    got, want = ns["model"].state_dict(), ref_model.state_dict()
    assert list(got.keys()) == list(want.keys()) and len(got) == 56
    for k in want:
        assert torch.equal(got[k], want[k]), k
    assert sum(p.numel() for p in ns["model"].parameters()) == sum(p.numel() for p in ref_model.parameters()) == 1351432
    # the hot-path call needs the GPU: on a CPU tensor the replacement raises instead of silently computing on the CPU
    with pytest.raises(RuntimeError):
        ns["model"](torch.zeros((1, 6, 128)), None, False)
    with pytest.raises(RuntimeError):
        ns["guard_mean_shift"](ns["ms"], torch.nn.functional.normalize(torch.randn(200, 128), dim=1), 0.015, 5)
