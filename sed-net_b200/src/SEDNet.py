"""SEDNet / DGCNNEncoderGn with the reference's constructor signature, attribute names and state_dict layout
(reference src/SEDNet.py:19-98, :216-342), whose forward() runs on the fused sm_100a kernels.

The torch.nn layers below are parameter containers only (so that ``load_state_dict`` of reference checkpoints works
unchanged, including the aliased ``encoder.convK.1`` GroupNorms, the unused ``encoder.bn4/bn5`` and the
``pos_enc.inv_freq`` buffer); no torch operator runs in forward().
"""
import torch
from torch import nn

from . import _lib


class _InvFreq(nn.Module):
    """Stand-in for positional_encodings.PositionalEncoding1D(256): only its ``inv_freq`` buffer exists in
    checkpoints (src/SEDNet.py:285); the module is never called by forward()."""

    def __init__(self, channels):
        super().__init__()
        self.register_buffer("inv_freq", 1.0 / (10000 ** (torch.arange(0, channels, 2).float() / channels)))


class DGCNNEncoderGn(nn.Module):
    def __init__(self, mode=0, input_channels=3, nn_nb=80, normal_metric_W=1.):
        super().__init__()
        if mode != 5:
            raise NotImplementedError("the B200 encoder implements mode 5 (xyz + normals), the inference driver's mode")
        self.k, self.mode, self.input_channels, self.normal_metric_W = nn_nb, mode, input_channels, normal_metric_W
        self.bn1, self.bn2, self.bn3 = nn.GroupNorm(2, 64), nn.GroupNorm(2, 64), nn.GroupNorm(2, 128)
        self.bn4, self.bn5 = nn.GroupNorm(4, 256), nn.GroupNorm(8, 1024)
        act = lambda: nn.LeakyReLU(negative_slope=0.2)
        self.conv1 = nn.Sequential(nn.Conv2d(input_channels * 2, 64, kernel_size=1, bias=False), self.bn1, act())
        self.conv2 = nn.Sequential(nn.Conv2d(128, 64, kernel_size=1, bias=False), self.bn2, act())
        self.conv3 = nn.Sequential(nn.Conv2d(128, 128, kernel_size=1, bias=False), self.bn3, act())
        self.mlp1 = nn.Conv1d(256, 1024, 1)
        self.bnmlp1 = nn.GroupNorm(8, 1024)

        self._ws = None

    @torch.no_grad()
    def forward(self, x):
        """src/SEDNet.py:78-98 (mode 5): (B,6,N) -> x4 (B,1024), x_features (B,256,N)."""
        x = _lib.require_cuda(x, name="x")
        B, Cc, N = x.shape
        if Cc != 6:
            raise RuntimeError("DGCNNEncoderGn (mode 5) expects an input of shape (B, 6, N)")
        dev = x.device
        table, keep = _lib.param_table(dict(self.named_parameters()), dev, prefix="encoder.", count=13)
        need = _lib.load().sed_sednet_workspace_bytes(B, N, self.k)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        x4 = torch.empty((B, 1024), dtype=torch.float32, device=dev)
        feats = torch.empty((B, 256, N), dtype=torch.float32, device=dev)
        _lib.call("sed_encoder_forward", table, _lib.ptr(x), B, N, self.k, float(self.normal_metric_W), _lib.ptr(x4),
                  _lib.ptr(feats), _lib.ptr(self._ws), need, _lib.stream())
        del keep
        return x4, feats


class SEDNet(nn.Module):
    def __init__(self, emb_size=50, num_primitives=8, primitives=False, embedding=False, mode=0, num_channels=3,
                 loss_function=None, nn_nb=80, combine_label_prim=False, edge_module=False, late_fusion=False,
                 w_pos_enc=0.2, normal_metric_W=1., predict_normal=False):
        super().__init__()
        if not (mode == 5 and primitives and embedding and combine_label_prim and edge_module and late_fusion) \
                or predict_normal or num_channels != 6:
            raise NotImplementedError(
                "the B200 build implements the inference driver's configuration (generate_predictions_aug.py:142-154): "
                "mode=5, num_channels=6, primitives, embedding, combine_label_prim, edge_module, late_fusion")
        if num_primitives + 2 != 8:
            raise NotImplementedError("prim_encoding is Conv1d(8, 256): num_primitives must be 6 (src/SEDNet.py:287)")
        self.mode, self.loss_function, self.w_pos_enc = mode, loss_function, w_pos_enc
        self.emb_size, self.num_primitives = emb_size, num_primitives
        self.primitives, self.embedding = primitives, embedding
        self.combine_label_prim, self.late_fusion, self.predict_normal = combine_label_prim, late_fusion, predict_normal
        self.encoder = DGCNNEncoderGn(mode=mode, input_channels=num_channels, nn_nb=nn_nb, normal_metric_W=normal_metric_W)
        self.conv1 = nn.Conv1d(1024 + 256, 512, 1)
        self.bn1 = nn.GroupNorm(8, 512)
        self.conv2 = nn.Conv1d(512, 256, 1)
        self.bn2 = nn.GroupNorm(4, 256)
        self.edge_module = nn.Sequential(nn.Conv1d(256, 128, 1), nn.GroupNorm(4, 128), nn.Conv1d(128, 2, 1))
        self.asis = nn.Sequential(nn.Conv1d(256, 256, 1), nn.GroupNorm(4, 256), nn.ReLU(True), nn.Dropout(0.0))
        self.mlp_seg_prob1 = nn.Conv1d(256, 256, 1)
        self.mlp_seg_prob2 = nn.Conv1d(256, emb_size, 1)
        self.bn_seg_prob1 = nn.GroupNorm(4, 256)
        self.mlp_prim_prob1 = nn.Conv1d(256, 256, 1)
        self.mlp_prim_prob2 = nn.Conv1d(256, num_primitives, 1)
        self.bn_prim_prob1 = nn.GroupNorm(4, 256)
        self.pos_enc = _InvFreq(256)
        self.prim_encoding = nn.Sequential(nn.Conv1d(8, 256, 1), nn.ReLU())
        self._ws = None

    def _run(self, points, want_encoder):
        points = _lib.require_cuda(points, name="points")
        B, Cc, N = points.shape
        if Cc != 6:
            raise RuntimeError("SEDNet (mode 5) expects points of shape (B, 6, N)")
        dev = points.device
        state = dict(self.named_parameters())
        table, keep = _lib.param_table(state, dev)
        k = self.encoder.k
        need = _lib.load().sed_sednet_workspace_bytes(B, N, k)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        emb = torch.empty((B, self.emb_size, N), dtype=torch.float32, device=dev)
        logp = torch.empty((B, self.num_primitives, N), dtype=torch.float32, device=dev)
        edges = torch.empty((B, 2, N), dtype=torch.float32, device=dev)
        x4 = torch.empty((B, 1024), dtype=torch.float32, device=dev) if want_encoder else None
        feats = torch.empty((B, 256, N), dtype=torch.float32, device=dev) if want_encoder else None
        _lib.call("sed_sednet_forward", table, _lib.ptr(points), B, N, k, float(self.encoder.normal_metric_W),
                  float(self.w_pos_enc), self.emb_size, self.num_primitives, _lib.ptr(emb), _lib.ptr(logp),
                  _lib.ptr(edges), _lib.ptr(x4), _lib.ptr(feats), _lib.ptr(self._ws), need, _lib.stream())
        del keep
        return emb, logp, edges, x4, feats

    @torch.no_grad()
    def encode(self, points):
        """DGCNNEncoderGn.forward (src/SEDNet.py:78-98): returns (x4 (B,1024), x_features (B,256,N))."""
        _, _, _, x4, feats = self._run(points, True)
        return x4, feats

    @torch.no_grad()
    def forward(self, points, labels=None, compute_loss=False):
        """src/SEDNet.py:292-342 -> [embedding (B,E,N), log_prob (B,P,N), embed_loss (1,), edges (B,2,N)]."""
        if compute_loss:
            raise NotImplementedError("training losses are outside the inference hot path")
        emb, logp, edges, _, _ = self._run(points, False)
        return [emb, logp, torch.zeros(1, device=emb.device), edges]
