#!/usr/bin/env python
"""Per-endpoint error of one training-mode forward pass against the fp64 oracle (where does a loss error come from?).
    python tools/parity_debug.py [B T C]        env: XV_GEMM_CG=1|2, XV_EPI_STATS=0|1"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import xvector_oracle as O
from tf_kaldi_speaker_b200.misc.utils import ParamsPlain
from tf_kaldi_speaker_b200.model.trainer import Trainer

B, T, C = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 64, 300)
D = 30
loss_type = "additive_angular_margin_softmax"
pd = dict(seed=0, network_type="tdnn", last_layer_no_bn=False, last_layer_linear=True, feature_norm=True,
          feature_scaling_factor=64, pooling_type="statistics_pooling", embedding_node="tdnn6_dense",
          weight_l2_regularizer=1e-2, batchnorm_momentum=0.99, clip_gradient=False, arcsoftmax_m=0.2,
          arcsoftmax_lambda_min=0, arcsoftmax_lambda_base=1000, arcsoftmax_lambda_gamma=1e-5, arcsoftmax_lambda_power=5)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 1, D, generator=g) + (0.5 + torch.rand(B, 1, D, generator=g)) * torch.randn(B, T, D, generator=g)
y = torch.randint(0, C, (B,), generator=g, dtype=torch.int32)
po = O.ParamsPlain(**dict(pd))
P = O.init_params(D, po, C, loss_type, seed=0)
with torch.no_grad():
    loss_o, total_o, ep_o = O.forward_loss(P, x.double(), y, po, loss_type, 1000, is_training=True, updates={})
    loss_e, _, ep_e = O.forward_loss(P, x.double(), y, po, loss_type, 1000, is_training=True, updates={}, emulate_bf16=True)
tr = Trainer(ParamsPlain(**dict(pd)), "/tmp/xv_dbg_model")
tr.build("train", D, loss_type, C)
tr.engine.epilogue_stats = os.environ.get("XV_EPI_STATS", "1") != "0"
tr.engine.store.load_tf({k: v.numpy() for k, v in P.items()})
l, _ = tr.forward_backward(x, y, 1000)
torch.cuda.synchronize()
lc = float(l.item())
print("loss cuda %.6f oracle %.6f rel %.2e | emulated-bf16 oracle rel %.2e" %
      (lc, loss_o.item(), abs(lc - loss_o.item()) / loss_o.item(), abs(loss_e.item() - loss_o.item()) / loss_o.item()))
for name in ("tdnn1_relu", "tdnn2_relu", "tdnn3_relu", "tdnn4_relu", "tdnn5_bn", "pooling", "tdnn6_dense", "tdnn6_relu",
             "tdnn7_dense", "tdnn7_bn"):
    ref = ep_o[name].numpy()
    emu = ep_e[name].numpy()
    node = tr.endpoints[name]
    if name == "tdnn5_bn":
        yv = node.dense().double().cpu().numpy()            # pre-BN (bias added back); compare BN output
        sc, sh = (t.double().cpu().numpy() for t in node.affine)
        bias = tr.engine.store.view("tdnn/tdnn5_dense/bias").double().cpu().numpy()
        got = (yv - bias[:yv.shape[-1]]) * sc[:yv.shape[-1]] + sh[:yv.shape[-1]]
    else:
        got = node.dense().double().cpu().numpy()
    e1 = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    e2 = np.linalg.norm(emu - ref) / np.linalg.norm(ref)
    print("%-12s rel-fro cuda vs fp64 %.3e | emulated-bf16 oracle vs fp64 %.3e | mean abs diff of column means %.3e"
          % (name, e1, e2, np.abs(got.reshape(-1, got.shape[-1]).mean(0) - ref.reshape(-1, ref.shape[-1]).mean(0)).mean()))
