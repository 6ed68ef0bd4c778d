import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from collections import OrderedDict
from oracle import xvector_oracle as O
from tests.xv_testlib import base_params, head_params, make_batch, rel_fro
from tf_kaldi_speaker_b200.misc.utils import ParamsPlain
from tf_kaldi_speaker_b200.model.trainer import Trainer

B, T, D, C = 12, 50, 30, 200
loss_type = "softmax"
pd = base_params(**head_params(loss_type)); pd.update(dict(last_layer_linear=False))
x, y = make_batch(B, T, D, C, seed=1)
po = O.ParamsPlain(**dict(pd))
P = O.init_params(D, po, C, loss_type, seed=3)
Pg = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in P.items())
loss, total, ep = O.forward_loss(Pg, x.double(), y, po, loss_type, 0, True, OrderedDict())
names = ["logits", "tdnn7_relu", "tdnn7_dense", "tdnn6_relu", "tdnn6_dense", "pooling", "tdnn5_relu", "tdnn5_dense", "tdnn4_relu", "tdnn4_dense", "tdnn3_relu", "tdnn3_conv"]
gs = torch.autograd.grad(loss, [ep[n] for n in names], allow_unused=True)
G = dict(zip(names, gs))

tr = Trainer(ParamsPlain(**dict(pd)), "/tmp/xv_diag")
tr.build("train", D, loss_type, C)
st = tr.engine.store
st.load_tf({k: v.numpy() for k, v in P.items()})
tr.forward_backward(x, y, 0)
torch.cuda.synchronize()
ws = {k[0]: v for k, v in tr.engine.ws.items()}
def cmp(tag, a, b):
    a = a.double().cpu().numpy(); b = b.detach().numpy()
    print("%-28s rel_fro %.4g   max|a| %.3g max|b| %.3g  mean(a-b) %.3g" % (tag, rel_fro(a, b), abs(a).max(), abs(b).max(), (a-b).mean()))
cmp("D (dlogits)", ws["head/d"][:, :C], G["logits"])
cmp("dx head (tdnn7/a/grad)", ws["tdnn7/a/grad"], G["tdnn7_relu"])
cmp("tdnn7/dy", ws["tdnn7/dy"], G["tdnn7_dense"])
cmp("tdnn6/a/grad", ws["tdnn6/a/grad"], G["tdnn6_relu"])
cmp("tdnn6/dy", ws["tdnn6/dy"], G["tdnn6_dense"])
Pp = 1500; cp = 1536
pg = ws["pool/grad"]; pg = torch.cat([pg[:, :Pp], pg[:, cp:cp+Pp]], 1)
cmp("pool/grad", pg, G["pooling"])
def fr(name, c, valid):
    t = ws[name]; return t.view(B, T, t.shape[1])[:, :valid, :c]
cmp("tdnn5/a/grad", fr("tdnn5/a/grad", 1500, 36), G["tdnn5_relu"])
cmp("tdnn5/dy", fr("tdnn5/dy", 1500, 36), G["tdnn5_dense"])
cmp("tdnn4/a/grad", fr("tdnn4/a/grad", 512, 36), G["tdnn4_relu"])
cmp("tdnn4/dy", fr("tdnn4/dy", 512, 36), G["tdnn4_dense"])
cmp("tdnn3/a/grad", fr("tdnn3/a/grad", 512, 36), G["tdnn3_relu"].squeeze(1) if G["tdnn3_relu"].dim()==4 else G["tdnn3_relu"])
cmp("tdnn3/dy", fr("tdnn3/dy", 512, 36), G["tdnn3_conv"])
# forward intermediates
cmp("fwd tdnn7/a", ws["tdnn7/a"], ep["tdnn7_relu"])
cmp("fwd pool", torch.cat([ws["pool/out"][:, :Pp], ws["pool/out"][:, cp:cp+Pp]], 1), ep["pooling"])
