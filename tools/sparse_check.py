"""Per-parameter relative L2 between the sparse and the dense anchor-head backward, with dense-vs-dense as noise floor."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import frcnn_b200 as F
from oracle import model as OM, anchors as OA, objective as OO

h, w, counts = 122, 192, [(10, 14), (0, 0), (3, 30)]
cfg = OM.CFG_DUPLO
m = F.vgg_small(F.duplo_cfg)
m.load_params(OM.init_params(OM.VGG_SMALL, cfg, seed=0, randomize_aux=True))
dims = m.output_dims(h, w)
oa = OA.Anchors(OM.VGG_SMALL["layers"], OM.VGG_SMALL["anchor_nets"], cfg["scales"])
frames, P, Q = [], [], []
for s, (np_, nn_) in enumerate(counts):
    pos, neg, _ = OO.synthetic_examples(oa, dims, w, h, max(np_, 1), max(nn_, 1), 3, cfg["class_count"], seed=70 + s)
    pos, neg = OO.clean_anchors(pos, dims)[:np_], OO.clean_anchors(neg, dims)[:nn_]
    frames.append(OM.synthetic_frame(h, w, seed=80 + s).cuda()); P.append(pos); Q.append(neg)
seeds = [3, 4, 5]
saved = m.weights.clone()
m.pnet.training(); m.cnet.training()
def run(mode):
    os.environ["FRCNN_HEAD_SPARSE"] = mode
    m.weights.copy_(saved); m.pack_weights(); m.zero_grad()
    m.train_batch(frames, P, Q, seeds=seeds)
    return m.gradient.clone()
gs, gd, gd2 = run("1"), run("0"), run("0")
off = 0
print("%-16s %10s %10s" % ("param", "sparse/dense", "dense/dense"))
for name, numel in zip(m.param_names, m.param_numel):
    a, b, c = gs[off:off + numel], gd[off:off + numel], gd2[off:off + numel]
    off += numel
    if name.endswith("weight"):
        n = b.norm().item() + 1e-30
        print("%-16s %10.2e %10.2e" % (name, (a - b).norm().item() / n, (c - b).norm().item() / n))
print("all", ((gs - gd).norm() / gd.norm()).item(), ((gd2 - gd).norm() / gd.norm()).item())
