"""CPU test of the host loop of lossAndGradient (faster-rcnn.torch_b200/objective.py, objective.lua:45-201) with a stub
model in place of the CUDA context: grouping of frames by size, cleanAnchors, use of pre-marshalled records, the dropout
seed of every frame (step, rank, frame index), counters and gradient:div(cls_count)."""
import torch


class _Net:
    def __init__(self):
        self.train = False

    def training(self):
        self.train = True


class _Anchor:
    def __init__(self, layer, y, x):
        self.layer, self.aspect, self.index = layer, 1, (1, y, x)


class _StubModel:
    host_only = True        # keeps create_objective off the NCCL path

    def __init__(self):
        self.gradient = torch.zeros(6)
        self.pnet, self.cnet = _Net(), _Net()
        self.calls = []

    def zero_grad(self):
        self.gradient.zero_()

    def output_dims(self, h, w):
        return [(18, h // 4, w // 4), (18, h // 8, w // 8), (18, h // 16, w // 16), (18, h // 16, w // 16), (384, h // 16, w // 16)]

    def train_batch(self, imgs, positives, negatives, seeds=None, packed=None):
        self.calls.append(dict(n=len(imgs), shape=tuple(imgs[0].shape), seeds=list(seeds), packed=packed,
                               counts=[(len(p), len(q)) for p, q in zip(positives, negatives)]))
        self.gradient += float(len(imgs))
        return [dict(cls=1.0 * len(p), reg=2.0 * len(p), creg=3.0, ccls=0.5) for p in positives]

    def train_image(self, img, positives, negatives, seed=0):
        return self.train_batch([img], [positives], [negatives], seeds=[seed])[0]


def _frame(h, w, n_pos, n_neg, outside=0):
    pos = [(_Anchor(1, 1 + i, 1), dict(rect=None, class_index=1)) for i in range(n_pos)]
    pos += [(_Anchor(1, h, w), dict(rect=None, class_index=1)) for _ in range(outside)]   # beyond the h/4 x w/4 map
    neg = [(_Anchor(2, 1, 1 + i),) for i in range(n_neg)]
    return dict(img=torch.zeros(3, h, w), positive=pos, negative=neg, packed=("P%d" % n_pos, "N%d" % n_neg))


def test_host_loop_groups_cleans_seeds_and_counts(F):
    m = _StubModel()
    batch = [_frame(64, 96, 3, 2), _frame(32, 48, 1, 1), _frame(64, 96, 2, 4, outside=1)]
    obj = F.create_objective(m, None, rank=2)
    loss, grad, stats = obj(batch, seed=5)
    assert m.pnet.train and m.cnet.train
    # two size groups, in first-appearance order; the 64x96 frames share one call
    assert [c["n"] for c in m.calls] == [2, 1]
    assert m.calls[0]["shape"] == (3, 64, 96) and m.calls[1]["shape"] == (3, 32, 48)
    # seed = step seed * 1000003 + rank * len(batch) + frame index in the batch
    base = 5 * 1000003 + 2 * 3
    assert m.calls[0]["seeds"] == [base + 0, base + 2] and m.calls[1]["seeds"] == [base + 1]
    # cleanAnchors dropped the out-of-map positive of frame 2: its pre-marshalled records no longer line up, so the
    # group falls back to marshalling (packed=None); the single-frame group keeps its records
    assert m.calls[0]["counts"] == [(3, 2), (2, 4)] and m.calls[0]["packed"] is None
    assert m.calls[1]["packed"] == [("P1", "N1")]
    # counters: cls_count = all listed anchors, reg_count = positives, one detection-stage mean per frame
    assert stats["cls_count"] == 3 + 2 + 1 + 1 + 2 + 4 and stats["reg_count"] == 6
    assert abs(stats["pcls"] - 6.0 / 13) < 1e-6 and abs(stats["preg"] - 12.0 / 6) < 1e-6
    assert abs(stats["dreg"] - 9.0 / 6) < 1e-6 and abs(stats["dcls"] - 0.5) < 1e-6
    assert abs(loss - (stats["pcls"] + stats["preg"])) < 1e-9
    # gradient:zero() at the start, gradient:div(cls_count) at the end (objective.lua:49,200)
    assert torch.allclose(grad, torch.full((6,), 3.0 / 13))
    # a second call zeroes the gradient again and, without an explicit seed, draws a new one from the step counter
    m.calls.clear()
    obj(batch)
    assert torch.allclose(m.gradient, torch.full((6,), 3.0 / 13))
    assert m.calls[0]["seeds"][0] == 2 * 1000003 + 2 * 3


def test_host_loop_deferred_division_and_default_rank(F):
    m = _StubModel()
    obj = F.create_objective(m, None, defer_div=True)
    _, grad, stats = obj([_frame(64, 96, 2, 2)], seed=1)
    assert torch.allclose(grad, torch.ones(6)) and stats["deferred_div"] == 4.0      # the optimiser pass divides
    assert m.calls[0]["seeds"] == [1000003]                                           # rank 0 without a process group
    assert m.calls[0]["packed"] == [("P2", "N2")]


def _dp_worker(rank, world, port, out):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import frcnn_b200 as F
    m = _StubModel()
    # this rank's share of the batch: rank 0 two frames, rank 1 two frames with other example counts
    batch = [_frame(64, 96, 1 + rank, 2), _frame(64, 96, 2, 1 + 2 * rank)]
    obj = F.create_objective(m, dist)
    loss, grad, stats = obj(batch, seed=3)
    torch.save(dict(loss=loss, grad=grad.clone(), stats=stats, seeds=m.calls[0]["seeds"]), out % rank)
    dist.barrier()
    dist.destroy_process_group()


def test_host_loop_two_ranks_gloo(tmp_path):
    """World size 2 on CPU (gloo): every rank runs its share, ONE all-reduce carries the flat gradient and the seven
    counters, both ranks end with the same gradient divided by the GLOBAL cls_count, and the dropout seeds are disjoint
    across ranks (rank * len(batch) + frame index)."""
    import os
    import torch.multiprocessing as mp
    out = str(tmp_path / "r%d.pt")
    mp.spawn(_dp_worker, args=(2, 33100 + os.getpid() % 2000, out), nprocs=2, join=True)
    r0, r1 = torch.load(out % 0, weights_only=False), torch.load(out % 1, weights_only=False)
    # stub gradient: +2 per rank (one call of two frames) -> 4 after the sum; cls_count = (1+2 + 2+1) + (2+2 + 2+3) = 15
    assert r0["stats"]["cls_count"] == 15 and r1["stats"]["cls_count"] == 15
    assert r0["stats"]["reg_count"] == 3 + 4
    assert torch.allclose(r0["grad"], torch.full((6,), 4.0 / 15)) and torch.equal(r0["grad"], r1["grad"])
    assert r0["loss"] == r1["loss"]
    base = 3 * 1000003
    assert r0["seeds"] == [base + 0, base + 1] and r1["seeds"] == [base + 2, base + 3]
