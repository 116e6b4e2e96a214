"""Torch7 .t7 snapshot compatibility (utilities.lua:113-134, main.lua:94-98,145-148).

The reference can neither run nor produce a snapshot here (no Lua / Torch7), and it ships no .t7 fixture, so the
reader / writer are pinned by a HAND-WRITTEN stream, tests/golden/snapshot_ascii.t7: the byte sequence
`torch.DiskFile('f','w'):writeObject{version=0, weights=CudaTensor{1.5,-2,0.25,1e-5,FLT_MAX}, options={lr=1e-4,
name='imgnet', seed=false}, stats={pcls={log 2, 0.5}, preg={}}}` produces according to torch7's File.lua /
generic/Tensor.c / generic/Storage.c / THDiskFile.c (grammar restated in t7.py's docstring), table fields in this
order."""
import io
import os

import numpy as np
import pytest

import frcnn_b200 as F

t7 = F.t7
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "snapshot_ascii.t7")


def test_reads_hand_written_reference_stream():
    obj = t7.load(GOLDEN)
    assert set(obj) == {"version", "weights", "options", "stats"}
    assert obj["version"] == 0
    w = obj["weights"]
    assert w.typename == "torch.CudaTensor" and w.dtype == np.float32 and w.shape == (5,)
    assert np.array_equal(np.asarray(w), np.array([1.5, -2, 0.25, 1e-5, np.finfo(np.float32).max], dtype=np.float32))
    assert obj["options"] == {"lr": 1e-4, "name": "imgnet", "seed": False}
    assert obj["stats"] == {"pcls": {1: float(np.log(2.0)), 2: 0.5}, "preg": {}}
    w2, opt, stats, bn = t7.load_model(GOLDEN)
    assert np.array_equal(w2, np.asarray(w)) and opt["name"] == "imgnet" and stats["preg"] == {} and bn is None


def test_writer_reproduces_the_stream_byte_for_byte():
    obj = t7.load(GOLDEN)
    assert t7.dumps(obj) == open(GOLDEN, "rb").read()


@pytest.mark.parametrize("ascii", [True, False])
def test_round_trip(ascii, tmp_path):
    rng = np.random.default_rng(0)
    w = rng.standard_normal(100_003).astype(np.float32)
    w[:4] = [np.float32(1) / 3, -0.0, np.finfo(np.float32).tiny, np.finfo(np.float32).max]
    shared = {"a": 1, "b": [1.5, "x", True, None]}
    obj = {"version": 0, "weights": w, "options": {"cfg": "config/imagenet.lua", "lr": 1e-4, "gpuid": 0, "restore": ""},
           "stats": {"pcls": [0.7, 0.6], "preg": {}, "m": rng.standard_normal((3, 4)), "ids": np.arange(5, dtype=np.int64)},
           "again": shared, "and_again": shared}
    path = str(tmp_path / "s.t7")
    t7.save(path, obj, ascii=ascii)
    back = t7.load(path)
    assert np.array_equal(np.asarray(back["weights"]), w)   # "%.9g" round-trips every float32 exactly
    assert back["weights"].typename == "torch.CudaTensor"   # float32 is what the reference flattens after :cuda()
    assert back["options"] == obj["options"]
    assert back["stats"]["pcls"] == {1: 0.7, 2: 0.6} and back["stats"]["preg"] == {}
    assert np.array_equal(np.asarray(back["stats"]["m"]), obj["stats"]["m"]) and back["stats"]["m"].typename == "torch.DoubleTensor"
    assert np.array_equal(np.asarray(back["stats"]["ids"]), np.arange(5)) and back["stats"]["ids"].typename == "torch.LongTensor"
    assert back["again"] is back["and_again"]                # shared references survive (File.lua object indices)
    assert back["again"]["b"] == {1: 1.5, 2: "x", 3: True, 4: None}   # a Python list is written as a 1-based Lua table


def test_round_trip_keeps_nil_values_out():
    back = t7.load(t7.dumps({"k": None, "v": 2}))
    assert back == {"k": None, "v": 2}


def test_non_contiguous_and_strided_tensor_bodies():
    """Torch writes size / stride / storageOffset and the WHOLE storage: a transposed view must come back as its values."""
    buf = io.BytesIO()
    w = t7._Writer(buf, ascii=True, cuda=False)
    # hand-built: a 2x3 FloatTensor viewing storage {0..7} with stride (1, 2) and storageOffset 2 (1-based)
    w.int(t7.TYPE_TORCH); w.int(1); w.string("V 1"); w.string("torch.FloatTensor")
    w.int(2); w.array(np.array([2, 3], dtype=np.int64), "%d"); w.array(np.array([1, 2], dtype=np.int64), "%d"); w.long(2)
    w.int(t7.TYPE_TORCH); w.int(2); w.string("V 1"); w.string("torch.FloatStorage"); w.long(8)
    w.array(np.arange(8, dtype=np.float32), "%.9g")
    t = t7.load(buf.getvalue())
    assert np.array_equal(np.asarray(t), np.array([[1, 3, 5], [2, 4, 6]], dtype=np.float32))


def test_snapshot_flat_layout_excludes_bn_running_stats():
    """`weights` = nn.Module.flatten over pnet:parameters() + cnet:parameters() (utilities.lua:136-147): learnable tensors
    only -- BatchNormalization running_mean / running_var are not parameters in Torch and are not part of the vector."""
    m = F.vgg_small(F.duplo_cfg, device=-1)  # host-only plan
    names = m.learnable_names()
    assert not any(n.endswith((".bn_mean", ".bn_var")) for n in names)
    assert [n for n in m.param_names if n not in names] == ["fc1.bn_mean", "fc1.bn_var"]
    numel = dict(zip(m.param_names, m.param_numel))
    # 26.78 M learnable parameters: conv trunk + heads + cnet
    assert sum(numel[n] for n in names) == sum(m.param_numel) - 2 * 1024
    assert names[0] == "b1_c1.weight" and names[-1] == "cls.bias"
    m.close()
