"""Host mirror of models/model_utilities.lua, models/vgg_small.lua, models/vgg_large.lua and config/*.lua:
the same description tables, turned into a plan of the CUDA library (frcnn_model_plan) instead of nn modules.

The returned `Model` plays the role of the reference's model table {cfg, layers, pnet, cnet}
(model_utilities.lua:128-134): `model.pnet.forward(img)` and `model.cnet.forward(x)` keep the Lua call shapes."""
import math

import numpy as np
import torch

from ._lib import check, ffi, lib

# config/duplo.lua, config/imagenet.lua (the parameters the detection path reads)
duplo_cfg = dict(class_count=16, target_smaller_side=450, scales=[32, 64, 128, 256], max_pixel_size=1000,
                 roi_pooling=dict(kw=6, kh=6), batch_size=256, positive_threshold=0.5, negative_threshold=0.25,
                 best_match=True, nearby_aversion=True)
imgnet_cfg = dict(class_count=200, target_smaller_side=480, scales=[48, 96, 192, 384], max_pixel_size=1000,
                  roi_pooling=dict(kw=6, kh=6), batch_size=300, positive_threshold=0.6, negative_threshold=0.25,
                  best_match=True, nearby_aversion=True)


class _Net:
    """Stands in for an nn.gModule: forward()/evaluate()/training() with the reference's call shapes."""

    def __init__(self, fwd, bwd=None):
        self._fwd = fwd
        self._bwd = bwd
        self.train = False

    def forward(self, *a, **k):
        return self._fwd(*a, **k)

    __call__ = forward

    def evaluate(self):
        self.train = False

    def training(self):
        self.train = True

    def backward(self, *a, **k):
        if self._bwd is None:
            raise NotImplementedError("backward of this net is not built yet (SURVEY 8a row P3)")
        return self._bwd(*a, **k)


class Model:
    def __init__(self, cfg, layers, anchor_nets, class_layers, device=0, dropout_eval_scale=-1.0, stream=None):
        self.cfg, self.layers, self.anchor_nets, self.class_layers = cfg, layers, anchor_nets, class_layers
        self._ctor = dict(device=device, dropout_eval_scale=dropout_eval_scale)
        # a caller-supplied stream is not ordered with torch's default stream: the training calls, which return before the
        # backward pass has finished, are followed by a synchronize() then (the library's own stream is a blocking one,
        # which the legacy default stream waits for implicitly)
        self._foreign_stream = bool(stream)
        self.host_only = device == -1  # plan + Localizer / Anchors geometry only; no compute entry point works
        if not self.host_only and not torch.cuda.is_available():
            # fail loudly: the product has no CPU path
            raise RuntimeError("frcnn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = None if self.host_only else torch.device("cuda", device)
        L = lib()
        pctx = ffi.new("frcnn_ctx**")
        check(None, L.frcnn_create(pctx, device, ffi.cast("void*", stream or 0)))
        self.ctx = pctx[0]
        blocks = ffi.new("frcnn_block_desc[]", len(layers))
        for i, l in enumerate(layers):
            blocks[i].filters, blocks[i].kW, blocks[i].kH = l["filters"], l["kW"], l["kH"]
            blocks[i].padW, blocks[i].padH, blocks[i].conv_steps = l["padW"], l["padH"], l["conv_steps"]
            blocks[i].dropout = l.get("dropout") or 0.0
        heads = ffi.new("frcnn_head_desc[]", len(anchor_nets))
        for i, a in enumerate(anchor_nets):
            heads[i].kW, heads[i].n, heads[i].input = a["kW"], a["n"], a["input"]
        fcs = ffi.new("frcnn_fc_desc[]", len(class_layers))
        for i, l in enumerate(class_layers):
            fcs[i].n, fcs[i].dropout, fcs[i].batch_norm = l["n"], l.get("dropout") or 0.0, 1 if l.get("batch_norm") else 0
        scales = ffi.new("double[]", [float(s) for s in cfg["scales"]])
        check(self.ctx, L.frcnn_model_plan(self.ctx, blocks, len(layers), heads, len(anchor_nets), fcs, len(class_layers),
                                           cfg["class_count"], cfg["roi_pooling"]["kh"], cfg["roi_pooling"]["kw"], scales,
                                           len(cfg["scales"]), dropout_eval_scale))
        self.param_names, self.param_numel = [], []
        name = ffi.new("char[64]")
        numel = ffi.new("int64_t*")
        for i in range(L.frcnn_param_count(self.ctx)):
            check(self.ctx, L.frcnn_param_info(self.ctx, i, name, 64, numel))
            self.param_names.append(ffi.string(name).decode())
            self.param_numel.append(int(numel[0]))
        self.pnet = _Net(self._pnet_forward, self._pnet_backward)
        self.cnet = _Net(self._cnet_forward, self._cnet_backward)
        self.n_heads = len(anchor_nets)
        if self.host_only:
            return
        # one flat fp32 buffer, like nn.Module.flatten (utilities.lua:136-147); params are views into it
        self.weights = torch.zeros(sum(self.param_numel), dtype=torch.float32, device=self.device)
        self.params, off = {}, 0
        for n, k in zip(self.param_names, self.param_numel):
            self.params[n] = self.weights[off:off + k]
            off += k
        ptrs = ffi.new("const float*[]", [ffi.cast("const float*", self.params[n].data_ptr()) for n in self.param_names])
        check(self.ctx, L.frcnn_bind_params(self.ctx, ptrs, len(self.param_names)))
        # the flat gradient buffer of combine_and_flatten_parameters (utilities.lua:136-147), same layout as `weights`
        self.gradient = torch.zeros_like(self.weights)
        self.grads, off = {}, 0
        for n, k in zip(self.param_names, self.param_numel):
            self.grads[n] = self.gradient[off:off + k]
            off += k
        gptrs = ffi.new("float*[]", [ffi.cast("float*", self.grads[n].data_ptr()) for n in self.param_names])
        check(self.ctx, L.frcnn_bind_grads(self.ctx, gptrs, len(self.param_names)))
        nl = ffi.new("int*")
        check(self.ctx, L.frcnn_dropout_layers(self.ctx, ffi.NULL, 0, nl))
        ch = ffi.new("int[]", max(nl[0], 1))
        check(self.ctx, L.frcnn_dropout_layers(self.ctx, ch, nl[0], nl))
        self.dropout_channels = [ch[i] for i in range(nl[0])]

    def close(self):
        if getattr(self, "ctx", None) is not None:
            lib().frcnn_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- parameters -------------------------------------------------------------------------------------------
    def load_params(self, params):
        """params: dict name -> array/tensor in Torch layouts (conv [Cout][Cin][kH][kW], linear [out][in])."""
        for n in self.param_names:
            v = params[n]
            v = torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v, dtype=torch.float32)
            self.params[n].copy_(v.reshape(-1).to(self.device))
        self.pack_weights()

    # -- snapshots: utilities.lua:126-134 / main.lua:94-98 (Torch7 .t7 object streams, see t7.py) ---------------------
    def learnable_names(self):
        """Parameter tensors in `nn.Module.flatten` order as combine_and_flatten_parameters builds it (utilities.lua:136-147):
        pnet's then cnet's learnable parameters.  BatchNormalization running statistics are not parameters in Torch.
        (Inside pnet the order follows nngraph's topological sort of {block1..4, AnchorNetwork1..4} -- blocks, then heads,
        is one valid order; the un-vendored nngraph's tie-breaking is unpinned.)"""
        return [n for n in self.param_names if not n.endswith((".bn_mean", ".bn_var"))]

    def get_flat_parameters(self):
        """The reference's flat `weights` vector (learnable parameters only) as a float32 numpy array."""
        return torch.cat([self.params[n].reshape(-1) for n in self.learnable_names()]).cpu().numpy()

    def set_flat_parameters(self, w):
        """weights:copy(stored.weights) (main.lua:97) + re-pack."""
        w = torch.as_tensor(np.asarray(w, dtype=np.float32).reshape(-1))
        names = self.learnable_names()
        total = sum(self.params[n].numel() for n in names)
        if w.numel() != total:
            raise ValueError("snapshot holds %d parameters, the model has %d" % (w.numel(), total))
        off = 0
        for n in names:
            k = self.params[n].numel()
            self.params[n].copy_(w[off:off + k].to(self.device))
            off += k
        self.pack_weights()

    def save_snapshot(self, file_name, options=None, stats=None, with_bn_running_stats=True, ascii=True):
        """save_model(file_name, weights, opt, training_stats) (utilities.lua:126-134, main.lua:147)."""
        from . import t7
        bn = None
        if with_bn_running_stats:
            bn = {n: self.params[n].cpu().numpy() for n in self.param_names if n.endswith((".bn_mean", ".bn_var"))}
        t7.save_model(file_name, t7.Tensor(self.get_flat_parameters(), "torch.CudaTensor"), options or {},
                      stats or {"pcls": {}, "preg": {}, "dcls": {}, "dreg": {}}, bn, ascii=ascii)

    def load_snapshot(self, file_name):
        """load_model's restore branch (main.lua:94-98): returns the stored training_stats."""
        from . import t7
        w, _, stats, bn = t7.load_model(file_name)
        self.set_flat_parameters(w)
        if bn:
            for n, v in bn.items():
                self.params[n].copy_(torch.as_tensor(np.asarray(v, dtype=np.float32)).reshape(-1).to(self.device))
            self.pack_weights()
        return stats

    def scale_frame(self, img, width, height):
        """image.scale(img, width, height) ('bilinear') on the GPU: [C][h][w] fp32 CUDA tensor -> [C][height][width]
        (BatchIterator.lua:49-52; frcnn_scale_frame)."""
        img = img.to(torch.float32).contiguous()
        out = torch.empty((img.shape[0], int(height), int(width)), dtype=torch.float32, device=img.device)
        check(self.ctx, lib().frcnn_scale_frame(self.ctx, ffi.cast("const float*", img.data_ptr()), img.shape[0], img.shape[1],
                                                img.shape[2], ffi.cast("float*", out.data_ptr()), int(height), int(width)))
        return out

    def normalize_frame(self, img, rgb2yuv=False, centering=True, scaling=True, contrastive_width=7):
        """The frame normalisation of BatchIterator:processImage / load_image (BatchIterator.lua:146-161,
        utilities.lua:211-212) on the GPU, in place on a [3][H][W] fp32 CUDA tensor (frcnn_normalize_frame)."""
        assert img.is_cuda and img.dtype == torch.float32 and img.dim() == 3 and img.is_contiguous()
        check(self.ctx, lib().frcnn_normalize_frame(self.ctx, ffi.cast("float*", img.data_ptr()), img.shape[1], img.shape[2],
                                                    1 if rgb2yuv else 0, 1 if centering else 0, 1 if scaling else 0, int(contrastive_width)))
        return img

    def set_schedule(self, schedule):
        """'latency' (default: every stage fills the machine with one frame batch) or 'throughput' (least SM time per
        frame: unsplit anchor heads with the tail fused; for several frames in flight).  frcnn_set_schedule."""
        L = lib()
        mode = {"latency": L.FRCNN_SCHED_LATENCY, "throughput": L.FRCNN_SCHED_THROUGHPUT}[schedule]
        rc = L.frcnn_set_schedule(self.ctx, mode)
        if rc != 0:
            raise RuntimeError("frcnn_set_schedule failed")

    def set_eval_precision(self, precision):
        """'fp16' (default) or 'bf16': 16-bit operand format of the evaluate-mode tensor-core convolutions / Linear layers
        (frcnn_set_eval_precision).  Training always uses bf16 operands."""
        L = lib()
        rc = L.frcnn_set_eval_precision(self.ctx, {"bf16": L.FRCNN_PREC_BF16, "fp16": L.FRCNN_PREC_FP16}[precision])
        if rc != 0:
            raise RuntimeError("frcnn_set_eval_precision failed")
        self._precision = precision

    def replicate(self):
        """A second context of the same architecture on the same device holding a copy of the current weights (own
        stream, workspaces and CUDA graph): one more frame in flight for DetectorPipeline."""
        r = Model(self.cfg, self.layers, self.anchor_nets, self.class_layers, **self._ctor)
        r.weights.copy_(self.weights)
        r.pack_weights()
        r.set_eval_precision(getattr(self, "_precision", "fp16"))
        return r

    def pack_weights(self):
        """Re-packs the flat fp32 weights into the tensor-core layouts; call after every change of `weights`."""
        torch.cuda.synchronize(self.device)
        check(self.ctx, lib().frcnn_pack_weights(self.ctx))

    # -- forward passes ---------------------------------------------------------------------------------------
    def output_dims(self, h, w):
        dims = ffi.new("int[]", 3 * (self.n_heads + 1))
        check(self.ctx, lib().frcnn_pnet_output_dims(self.ctx, h, w, dims))
        return [tuple(dims[3 * i + j] for j in range(3)) for i in range(self.n_heads + 1)]

    def _pnet_forward(self, img, dropout_masks=None, seed=0):
        """pnet:forward(img) (Detector.lua:33, objective.lua:71): img [3][H][W] (or [N][3][H][W]) fp32 CUDA tensor ->
        list of the 4 anchor-head maps [18][h][w] and the last conv-block map [C][h][w] (leading N if batched).
        After pnet.training() the forward runs in training mode (SpatialDropout masks: `dropout_masks` = list of
        [N][C] 0/1 tensors per dropout layer, or drawn from `seed`) and keeps the state pnet.backward needs."""
        batched = img.dim() == 4
        x = (img if batched else img.unsqueeze(0)).to(self.device, torch.float32).contiguous()
        n, _, h, w = x.shape
        outs = [torch.empty((n,) + d, dtype=torch.float32, device=self.device) for d in self.output_dims(h, w)]
        ptrs = ffi.new("float*[]", [ffi.cast("float*", o.data_ptr()) for o in outs])
        if self.pnet.train:
            self._train_img = x  # pnet:backward re-reads the frame (first-layer weight gradient)
            keep, mp = [], ffi.NULL
            if dropout_masks is not None:
                keep = [m.to(self.device, torch.float32).reshape(n, -1).contiguous() for m in dropout_masks]
                mp = ffi.new("const float*[]", [ffi.cast("const float*", m.data_ptr()) for m in keep])
            check(self.ctx, lib().frcnn_pnet_forward_train(self.ctx, ffi.cast("const float*", x.data_ptr()), n, h, w, ptrs, mp, seed))
            torch.cuda.synchronize(self.device)  # `keep` may be released afterwards
        else:
            check(self.ctx, lib().frcnn_pnet_forward(self.ctx, ffi.cast("const float*", x.data_ptr()), n, h, w, ptrs))
        return outs if batched else [o[0] for o in outs]

    def _pnet_backward(self, img, delta_outputs):
        """pnet:backward(img, delta_outputs) (objective.lua:189): accumulates the parameter gradients into
        `self.gradient`; entries of delta_outputs may be None (= zero)."""
        keep = [None if d is None else (d if d.dim() == 4 else d.unsqueeze(0)).to(self.device, torch.float32).contiguous()
                for d in delta_outputs]
        ptrs = ffi.new("const float*[]", [ffi.NULL if d is None else ffi.cast("const float*", d.data_ptr()) for d in keep])
        check(self.ctx, lib().frcnn_pnet_backward(self.ctx, ptrs))
        torch.cuda.synchronize(self.device)
        return None  # the reference ignores the returned input gradient (objective.lua:189)

    def zero_grad(self):
        self.gradient.zero_()

    def synchronize(self):
        """Waits for the context's stream: train_image / train_batch return when the losses are on the host, while the
        backward pass may still be accumulating into `self.gradient` (torch ops on the default stream are ordered behind
        it anyway; this is for host-side readers and timing)."""
        check(self.ctx, lib().frcnn_synchronize(self.ctx))

    def train_image(self, img, positives, negatives, pnet_masks=None, cnet_masks=None, seed=0):
        """The body of the per-image loop of lossAndGradient (objective.lua:65-198) for one frame: forward, criteria,
        backward; gradients accumulate into `self.gradient`.  positives: [(anchor, roi)], negatives: [(anchor,)] as
        BatchIterator:nextTraining yields them (anchor = Anchors:get result, roi = {rect, class_index}); the lists
        must be cleaned with cleanAnchors first.  Returns {cls, reg, creg, ccls} loss sums of the frame."""
        from .geometry import Anchors
        x = img.to(self.device, torch.float32).contiguous()
        _, h, w = x.shape

        def fill(arr, i, anchor, roi):
            e = arr[i]
            for k, v in enumerate(anchor.unpack()):
                e.anchor[k] = v
            e.layer, e.aspect, e.y, e.x = anchor.layer, anchor.aspect, anchor.index[1], anchor.index[2]
            if roi is not None:
                for k, v in enumerate(roi["rect"].unpack()):
                    e.roi[k] = v
                t = Anchors.inputToAnchor(anchor, roi["rect"])
                for k in range(4):
                    e.reg_target[k] = float(t[k])
                e.class_index = int(roi["class_index"])

        pos = ffi.new("frcnn_example[]", max(len(positives), 1))
        neg = ffi.new("frcnn_example[]", max(len(negatives), 1))
        for i, (a, roi) in enumerate(positives):
            fill(pos, i, a, roi)
        for i, ex in enumerate(negatives):
            fill(neg, i, ex[0], None)
        keep = []

        def ptrs(masks):
            if masks is None:
                return ffi.NULL
            ts = [m.to(self.device, torch.float32).contiguous() for m in masks]
            keep.extend(ts)
            return ffi.new("const float*[]", [ffi.cast("const float*", t.data_ptr()) for t in ts])

        pm, cm = ptrs(pnet_masks), ptrs(cnet_masks)
        losses = ffi.new("float[4]")
        torch.cuda.synchronize(self.device)
        check(self.ctx, lib().frcnn_train_image(self.ctx, ffi.cast("const float*", x.data_ptr()), h, w, pos, len(positives), neg,
                                                len(negatives), pm, cm, seed, losses))
        if self._foreign_stream:
            self.synchronize()      # `keep` (the injected masks) is released on return
        return dict(cls=losses[0], reg=losses[1], creg=losses[2], ccls=losses[3])

    # frcnn_example as a numpy record: the example lists are marshalled with array operations, not per-field cffi writes
    _EXAMPLE_DTYPE = np.dtype([("anchor", np.float64, 4), ("roi", np.float64, 4), ("reg_target", np.float32, 4), ("layer", np.int32),
                               ("aspect", np.int32), ("y", np.int32), ("x", np.int32), ("class_index", np.int32), ("pad_", np.int32)])

    @classmethod
    def pack_examples(cls, examples):
        """[(anchor, roi)] or [(anchor,)] -> numpy record array laid out as frcnn_example[]."""
        from .geometry import Anchors
        arr = np.zeros(max(len(examples), 1), dtype=cls._EXAMPLE_DTYPE)
        assert arr.dtype.itemsize == ffi.sizeof("frcnn_example")
        for i, ex in enumerate(examples):
            a = ex[0]
            e = arr[i]
            e["anchor"] = a.unpack()
            e["layer"], e["aspect"], e["y"], e["x"] = a.layer, a.aspect, a.index[1], a.index[2]
            if len(ex) > 1 and ex[1] is not None:
                roi = ex[1]
                e["roi"] = roi["rect"].unpack()
                e["reg_target"] = Anchors.inputToAnchor(a, roi["rect"])
                e["class_index"] = int(roi["class_index"])
        return arr

    def train_batch(self, imgs, positives, negatives, pnet_masks=None, seeds=None, packed=None):
        """lossAndGradient's per-image loop (objective.lua:65-198) for a list / stack of frames of ONE size in one call
        (frcnn_train_batch): pnet forward and backward once over all frames, the per-image stages in between.
        positives / negatives: one example list per frame (cleaned).  packed: optional pre-marshalled
        (pos_records, neg_records) per frame from pack_examples.  Returns a list of {cls, reg, creg, ccls} per frame."""
        x = (imgs if torch.is_tensor(imgs) else torch.stack(list(imgs))).to(self.device, torch.float32).contiguous()
        n, _, h, w = x.shape
        if packed is None:
            packed = [(self.pack_examples(p), self.pack_examples(q)) for p, q in zip(positives, negatives)]
        n_pos = ffi.new("int[]", [len(p) for p in positives])
        n_neg = ffi.new("int[]", [len(q) for q in negatives])
        pp = ffi.new("const frcnn_example*[]", [ffi.cast("const frcnn_example*", a.ctypes.data) for a, _ in packed])
        qq = ffi.new("const frcnn_example*[]", [ffi.cast("const frcnn_example*", b.ctypes.data) for _, b in packed])
        sd = ffi.new("uint64_t[]", [int(v) for v in (seeds if seeds is not None else range(n))])
        keep, pm = [], ffi.NULL
        if pnet_masks is not None:
            keep = [m.to(self.device, torch.float32).reshape(n, -1).contiguous() for m in pnet_masks]
            pm = ffi.new("const float*[]", [ffi.cast("const float*", t.data_ptr()) for t in keep])
        losses = ffi.new("float[]", 4 * n)
        # the library's stream must see what torch's current stream produced (the frame stack); NOT a device-wide
        # synchronize: uploads of the next step's frames on a prefetcher's stream keep running
        torch.cuda.current_stream(self.device).synchronize()
        check(self.ctx, lib().frcnn_train_batch(self.ctx, ffi.cast("const float*", x.data_ptr()), n, h, w, pp, n_pos, qq, n_neg, pm, sd, losses))
        if self._foreign_stream:
            self.synchronize()
        return [dict(cls=losses[4 * i], reg=losses[4 * i + 1], creg=losses[4 * i + 2], ccls=losses[4 * i + 3]) for i in range(n)]

    def _cnet_forward(self, x, dropout_masks=None, seed=0):
        """cnet:forward(cinput) (Detector.lua:101, objective.lua:164): x [R][kh*kw*C] fp32 -> (bbox [R][4], log-softmax
        [R][classes+1]).  After cnet.training(): BatchNormalization batch statistics (running statistics updated), Dropout
        masks drawn from `seed` or injected (`dropout_masks`: one [R][n] 0/1 tensor per class layer); the state
        cnet.backward needs stays in the context."""
        x = x.to(self.device, torch.float32).contiguous()
        R = x.shape[0]
        reg = torch.empty((R, 4), dtype=torch.float32, device=self.device)
        cls = torch.empty((R, self.cfg["class_count"] + 1), dtype=torch.float32, device=self.device)
        if self.cnet.train:
            keep = [m.to(self.device, torch.float32).contiguous() for m in dropout_masks] if dropout_masks is not None else []
            mp = ffi.new("const float*[]", [ffi.cast("const float*", m.data_ptr()) for m in keep]) if keep else ffi.NULL
            check(self.ctx, lib().frcnn_cnet_forward_train(self.ctx, ffi.cast("const float*", x.data_ptr()), R, mp, seed,
                                                           ffi.cast("float*", reg.data_ptr()), ffi.cast("float*", cls.data_ptr())))
            torch.cuda.synchronize(self.device)  # `keep` / `x` may be released afterwards
            return reg, cls
        check(self.ctx, lib().frcnn_cnet_forward(self.ctx, ffi.cast("const float*", x.data_ptr()), R,
                                                 ffi.cast("float*", reg.data_ptr()), ffi.cast("float*", cls.data_ptr())))
        return reg, cls

    def _cnet_backward(self, cinput, deltas):
        """cnet:backward(cinput, {crdelta, ccdelta}) (objective.lua:179) after a training-mode cnet.forward(cinput): returns
        post_roi_delta [R][kh*kw*C]; parameter gradients accumulate in `self.gradient`."""
        crdelta, ccdelta = deltas
        dr = crdelta.to(self.device, torch.float32).contiguous()
        dc = ccdelta.to(self.device, torch.float32).contiguous()
        dx = torch.empty((dr.shape[0], cinput.shape[1]), dtype=torch.float32, device=self.device)
        check(self.ctx, lib().frcnn_cnet_backward(self.ctx, ffi.cast("const float*", dr.data_ptr()),
                                                  ffi.cast("const float*", dc.data_ptr()), ffi.cast("float*", dx.data_ptr())))
        torch.cuda.synchronize(self.device)
        return dx

    def cnet_train_step(self, x, n_pos, crtarget, cctarget, masks=None, seed=0):
        """cnet:forward (training) + detection-stage criteria + cnet:backward (objective.lua:164-179) on example rows
        x [R][kh*kw*C]; returns (post_roi_delta [R][kh*kw*C], {creg, ccls}); gradients accumulate in self.gradient."""
        x = x.to(self.device, torch.float32).contiguous()
        R = x.shape[0]
        crt = crtarget.to(self.device, torch.float32).contiguous()
        cct = cctarget.to(self.device, torch.int32).contiguous()
        dx = torch.empty_like(x)
        keep = [m.to(self.device, torch.float32).contiguous() for m in masks] if masks is not None else []
        mp = ffi.new("const float*[]", [ffi.cast("const float*", m.data_ptr()) for m in keep]) if keep else ffi.NULL
        losses = ffi.new("float[2]")
        torch.cuda.synchronize(self.device)
        check(self.ctx, lib().frcnn_cnet_train_step(self.ctx, ffi.cast("const float*", x.data_ptr()), R, n_pos,
                                                    ffi.cast("const float*", crt.data_ptr()), ffi.cast("const int32_t*", cct.data_ptr()),
                                                    mp, seed, ffi.cast("float*", dx.data_ptr()), losses))
        return dx, dict(creg=losses[0], ccls=losses[1])

    def launch_count(self):
        return int(lib().frcnn_launch_count(self.ctx))

    def block_outputs(self, n=1):
        """Diagnostic: the pooled output of every conv block of the last pnet forward over `n` frames, fp32 [n][C][h][w]
        (frcnn_block_output)."""
        outs, d = [], ffi.new("int[3]")
        for b in range(1, len(self.layers) + 1):
            check(self.ctx, lib().frcnn_block_output(self.ctx, b, ffi.NULL, d))
            t = torch.empty((n, d[0], d[1], d[2]), dtype=torch.float32, device=self.device)
            check(self.ctx, lib().frcnn_block_output(self.ctx, b, ffi.cast("float*", t.data_ptr()), d))
            outs.append(t)
        torch.cuda.synchronize(self.device)
        return outs

    def dp_info(self):
        """frcnn_dp_info: the context's NCCL communicator (nranks 0 = none), NCCL version, bytes all-reduced so far."""
        r, n, v, b = ffi.new("int*"), ffi.new("int*"), ffi.new("int*"), ffi.new("int64_t*")
        check(self.ctx, lib().frcnn_dp_info(self.ctx, r, n, v, b))
        return dict(comm_rank=int(r[0]), comm_nranks=int(n[0]), nccl_version=int(v[0]), bytes_reduced=int(b[0]))


def find_target_size(orig_w, orig_h, target_smaller_side, max_pixel_size):  # utilities.lua:188-204
    w, h = ffi.new("int*"), ffi.new("int*")
    check(None, lib().frcnn_find_target_size(int(orig_w), int(orig_h), float(target_smaller_side), float(max_pixel_size), w, h))
    return int(w[0]), int(h[0])


def create_model(cfg, layers, anchor_nets, class_layers, **kw):  # model_utilities.lua:126-136
    return Model(cfg, layers, anchor_nets, class_layers, **kw)


def vgg_small(cfg, **kw):  # models/vgg_small.lua:3-25
    layers = [dict(filters=64, kW=3, kH=3, padW=1, padH=1, dropout=0.0, conv_steps=1),
              dict(filters=128, kW=3, kH=3, padW=1, padH=1, dropout=0.4, conv_steps=2),
              dict(filters=256, kW=3, kH=3, padW=1, padH=1, dropout=0.4, conv_steps=2),
              dict(filters=384, kW=3, kH=3, padW=1, padH=1, dropout=0.4, conv_steps=2)]
    anchor_nets = [dict(kW=3, n=256, input=3), dict(kW=3, n=256, input=4), dict(kW=5, n=256, input=4),
                   dict(kW=7, n=256, input=4)]
    class_layers = [dict(n=1024, dropout=0.5, batch_norm=True), dict(n=512, dropout=0.5)]
    return create_model(cfg, layers, anchor_nets, class_layers, **kw)


def vgg_large(cfg, **kw):  # models/vgg_large.lua:3-25
    layers = [dict(filters=64, kW=3, kH=3, padW=1, padH=1, dropout=0.0, conv_steps=2),
              dict(filters=128, kW=3, kH=3, padW=1, padH=1, dropout=0.4, conv_steps=2),
              dict(filters=256, kW=3, kH=3, padW=1, padH=1, dropout=0.4, conv_steps=3),
              dict(filters=512, kW=3, kH=3, padW=1, padH=1, dropout=0.4, conv_steps=3)]
    anchor_nets = [dict(kW=3, n=256, input=3), dict(kW=3, n=256, input=4), dict(kW=5, n=256, input=4),
                   dict(kW=7, n=256, input=4)]
    class_layers = [dict(n=1024, dropout=0.5, batch_norm=True), dict(n=512, dropout=0.5)]
    return create_model(cfg, layers, anchor_nets, class_layers, **kw)
