"""Drop-in for the reference's unet/unet.py::AtomUnet (lines 223-390) on the B200 engine.

Same constructor, `.model` facade (`predict`, `train_on_batch`, `test_on_batch`, `save_weights`, `load_weights`,
`save`, `predict_generator`), `train_generator`, `save_`, and the module-level names the rest of the reference
imports (`custom_objects`, `weighted_categorical_crossentropy`, `f1_m`, `wr_m`, `p_m`, `r_m` — unet.py:159-221,
393-399).  In the reference those are Keras loss/metric closures; here they are descriptors that the fused head
kernel (csrc/heads.cu) implements.  The TrainingPlot callback (unet.py:39-157) is out of scope (matplotlib).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from ..engine import Dist
from ..optimizers import Adam
from ..params import ParamStore, unet_specs
from ..unet_engine import UNetEngine
from ..weights_io import load_weights_file, save_weights_file


class _LossSpec:
    def __init__(self, weights):
        self.weights = weights
        self.__name__ = "loss"


def weighted_categorical_crossentropy(weights):
    """unet.py:196-221.  `weights` may be a (C,) array or — as the reference's own compile() call passes it
    (unet.py:254) — the scalar number of classes."""
    return _LossSpec(weights)


K_EPSILON = 1e-7  # keras.backend.epsilon()


def _counts(y_true, y_pred):
    """The K.round(K.clip(., 0, 1)) sums of unet.py:159-193 on the device (csrc/post.cu::metric_counts_kernel):
    [tp, possible, predicted, tp without class 0, possible without class 0]."""
    from .. import ops
    dev = f"cuda:{torch.cuda.current_device()}" if torch.cuda.is_available() else "cuda"
    t = _to_dev(y_true, dev, torch.float32).contiguous()
    p = _to_dev(y_pred, dev, torch.float32).contiguous()
    return ops.metric_counts(t, p).cpu().tolist()


def r_m(y_true, y_pred):
    """unet.py:159-167 recall over all classes."""
    tp, possible, _, _, _ = _counts(y_true, y_pred)
    return tp / (possible + K_EPSILON)


def wr_m(y_true, y_pred):
    """unet.py:170-179 recall without the zero (background) class (hard-coded np.ones(95) weights with w[0] = 0)."""
    _, _, _, tp_w, possible_w = _counts(y_true, y_pred)
    return tp_w / (possible_w + K_EPSILON)


def p_m(y_true, y_pred):
    """unet.py:182-187 precision."""
    tp, _, predicted, _, _ = _counts(y_true, y_pred)
    return tp / (predicted + K_EPSILON)


def f1_m(y_true, y_pred):
    """unet.py:189-193."""
    tp, possible, predicted, _, _ = _counts(y_true, y_pred)
    precision, recall = tp / (predicted + K_EPSILON), tp / (possible + K_EPSILON)
    return 2 * ((precision * recall) / (precision + recall + K_EPSILON))


def _to_dev(a, dev, dtype):
    if torch.is_tensor(a):
        return a.to(device=dev, dtype=dtype, non_blocking=True)
    return torch.as_tensor(np.ascontiguousarray(a)).to(device=dev, dtype=dtype, non_blocking=True)


class _Model:
    def __init__(self, owner):
        self._o = owner

    def _labels(self, y):
        """Accept the reference's one-hot float labels (unet/data.py:89) or uint8 species directly."""
        ys = y[0] if isinstance(y, (list, tuple)) else y
        t = ys if torch.is_tensor(ys) else torch.as_tensor(np.asarray(ys))
        if t.dim() == 5 and t.shape[-1] == self._o.num_classes:
            t = t.argmax(dim=-1)
        elif t.dim() == 5 and t.shape[-1] == 1:
            t = t[..., 0]
        return t.to(torch.uint8)

    def train_on_batch(self, x, y):
        """-> [loss, soft_loss, sig_loss, soft_f1_m, soft_wr_m] (the order Keras reports for unet.py:252-259)."""
        return self._o._step(x, self._labels(y), train=True)

    def test_on_batch(self, x, y):
        return self._o._step(x, self._labels(y), train=False)

    def predict(self, x, batch_size=None):
        """-> [soft (n,d,d,d,classes), sig (n,d,d,d,1)] float32 (generate.py:220)."""
        soft, sig, _ = self._o._predict(x, want_probs=True)
        return [soft, sig]

    def predict_generator(self, gen):
        outs = [self.predict(gen[i][0]) for i in range(len(gen))]
        return [np.concatenate([o[0] for o in outs]), np.concatenate([o[1] for o in outs])]

    def save_weights(self, path):
        """Keras HDF5 for .h5/.hdf5 paths (one group per layer, readable by the reference's `load_weights`), else .npz.
        Data parallel: rank 0 alone writes."""
        o = self._o
        if o.dist is not None and o.dist.world > 1 and o.dist.rank != 0:
            return
        save_weights_file(path, o.params.to_dict(), o.params.specs, model="unet")

    def load_weights(self, path):
        self._o.params.load_dict(load_weights_file(path, self._o.params.specs))

    def save(self, path):
        self.save_weights(path)


class AtomUnet:
    def __init__(self, num_classes=95, class_weights=None, weights=None, input_shape=(32, 32, 32, 4), lr=1e-6, device=None,
                 dist: Dist | None = None, seed=2, loss_weight=None, use_cuda_graph=True, dtype="bf16"):
        if dtype not in ("bf16", "fp32"):
            raise ValueError("dtype must be 'bf16' (throughput mode) or 'fp32' (fp32-class split operands, parity mode)")
        self.dtype = dtype
        self._x3, self._x3t = {}, {}
        self.class_weights = class_weights
        self.input_shape = tuple(input_shape)
        self.optimizer = Adam(lr)
        self.num_classes = num_classes
        self.device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
        self.dist = dist
        self.use_cuda_graph = use_cuda_graph
        # reference quirk: the compiled loss uses the scalar `num_classes` as weight; `class_weights` is unused (unet.py:243,254)
        self.loss_weight = loss_weight
        self.params = ParamStore(unet_specs(self.input_shape[-1], num_classes), self.device).init(seed)
        self.model = _Model(self)
        self.metric_names = ["Loss", "lsoft", "lsig", "f1", "wr"]
        self._engines = {}
        if weights and os.path.exists(weights):
            self.model.load_weights(weights)
            print("loaded weights")
            self.filepath = weights
        elif weights and not os.path.exists(weights):
            self.filepath = weights
        else:
            self.filepath = "./saved_models/unet_%d_channel_weights.best.hdf5" % self.input_shape[-1]

    def engine(self, batch, train=True) -> UNetEngine:
        """train=False: inference-only engine (no gradient / Adam buffers, no dgrad weight copies)."""
        key = (batch, bool(train))
        eng = self._engines.pop(key, None)
        if eng is None and not train:  # a training engine of that batch can serve inference too
            eng = self._engines.pop((batch, True), None)
            key = (batch, True) if eng is not None else key
        if eng is None:
            while len(self._engines) >= 2:  # keep at most two engines (buffers + captured graph) alive, evict the LRU one
                self._engines.pop(next(iter(self._engines)))
            eng = UNetEngine(batch, d=self.input_shape[0], channels=self.input_shape[-1], classes=self.num_classes,
                             device=self.device, params=self.params, lr=self.optimizer.lr, class_weight=self.loss_weight,
                             dist=self.dist, train=train)
        self._engines[key] = eng
        return eng

    def _step_x3(self, x, labels, train):
        from ..engine_x3 import UNetTrainX3
        B, d = len(x), self.input_shape[0]
        eng = self._x3t.get(B)
        if eng is None:
            self._x3t.clear()
            eng = self._x3t[B] = UNetTrainX3(B, d=d, channels=self.input_shape[-1], classes=self.num_classes, device=self.device,
                                             params=self.params, lr=self.optimizer.lr, class_weight=self.loss_weight)
        x = _to_dev(x, self.device, torch.float32).reshape(B, d, d, d, -1)
        if train:
            return eng.train_step(x, labels).cpu().tolist()
        # test_on_batch: learning phase 0 forward + the losses / metrics on the labels
        from .. import ops
        logits = eng.forward(x, training=False)
        Mv = B * d ** 3
        part = torch.zeros(ops.heads_loss_nparts(Mv), 6, dtype=torch.float64, device=self.device)
        out = torch.zeros(5, dtype=torch.float32, device=self.device)
        sp = labels.to(self.device).to(torch.uint8).reshape(B, d, d, d).contiguous()
        ops.heads_loss(logits, self.num_classes, sp, eng.class_w, 1.0 / Mv, part)
        ops.heads_loss_finalize(part, float(Mv), out)
        return out.cpu().tolist()

    def _step(self, x, labels, train):
        if self.dtype == "fp32":
            if self.dist is not None and self.dist.world > 1:
                raise NotImplementedError("the fp32-class mode is single-process (parity mode)")
            return self._step_x3(x, labels, train)
        B = len(x)
        eng = self.engine(B)
        eng.set_inputs(_to_dev(x, self.device, torch.float32), labels.to(self.device))
        if train:
            if self.use_cuda_graph and not eng.use_graph:
                eng.capture_train_graph()
            eng.train_step()
        else:
            eng.eval_step()
        return eng.metrics_host()

    def _predict(self, x, want_probs=False, batch=None, threshold=0.8):
        n = len(x)
        B = min(n, batch or 8)
        eng = self.engine(B, train=False)
        d, C = self.input_shape[0], self.num_classes
        x = _to_dev(x, self.device, torch.float32)
        probs = torch.empty(B, d, d, d, C, dtype=torch.float32, device=self.device) if want_probs else None
        soft, sig, lab, msk = [], [], [], []
        for s in range(0, n, B):
            k = min(s + B, n) - s
            eng.X[:k].copy_(x[s:s + k].reshape(k, d, d, d, -1))
            eng.predict(probs, threshold=threshold)
            if want_probs:
                soft.append(probs[:k].cpu().numpy().copy())
            sig.append(eng.sigp[:k].cpu().numpy().copy()[..., None])
            lab.append(eng.argmax[:k].cpu().numpy().copy())
            msk.append(eng.mask[:k].cpu().numpy().copy())
        self._last_mask = np.concatenate(msk)
        return (np.concatenate(soft) if want_probs else None), np.concatenate(sig), np.concatenate(lab)

    def predict_labels(self, x, threshold=0.8, batch=None):
        """Fused post-processing of generate.py:221-225: argmax species labels (uint8) and the thresholded atom mask
        (bool), both produced on the device (csrc/post.cu::heads_predict_kernel) without materialising the
        (n,d,d,d,95) probability tensor.  With AtomUnet(dtype="fp32") the pass runs on fp32-class split operands
        (engine_x3.py), the mode in which the labels are bit-exact against the fp32 reference graph."""
        if self.dtype == "fp32":
            return self._predict_labels_x3(x, threshold, batch)
        _, _, lab = self._predict(x, want_probs=False, batch=batch, threshold=threshold)
        return lab, self._last_mask.astype(bool)

    def _predict_labels_x3(self, x, threshold, batch):
        from .. import ops
        from ..engine_x3 import UNetForwardX3
        n, d = len(x), self.input_shape[0]
        B = min(n, batch or 8)
        x = _to_dev(x, self.device, torch.float32).reshape(n, d, d, d, -1)
        lab, msk = [], []
        for s in range(0, n, B):
            k = min(s + B, n) - s
            un = self._x3.get(k)
            if un is None:
                un = self._x3[k] = UNetForwardX3(k, d=d, channels=self.input_shape[-1], classes=self.num_classes,
                                                 device=self.device, params=self.params)
            logits, _, _ = un.predict(x[s:s + k])
            am = torch.empty(k, d, d, d, dtype=torch.uint8, device=self.device)
            mk = torch.empty(k, d, d, d, dtype=torch.uint8, device=self.device)
            ops.heads_predict(logits, self.num_classes, float(threshold), argmax=am, mask=mk)
            lab.append(am.cpu().numpy())
            msk.append(mk.cpu().numpy().astype(bool))
        return np.concatenate(lab), np.concatenate(msk)

    def fit_epoch(self, gen, steps=None, train=True):
        """The batch loop of fit_generator (unet.py:357-381): one train_on_batch (test_on_batch) per gen[b] -> (x, y),
        returned as a [steps, 5] array of [loss, soft_loss, sig_loss, soft_f1_m, soft_wr_m].  Pipelined like
        LatticeDFCVAE.fit_epoch: the next batch's host -> device copy runs on a copy stream under the current step, the
        metrics return through a pinned ring, the host waits once at the end."""
        steps = len(gen) if steps is None else int(steps)
        if steps <= 0:
            return np.zeros((0, 5))
        if self.dtype == "fp32":
            return np.array([self._step(gen[b][0], self.model._labels(gen[b][1]), train=train) for b in range(steps)])
        x0, y0 = gen[0]
        B = len(x0)
        eng = self.engine(B)
        h = lambda a: a if torch.is_tensor(a) else torch.as_tensor(np.ascontiguousarray(a))
        f32 = lambda a: a if a.dtype == torch.float32 else a.float()
        if train and self.use_cuda_graph and not eng.use_graph:
            eng.set_inputs(_to_dev(x0, self.device, torch.float32), self.model._labels(y0).to(self.device))
            eng.capture_train_graph()
        main = torch.cuda.current_stream()
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream, self._stage = torch.cuda.Stream(), {}
        if B not in self._stage:
            self._stage = {B: [(torch.empty_like(eng.X), torch.empty_like(eng.species), torch.cuda.Event(), torch.cuda.Event())
                               for _ in range(2)]}
        stage, cs = self._stage[B], self._copy_stream
        ring = torch.empty(steps, 5, dtype=torch.float32).pin_memory()
        keep = []

        def prefetch(b):
            xb, yb = (x0, y0) if b == 0 else gen[b]
            xh, yh = f32(h(xb)).reshape(eng.X.shape), self.model._labels(yb).reshape(eng.species.shape)
            keep.append((xh, yh))
            sx, sy, ready, free = stage[b & 1]
            with torch.cuda.stream(cs):
                cs.wait_event(free)
                sx.copy_(xh, non_blocking=True)
                sy.copy_(yh, non_blocking=True)
                ready.record(cs)

        cs.wait_stream(main)
        prefetch(0)
        for b in range(steps):
            if b + 1 < steps:
                prefetch(b + 1)
            sx, sy, ready, free = stage[b & 1]
            main.wait_event(ready)
            eng.set_inputs(sx, sy)
            free.record(main)
            if train:
                eng.train_step()
            else:
                eng.eval_step()
            ring[b].copy_(eng.metrics, non_blocking=True)
            if len(keep) > 4:
                del keep[0]
        main.synchronize()
        out = ring.double()
        if self.dist is not None and self.dist.world > 1:
            t = out.to(self.device)
            self.dist.all_reduce_sum(t)
            out = (t / self.dist.world).cpu()
        return out.numpy()

    def train_generator(self, train_gen, val_gen, epochs=100, output_dir="output/unet/"):
        """unet.py:357-381: fit_generator + ModelCheckpoint(save_best_only on val_loss); plotting callback omitted."""
        print("Training...")
        best = np.inf
        for e in range(epochs):
            tm = self.fit_epoch(train_gen, train=True)
            vm = self.fit_epoch(val_gen, train=False)
            tm, vm = np.mean(tm, axis=0), (np.mean(vm, axis=0) if len(vm) else np.mean(tm, axis=0))
            print("Epoch %d/%d - loss: %.4f - soft_loss: %.4f - sig_loss: %.4f - soft_f1_m: %.4f - soft_wr_m: %.4f - "
                  "val_loss: %.4f" % (e + 1, epochs, tm[0], tm[1], tm[2], tm[3], tm[4], vm[0]))
            if vm[0] < best:
                best = vm[0]
                self.model.save_weights(self.filepath)
            if hasattr(train_gen, "on_epoch_end"):
                train_gen.on_epoch_end()
        self.model.load_weights(self.filepath)
        self.model.save(os.path.splitext(self.filepath)[0] + ".h5")
        print("Model saved")

    def predict_generator(self, test_gen):
        return self.model.predict_generator(test_gen)

    def save_(self, weights, model="saved_models/unet.h5"):
        self.model.load_weights(weights)
        self.model.save(model)


def get_weights(path="", training_ids=(), n_classes=95):
    """unet/get_weights.py:19-33 with path='' (what custom_objects uses): ones(n_classes)."""
    if path:
        raise NotImplementedError("class-frequency weights from .npy species grids: use the reference's get_weights.py")
    return np.ones(n_classes)


class_weights = get_weights()
custom_objects = {"loss": weighted_categorical_crossentropy(class_weights), "f1_m": f1_m, "wr_m": wr_m}
