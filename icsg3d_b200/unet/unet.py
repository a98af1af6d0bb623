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
from ..weights_io import load_npz, save_npz


class _LossSpec:
    def __init__(self, weights):
        self.weights = weights
        self.__name__ = "loss"


def weighted_categorical_crossentropy(weights):
    """unet.py:196-221.  `weights` may be a (C,) array or — as the reference's own compile() call passes it
    (unet.py:254) — the scalar number of classes."""
    return _LossSpec(weights)


def _metric(name):
    def f(y_true, y_pred):
        raise RuntimeError(f"{name} is computed inside the fused head kernel (metrics of train/test_on_batch)")
    f.__name__ = name
    return f


r_m, wr_m, p_m, f1_m = _metric("r_m"), _metric("wr_m"), _metric("p_m"), _metric("f1_m")


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
        save_npz(path, self._o.params.to_dict())

    def load_weights(self, path):
        self._o.params.load_dict(load_npz(path))

    def save(self, path):
        self.save_weights(path)


class AtomUnet:
    def __init__(self, num_classes=95, class_weights=None, weights=None, input_shape=(32, 32, 32, 4), lr=1e-6, device=None,
                 dist: Dist | None = None, seed=2, loss_weight=None, use_cuda_graph=True):
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

    def engine(self, batch) -> UNetEngine:
        eng = self._engines.get(batch)
        if eng is None:
            eng = UNetEngine(batch, d=self.input_shape[0], channels=self.input_shape[-1], classes=self.num_classes,
                             device=self.device, params=self.params, lr=self.optimizer.lr, class_weight=self.loss_weight,
                             dist=self.dist)
            self._engines[batch] = eng
        return eng

    def _step(self, x, labels, train):
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

    def _predict(self, x, want_probs=False, batch=None):
        n = len(x)
        B = min(n, batch or 8)
        eng = self.engine(B)
        d, C = self.input_shape[0], self.num_classes
        x = _to_dev(x, self.device, torch.float32)
        probs = torch.empty(B, d, d, d, C, dtype=torch.float32, device=self.device) if want_probs else None
        soft, sig, lab = [], [], []
        for s in range(0, n, B):
            k = min(s + B, n) - s
            eng.X[:k].copy_(x[s:s + k].reshape(k, d, d, d, -1))
            eng.predict(probs)
            if want_probs:
                soft.append(probs[:k].cpu().numpy().copy())
            sig.append(eng.sigp[:k].cpu().numpy().copy()[..., None])
            lab.append(eng.argmax[:k].cpu().numpy().copy())
        return (np.concatenate(soft) if want_probs else None), np.concatenate(sig), np.concatenate(lab)

    def predict_labels(self, x, threshold=0.8, batch=None):
        """Fused post-processing of generate.py:221-225: argmax species labels (uint8) and the thresholded atom mask,
        without materialising the (n,d,d,d,95) probability tensor."""
        _, sig, lab = self._predict(x, want_probs=False, batch=batch)
        return lab, (sig[..., 0] >= threshold)

    def train_generator(self, train_gen, val_gen, epochs=100, output_dir="output/unet/"):
        """unet.py:357-381: fit_generator + ModelCheckpoint(save_best_only on val_loss); plotting callback omitted."""
        print("Training...")
        best = np.inf
        for e in range(epochs):
            tm = [self.model.train_on_batch(*train_gen[i]) for i in range(len(train_gen))]
            vm = [self.model.test_on_batch(*val_gen[i]) for i in range(len(val_gen))]
            tm, vm = np.mean(tm, axis=0), (np.mean(vm, axis=0) if vm else np.mean(tm, axis=0))
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
