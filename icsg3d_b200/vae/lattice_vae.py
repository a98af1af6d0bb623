"""Drop-in for the reference's vae/lattice_vae.py::LatticeDFCVAE (lines 69-357) on the B200 engine.

Same constructor arguments, same methods (`_set_model`, `train`, `sample_vae`, `save_`), and Keras-like
`.encoder` / `.decoder` / `.model` facades with `predict`, `train_on_batch`, `test_on_batch`, `save_weights`,
`load_weights`, `save` — the calls train_vae.py:137-142, generate.py:172-208, eval.py:144-163 and
interpolate.py:50-57 make.  All compute goes through icsg3d_b200.engine -> libicsg3d.so; the plotting
methods of the reference (lattice_vae.py:359-436) are out of scope.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from ..engine import Dist, VAEEngine
from ..optimizers import Adam
from ..params import ParamStore, unet_specs, vae_specs
from ..weights_io import load_weights_file, save_weights_file


def _to_dev(a, dev, dtype=torch.float32):
    if torch.is_tensor(a):
        return a.to(device=dev, dtype=dtype, non_blocking=True)
    return torch.as_tensor(np.ascontiguousarray(a)).to(device=dev, dtype=dtype, non_blocking=True)


def _as_f32(a):
    """numpy / torch (host or device) -> fp32 torch tensor without touching the device (Keras casts float64 batches)."""
    if torch.is_tensor(a):
        return a if a.dtype == torch.float32 else a.float()
    a = np.ascontiguousarray(a)
    return torch.from_numpy(a if a.dtype == np.float32 else a.astype(np.float32))


class _Facade:
    def __init__(self, owner):
        self._o = owner


class _Encoder(_Facade):
    def predict(self, inputs, batch_size=None):
        """-> [z_mean, z_log_var, z]; z is SAMPLED (the Lambda has no learning-phase switch)."""
        M, cond = inputs
        return self._o._predict(M, cond, what="encode")


class _Decoder(_Facade):
    def predict(self, inputs, batch_size=None):
        z, cond = inputs
        return self._o._predict(z, cond, what="decode")


class _Model(_Facade):
    def predict(self, inputs, batch_size=None):
        M, cond = inputs
        return self._o._predict(M, cond, what="reconstruct")

    def train_on_batch(self, x, y=None):
        """-> [loss, perceptual_loss, mse_loss, kld_loss] (lattice_vae.py:124-125, 296-298)."""
        M, cond = x
        return self._o._step(M, cond, train=True)

    def test_on_batch(self, x, y=None):
        M, cond = x
        return self._o._step(M, cond, train=False)

    def save_weights(self, path):
        """Keras HDF5 when the path says so (.h5/.hdf5: `encoder`/`decoder` groups in Keras' nested-model weight order,
        readable by the reference's `model.load_weights`), the native .npz container otherwise.  Data parallel: every
        rank holds identical weights, rank 0 alone writes."""
        o = self._o
        if o.dist is not None and o.dist.world > 1 and o.dist.rank != 0:
            return
        save_weights_file(path, o.params.to_dict(), o.params.specs, model="vae")

    def load_weights(self, path):
        self._o.params.load_dict(load_weights_file(path, self._o.params.specs))

    def save(self, path):
        self.save_weights(path)


class LatticeDFCVAE:
    def __init__(self, input_shape=(32, 32, 32, 4), kernel_size=(3, 3, 3), pool_size=(2, 2, 2), filters=[16, 32, 64, 128],
                 latent_dim=256, beta=3e-4, alpha=0.5, optimizer=None, perceptual_model="saved_models/unet.h5",
                 pm_layers=["re_lu_2", "re_lu_4", "re_lu_6", "re_lu_8"], pm_layer_weights=[1.0, 1.0, 1.0, 1.0],
                 cond_shape=10, custom_objects=None, output_dir="output", device=None, dist: Dist | None = None,
                 seed=1, use_cuda_graph=True, dtype="bf16"):
        """dtype: "bf16" = the throughput mode (bf16 operands, fp32 accumulation; split-operand encoder forward);
        "fp32" = fp32-class split operands for every conv, forward AND backward (icsg3d_b200/engine_x3.py) — the parity
        mode of north_star's 1e-4 tier (activations, gradients), ~3x the tensor-core work, single process."""
        if dtype not in ("bf16", "fp32"):
            raise ValueError("dtype must be 'bf16' or 'fp32'")
        if dtype == "fp32" and dist is not None and dist.world > 1:
            raise NotImplementedError("the fp32-class mode is single-process (parity mode)")
        self.dtype = dtype
        self._x3 = {}
        if tuple(kernel_size) != (3, 3, 3) or tuple(pool_size) != (2, 2, 2):
            raise NotImplementedError("the B200 path implements the reference's 3x3x3 conv / 2x2x2 pool only")
        if list(pm_layers) != ["re_lu_2", "re_lu_4", "re_lu_6", "re_lu_8"]:
            raise NotImplementedError("perceptual taps are the ReLUs of c2,c4,c6,c10 (lattice_vae.py:100)")
        self.input_shape = tuple(input_shape)
        self.kernel_size, self.pool_size = kernel_size, pool_size
        self.filters = list(filters)
        self.latent_dim = latent_dim
        self.channels = self.input_shape[-1]
        self.optimizer = optimizer or Adam(5e-4)
        self.beta, self.alpha = beta, alpha
        self.batch_size = None
        self.cond_shape = cond_shape
        self.losses = []
        self.sdir = output_dir
        self.pm_layers, self.pm_layer_weights = pm_layers, pm_layer_weights
        self.metric_names = ["Loss", "PM", "MSE", "KLD"]
        self.device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
        self.dist = dist
        self.seed = seed
        self.use_cuda_graph = use_cuda_graph
        # --- perceptual model: the pre-trained U-Net (lattice_vae.py:120 load_model) ---
        new_pm = lambda: ParamStore(unet_specs(self.channels), self.device, with_grads=False, with_adam=False)
        if isinstance(perceptual_model, ParamStore):
            self.pm = perceptual_model
        elif hasattr(perceptual_model, "params"):
            self.pm = perceptual_model.params
        elif perceptual_model is None:
            self.pm = new_pm().init(seed + 1)  # synthetic run: seeded Glorot U-Net (no pre-trained blobs exist)
        else:
            if not os.path.exists(perceptual_model):
                raise OSError(f"perceptual model weights not found: {perceptual_model}")
            # Keras HDF5 (`unet.model.save`, unet.py:378-379; weights-only or full-model file) or the native .npz
            self.pm = new_pm().load_dict(load_weights_file(perceptual_model, unet_specs(self.channels)))
        self.params = None
        self._engines = {}
        self.max_engines = 2  # engines (activation buffers + captured graph) kept alive, least recently used evicted
        self.encoder = self.decoder = self.model = None

    # ---- reference API -------------------------------------------------------------------------
    def _set_model(self, weights=None, batch_size=20):
        """lattice_vae.py:127-158: build, 'compile', load weights if the file exists."""
        d = self.input_shape[0]
        self.params = ParamStore(vae_specs(self.channels, self.cond_shape, d, self.latent_dim, self.filters), self.device)
        self.params.init(self.seed)
        self._engines = {}
        self.encoder, self.decoder, self.model = _Encoder(self), _Decoder(self), _Model(self)
        self.batch_size = batch_size
        if weights and os.path.exists(weights):
            self.model.load_weights(weights)
            self.filepath = weights
        elif weights and not os.path.exists(weights):
            self.filepath = weights
        else:
            self.filepath = "saved_models/lattice_dfc_vae_weights.best.hdf5"

    def train(self, train_gen, val_gen, epochs, weights=None):
        """lattice_vae.py:272-342 without the plotting calls."""
        best_loss = np.inf
        self.train_batch_size = train_gen.batch_size
        self.val_batch_size = val_gen.batch_size
        self.batch_size = self.train_batch_size
        self.num_epochs = epochs
        train_steps = int(len(train_gen.list_IDs) / self.train_batch_size)
        val_steps = int(len(val_gen.list_IDs) / self.val_batch_size)
        print("Data size %d,    batch_size %d    steps per epoch %d" % (len(train_gen.list_IDs), self.train_batch_size, train_steps))
        self._set_model(weights)
        self.losses = np.empty((self.num_epochs, 2))
        for e in range(self.num_epochs):
            print("Epoch %s:" % e)
            t0 = time.time()
            tm = np.mean(self.fit_epoch(train_gen, train_steps, train=True), axis=0)
            vm = np.mean(self.fit_epoch(val_gen, val_steps, train=False), axis=0) if val_steps else tm
            s = "Time: %.3f s   " % (time.time() - t0)
            for m in range(4):
                s += "Train %s: %.3f    " % (self.metric_names[m], tm[m])
            for m in range(4):
                s += "Val %s: %.3f    " % (self.metric_names[m], vm[m])
            print(s)
            self.losses[e] = [tm[0], vm[0]]
            if vm[0] < best_loss:
                best_loss = vm[0]
                print("Saving Model")
                self.model.save_weights(self.filepath)
            if hasattr(train_gen, "on_epoch_end"):
                train_gen.on_epoch_end()
        self.model.load_weights(self.filepath)
        self.model.save(os.path.splitext(self.filepath)[0] + ".h5")
        print("Model saved")

    def fit_epoch(self, gen, steps=None, train=True):
        """The batch loop of train() (lattice_vae.py:289-299, 301-305): one train_on_batch (test_on_batch) per gen[b] ->
        (batch, cond), returned as a [steps, 4] array of [loss, perceptual, mse, kld].  Unlike a loop of train_on_batch
        calls it does not stop the GPU between batches: batch b+1 travels host -> device on a copy stream into a second
        staging buffer while step b computes, every step's metrics come back through a pinned ring, and the host waits
        once at the end.  Pinned host batches (torch tensors) are copied asynchronously; numpy / pageable batches work
        and simply copy synchronously."""
        steps = len(gen) if steps is None else int(steps)
        if steps <= 0:
            return np.zeros((0, 4))
        if self.dtype == "fp32":
            return np.array([self._step(*gen[b], train=train) for b in range(steps)])
        M0, c0 = gen[0]
        B = len(M0)
        eng = self.engine(B)
        if train and self.use_cuda_graph and not eng.use_graph:
            eng.M.copy_(_as_f32(M0))         # capture needs valid inputs; no eps draw here, so the random stream of the
            eng.cond.copy_(_as_f32(c0))      # epoch is the one a loop of train_on_batch calls consumes
            eng.capture_train_graph()
        main = torch.cuda.current_stream()
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream, self._stage = torch.cuda.Stream(), {}
        if B not in self._stage:
            self._stage = {B: [(torch.empty_like(eng.M), torch.empty_like(eng.cond), torch.cuda.Event(), torch.cuda.Event())
                               for _ in range(2)]}
        stage, cs = self._stage[B], self._copy_stream
        ring = torch.empty(steps, 4, dtype=torch.float32).pin_memory()
        keep = []  # host tensors of copies in flight

        def prefetch(b):
            Mb, cb = (M0, c0) if b == 0 else gen[b]
            Mh, ch = _as_f32(Mb), _as_f32(cb)
            keep.append((Mh, ch))
            sm, sc, ready, free = stage[b & 1]
            with torch.cuda.stream(cs):
                cs.wait_event(free)          # the step two batches back has consumed this staging buffer
                sm.copy_(Mh, non_blocking=True)
                sc.copy_(ch, non_blocking=True)
                ready.record(cs)

        cs.wait_stream(main)
        prefetch(0)
        for b in range(steps):
            if b + 1 < steps:
                prefetch(b + 1)
            sm, sc, ready, free = stage[b & 1]
            main.wait_event(ready)
            eng.set_inputs(sm, sc)           # device -> device into the engine's static buffers (+ fresh eps)
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

    def save_(self, weights, model="saved_models/vae.h5"):
        self.model.load_weights(weights)
        self.model.save(model)

    def sample_vae(self, n_samples, cond=None, var=1.0):
        """lattice_vae.py:349-357 (incl. its np.tile of the one-hot block)."""
        if cond is None:
            cond = np.random.randint(low=0, high=self.cond_shape, size=n_samples)
        cond_tensor = np.eye(self.cond_shape, dtype=np.float32)[np.atleast_1d(cond)]
        cond_tensor = np.tile(cond_tensor, (n_samples, 1))[: n_samples if np.ndim(cond) else None]
        z_sample = np.random.normal(0, var, size=(len(cond_tensor), self.latent_dim))
        return z_sample, self.decoder.predict([z_sample, cond_tensor])

    # ---- engine plumbing ------------------------------------------------------------------------
    def engine(self, batch) -> VAEEngine:
        if self.params is None:
            self._set_model()
        eng = self._engines.pop(batch, None)
        if eng is None:
            while len(self._engines) >= self.max_engines:  # dicts keep insertion order: the first key is the LRU engine
                self._engines.pop(next(iter(self._engines)))
            eng = VAEEngine(batch, d=self.input_shape[0], channels=self.channels, ncond=self.cond_shape, latent=self.latent_dim,
                            filters=self.filters, device=self.device, vae_params=self.params, pm_params=self.pm,
                            alpha=self.alpha, beta=self.beta, pm_layer_weights=self.pm_layer_weights,
                            lr=self.optimizer.lr, dist=self.dist)
        self._engines[batch] = eng  # (re-)inserted last = most recently used
        return eng

    def _step_x3(self, M, cond, train):
        """train_on_batch / test_on_batch in the fp32-class mode (eps drawn on the device like the bf16 engine)."""
        from ..engine_x3 import VAETrainX3
        if self.params is None:
            self._set_model()
        B = len(M)
        eng = self._x3.get(B)
        if eng is None:
            self._x3.clear()
            eng = self._x3[B] = VAETrainX3(B, d=self.input_shape[0], ncond=self.cond_shape, latent=self.latent_dim,
                                           filters=self.filters, device=self.device, vae_params=self.params, pm_params=self.pm,
                                           alpha=self.alpha, beta=self.beta, pm_layer_weights=self.pm_layer_weights,
                                           lr=self.optimizer.lr)
        eps = torch.randn(B, self.latent_dim, device=self.device)
        M, cond = _to_dev(M, self.device), _to_dev(cond, self.device)
        m = eng.train_step(M, cond, eps) if train else eng.forward(M, cond, eps, training=False)
        return m.cpu().tolist()

    def _step(self, M, cond, train):
        if self.dtype == "fp32":
            return self._step_x3(M, cond, train)
        B = len(M)
        eng = self.engine(B)
        # host batches go straight into the engine's static input buffers (one H2D DMA, no staging allocation)
        eng.set_inputs(_as_f32(M), _as_f32(cond))
        if train:
            if self.use_cuda_graph and not eng.use_graph:
                eng.capture_train_graph()
            eng.train_step()
        else:
            eng.eval_step()
        return eng.metrics_host()

    def _predict(self, a, cond, what):
        n = len(a)
        B = min(n, self.batch_size or 20)
        # partial batches run on an existing larger engine (learning phase 0: samples are independent) instead of
        # allocating a new set of activation buffers per distinct batch size
        fits = [b for b in self._engines if b >= n]
        eng = self.engine(min(fits) if fits else B)
        B = eng.B
        outs = []
        a = _to_dev(a, self.device)
        cond = _to_dev(cond, self.device)
        for s in range(0, n, B):
            e = min(s + B, n)
            k = e - s
            if what == "decode":
                eng.z[:k].copy_(a[s:e])
                eng.cond[:k].copy_(cond[s:e])
                eng.pack_weights()
                eng.decode(False)
                outs.append(eng.xhat[:k].cpu().numpy().copy())
            else:
                eng.M[:k].copy_(a[s:e])
                eng.cond[:k].copy_(cond[s:e])
                eng.eps.normal_()
                from .. import ops
                eng.pack_weights()
                eng.pack_inputs()
                eng.encode(False)
                if what == "encode":
                    outs.append([t[:k].cpu().numpy().copy() for t in (eng.mu, eng.lv, eng.z)])
                else:
                    eng.decode(False)
                    outs.append(eng.xhat[:k].cpu().numpy().copy())
        if what == "encode":
            return [np.concatenate([o[i] for o in outs], axis=0) for i in range(3)]
        return np.concatenate(outs, axis=0)
