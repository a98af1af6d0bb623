"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol include/icsg3d.h declares,
parameter tables reproduce the reference's counts, the product fails loudly without a GPU, the generators keep the
reference's array contract, the weight container round-trips."""
import ctypes
import os

import numpy as np
import pytest
import torch


def test_library_exports_every_declared_symbol():
    from icsg3d_b200 import _lib
    names = _lib.declared_symbols()
    assert len(names) >= 45
    L = _lib.lib()
    for n in names:
        assert hasattr(L, n), n
    assert L.icsg3d_version() >= 100
    assert isinstance(_lib.last_error(), str)


def test_conv_plan_diagnostic_runs_without_gpu():
    from icsg3d_b200 import _lib
    out = (ctypes.c_int * 10)()
    assert _lib.lib().icsg3d_conv3d_k3_plan(32, 32, 32, 32, 32, 64, 148, out) == 0
    impl, R, TH, T, C, stages = out[0], out[1], out[2], out[3], out[4], out[5]
    # c2 (32->64 @32^3): plane-streaming kd-folded kernel; ring slots x tiles x Cout columns must fit TMEM
    assert impl == 2 and C == 64 and R >= 4 and R * T * C <= 512 and T * 128 >= TH * 33 and stages >= 2
    # 64->64 @16^3 (weights too large to stay resident): the streaming kernel with two 32-channel slices over blockIdx.y
    assert _lib.lib().icsg3d_conv3d_k3_plan(32, 16, 16, 16, 64, 64, 148, out) == 0
    assert out[0] == 2 and out[4] == 32
    # a deep-K layer (128->64 @16^3) stays on the halo-reuse kernel
    assert _lib.lib().icsg3d_conv3d_k3_plan(32, 16, 16, 16, 128, 64, 148, out) == 0
    assert out[0] == 1 and out[4] == 64 and 2 * out[3] * out[4] <= 512
    # invalid arguments are reported through the error string, not a crash
    assert _lib.lib().icsg3d_conv3d_k3_plan(32, 32, 32, 32, 32, 64, 0, out) != 0
    assert "plan" in _lib.last_error()


def test_parameter_tables_match_reference_counts():
    """838,832 trainable VAE parameters (+968 BN moving stats), 31,156,800 U-Net, 12,336,160 in c1..c10."""
    from icsg3d_b200.params import ParamStore, unet_specs, vae_specs
    v = ParamStore(vae_specs(), "cpu", with_grads=False, with_adam=False).init(1)
    u = ParamStore(unet_specs(), "cpu", with_grads=False, with_adam=False).init(2)
    assert v.count_trainable() == 838832 and u.count_trainable() == 31156800
    assert sum(v.p[k].numel() for k in v.names(trainable=False)) == 968
    pm = sum(u.p[f"{n}/{w}"].numel() for n in ("c1", "c2", "c3", "c4", "c5", "c6", "c9", "c10") for w in ("kernel", "bias"))
    assert pm == 12336160
    assert tuple(v.p["enc_conv1/kernel"].shape) == (3, 3, 3, 44, 16)       # K.tile -> 4 + 4*10 input channels (R1)
    assert tuple(v.p["dec_dense/kernel"].shape) == (266, 256)
    # glorot_uniform limits (R5, R8)
    k = v.p["enc_conv1/kernel"]
    assert float(k.abs().max()) <= (6.0 / (27 * 44 + 27 * 16)) ** 0.5 + 1e-7
    assert float(v.p["enc_bn1/gamma"].min()) == 1.0 and float(v.p["enc_bn1/moving_variance"].min()) == 1.0


def test_oracle_and_product_param_names_agree():
    from icsg3d_b200.params import unet_specs, vae_specs
    from oracle import nets
    pv, pu = nets.init_vae_params(1), nets.init_unet_params(1)
    assert [n for n, *_ in vae_specs()] == list(pv.keys())
    assert [n for n, *_ in unet_specs()] == list(pu.keys())
    for n, shape, *_ in vae_specs():
        assert tuple(pv[n].shape) == tuple(shape)


def test_ops_fail_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from icsg3d_b200 import _lib, ops
    x = torch.zeros(1, 2, 2, 2, 16, dtype=torch.bfloat16)
    w = torch.zeros(27, 16, 16, dtype=torch.bfloat16)
    with pytest.raises(_lib.Icsg3dError):
        ops.conv3d_k3(x, w)


def test_weight_container_roundtrip(tmp_path):
    from icsg3d_b200.weights_io import load_npz, save_npz
    d = {"enc_conv1/kernel": np.random.rand(3, 3, 3, 4, 2).astype(np.float32), "enc_bn1/gamma": np.ones(2, np.float32)}
    path = str(tmp_path / "sub" / "vae_weights_x.best.hdf5")
    save_npz(path, d)
    assert os.path.exists(path)  # exact path, no extension appended
    back = load_npz(path)
    assert set(back) == set(d) and all(np.array_equal(back[k], d[k]) for k in d)


def test_generators_keep_reference_contract(tmp_path):
    import pandas as pd
    from icsg3d_b200.datasplit import data_split
    from icsg3d_b200.unet.data import UnetDataGenerator
    from icsg3d_b200.vae.data import VAEDataGenerator
    root = tmp_path / "matrices"
    for sub in ("density_matrices", "coordinate_grids", "species_matrices"):
        (root / sub).mkdir(parents=True)
    ids = [f"mp-{i}" for i in range(6)]
    d = 8
    for i in ids:
        for r in [""] + [f"_rot_{k}" for k in range(2)]:
            np.save(root / "density_matrices" / f"{i}{r}.npy", np.random.rand(d, d, d))
            np.save(root / "coordinate_grids" / f"{i}{r}.npy", np.random.rand(d, d, d, 3))
            np.save(root / "species_matrices" / f"{i}{r}.npy", np.random.randint(0, 95, (d, d, d)).astype(np.float64))
    csv = tmp_path / "props.csv"
    pd.DataFrame({"task_id": ids, "formation_energy_per_atom": np.linspace(-3, 0, 6), "nsites": 5}).to_csv(csv, index=False)
    tr, va = data_split(str(root), None, frac=0.5, n_rot=2)
    assert len(tr) == 9 and len(va) == 9 and not set(tr) & set(va)
    g = VAEDataGenerator(tr, str(root), batch_size=3, dim=(d, d, d), property_csv=str(csv), n_bins=3)
    M, cond = g[0]
    assert len(g) == 3 and M.shape == (3, d, d, d, 4) and M.dtype == np.float64 and cond.shape == (3, 3)
    assert np.all(cond.sum(1) == 1) and len(g.list_IDs_temp) == 3
    gs = VAEDataGenerator(tr, str(root), batch_size=3, dim=(d, d, d), property_csv=str(csv), n_bins=3, return_S=True)
    M2, (cond2, S1h, Sb) = gs[0]  # vae/data.py:72-86: [cond, to_categorical(S), S != 0]
    assert np.array_equal(M2, M) and np.array_equal(cond2, cond)
    assert S1h.shape == (3, d, d, d, 95) and Sb.shape == (3, d, d, d, 1)
    assert np.array_equal(Sb[..., 0], (S1h.argmax(-1) != 0).astype(np.float64))
    u = UnetDataGenerator(va, str(root), batch_size=2, dim=(d, d, d), n_channels=4)
    X, (y, b) = u[0]
    assert X.shape == (2, d, d, d, 4) and y.shape == (2, d, d, d, 95) and b.shape == (2, d, d, d, 1)
    assert np.array_equal(b[..., 0], (y.argmax(-1) != 0).astype(np.float32))


def test_rot90_transform_composes_like_numpy_rot90():
    """Host logic of the rotation augmentation (utils.py:193-222): the composed signed permutation the kernel applies
    equals chained np.rot90(., 1, axes) = scipy.ndimage.rotate(., 90, axes, reshape=False) for all 27 axis sequences."""
    import itertools
    from icsg3d_b200 import utils
    d = 5
    X = np.random.default_rng(0).random((d, d, d))
    o = np.indices((d, d, d))
    for seq in itertools.product(utils.ROT_AXES, repeat=3):
        Y = X
        for ax in seq:
            Y = np.rot90(Y, 1, axes=ax)
        perm, flip = utils.rot90_transform(seq)
        s = [(d - 1 - o[perm[x]]) if flip[x] else o[perm[x]] for x in range(3)]
        assert np.array_equal(Y, X[s[0], s[1], s[2]]), seq


def test_conv_dispatch_plans_respect_hardware_limits():
    """Host-only sweep of the conv dispatcher's planner (icsg3d_conv3d_k3_plan) over the layer shapes of both networks
    at several batch sizes / grid edges: every plan must fit TMEM (512 columns), shared memory (227 KB) and the SM count."""
    from icsg3d_b200 import _lib
    L = _lib.lib()
    out = (ctypes.c_int * 10)()
    shapes = [(16, 16), (16, 32), (32, 16), (32, 64), (64, 32), (64, 64), (64, 128), (128, 64), (128, 128), (128, 256),
              (256, 128), (256, 512), (512, 512), (48, 16), (96, 64), (192, 128), (768, 512), (384, 256)]
    seen = set()
    for B in (1, 2, 8, 32, 128):
        for D in (4, 8, 16, 32, 64):
            for cin, cout in shapes:
                assert L.icsg3d_conv3d_k3_plan(B, D, D, D, cin, cout, 148, out) == 0, _lib.last_error()
                impl = out[0]
                seen.add(impl)
                if impl == 2:  # plane-streaming: {2, R, TH, T, C, stages, issuers, grid, kc, smem}
                    R, TH, T, C, stages, issuers, grid, kc, smem = (out[i] for i in range(1, 10))
                    assert R in (4, 8) and R * T * C <= 512 and T * 128 >= TH * (D + 1) and 1 <= issuers <= 3
                    assert stages in (2, 4) and 1 <= grid <= 148 and smem <= 225 * 1024 and D >= 16 and C in (16, 32, 64)
                elif impl == 1:  # halo reuse: {1, TD, TH, G, NT, a_bufs, b_stages, items, kc, smem}
                    TD, TH, G, NT, a_bufs, b_stages, items, kc, smem = (out[i] for i in range(1, 10))
                    assert 2 * G * NT <= 512 and smem <= 225 * 1024 and D % TH == 0 and items >= 1 and D >= 8
    assert seen == {0, 1, 2}


def test_bn_backward_grids_are_one_wave():
    """Host-only contract of the BatchNorm backward partial counts (one wave of resident blocks on 148 SMs; no device
    needed): bf16 reduce = 3 blocks/SM (plain, max-pool), apply = 2 blocks/SM, upsample / fp32 keep the 4-per-SM cap."""
    from icsg3d_b200 import _lib
    L = _lib.lib()
    from icsg3d_b200 import ops
    BF16, F32, NONE, POOL, UP = ops.DT_BF16, ops.DT_F32, ops.POST_NONE, ops.POST_POOL2, ops.POST_UP2
    assert L.icsg3d_bn_bwd_nparts(32, 32, 32, 32, 64, BF16, POOL) == 148 * 3      # pm.c2
    assert L.icsg3d_bn_bwd_nparts(32, 32, 32, 32, 32, BF16, NONE) == 148 * 3      # pm.c1
    assert L.icsg3d_bn_bwd_apply_nblocks(32, 32, 32, 32, 64, BF16, POOL) == 148 * 2
    assert L.icsg3d_bn_bwd_apply_nblocks(32, 32, 32, 32, 32, BF16, NONE) == 148 * 2
    assert L.icsg3d_bn_bwd_nparts(32, 16, 16, 16, 32, BF16, UP) == 148 * 4        # dec3: generic kernel
    assert L.icsg3d_bn_bwd_nparts(32, 32, 32, 32, 4, F32, NONE) == 148 * 4        # dec_bn5 (fp32 logits)
    # small layers: one block per rows-per-block group, never more than the rows
    assert L.icsg3d_bn_bwd_nparts(2, 4, 4, 4, 512, BF16, NONE) == 2 * 64 // (256 // 64)
    assert L.icsg3d_bn_bwd_nparts(2, 4, 4, 4, 12, BF16, NONE) < 0                 # C must be a multiple of the vector width
