"""utils.data_split of the reference (utils.py:36-61): seeded train/validation split of the matrices folder,
keeping every `_rot_k` augmentation on the same side as its parent."""
import os
import random


def data_split(path, n=None, frac=0.80, n_rot=10, shuffle=True, seed=28):
    ids = sorted(x for x in os.listdir(path + "/density_matrices") if x.endswith(".npy"))
    plain = [x for x in ids if "_rot_" not in x][:n]
    if shuffle:
        if seed is not None:
            random.seed(seed)
        random.shuffle(plain)
    k = int(frac * len(plain))
    tr_plain, va_plain = plain[:k], plain[k:]
    assert not set(tr_plain) & set(va_plain)
    expand = lambda lst: [y for i in lst for y in [i] + [i[:-4] + "_rot_" + str(r) + ".npy" for r in range(n_rot)]]
    tr, va = expand(tr_plain), expand(va_plain)
    assert not set(tr) & set(va)
    return tr, va
