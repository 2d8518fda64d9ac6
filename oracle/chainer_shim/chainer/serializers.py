"""chainer.serializers.load_npz stand-in: '/'-separated keys -> child links."""
import numpy as np


def load_npz(file, obj, path="", strict=True):
    with np.load(file) as z:
        keys = set(z.files)
        for name, link, attr in obj.namedparams():
            key = path + name.lstrip("/")
            if key not in keys:
                if strict:
                    raise KeyError(key)
                continue
            setattr(link, attr, np.ascontiguousarray(z[key]))


def save_npz(file, obj, compression=True):
    d = {name.lstrip("/"): getattr(link, attr) for name, link, attr in obj.namedparams()}
    (np.savez_compressed if compression else np.savez)(file, **d)
