#!/usr/bin/env python
"""Convert the MNIST idx files into the dense text format of the LogisticRegression
application: one sample per line, `label v0 v1 ... v783` with pixel values scaled to [0, 1].

    python convert.py /path/to/mnist            # expects train-images-idx3-ubyte[.gz] etc.
"""
import gzip
import os
import struct
import sys

import numpy as np


def read_idx(path):
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        _, dtype, ndim = struct.unpack(">HBB", f.read(4))
        assert dtype == 8, "MNIST idx files hold unsigned bytes"
        shape = struct.unpack(">" + "I" * ndim, f.read(4 * ndim))
        return np.frombuffer(f.read(), dtype=np.uint8).reshape(shape)


def find(d, stem):
    for name in (stem, stem + ".gz", stem.replace("-idx", ".idx"), stem.replace("-idx", ".idx") + ".gz"):
        if os.path.exists(os.path.join(d, name)):
            return os.path.join(d, name)
    raise FileNotFoundError(f"{stem}[.gz] not found in {d}")


def convert(images, labels, out):
    x = read_idx(images).reshape(-1, 28 * 28).astype(np.float32) / 255.0
    y = read_idx(labels)
    with open(out, "w") as f:
        for xi, yi in zip(x, y):
            f.write(str(int(yi)) + " " + " ".join("%g" % v for v in xi) + "\n")
    print(f"{out}: {len(y)} samples")


if __name__ == "__main__":
    d = sys.argv[1] if len(sys.argv) > 1 else "."
    convert(find(d, "train-images-idx3-ubyte"), find(d, "train-labels-idx1-ubyte"), "train.data")
    convert(find(d, "t10k-images-idx3-ubyte"), find(d, "t10k-labels-idx1-ubyte"), "test.data")
