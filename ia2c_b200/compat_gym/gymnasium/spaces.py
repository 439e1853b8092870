import numpy as np


class Discrete:
    def __init__(self, n, start=0):
        self.n, self.start, self.shape, self.dtype = int(n), int(start), (), np.dtype(np.int64)


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        low, high = np.asarray(low), np.asarray(high)
        self.shape = tuple(low.shape if shape is None else shape)
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(low, self.shape).astype(self.dtype)
        self.high = np.broadcast_to(high, self.shape).astype(self.dtype)
