"""Mirror of the reference's model.py (inference side): `ResidualCNN().load_weights(path)` then
`.predict(x)` -> (p[294] float64 soft-maxed, v) exactly like Model.predict (model.py:21-24), plus
`predict_batch` for (B,7,7,7) inputs.  The arithmetic runs in libccx.so's CUDA kernels; there is no CPU path.

Weights: Keras `save_weights` HDF5 files (good_model.h5 ...) are read with h5lite (no h5py needed);
`.npz` files holding the same '<layer>/<param>' arrays are accepted too.  BatchNormalization
(eps 1e-3, Keras default) is folded into the preceding conv at load time."""
import ctypes

import numpy as np
import torch

from . import _lib, h5lite
from .config import DTYPE_BF16, DTYPE_F32, DTYPE_U8, INPUT_DIM, NUM_ACTIONS, NUM_FILTERS

BN_EPS = 1e-3


def read_weight_file(path):
    if str(path).endswith(".npz"):
        return {k: np.asarray(v, dtype=np.float32) for k, v in np.load(path).items()}
    return h5lite.read_weights(path)


def _fold(w, i):
    """conv2d_i + batch_normalization_i -> (K x Cout matrix, bias), float64 folding then float32."""
    k = w["conv2d_%d/kernel" % i].astype(np.float64)
    b = w["conv2d_%d/bias" % i].astype(np.float64)
    g, be = w["batch_normalization_%d/gamma" % i].astype(np.float64), w["batch_normalization_%d/beta" % i].astype(np.float64)
    mu, var = (w["batch_normalization_%d/moving_mean" % i].astype(np.float64),
               w["batch_normalization_%d/moving_variance" % i].astype(np.float64))
    s = g / np.sqrt(var + BN_EPS)
    kh, kw, cin, cout = k.shape
    return (k * s).reshape(kh * kw * cin, cout), (b - mu) * s + be


def pack_weights(w):
    """Keras tensors -> the flat fp32 blob ccx_net_load expects (layout: csrc/ccx_net.cu header)."""
    parts = []

    def add(i):
        m, b = _fold(w, i)
        parts.extend([m.ravel(), b.ravel()])
    add(1)                                     # model.py:62
    for blk in range(9):                       # model.py:66-76
        add(2 + 3 * blk); add(3 + 3 * blk); add(4 + 3 * blk)
    add(29)                                    # policy conv, model.py:108
    parts.extend([w["policy_head/kernel"].ravel(), w["policy_head/bias"].ravel()])
    add(30)                                    # value conv, model.py:91
    parts.extend([w["dense_1/kernel"].ravel(), w["dense_1/bias"].ravel(),
                  w["value_head/kernel"].ravel(), w["value_head/bias"].ravel()])
    return np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64) for p in parts]).astype(np.float32))


class Model:
    """model.py:15-47"""

    def __init__(self, input_dim, filters, version=0):
        self.input_dim, self.filters, self.version = input_dim, filters, version


class ResidualCNN(Model):
    """model.py:52-145 (inference).  `engine` is an engine.Engine; one is created on cuda:0 if omitted."""

    def __init__(self, input_dim=INPUT_DIM, filters=NUM_FILTERS, engine=None):
        Model.__init__(self, input_dim, filters)
        from .engine import Engine
        self.eng = engine or Engine(0)
        self.loaded = False

    def load_weights(self, filepath):
        packed = pack_weights(read_weight_file(filepath))
        if packed.size != self.eng.L.ccx_net_num_weights():
            raise _lib.CcxError("weight file does not describe the 9-block ResidualCNN")
        self.eng.call("ccx_net_load", ctypes.c_void_p(packed.ctypes.data), packed.size)
        self.loaded = True
        return self

    # -- batched inference ---------------------------------------------------------------------------
    def forward(self, planes):
        """planes: (B,7,7,7) torch tensor on the engine's device, dtype uint8 / bfloat16 / float32.
        Returns (logits float32 (B,294), value float32 (B,))."""
        dt = {torch.uint8: DTYPE_U8, torch.bfloat16: DTYPE_BF16, torch.float32: DTYPE_F32}[planes.dtype]
        planes = planes.contiguous()
        n = planes.shape[0]
        logits = self.eng.empty((n, NUM_ACTIONS), torch.float32)
        value = self.eng.empty((n,), torch.float32)
        self.eng.call("ccx_net_forward", n, ctypes.c_void_p(planes.data_ptr()), dt,
                      ctypes.c_void_p(logits.data_ptr()), ctypes.c_void_p(value.data_ptr()))
        return logits, value

    def predict_batch(self, planes):
        """Model.predict for a batch: (p float64 (B,294) soft-maxed over all logits, v float64 (B,))."""
        logits, value = self.forward(planes)
        n = logits.shape[0]
        p = self.eng.empty((n, NUM_ACTIONS), torch.float64)
        v = self.eng.empty((n,), torch.float64)
        self.eng.call("ccx_softmax_f64", n, ctypes.c_void_p(logits.data_ptr()), ctypes.c_void_p(value.data_ptr()),
                      ctypes.c_void_p(p.data_ptr()), ctypes.c_void_p(v.data_ptr()))
        return p, v

    def predict(self, input_board):
        """model.py:21-24 — one (7,7,7) array in, (p[294], v) numpy out."""
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(input_board, dtype=np.float32))[None]).to(self.eng.device)
        p, v = self.predict_batch(x)
        return p[0].cpu().numpy(), v[0].cpu().numpy()

    def evaluate_states(self, leaf_state):
        """to_model_input + predict for packed states (5+, n) int64 on the device (MCTS.py:93)."""
        n = leaf_state.shape[1]
        p = self.eng.empty((n, NUM_ACTIONS), torch.float64)
        v = self.eng.empty((n,), torch.float64)
        self.eng.call("ccx_net_eval", n, ctypes.c_void_p(leaf_state.data_ptr()), ctypes.c_void_p(p.data_ptr()),
                      ctypes.c_void_p(v.data_ptr()))
        return p, v
