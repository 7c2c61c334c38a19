"""NumPy restatement of the reference's policy/value net forward pass — TEST INFRASTRUCTURE.

PARITY UNPINNED: the arithmetic lives in Keras 2.1.6 / TensorFlow >= 1.6 (README.md:15-20; un-vendored,
not installable here) and the reference ships no golden outputs for it.  This file restates the graph
built by model.py:58-145 with Keras inference semantics:
  Conv2D = cross-correlation, kernel (kh, kw, cin, cout), bias, channels_last; first conv padding
  'valid' (model.py:62), 3x3 convs in the blocks 'same' (model.py:131-134), 1x1 otherwise;
  BatchNormalization(axis=-1, eps=1e-3): gamma*(x-mean)/sqrt(var+eps)+beta;
  Flatten in (H, W, C) order; Dense x@W+b; relu / tanh; Model.predict applies utils.softmax in float64
  over all 294 logits with no legality mask (model.py:21-24, utils.py:187-192).
Anchors: layer/weight mapping verified in SURVEY.md §8c (start position v = -0.043674, top-5 policy
indices [76,117,101,60,143] for good_model.h5; the agent built on it beats the greedy player 11/12).
"""
import numpy as np

EPS = 1e-3


def _conv(x, k, b, padding):
    """x (B,H,W,Cin); k (kh,kw,Cin,Cout)."""
    kh, kw, cin, cout = k.shape
    if padding == "same" and kh > 1:
        p = kh // 2
        x = np.pad(x, ((0, 0), (p, p), (p, p), (0, 0)))
    B, H, W, _ = x.shape
    oh, ow = H - kh + 1, W - kw + 1
    out = np.zeros((B, oh, ow, cout), dtype=x.dtype)
    for dy in range(kh):
        for dx in range(kw):
            out += x[:, dy:dy + oh, dx:dx + ow, :] @ k[dy, dx]
    return out + b


def _bn(x, w, name):
    g, be, mu, var = (w[name + "/gamma"], w[name + "/beta"], w[name + "/moving_mean"], w[name + "/moving_variance"])
    return g * (x - mu) / np.sqrt(var + EPS) + be


def forward(weights, planes, dtype=np.float64):
    """planes (B,7,7,7) -> (logits (B,294), value (B,)) in `dtype` arithmetic."""
    w = {k: v.astype(dtype) for k, v in weights.items()}
    relu = lambda t: np.maximum(t, 0)

    def cbr(x, i, padding="valid", act=True):
        x = _bn(_conv(x, w["conv2d_%d/kernel" % i], w["conv2d_%d/bias" % i], padding), w, "batch_normalization_%d" % i)
        return relu(x) if act else x

    x = cbr(planes.astype(dtype), 1)                               # model.py:62-64
    for b in range(9):                                             # model.py:66-76, 120-145
        y = cbr(x, 2 + 3 * b)
        y = cbr(y, 3 + 3 * b, padding="same")
        y = cbr(y, 4 + 3 * b, act=False)
        x = relu(y + x)
    p = cbr(x, 29)                                                 # model.py:107-117
    logits = p.reshape(p.shape[0], -1) @ w["policy_head/kernel"] + w["policy_head/bias"]
    v = cbr(x, 30)                                                 # model.py:90-104
    v = relu(v.reshape(v.shape[0], -1) @ w["dense_1/kernel"] + w["dense_1/bias"])
    v = np.tanh(v @ w["value_head/kernel"] + w["value_head/bias"])
    return logits, v[:, 0]


def softmax64(logits):
    """utils.py:187-192"""
    x = logits.astype(np.float64)
    x = x - x.max(axis=-1, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=-1, keepdims=True)


def predict(weights, planes, dtype=np.float32):
    """Model.predict on a batch (model.py:21-24): (p float64 (B,294), v (B,))."""
    logits, v = forward(weights, planes, dtype)
    return softmax64(logits), v
