"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  ctypes access to oracle/_build/libref_ops.so
(plain-C, double-accumulating restatement of the primitive ops; see ref_ops.c)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libref_ops.so"
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB.exists():
            subprocess.run(["make", "-s", "-C", str(HERE)], check=True)
        _lib = C.CDLL(str(LIB))
        _lib.ref_weighted_ce.restype = C.c_double
        _lib.ref_argmax_confusion.restype = C.c_int64
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def conv2d(x, w, b, stride, pad, dil):
    x, w = _f(x), _f(w)
    b = None if b is None else _f(b)
    n, ci, h, wd = x.shape
    co, _, k, _ = w.shape
    ho = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1
    wo = (wd + 2 * pad - dil * (k - 1) - 1) // stride + 1
    y = np.empty((n, co, ho, wo), np.float32)
    lib().ref_conv2d(_p(x), _p(w), _p(b), _p(y), n, ci, h, wd, co, k, stride, pad, dil)
    return y


def conv_transpose2d(x, w, b):
    x, w = _f(x), _f(w)
    b = None if b is None else _f(b)
    n, ci, h, wd = x.shape
    co = w.shape[1]
    y = np.empty((n, co, 2 * h, 2 * wd), np.float32)
    lib().ref_conv_transpose2d(_p(x), _p(w), _p(b), _p(y), n, ci, h, wd, co)
    return y


def bn_eval(x, gamma, beta, mean, var, eps=1e-5):
    x = _f(x)
    n, c = x.shape[:2]
    hw = x.size // (n * c)
    y = np.empty_like(x)
    lib().ref_bn_eval(_p(x), _p(y), n, c, C.c_int64(hw), _p(_f(gamma)), _p(_f(beta)), _p(_f(mean)), _p(_f(var)),
                      C.c_float(eps))
    return y


def bn_train(x, gamma, beta, running_mean, running_var, momentum=0.1, eps=1e-5):
    x = _f(x)
    n, c = x.shape[:2]
    hw = x.size // (n * c)
    y = np.empty_like(x)
    rm, rv = _f(running_mean).copy(), _f(running_var).copy()
    sm, si = np.empty(c, np.float32), np.empty(c, np.float32)
    lib().ref_bn_train(_p(x), _p(y), n, c, C.c_int64(hw), _p(_f(gamma)), _p(_f(beta)), _p(rm), _p(rv),
                       C.c_float(momentum), C.c_float(eps), _p(sm), _p(si))
    return y, rm, rv, sm, si


def maxpool2x2(x):
    x = _f(x)
    n, c, h, w = x.shape
    y = np.empty((n, c, h // 2, w // 2), np.float32)
    idx = np.empty(y.shape, np.int64)
    lib().ref_maxpool2x2(_p(x), _p(y), _p(idx), n, c, h, w)
    return y, idx


def weighted_ce(logits, target, w=None, want_grad=False):
    logits = _f(logits)
    target = np.ascontiguousarray(target, dtype=np.int64)
    n, c = logits.shape[:2]
    hw = logits.size // (n * c)
    d = np.empty_like(logits) if want_grad else None
    w = None if w is None else _f(w)
    loss = lib().ref_weighted_ce(_p(logits), _p(target), _p(w), n, c, C.c_int64(hw), _p(d))
    return (loss, d) if want_grad else loss


def argmax_confusion(logits, target):
    logits = _f(logits)
    target = np.ascontiguousarray(target, dtype=np.int64)
    n, c = logits.shape[:2]
    hw = logits.size // (n * c)
    am = np.empty((n,) + logits.shape[2:], np.int64)
    conf = np.empty((n, c, c), np.int64)
    correct = lib().ref_argmax_confusion(_p(logits), _p(target), n, c, C.c_int64(hw), _p(am), _p(conf))
    return am, conf, int(correct)
