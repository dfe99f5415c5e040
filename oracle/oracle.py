"""ctypes loader for the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module (see oracle/oracle.h).  The product path (vit-fpga_b200/netcuda) never does.

`Oracle`    wraps oracle/liboracle.so  (C restatement: oracle_mlp.c, oracle_vit.c)
`Reference` wraps oracle/_ref/libnetfpga_ref.so (the reference's own unmodified src/netFPGA.cpp
            compiled over the OpenCL shim; present only when it was built in the CPU container)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ACT_RELU_HIDDEN, ACT_RELU_ALL, ACT_NONE = 0, 1, 2


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    if force or not os.path.exists(os.path.join(HERE, "liboracle.so")) or os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class VitCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("image_size", "patch_size", "dim", "depth", "heads", "mlp_dim", "n_classes")]


class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        self.lib.oracle_vit_param_count.restype = C.c_size_t
        self.lib.oracle_gelu.restype = C.c_float
        self.lib.oracle_gelu.argtypes = [C.c_float]
        self.threads = int(self.lib.oracle_max_threads())

    # ---- MLP fp32 -------------------------------------------------------------------------
    def mlp_forward_one(self, x, w_flat, b_flat, npl, n_ins, act=ACT_RELU_HIDDEN):
        x, w_flat, b_flat = _f32(x), _f32(w_flat), _f32(b_flat)
        npl = np.ascontiguousarray(npl, dtype=np.int32)
        out = np.empty(int(npl[-1]), dtype=np.float32)
        self.lib.oracle_mlp_forward_one(_p(x, C.c_float), _p(w_flat, C.c_float), _p(b_flat, C.c_float),
                                        _p(out, C.c_float), _p(npl, C.c_int), C.c_int(len(npl)), C.c_int(n_ins),
                                        C.c_int(act))
        return out

    def mlp_forward(self, x, w_flat, b_flat, npl, n_ins, act=ACT_RELU_HIDDEN, threads=None):
        x, w_flat, b_flat = _f32(x).reshape(-1, n_ins), _f32(w_flat), _f32(b_flat)
        npl = np.ascontiguousarray(npl, dtype=np.int32)
        out = np.empty((x.shape[0], int(npl[-1])), dtype=np.float32)
        self.lib.oracle_mlp_forward(_p(x, C.c_float), C.c_size_t(x.shape[0]), _p(w_flat, C.c_float),
                                    _p(b_flat, C.c_float), _p(out, C.c_float), _p(npl, C.c_int), C.c_int(len(npl)),
                                    C.c_int(n_ins), C.c_int(act), C.c_int(threads or self.threads))
        return out

    def rand_init(self, seed, n_params, n_neurons):
        w = np.empty(n_params, dtype=np.float32)
        b = np.empty(n_neurons, dtype=np.float32)
        self.lib.oracle_rand_init(C.c_uint(seed), _p(w, C.c_float), C.c_size_t(n_params), _p(b, C.c_float),
                                  C.c_size_t(n_neurons))
        return w, b

    # ---- MLP int8 -------------------------------------------------------------------------
    def quantize_q17(self, x):
        x = _f32(x)
        q = np.empty(x.shape, dtype=np.int8)
        self.lib.oracle_quantize_q17(_p(x, C.c_float), C.c_size_t(x.size), _p(q, C.c_int8))
        return q

    def quantize_bias(self, b):
        b = _f32(b)
        q = np.empty(b.shape, dtype=np.int32)
        self.lib.oracle_quantize_bias_q214(_p(b, C.c_float), C.c_size_t(b.size), _p(q, C.c_int32))
        return q

    def mlp_forward_i8(self, xq, wq, bq, npl, n_ins, act=ACT_RELU_HIDDEN, threads=None):
        xq = np.ascontiguousarray(xq, dtype=np.int8).reshape(-1, n_ins)
        wq = np.ascontiguousarray(wq, dtype=np.int8)
        bq = np.ascontiguousarray(bq, dtype=np.int32)
        npl = np.ascontiguousarray(npl, dtype=np.int32)
        out = np.empty((xq.shape[0], int(npl[-1])), dtype=np.int32)
        self.lib.oracle_mlp_forward_i8(_p(xq, C.c_int8), C.c_size_t(xq.shape[0]), _p(wq, C.c_int8), _p(bq, C.c_int32),
                                       _p(out, C.c_int32), _p(npl, C.c_int), C.c_int(len(npl)), C.c_int(n_ins),
                                       C.c_int(act), C.c_int(threads or self.threads))
        return out

    # ---- building blocks --------------------------------------------------------------------
    def linear(self, a, w, bias=None, threads=None):
        a, w = _f32(a), _f32(w)
        m, k = a.shape
        n = w.shape[0]
        out = np.empty((m, n), dtype=np.float32)
        bp = _p(_f32(bias), C.c_float) if bias is not None else None
        self.lib.oracle_linear(_p(a, C.c_float), C.c_size_t(m), C.c_int(k), _p(w, C.c_float), bp, C.c_int(n),
                               _p(out, C.c_float), C.c_int(threads or self.threads))
        return out

    def layernorm(self, x, gamma, beta, eps=1e-6):
        x, gamma, beta = _f32(x), _f32(gamma), _f32(beta)
        y = np.empty_like(x)
        self.lib.oracle_layernorm(_p(x, C.c_float), C.c_size_t(x.shape[0]), C.c_int(x.shape[1]), _p(gamma, C.c_float),
                                  _p(beta, C.c_float), C.c_float(eps), _p(y, C.c_float))
        return y

    def attention(self, qkv, batch, tokens, heads, head_dim=64, threads=None):
        qkv = _f32(qkv)
        out = np.empty((batch * tokens, heads * head_dim), dtype=np.float32)
        self.lib.oracle_attention(_p(qkv, C.c_float), C.c_size_t(batch), C.c_int(tokens), C.c_int(heads),
                                  C.c_int(head_dim), _p(out, C.c_float), C.c_int(threads or self.threads))
        return out

    def gelu(self, x):
        x = _f32(x)
        return np.array([self.lib.oracle_gelu(float(v)) for v in x.ravel()], dtype=np.float32).reshape(x.shape)

    # ---- image side channel ------------------------------------------------------------------
    def filter3x3(self, img):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape
        out = np.empty_like(img)
        self.lib.oracle_filter3x3(_p(img, C.c_uint8), _p(out, C.c_uint8), C.c_int(h), C.c_int(w))
        return out

    # ---- ViT --------------------------------------------------------------------------------
    def vit_param_count(self, cfg: dict) -> int:
        c = VitCfg(**cfg)
        return int(self.lib.oracle_vit_param_count(C.byref(c)))

    def vit_forward(self, cfg: dict, flat, images, threads=None):
        c = VitCfg(**cfg)
        flat, images = _f32(flat), _f32(images)
        assert flat.size == self.vit_param_count(cfg), (flat.size, self.vit_param_count(cfg))
        batch = images.size // (3 * cfg["image_size"] ** 2)
        out = np.empty((batch, cfg["n_classes"]), dtype=np.float32)
        rc = self.lib.oracle_vit_forward(C.byref(c), _p(flat, C.c_float), _p(images, C.c_float), C.c_size_t(batch),
                                         _p(out, C.c_float), C.c_int(threads or self.threads))
        if rc != 0:
            raise ValueError("oracle_vit_forward: bad config")
        return out


class Reference:
    """The reference's own fpga::net_fpga (unmodified source) over the OpenCL shim."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(os.path.join(HERE, "_ref", "libnetfpga_ref.so"))

    def __init__(self):
        self.lib = C.CDLL(os.path.join(HERE, "_ref", "libnetfpga_ref.so"))
        self.lib.ref_net_create.restype = C.c_void_p
        self.lib.ref_net_forward_us.restype = C.c_long
        self.lib.shim_task_count.restype = C.c_ulong

    def create(self, npl, n_ins, w_flat=None, b_flat=None, random=False, seed=1):
        npl = np.ascontiguousarray(npl, dtype=np.int32)
        wp = _p(_f32(w_flat), C.c_float) if w_flat is not None else None
        bp = _p(_f32(b_flat), C.c_float) if b_flat is not None else None
        h = self.lib.ref_net_create(_p(npl, C.c_int), C.c_int(len(npl)), C.c_int(n_ins), wp, bp, C.c_int(int(random)),
                                    C.c_uint(seed))
        if not h:
            raise RuntimeError("ref_net_create failed")
        return C.c_void_p(h)

    def flat(self, h):
        n_params, n_neurons, n_out = C.c_int(), C.c_int(), C.c_int()
        self.lib.ref_net_sizes(h, C.byref(n_params), C.byref(n_neurons), C.byref(n_out))
        w = np.empty(n_params.value, dtype=np.float32)
        b = np.empty(n_neurons.value, dtype=np.float32)
        self.lib.ref_net_flat(h, _p(w, C.c_float), _p(b, C.c_float))
        return w, b, n_out.value

    def forward(self, h, x, n_ins, n_out):
        x = _f32(x).reshape(-1, n_ins)
        out = np.empty((x.shape[0], n_out), dtype=np.float32)
        rc = self.lib.ref_net_forward(h, _p(x, C.c_float), C.c_size_t(x.shape[0]), _p(out, C.c_float))
        if rc != 0:
            raise RuntimeError(f"ref_net_forward rc={rc}")
        return out

    def forward_us(self, h) -> int:
        return int(self.lib.ref_net_forward_us(h))

    def set_activation(self, act: int) -> None:
        self.lib.shim_set_activation(C.c_int(act))

    def destroy(self, h) -> None:
        self.lib.ref_net_destroy(h)
