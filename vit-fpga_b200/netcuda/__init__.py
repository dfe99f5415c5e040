"""ctypes binding of the netCUDA C ABI (include/netcuda.h) and of the C++ class driver.

This module is plumbing for tests/ and bench.py: the product is the native library.  It never
imports the CPU oracle and has no fallback -- if libnetcuda.so is missing, importing fails; if no
sm_100 GPU is visible, `Net(...)` raises `NetcudaError` carrying the library's own message.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
LIB_DIR = os.environ.get("NETCUDA_LIB_DIR") or os.path.join(PKG, "lib")  # (override: A/B of two builds inside one GPU call)
HEADER = os.path.join(ROOT, "include", "netcuda.h")

KIND_MLP, KIND_VIT = 0, 1
PREC_FP32, PREC_TF32, PREC_BF16, PREC_INT8 = 0, 1, 2, 3
ACT_RELU_HIDDEN, ACT_RELU_ALL, ACT_NONE = 0, 1, 2
OUT_F32, OUT_BF16, OUT_S8, OUT_S32 = 0, 1, 2, 3
EPI_NONE, EPI_RELU, EPI_GELU, EPI_RESIDUAL, EPI_REQUANT = 0, 1, 2, 3, 4
OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_KERNEL, ERR_RING_FULL, ERR_RING_EMPTY = 0, 1, 2, 3, 4, 5, 6, 7
RING_DEPTH = 24
PRECISIONS = {"fp32": PREC_FP32, "tf32": PREC_TF32, "bf16": PREC_BF16, "int8": PREC_INT8}


class NetcudaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"netcuda error {code}: {message}")
        self.code = code


class Desc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("precision", C.c_int32), ("device", C.c_int32), ("activation", C.c_int32),
                ("max_batch", C.c_int32), ("n_ins", C.c_int32), ("n_layers", C.c_int32), ("n_p_l", C.POINTER(C.c_int32)),
                ("image_size", C.c_int32), ("patch_size", C.c_int32), ("dim", C.c_int32), ("depth", C.c_int32),
                ("heads", C.c_int32), ("mlp_dim", C.c_int32), ("n_classes", C.c_int32)]


FILE_F32, FILE_Q17, FILE_MAX_LAYERS = 0, 1, 64


class FileInfo(C.Structure):
    _fields_ = [("desc", Desc), ("n_p_l", C.c_int32 * FILE_MAX_LAYERS), ("dtype", C.c_int32), ("n_weights", C.c_uint64),
                ("n_biases", C.c_uint64)]


class KernelStat(C.Structure):
    _fields_ = [("label", C.c_char * 32), ("launches", C.c_uint64), ("ms", C.c_double), ("flops", C.c_double), ("bytes", C.c_double)]


def declared_symbols() -> list[str]:
    """Every function include/netcuda.h declares (used by the CPU test that checks the exports)."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(netcuda_[a-z0-9_]+)\s*\(", text)))


def _load(name: str) -> C.CDLL:
    path = os.path.join(LIB_DIR, name)
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `python vit-fpga_b200/build.py` (there is no fallback path)")
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


lib = _load("libnetcuda.so")
lib.netcuda_last_error.restype = C.c_char_p
_hostcls = _load("libnetcuda_host.so")   # cuda::net_cuda (the shipped class library)
hostlib = _load("libnetcuda_hostdrv.so")  # its ctypes driver (test / bench plumbing only)
hostlib.nch_last_error.restype = C.c_char_p
hostlib.nch_mlp_create.restype = C.c_void_p
hostlib.nch_vit_create.restype = C.c_void_p
hostlib.nch_load.restype = C.c_void_p
hostlib.nch_launch_forward.restype = C.c_longlong
hostlib.nch_launch_forward_frame.restype = C.c_longlong
hostlib.nch_get_filtered_image.restype = C.c_longlong
hostlib.nch_forward_us.restype = C.c_long
hostlib.nch_gradient_us.restype = C.c_long


def _check(rc: int) -> None:
    if rc != OK:
        raise NetcudaError(rc, lib.netcuda_last_error().decode())


def device_count() -> int:
    n = C.c_int(0)
    rc = lib.netcuda_device_count(C.byref(n))
    return n.value if rc == OK else 0


def host_register(a: np.ndarray) -> None:
    """Page-lock a numpy array's buffer (netcuda_host_register): forward / submit then DMA straight from / into it."""
    rc = lib.netcuda_host_register(C.c_void_p(a.ctypes.data), C.c_size_t(a.nbytes))
    if rc != OK:
        raise NetcudaError(rc, lib.netcuda_last_error().decode())


def host_unregister(a: np.ndarray) -> None:
    rc = lib.netcuda_host_unregister(C.c_void_p(a.ctypes.data))
    if rc != OK:
        raise NetcudaError(rc, lib.netcuda_last_error().decode())


def _ptr(t) -> C.c_void_p:
    """Device/host pointer of a torch tensor or numpy array."""
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, np.ndarray):
        return C.c_void_p(t.ctypes.data)
    return C.c_void_p(t.data_ptr())


CUDA_STREAM_LEGACY = 1  # cudaStreamLegacy: names the default stream explicitly (a NULL stream argument means "the handle's own stream")


def _stream(stream) -> C.c_void_p:
    """None -> NULL (the handle's own stream); a torch stream / raw handle -> that stream.  torch's default stream has the raw
    handle 0, which the C ABI would read as NULL, so it is passed as cudaStreamLegacy."""
    if stream is None:
        return C.c_void_p(0)
    raw = int(getattr(stream, "cuda_stream", stream))
    return C.c_void_p(raw if raw != 0 else CUDA_STREAM_LEGACY)


import sys as _sys

if PKG not in _sys.path:
    _sys.path.insert(0, PKG)
import vit_presets as _presets  # pure Python (no dlopen): shared with bench.py's reference arm
from vit_presets import VIT_PRESETS  # noqa: E402,F401


def vit_param_count(cfg: dict) -> int:
    d = Desc(kind=KIND_VIT, precision=PREC_BF16, **cfg)
    n = C.c_size_t(0)
    _check(lib.netcuda_vit_param_count(C.byref(d), C.byref(n)))
    return n.value


def vit_random_params(cfg: dict, seed: int = 0) -> np.ndarray:
    """vit_presets.vit_random_params, with the size checked against the library's own count."""
    flat = _presets.vit_random_params(cfg, seed)
    assert flat.size == vit_param_count(cfg)
    return flat


# ---- weight files (host only: no GPU needed) -----------------------------------------------------------

def _npl_arr(npl):
    return (C.c_int32 * len(npl))(*[int(v) for v in npl])


def file_write_mlp(path, npl, n_ins, w_flat, b_flat, activation=ACT_RELU_HIDDEN) -> None:
    w = np.ascontiguousarray(w_flat, dtype=np.float32)
    b = np.ascontiguousarray(b_flat, dtype=np.float32)
    _check(lib.netcuda_file_write_mlp(os.fsencode(path), _npl_arr(npl), C.c_int(len(npl)), C.c_int(int(n_ins)), C.c_int(activation),
                                      _ptr(w), _ptr(b)))


def file_write_mlp_i8(path, npl, n_ins, wq, bq, activation=ACT_RELU_HIDDEN) -> None:
    w = np.ascontiguousarray(wq, dtype=np.int8)
    b = np.ascontiguousarray(bq, dtype=np.int32)
    _check(lib.netcuda_file_write_mlp_i8(os.fsencode(path), _npl_arr(npl), C.c_int(len(npl)), C.c_int(int(n_ins)), C.c_int(activation),
                                         _ptr(w), _ptr(b)))


def file_write_vit(path, cfg: dict, flat) -> None:
    f = np.ascontiguousarray(flat, dtype=np.float32)
    d = Desc(kind=KIND_VIT, precision=PREC_BF16, **cfg)
    _check(lib.netcuda_file_write_vit(os.fsencode(path), C.byref(d), _ptr(f), C.c_size_t(f.size)))


def file_info(path) -> dict:
    info = FileInfo()
    _check(lib.netcuda_file_info_read(os.fsencode(path), C.byref(info)))
    d = info.desc
    out = dict(kind=d.kind, dtype=info.dtype, precision=d.precision, activation=d.activation, n_ins=d.n_ins,
               npl=[int(info.n_p_l[i]) for i in range(d.n_layers)], n_weights=int(info.n_weights), n_biases=int(info.n_biases))
    if d.kind == KIND_VIT:
        out["cfg"] = {k: int(getattr(d, k)) for k in ("image_size", "patch_size", "dim", "depth", "heads", "mlp_dim", "n_classes")}
    return out


def file_read(path):
    """(weights, biases) as numpy arrays of the file's dtype (biases is empty for a ViT)."""
    info = file_info(path)
    q = info["dtype"] == FILE_Q17
    w = np.empty(info["n_weights"], dtype=np.int8 if q else np.float32)
    b = np.empty(info["n_biases"], dtype=np.int32 if q else np.float32)
    _check(lib.netcuda_file_read(os.fsencode(path), _ptr(w), C.c_size_t(w.nbytes), _ptr(b), C.c_size_t(b.nbytes)))
    return w, b


class Net:
    """One net on one GPU through the C ABI."""

    def __init__(self, desc: Desc, keepalive=None, handle=None):
        self._h = C.c_void_p(0)
        self._keep = keepalive
        if handle is not None:
            self._h = handle
        else:
            _check(lib.netcuda_create(C.byref(desc), C.byref(self._h)))
        n = C.c_size_t(0)
        _check(lib.netcuda_n_in(self._h, C.byref(n)))
        self.n_in = n.value
        _check(lib.netcuda_n_out(self._h, C.byref(n)))
        self.n_out = n.value
        self.precision = desc.precision
        self.kind = desc.kind

    @classmethod
    def mlp(cls, npl, n_ins, precision=PREC_BF16, device=0, activation=ACT_RELU_HIDDEN, max_batch=0) -> "Net":
        arr = (C.c_int32 * len(npl))(*[int(v) for v in npl])
        d = Desc(kind=KIND_MLP, precision=precision, device=device, activation=activation, max_batch=max_batch, n_ins=int(n_ins),
                 n_layers=len(npl), n_p_l=C.cast(arr, C.POINTER(C.c_int32)))
        return cls(d, keepalive=arr)

    @classmethod
    def vit(cls, cfg: dict, device=0, max_batch=0, precision=PREC_BF16) -> "Net":
        """precision: PREC_BF16 (default) or PREC_TF32 (fp32 weights / activations, kind::tf32 MMAs in every linear layer)."""
        d = Desc(kind=KIND_VIT, precision=precision, device=device, max_batch=max_batch, **cfg)
        return cls(d)

    @classmethod
    def from_file(cls, path, precision=-1, device=0, max_batch=0) -> "Net":
        """netcuda_create_from_file: build the net a weight file describes and upload its weights."""
        h = C.c_void_p(0)
        _check(lib.netcuda_create_from_file(os.fsencode(path), C.c_int(precision), C.c_int(device), C.c_int(max_batch), C.byref(h)))
        info = file_info(path)
        d = Desc(kind=info["kind"], precision=info["precision"] if precision < 0 else precision)
        return cls(d, handle=h)

    def close(self) -> None:
        if self._h:
            lib.netcuda_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights
    def upload_mlp(self, w_flat, b_flat) -> None:
        w = np.ascontiguousarray(w_flat, dtype=np.float32)
        b = np.ascontiguousarray(b_flat, dtype=np.float32)
        _check(lib.netcuda_upload_mlp(self._h, _ptr(w), _ptr(b)))

    def upload_mlp_i8(self, wq, bq) -> None:
        w = np.ascontiguousarray(wq, dtype=np.int8)
        b = np.ascontiguousarray(bq, dtype=np.int32)
        _check(lib.netcuda_upload_mlp_i8(self._h, _ptr(w), _ptr(b)))

    def upload_vit(self, flat) -> None:
        f = np.ascontiguousarray(flat, dtype=np.float32)
        _check(lib.netcuda_upload_vit(self._h, _ptr(f), C.c_size_t(f.size)))

    # ---- forward
    def forward(self, x) -> np.ndarray:
        """Host fp32 in -> host fp32 out.  Accepts numpy arrays or (pinned) CPU torch tensors."""
        if isinstance(x, np.ndarray):
            x = np.ascontiguousarray(x, dtype=np.float32)
            batch = x.size // self.n_in
            assert batch * self.n_in == x.size
            out = np.empty((batch, self.n_out), dtype=np.float32)
            _check(lib.netcuda_forward(self._h, _ptr(x), C.c_size_t(batch), _ptr(out)))
            return out
        raise TypeError("forward expects a numpy array; use forward_into for torch tensors")

    def forward_into(self, x, out) -> None:
        """Host torch tensors (ideally pinned): x fp32 [batch, n_in] -> out fp32 [batch, n_out]."""
        batch = x.numel() // self.n_in
        _check(lib.netcuda_forward(self._h, _ptr(x), C.c_size_t(batch), _ptr(out)))

    def forward_i8(self, xq, out=None) -> np.ndarray:
        xq = np.ascontiguousarray(xq, dtype=np.int8)
        batch = xq.size // self.n_in
        if out is None:
            out = np.empty((batch, self.n_out), dtype=np.int32)
        assert out.dtype == np.int32 and out.size == batch * self.n_out and out.flags.c_contiguous
        _check(lib.netcuda_forward_i8(self._h, _ptr(xq), C.c_size_t(batch), _ptr(out)))
        return out

    # ---- u8 frames (ViT)
    def set_u8_normalization(self, mean, std) -> None:
        m = (C.c_float * 3)(*[float(v) for v in mean])
        sd = (C.c_float * 3)(*[float(v) for v in std])
        _check(lib.netcuda_set_u8_normalization(self._h, m, sd))

    def forward_u8(self, frames) -> np.ndarray:
        """frames: uint8 [batch, H, W, 3] (numpy or CPU torch) -> fp32 logits."""
        f = np.ascontiguousarray(frames, dtype=np.uint8) if isinstance(frames, np.ndarray) else frames
        batch = int(f.shape[0])
        out = np.empty((batch, self.n_out), dtype=np.float32)
        _check(lib.netcuda_forward_u8(self._h, _ptr(f), C.c_size_t(batch), _ptr(out)))
        return out

    def submit_u8(self, frames, out) -> int:
        t = C.c_uint64(0)
        _check(lib.netcuda_submit_u8(self._h, _ptr(frames), C.c_size_t(int(frames.shape[0])), _ptr(out), C.byref(t)))
        return t.value

    def forward_device_u8(self, d_frames, d_out, batch: int, stream=None) -> None:
        _check(lib.netcuda_forward_device_u8(self._h, _ptr(d_frames), C.c_size_t(batch), _ptr(d_out), _stream(stream)))

    def submit(self, x, out) -> int:
        """netcuda_submit: non-blocking host-buffer forward; `x` / `out` (numpy or CPU torch, ideally pinned) must stay alive
        until wait(ticket)."""
        t = C.c_uint64(0)
        n = (x.size if isinstance(x, np.ndarray) else x.numel()) // self.n_in
        _check(lib.netcuda_submit(self._h, _ptr(x), C.c_size_t(n), _ptr(out), C.byref(t)))
        return t.value

    def wait(self, ticket: int) -> None:
        _check(lib.netcuda_wait(self._h, C.c_uint64(ticket)))

    def query(self, ticket: int) -> bool:
        d = C.c_int(0)
        _check(lib.netcuda_query(self._h, C.c_uint64(ticket), C.byref(d)))
        return bool(d.value)

    def forward_device(self, d_in, d_out, batch: int, stream=None) -> None:
        """Device tensors; asynchronous on `stream` (torch stream or raw handle)."""
        _check(lib.netcuda_forward_device(self._h, _ptr(d_in), C.c_size_t(batch), _ptr(d_out), _stream(stream)))

    def forward_device_i8(self, d_in, d_out, batch: int, stream=None) -> None:
        _check(lib.netcuda_forward_device_i8(self._h, _ptr(d_in), C.c_size_t(batch), _ptr(d_out), _stream(stream)))

    # ---- introspection
    @property
    def launches(self) -> int:
        n = C.c_uint64(0)
        _check(lib.netcuda_launch_count(self._h, C.byref(n)))
        return n.value

    @property
    def flops_per_sample(self) -> float:
        f = C.c_double(0)
        _check(lib.netcuda_flops_per_sample(self._h, C.byref(f)))
        return f.value

    @property
    def last_forward_us(self) -> int:
        n = C.c_int64(0)
        _check(lib.netcuda_last_forward_us(self._h, C.byref(n)))
        return n.value

    def profile_enable(self, on: bool) -> None:
        _check(lib.netcuda_profile_enable(self._h, C.c_int(int(on))))

    def profile_read(self) -> dict:
        """{label: {"launches", "ms", "flops", "bytes"}} for everything launched since the last read."""
        buf = (KernelStat * 64)()
        n = C.c_int(0)
        _check(lib.netcuda_profile_read(self._h, buf, C.c_int(64), C.byref(n)))
        return {buf[i].label.decode(): dict(launches=int(buf[i].launches), ms=buf[i].ms, flops=buf[i].flops, bytes=buf[i].bytes)
                for i in range(min(n.value, 64))}

    def set_gemm_variant(self, variant: int) -> None:
        _check(lib.netcuda_set_gemm_variant(self._h, C.c_int(variant)))


# ---- single-kernel entry points (torch CUDA tensors) -------------------------------------------------

def op_gemm(a, w, bias, out, precision, out_type, epilogue=EPI_NONE, variant=0, m=None, n=None, k=None, lda=None, ldw=None,
            ldc=None, device=0, stream=None) -> None:
    m = a.shape[0] if m is None else m
    n = w.shape[0] if n is None else n
    k = a.shape[1] if k is None else k
    lda = a.stride(0) if lda is None else lda
    ldw = w.stride(0) if ldw is None else ldw
    ldc = out.stride(0) if ldc is None else ldc
    _check(lib.netcuda_op_gemm(C.c_int(device), C.c_int(precision), C.c_int(variant), _ptr(a), C.c_int(lda), _ptr(w), C.c_int(ldw),
                               _ptr(bias), _ptr(out), C.c_int(ldc), C.c_int(out_type), C.c_int(epilogue), C.c_int(m), C.c_int(n),
                               C.c_int(k), _stream(stream)))


def op_layernorm(x, gamma, beta, y, eps=1e-6, rows=None, dim=None, ldx=None, ldy=None, device=0, stream=None) -> None:
    rows = x.shape[0] if rows is None else rows
    dim = x.shape[1] if dim is None else dim
    ldx = x.stride(0) if ldx is None else ldx
    ldy = y.stride(0) if ldy is None else ldy
    _check(lib.netcuda_op_layernorm(C.c_int(device), _ptr(x), C.c_int(ldx), _ptr(gamma), _ptr(beta), _ptr(y), C.c_int(ldy),
                                    C.c_int(rows), C.c_int(dim), C.c_float(eps), _stream(stream)))


def op_attention(qkv, out, batch, tokens, heads, device=0, stream=None) -> None:
    _check(lib.netcuda_op_attention(C.c_int(device), _ptr(qkv), _ptr(out), C.c_int(batch), C.c_int(tokens), C.c_int(heads),
                                    _stream(stream)))


def op_filter3x3(d_in, d_out, h, w, device=0, stream=None) -> None:
    _check(lib.netcuda_op_filter3x3(C.c_int(device), _ptr(d_in), _ptr(d_out), C.c_int(h), C.c_int(w), _stream(stream)))


class FrameRing:
    """netcuda_ring_*: the image side channel's ring of in-flight frames (src/netFPGA.cpp:292-365)."""

    def __init__(self, max_pixels, depth=RING_DEPTH, device=0):
        self._r = C.c_void_p(0)
        _check(lib.netcuda_ring_create(C.c_int(device), C.c_int(depth), C.c_size_t(max_pixels), C.byref(self._r)))

    def push(self, frame) -> None:
        f = np.ascontiguousarray(frame, dtype=np.uint8)
        _check(lib.netcuda_ring_push(self._r, _ptr(f), C.c_size_t(f.shape[0]), C.c_size_t(f.shape[1])))

    def pop(self) -> np.ndarray:
        h, w = C.c_size_t(0), C.c_size_t(0)
        _check(lib.netcuda_ring_peek(self._r, C.byref(h), C.byref(w)))
        out = np.empty((h.value, w.value), dtype=np.uint8)
        _check(lib.netcuda_ring_pop(self._r, _ptr(out), C.c_size_t(out.size), C.byref(h), C.byref(w)))
        return out

    @property
    def in_flight(self) -> int:
        n, d = C.c_int(0), C.c_uint64(0)
        _check(lib.netcuda_ring_in_flight(self._r, C.byref(n), C.byref(d)))
        return n.value

    @property
    def dropped(self) -> int:
        n, d = C.c_int(0), C.c_uint64(0)
        _check(lib.netcuda_ring_in_flight(self._r, C.byref(n), C.byref(d)))
        return d.value

    def close(self) -> None:
        if self._r:
            lib.netcuda_ring_destroy(self._r)
            self._r = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


ATT_KERNEL_MMA_SYNC = 100


def op_attention_ex(qkv, out, batch, tokens, heads, kernel=-1, out_f32=False, device=0, stream=None) -> None:
    _check(lib.netcuda_op_attention_ex(C.c_int(device), _ptr(qkv), _ptr(out), C.c_int(batch), C.c_int(tokens), C.c_int(heads),
                                       C.c_int(kernel), C.c_int(int(out_f32)), _stream(stream)))


def op_patchify(img, patches, batch, image_size, patch_size, device=0, stream=None) -> None:
    _check(lib.netcuda_op_patchify(C.c_int(device), _ptr(img), _ptr(patches), C.c_int(batch), C.c_int(image_size),
                                   C.c_int(patch_size), _stream(stream)))


# ---- the C++ class, driven through net::net_abstract* --------------------------------------------------

class HostNet:
    """cuda::net_cuda behind a net::net_abstract pointer (vit-fpga_b200/host/host_capi.cpp)."""

    def __init__(self, handle, n_in, n_out):
        if not handle:
            raise RuntimeError("net_cuda construction failed: " + hostlib.nch_last_error().decode())
        self._h = C.c_void_p(handle)
        self.n_in, self.n_out = n_in, n_out

    @classmethod
    def mlp(cls, npl, n_ins, w=None, b=None, random=False, seed=1, precision=-1, device=0, activation=ACT_RELU_HIDDEN,
            max_batch=0) -> "HostNet":
        arr = np.ascontiguousarray(npl, dtype=np.int32)
        wp = _ptr(np.ascontiguousarray(w, dtype=np.float32)) if w is not None else C.c_void_p(0)
        bp = _ptr(np.ascontiguousarray(b, dtype=np.float32)) if b is not None else C.c_void_p(0)
        h = hostlib.nch_mlp_create(_ptr(arr), C.c_int(len(arr)), C.c_int(n_ins), wp, bp, C.c_int(int(random)), C.c_uint(seed),
                                   C.c_int(precision), C.c_int(device), C.c_int(activation), C.c_int(max_batch))
        return cls(h, int(n_ins), int(arr[-1]))

    @classmethod
    def vit(cls, cfg: dict, flat, device=0, max_batch=0, precision=PREC_BF16) -> "HostNet":
        f = np.ascontiguousarray(flat, dtype=np.float32)
        h = hostlib.nch_vit_create(C.c_int(cfg["image_size"]), C.c_int(cfg["patch_size"]), C.c_int(cfg["dim"]), C.c_int(cfg["depth"]),
                                   C.c_int(cfg["heads"]), C.c_int(cfg["mlp_dim"]), C.c_int(cfg["n_classes"]), _ptr(f),
                                   C.c_size_t(f.size), C.c_int(device), C.c_int(max_batch), C.c_int(precision))
        return cls(h, 3 * cfg["image_size"] ** 2, cfg["n_classes"])

    @classmethod
    def load(cls, path, precision=-1, device=0, max_batch=0) -> "HostNet":
        """cuda::net_cuda::load: a net from a weight file."""
        ni, no = C.c_size_t(0), C.c_size_t(0)
        h = hostlib.nch_load(os.fsencode(path), C.c_int(precision), C.c_int(device), C.c_int(max_batch), C.byref(ni), C.byref(no))
        return cls(h, ni.value, no.value)

    def save(self, path) -> None:
        if hostlib.nch_save(self._h, os.fsencode(path)) != 0:
            raise RuntimeError("net_cuda::save failed: " + hostlib.nch_last_error().decode())

    def launch_forward(self, x) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32).ravel()
        cap = (x.size // max(self.n_in, 1) + 1) * self.n_out
        out = np.empty(cap, dtype=np.float32)
        n = hostlib.nch_launch_forward(self._h, _ptr(x), C.c_size_t(x.size), _ptr(out), C.c_size_t(cap))
        if n < 0:
            raise ValueError(hostlib.nch_last_error().decode())
        return out[:n].reshape(-1, self.n_out)

    def time_launch_forward(self, x, reps=3):
        """(seconds per call, outputs of the last call) of launch_forward(std::vector) with the input vector built beforehand."""
        x = np.ascontiguousarray(x, dtype=np.float32).ravel()
        out = np.empty((x.size // self.n_in) * self.n_out, dtype=np.float32)
        hostlib.nch_time_launch_forward.restype = C.c_double
        s = hostlib.nch_time_launch_forward(self._h, _ptr(x), C.c_size_t(x.size), C.c_int(reps), _ptr(out), C.c_size_t(out.size))
        if s < 0:
            raise RuntimeError(hostlib.nch_last_error().decode())
        return s, out.reshape(-1, self.n_out)

    def launch_forward_frame(self, frame, h=0, w=0) -> np.ndarray:
        """launch_forward(const net::image_set&): one uint8 [H, W, 3] frame."""
        f = np.ascontiguousarray(frame, dtype=np.uint8)
        out = np.empty(self.n_out, dtype=np.float32)
        n = hostlib.nch_launch_forward_frame(self._h, _ptr(f), C.c_size_t(f.size), C.c_size_t(h), C.c_size_t(w), _ptr(out), C.c_size_t(out.size))
        if n < 0:
            raise RuntimeError("launch_forward(image_set) failed: " + hostlib.nch_last_error().decode())
        return out[:n]

    def get_net_data(self, n_params, n_neurons):
        w = np.empty(n_params, dtype=np.float32)
        b = np.empty(n_neurons, dtype=np.float32)
        n_ins, n_layers = C.c_size_t(0), C.c_size_t(0)
        rc = hostlib.nch_get_net_data(self._h, _ptr(w), C.c_size_t(n_params), _ptr(b), C.c_size_t(n_neurons), C.byref(n_ins),
                                      C.byref(n_layers))
        if rc != 0:
            raise RuntimeError(f"get_net_data rc={rc}: " + hostlib.nch_last_error().decode())
        return w, b, n_ins.value, n_layers.value

    def filter_image(self, frame) -> None:
        """net_abstract::filter_image: enqueue one single-channel u8 frame [h, w] (dropped when 24 frames are in flight)."""
        f = np.ascontiguousarray(frame, dtype=np.uint8)
        if hostlib.nch_filter_image(self._h, _ptr(f), C.c_size_t(f.shape[0]), C.c_size_t(f.shape[1])) != 0:
            raise RuntimeError("filter_image failed: " + hostlib.nch_last_error().decode())

    def get_filtered_image(self, capacity=1920 * 1080):
        """net_abstract::get_filtered_image: (pixels [h, w] or None for an empty ring, (original_h, original_w))."""
        out = np.empty(capacity, dtype=np.uint8)
        h, w = C.c_size_t(0), C.c_size_t(0)
        n = hostlib.nch_get_filtered_image(self._h, _ptr(out), C.c_size_t(capacity), C.byref(h), C.byref(w))
        if n < 0:
            raise RuntimeError("get_filtered_image failed: " + hostlib.nch_last_error().decode())
        return (out[:n].reshape(h.value, w.value).copy() if n else None), (h.value, w.value)

    def forward_us(self) -> int:
        return int(hostlib.nch_forward_us(self._h))

    def check_stubs(self) -> int:
        return int(hostlib.nch_check_stubs(self._h))

    def check_move_copy(self, x, expect) -> int:
        x = np.ascontiguousarray(x, dtype=np.float32)
        e = np.ascontiguousarray(expect, dtype=np.float32)
        return int(hostlib.nch_check_move_copy(self._h, _ptr(x), C.c_size_t(x.size), _ptr(e), C.c_size_t(e.size)))

    def close(self) -> None:
        if self._h:
            hostlib.nch_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
