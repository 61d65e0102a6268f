"""ctypes access to the CPU checkers. TEST INFRASTRUCTURE ONLY.

Two things live here:
  * `Oracle`   -- oracle/liboracle.so, our C restatement (oracle/qwen3_oracle.c).
  * `RefLib`   -- oracle/_ref/libqwen3_ref_*.so, the reference's own sources compiled
                  unchanged (oracle/Makefile); present only if built in the container
                  that has /root/reference (the .so files then travel to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this
module; the product package (qwen3.c_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

c_float_p = C.POINTER(C.c_float)
c_int8_p = C.POINTER(C.c_int8)
c_int32_p = C.POINTER(C.c_int32)


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(c_float_p)


def _i8(a: np.ndarray):
    assert a.dtype == np.int8 and a.flags.c_contiguous
    return a.ctypes.data_as(c_int8_p)


def build_oracle() -> str:
    """Compile liboracle.so if missing or stale; returns its path."""
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "qwen3_oracle.c")
    hdr = os.path.join(HERE, "qwen3_oracle.h")
    if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    return so


def build_ref() -> bool:
    """Compile oracle/_ref from /root/reference when that tree exists. Returns availability."""
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)
    return os.path.exists(os.path.join(REF_DIR, "libqwen3_ref_strict.so"))


class _OrcQ8(C.Structure):
    _fields_ = [("q", c_int8_p), ("s", c_float_p)]


class _OrcModel(C.Structure):
    _fields_ = (
        [(n, C.c_int) for n in ("dim", "hidden_dim", "n_layers", "n_heads", "n_kv_heads", "vocab_size",
                                "seq_len", "head_dim", "shared_classifier", "group_size")]
        + [(n, c_float_p) for n in ("att_norm", "ffn_norm", "out_norm", "q_norm", "k_norm")]
        + [("emb", _OrcQ8), ("cls", _OrcQ8)]
        + [(n, C.POINTER(_OrcQ8)) for n in ("wq", "wk", "wv", "wo", "w1", "w2", "w3")]
        + [(n, c_float_p) for n in ("x", "xb", "q", "att", "h1", "h3", "scores", "logits", "k_cache", "v_cache")]
        + [("aq", c_int8_p), ("as_", c_float_p), ("map", C.c_void_p), ("map_len", C.c_size_t),
           ("trace_on", C.c_int), ("tr_qkv_in_q", c_int8_p)]
        + [(n, c_float_p) for n in ("tr_qkv_in_s", "tr_q_rot", "tr_att_out", "tr_x_after_att",
                                    "tr_x_after_ffn", "tr_h")]
        + [("tr_wo_in_q", c_int8_p), ("tr_wo_in_s", c_float_p), ("tr_ffn_in_q", c_int8_p), ("tr_ffn_in_s", c_float_p),
           ("tr_w2_in_q", c_int8_p), ("tr_w2_in_s", c_float_p), ("tr_cls_in_q", c_int8_p), ("tr_cls_in_s", c_float_p),
           ("tr_x_final", c_float_p), ("tr_x_normed", c_float_p)]
    )


class Oracle:
    """Our restatement. Array-in / array-out numpy wrappers plus a model handle."""

    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.orc_model_open.restype = C.POINTER(_OrcModel)
        L.orc_model_open.argtypes = [C.c_char_p, C.c_int]
        L.orc_model_close.argtypes = [C.POINTER(_OrcModel)]
        L.orc_forward.restype = c_float_p
        L.orc_forward.argtypes = [C.POINTER(_OrcModel), C.c_int, C.c_int]
        L.orc_forward_ex.restype = c_float_p
        L.orc_forward_ex.argtypes = [C.POINTER(_OrcModel), C.c_int, C.c_int, C.c_int]
        L.orc_trace_enable.argtypes = [C.POINTER(_OrcModel)]
        L.orc_argmax.restype = C.c_int
        L.orc_argmax.argtypes = [c_float_p, C.c_int, c_float_p]
        L.orc_sample.restype = C.c_int
        L.orc_sample.argtypes = [c_float_p, C.c_int, C.c_float, C.c_float, C.c_float, c_float_p]
        L.orc_xorshift_float.restype = C.c_float
        L.orc_xorshift_float.argtypes = [C.POINTER(C.c_uint64)]
        L.orc_xorshift_int32.restype = C.c_uint32
        L.orc_xorshift_int32.argtypes = [C.POINTER(C.c_uint64)]
        L.orc_sampler_clamp.argtypes = [c_float_p, c_float_p]
        L.orc_sigmoid.restype = C.c_float
        L.orc_sigmoid.argtypes = [C.c_float]
        L.orc_silu.restype = C.c_float
        L.orc_silu.argtypes = [C.c_float]

    # -- ops ------------------------------------------------------------
    def q8_quantize(self, x: np.ndarray, gs: int = 64):
        x = np.ascontiguousarray(x, np.float32)
        q = np.zeros(x.size, np.int8)
        s = np.zeros(max(x.size // gs, 1), np.float32)
        self.lib.orc_q8_quantize(_i8(q), _fp(s), _fp(x), C.c_int(x.size), C.c_int(gs))
        return q, s[: x.size // gs]

    def q8_dequantize(self, q, s, gs: int = 64):
        x = np.zeros(q.size, np.float32)
        self.lib.orc_q8_dequantize(_fp(x), _i8(q), _fp(s), C.c_int(q.size), C.c_int(gs))
        return x

    def group_dots(self, xq, wq, n, d, gs=64):
        dots = np.zeros(d * (n // gs), np.int32)
        self.lib.orc_group_dots(dots.ctypes.data_as(c_int32_p), _i8(xq), _i8(wq), C.c_int(n), C.c_int(d), C.c_int(gs))
        return dots.reshape(d, n // gs)

    def matmul(self, xq, xs, wq, ws, n, d, gs=64):
        out = np.zeros(d, np.float32)
        self.lib.orc_matmul(_fp(out), _i8(xq), _fp(xs), _i8(wq), _fp(ws), C.c_int(n), C.c_int(d), C.c_int(gs))
        return out

    def rmsnorm(self, x, w):
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros_like(x)
        self.lib.orc_rmsnorm(_fp(out), _fp(x), _fp(np.ascontiguousarray(w, np.float32)), C.c_int(x.size))
        return out

    def softmax(self, x):
        y = np.array(x, np.float32, copy=True)
        self.lib.orc_softmax(_fp(y), C.c_int(y.size))
        return y

    def rotary(self, x, head_dim, pos):
        y = np.array(x, np.float32, copy=True)
        self.lib.orc_rotary(_fp(y), C.c_int(head_dim), C.c_int(pos))
        return y

    def swiglu(self, x1, x3):
        y = np.array(x1, np.float32, copy=True)
        self.lib.orc_swiglu(_fp(y), _fp(np.ascontiguousarray(x3, np.float32)), C.c_int(y.size))
        return y

    def attention(self, q, k_layer, v_layer, n_heads, n_kv_heads, head_dim, seq_len, pos):
        out = np.zeros(n_heads * head_dim, np.float32)
        scores = np.zeros(n_heads * seq_len, np.float32)
        self.lib.orc_attention(_fp(out), _fp(q), _fp(k_layer), _fp(v_layer), _fp(scores), C.c_int(n_heads),
                               C.c_int(n_kv_heads), C.c_int(head_dim), C.c_int(seq_len), C.c_int(pos))
        return out

    def argmax(self, v):
        m = C.c_float(0)
        i = self.lib.orc_argmax(_fp(v), C.c_int(v.size), C.byref(m))
        return i, m.value

    # -- sampler (reference src/sampler.c, src/xorshift.c) ---------------
    def sampler_clamp(self, temperature: float, top_p: float):
        t, p = C.c_float(temperature), C.c_float(top_p)
        self.lib.orc_sampler_clamp(C.byref(t), C.byref(p))
        return t.value, p.value

    def sample(self, logits, temperature: float, top_p: float, coin: float):
        """sample() on a copy of `logits` with already clamped temperature / top_p and a given coin;
        returns (token, gap) -- gap: relative distance of the decision from its nearest boundary."""
        x = np.array(logits, np.float32, copy=True)
        g = C.c_float(0)
        tok = self.lib.orc_sample(_fp(x), C.c_int(x.size), C.c_float(temperature), C.c_float(top_p), C.c_float(coin), C.byref(g))
        return int(tok), float(g.value)

    def xorshift_floats(self, seed: int, n: int):
        st = C.c_uint64(seed)
        out = [float(self.lib.orc_xorshift_float(C.byref(st))) for _ in range(n)]
        return out, int(st.value)

    # -- model ----------------------------------------------------------
    def open(self, path: str, seq_len: int = 0, trace: bool = False) -> "OracleModel":
        return OracleModel(self, path, seq_len, trace)


class OracleModel:
    def __init__(self, orc: Oracle, path: str, seq_len: int, trace: bool):
        self.orc = orc
        self.h = orc.lib.orc_model_open(path.encode(), seq_len)
        if not self.h:
            raise RuntimeError(f"oracle could not open {path}")
        self.p = self.h.contents
        if trace:
            assert orc.lib.orc_trace_enable(self.h) == 0

    def forward(self, token: int, pos: int) -> np.ndarray:
        ptr = self.orc.lib.orc_forward(self.h, token, pos)
        return np.ctypeslib.as_array(ptr, shape=(self.p.vocab_size,)).copy()

    def forward_no_logits(self, token: int, pos: int) -> None:
        """A prompt position: KV rows and residual stream as forward(), without the classifier matmul."""
        self.orc.lib.orc_forward_ex(self.h, token, pos, 0)

    def _arr(self, ptr, n, dtype=np.float32):
        return np.ctypeslib.as_array(ptr, shape=(n,)).copy().astype(dtype, copy=False)

    def kv(self):
        """(k, v) caches as [L][seq_len][kv_dim] copies (reference layout)."""
        p = self.p
        n = p.n_layers * p.seq_len * p.n_kv_heads * p.head_dim
        shape = (p.n_layers, p.seq_len, p.n_kv_heads * p.head_dim)
        return self._arr(p.k_cache, n).reshape(shape), self._arr(p.v_cache, n).reshape(shape)

    def set_kv(self, k: np.ndarray, v: np.ndarray):
        p = self.p
        n = p.n_layers * p.seq_len * p.n_kv_heads * p.head_dim
        C.memmove(p.k_cache, np.ascontiguousarray(k, np.float32).ctypes.data, 4 * n)
        C.memmove(p.v_cache, np.ascontiguousarray(v, np.float32).ctypes.data, 4 * n)

    def trace(self):
        p = self.p
        L, D, Hd, P = p.n_layers, p.dim, p.hidden_dim, p.n_heads * p.head_dim
        return {
            "qkv_in_q": np.ctypeslib.as_array(p.tr_qkv_in_q, shape=(L * D,)).copy().reshape(L, D),
            "qkv_in_s": self._arr(p.tr_qkv_in_s, L * (D // p.group_size)).reshape(L, -1),
            "q_rot": self._arr(p.tr_q_rot, L * P).reshape(L, P),
            "att_out": self._arr(p.tr_att_out, L * P).reshape(L, P),
            "x_after_att": self._arr(p.tr_x_after_att, L * D).reshape(L, D),
            "x_after_ffn": self._arr(p.tr_x_after_ffn, L * D).reshape(L, D),
            "h": self._arr(p.tr_h, L * Hd).reshape(L, Hd),
            "wo_in_q": np.ctypeslib.as_array(p.tr_wo_in_q, shape=(L * P,)).copy().reshape(L, P),
            "wo_in_s": self._arr(p.tr_wo_in_s, L * (P // p.group_size)).reshape(L, -1),
            "ffn_in_q": np.ctypeslib.as_array(p.tr_ffn_in_q, shape=(L * D,)).copy().reshape(L, D),
            "ffn_in_s": self._arr(p.tr_ffn_in_s, L * (D // p.group_size)).reshape(L, -1),
            "w2_in_q": np.ctypeslib.as_array(p.tr_w2_in_q, shape=(L * Hd,)).copy().reshape(L, Hd),
            "w2_in_s": self._arr(p.tr_w2_in_s, L * (Hd // p.group_size)).reshape(L, -1),
            "cls_in_q": np.ctypeslib.as_array(p.tr_cls_in_q, shape=(D,)).copy(),
            "cls_in_s": self._arr(p.tr_cls_in_s, D // p.group_size),
            "x_final": self._arr(p.tr_x_final, D),
            "x_normed": self._arr(p.tr_x_normed, D),
        }

    def close(self):
        if self.h:
            self.orc.lib.orc_model_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


# ---------------------------------------------------------------------------
# The compiled reference. Struct mirrors of /root/reference/include/{q8,model}.h.
# ---------------------------------------------------------------------------
class Q8Tensor(C.Structure):
    _fields_ = [("s", c_float_p), ("q", c_int8_p)]


class ModelParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("magic", "version", "dim", "hidden_dim", "n_layers", "n_heads",
                                       "n_kv_heads", "vocab_size", "seq_len", "head_dim",
                                       "shared_classifier", "block_size")]


class ModelWeights(C.Structure):
    _fields_ = ([(n, C.POINTER(Q8Tensor)) for n in ("wq", "wk", "wv", "wo", "w1", "w2", "w3", "cls", "qe")]
                + [(n, c_float_p) for n in ("fe", "att_rms_norm", "ffn_rms_norm", "out_rms_norm",
                                            "q_rms_norm", "k_rms_norm")])


class ForwardState(C.Structure):
    _fields_ = ([(n, c_float_p) for n in ("x", "x_rms_norm", "q", "k", "v", "scores", "mlp_in", "mlp_gate",
                                          "logits", "k_cache", "v_cache")]
                + [("qx", Q8Tensor), ("qh", Q8Tensor)])


class Model(C.Structure):
    _fields_ = [("params", ModelParams), ("weights", ModelWeights), ("state", ForwardState),
                ("data", C.c_void_p), ("size", C.c_ssize_t)]


def bind_reference_api(lib: C.CDLL) -> C.CDLL:
    """Attach the forward.h / q8.h / model.h prototypes to `lib` (works for the compiled
    reference and for this repo's drop-in libqwen3.so alike: same ABI)."""
    lib.model_create.restype = C.POINTER(Model)
    lib.model_create.argtypes = [C.c_char_p, C.c_int]
    lib.model_free.argtypes = [C.POINTER(Model)]
    lib.model_free.restype = None
    lib.forward.restype = c_float_p
    lib.forward.argtypes = [C.POINTER(Model), C.c_int, C.c_int]
    lib.attention.argtypes = [C.POINTER(Model), C.c_int, C.c_int]
    lib.attention.restype = None
    lib.rmsnorm.argtypes = [c_float_p, c_float_p, c_float_p, C.c_int]
    lib.rmsnorm.restype = None
    lib.softmax.argtypes = [c_float_p, C.c_int]
    lib.softmax.restype = None
    lib.matmul.argtypes = [c_float_p, C.POINTER(Q8Tensor), C.POINTER(Q8Tensor), C.c_int, C.c_int, C.c_int]
    lib.matmul.restype = None
    lib.rotary.argtypes = [c_float_p, C.c_int, C.c_int]
    lib.rotary.restype = None
    lib.swiglu.argtypes = [c_float_p, c_float_p, C.c_int]
    lib.swiglu.restype = None
    lib.sigmoid.argtypes = [C.c_float]
    lib.sigmoid.restype = C.c_float
    lib.silu.argtypes = [C.c_float]
    lib.silu.restype = C.c_float
    lib.q8_quantize.argtypes = [C.POINTER(Q8Tensor), c_float_p, C.c_int, C.c_int]
    lib.q8_quantize.restype = None
    lib.q8_dequantize.argtypes = [C.POINTER(Q8Tensor), c_float_p, C.c_int, C.c_int]
    lib.q8_dequantize.restype = None
    return lib


def cpu_flags() -> set:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def pick_fast_ref() -> str | None:
    """Best `-Ofast -fopenmp` build of the reference this host can execute."""
    flags = cpu_flags()
    native = os.path.join(REF_DIR, "libqwen3_ref_fast_native.so")
    tag = os.path.join(REF_DIR, "native_cpu.txt")
    if os.path.exists(native) and os.path.exists(tag) and open(tag).read().strip() == cpu_model():
        return native
    v4 = {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"}
    if v4 <= flags and os.path.exists(os.path.join(REF_DIR, "libqwen3_ref_fast_v4.so")):
        return os.path.join(REF_DIR, "libqwen3_ref_fast_v4.so")
    v3 = {"avx2", "fma", "bmi2", "f16c"}
    if v3 <= flags and os.path.exists(os.path.join(REF_DIR, "libqwen3_ref_fast_v3.so")):
        return os.path.join(REF_DIR, "libqwen3_ref_fast_v3.so")
    return None


class RefLib:
    """The reference's own q8.c/model.c/forward.c, compiled unchanged."""

    def __init__(self, kind: str = "strict"):
        if kind == "strict":
            path = os.path.join(REF_DIR, "libqwen3_ref_strict.so")
        elif kind == "fast":
            path = pick_fast_ref()
        else:
            path = kind
        if not path or not os.path.exists(path):
            raise FileNotFoundError("compiled reference not available (oracle/_ref)")
        self.path = path
        self.lib = bind_reference_api(C.CDLL(path))

    @staticmethod
    def available(kind: str = "strict") -> bool:
        if kind == "strict":
            return os.path.exists(os.path.join(REF_DIR, "libqwen3_ref_strict.so"))
        return pick_fast_ref() is not None

    def open(self, path: str, seq_len: int = 0):
        # the reference prints its banners on stderr; keep them out of test output
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(2)
        os.dup2(devnull, 2)
        try:
            m = self.lib.model_create(path.encode(), seq_len)
        finally:
            os.dup2(saved, 2)
            os.close(saved)
            os.close(devnull)
        if not m:
            raise RuntimeError(f"reference model_create failed for {path}")
        return m

    def forward(self, m, token: int, pos: int) -> np.ndarray:
        ptr = self.lib.forward(m, token, pos)
        return np.ctypeslib.as_array(ptr, shape=(m.contents.params.vocab_size,)).copy()

    def close(self, m):
        self.lib.model_free(m)


class RefSampler:
    """The reference's own sampler.c + xorshift.c, compiled unchanged (oracle/_ref/libqwen3_ref_sampler.so)."""

    class _S(C.Structure):  # reference include/sampler.h:20-26
        _fields_ = [("dist", C.c_void_p), ("seed", C.c_uint64), ("temperature", C.c_float), ("top_p", C.c_float),
                    ("vocab_size", C.c_int)]

    PATH = os.path.join(REF_DIR, "libqwen3_ref_sampler.so")

    @staticmethod
    def available() -> bool:
        return os.path.exists(RefSampler.PATH)

    def __init__(self):
        L = C.CDLL(self.PATH)
        L.sampler_create.restype = C.POINTER(self._S)
        L.sampler_create.argtypes = [C.c_int, C.c_float, C.c_float, C.c_uint64]
        L.sampler_free.argtypes = [C.POINTER(self._S)]
        L.sample.restype = C.c_int
        L.sample.argtypes = [C.POINTER(self._S), c_float_p]
        L.xorshift_float.restype = C.c_float
        L.xorshift_float.argtypes = [C.POINTER(C.c_uint64)]
        self.lib = L

    def create(self, vocab: int, temperature: float, top_p: float, seed: int):
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(2)
        os.dup2(devnull, 2)  # the reference prints its settings on stderr
        try:
            return self.lib.sampler_create(vocab, temperature, top_p, seed)
        finally:
            os.dup2(saved, 2)
            os.close(saved)
            os.close(devnull)

    def sample(self, s, logits) -> int:
        x = np.array(logits, np.float32, copy=True)
        return int(self.lib.sample(s, _fp(x)))

    def free(self, s):
        self.lib.sampler_free(s)
