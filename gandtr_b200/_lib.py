"""ctypes binding of libgandtr_b200.so -- the C ABI declared in include/gandtr_b200.h.

The reference is pure Python, so this file *is* the reference-side FFI stub.
PyTorch is used only for device memory and streams: every function takes torch CUDA tensors, checks
dtype / contiguity / device, and passes raw pointers and the current CUDA stream to the library.

There is no CPU fallback: if the shared library is missing or no CUDA device is visible, calls raise.
"""
import ctypes
import os
import threading

import numpy as np
import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libgandtr_b200.so")
LUT_PATH = os.path.join(_PKG_DIR, "data", "rgb2lab_lut_s16.bin")

GDT_OK = 0
GDT_ERR_CANDIDATE_OVERFLOW = -7
GDT_MAX_SCALES = 8
GDT_GEM_AGGREGATE = 1
GDT_GEM_MSP_IS_P = 2

# every symbol include/gandtr_b200.h declares: (name, restype, argtypes)
_c = ctypes
_P = _c.c_void_p
SYMBOLS = [
    ("gdt_abi_version", _c.c_int, []),
    ("gdt_status_string", _c.c_char_p, [_c.c_int]),
    ("gdt_last_cuda_error", _c.c_char_p, []),
    ("gdt_init", _c.c_int, [_P]),
    ("gdt_is_initialised", _c.c_int, []),
    ("gdt_debug_get_spline_table", _c.c_int, [_P]),
    ("gdt_debug_div_check", _c.c_int, [_c.c_float, _c.c_uint32, _c.c_uint32, _P, _P]),
    ("gdt_debug_k1_config", _c.c_int, [_c.c_int] * 5),
    ("gdt_debug_k1_rows", _c.c_int, [_c.c_int]),
    ("gdt_debug_k1_pack", _c.c_int, [_c.c_int]),
    ("gdt_debug_k1_persist", _c.c_int, [_c.c_int]),
    ("gdt_debug_k1_div1", _c.c_int, [_c.c_int]),
    ("gdt_debug_k1_chroma_f", _c.c_int, [_c.c_int]),
    ("gdt_debug_k1_div1_verified", _c.c_int, [_c.c_float]),
    ("gdt_debug_k1_rec32", _c.c_int, [_c.c_int]),
    ("gdt_debug_k1_chunk", _c.c_int, [_c.c_int]),
    ("gdt_clahe_workspace_bytes", _c.c_size_t, [_c.c_int, _c.c_int, _c.c_int, _c.c_int]),
    ("gdt_clahe_u8", _c.c_int, [_P, _c.c_int, _c.c_int, _c.c_int, _c.c_double, _c.c_int, _P, _P, _P, _P, _c.c_size_t, _P]),
    ("gdt_clahe_f32", _c.c_int, [_P, _c.c_int, _c.c_int, _c.c_int, _c.c_double, _c.c_int, _P, _P, _P, _P, _P, _P,
                                 _c.c_size_t, _P]),
    ("gdt_meanstd_adapt", _c.c_int, [_P, _c.c_longlong, _c.c_longlong, _P, _P, _P, _P, _P, _P]),
    ("gdt_gem_whiten_workspace_bytes", _c.c_size_t, [_c.c_int, _c.c_int, _c.c_int, _c.c_int]),
    ("gdt_gem_whiten", _c.c_int, [_P, _P, _P, _c.c_int, _c.c_int, _c.c_int, _P, _c.c_float, _c.c_int, _P, _c.c_int, _P, _P,
                                  _c.c_int, _P, _P, _c.c_size_t, _P]),
    ("gdt_whiten_prepare", _c.c_int, [_P, _c.c_int, _c.c_int, _c.c_int, _P, _P]),
    ("gdt_gem_pool", _c.c_int, [_P, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _P, _c.c_float, _P, _P]),
    ("gdt_l2n_rows", _c.c_int, [_P, _c.c_int, _c.c_int, _c.c_float, _P, _P]),
    ("gdt_desc_post_workspace_bytes", _c.c_size_t, [_c.c_int, _c.c_int, _c.c_int]),
    ("gdt_desc_post", _c.c_int, [_P, _c.c_int, _c.c_int, _c.c_int, _P, _c.c_float, _c.c_int, _P, _c.c_int, _P, _P, _c.c_int,
                                 _P, _P, _c.c_size_t, _P]),
    ("gdt_db_prepare_workspace_bytes", _c.c_size_t, [_c.c_longlong, _c.c_int]),
    ("gdt_db_prepare", _c.c_int, [_P, _c.c_longlong, _c.c_int, _P, _P, _P, _c.c_size_t, _P]),
    ("gdt_db_prepare_norm", _c.c_int, [_P, _c.c_longlong, _c.c_int, _P, _P]),
    ("gdt_db_prepare_convert", _c.c_int, [_P, _c.c_longlong, _c.c_int, _P, _P, _P]),
    ("gdt_score_topk_exchange_layout", _c.c_int, [_c.c_int, _c.c_longlong, _c.c_int, _c.c_int, _P, _P]),
    ("gdt_score_topk_filter", _c.c_int, [_P, _P, _P, _c.c_int, _c.c_longlong, _c.c_int, _c.c_int, _P, _P, _c.c_size_t, _P]),
    ("gdt_score_topk_finalize", _c.c_int, [_P, _P, _c.c_int, _c.c_longlong, _c.c_int, _c.c_int, _c.c_longlong, _P, _P, _P,
                                           _P, _c.c_size_t, _P]),
    ("gdt_debug_k3_coarse_scores", _c.c_int, [_P, _P, _P, _c.c_int, _c.c_longlong, _c.c_int, _c.c_int, _P, _P, _P, _P,
                                              _c.c_size_t, _P]),
    ("gdt_score_topk_workspace_bytes", _c.c_size_t, [_c.c_int, _c.c_longlong, _c.c_int, _c.c_int]),
    ("gdt_score_topk", _c.c_int, [_P, _P, _P, _P, _c.c_int, _c.c_longlong, _c.c_int, _c.c_int, _c.c_longlong, _P, _P,
                                  _P, _P, _c.c_size_t, _P]),
    ("gdt_score_topk_exact_workspace_bytes", _c.c_size_t, [_c.c_int, _c.c_longlong, _c.c_int, _c.c_int]),
    ("gdt_score_topk_exact", _c.c_int, [_P, _P, _c.c_int, _c.c_longlong, _c.c_int, _c.c_int, _c.c_longlong, _P, _P, _P,
                                        _c.c_size_t, _P]),
    ("gdt_topk_merge", _c.c_int, [_P, _P, _c.c_int, _c.c_int, _c.c_int, _P, _P, _P]),
    ("gdt_topk_pack", _c.c_int, [_P, _P, _c.c_longlong, _P, _P]),
    ("gdt_topk_merge_packed", _c.c_int, [_P, _c.c_int, _c.c_int, _c.c_int, _P, _P, _P]),
    ("gdt_topk_merge_packed_keys", _c.c_int, [_P, _c.c_int, _c.c_int, _c.c_int, _P, _P]),
    ("gdt_topk_unpack", _c.c_int, [_P, _c.c_longlong, _P, _P, _P]),
    ("gdt_probe_scores", _c.c_int, [_P, _P, _c.c_int, _c.c_longlong, _c.c_int, _c.c_longlong, _P, _c.c_int, _P, _P]),
    ("gdt_rank_counts_workspace_bytes", _c.c_size_t, [_c.c_int, _c.c_int]),
    ("gdt_debug_k4_exact", _c.c_int, [_c.c_int]),
    ("gdt_jpeg_available", _c.c_int, []),
    ("gdt_jpeg_dims", _c.c_int, [_P, _c.c_size_t, _P, _P]),
    ("gdt_jpeg_decode_batch", _c.c_int, [_P, _P, _c.c_int, _P, _P, _P, _P]),
    ("gdt_debug_jpeg_last_backend", _c.c_int, []),
    ("gdt_debug_jpeg_status", _c.c_int, [_P]),
    ("gdt_rank_counts", _c.c_int, [_P, _P, _c.c_int, _c.c_longlong, _c.c_int, _c.c_longlong, _P, _P, _c.c_int, _P, _P,
                                   _c.c_size_t, _P]),
    ("gdt_map_eval", _c.c_int, [_P, _c.c_int, _P, _c.c_int, _P, _P, _P, _c.c_int, _P, _c.c_int, _P, _P, _P]),
    ("gdt_diverse_anchors_workspace_bytes", _c.c_size_t, [_c.c_int]),
    ("gdt_diverse_anchors", _c.c_int, [_P, _c.c_int, _c.c_int, _P, _c.c_int, _c.c_int, _P, _P, _P, _c.c_size_t, _P]),
    ("gdt_syrk_f64_workspace_bytes", _c.c_size_t, [_c.c_int, _c.c_longlong]),
    ("gdt_syrk_f64", _c.c_int, [_P, _c.c_int, _c.c_longlong, _c.c_longlong, _c.c_double, _P, _P, _c.c_size_t, _P]),
    ("gdt_thumbnail_geometry", _c.c_int, [_c.c_int, _c.c_int, _c.c_double, _P, _P, _P, _P]),
    ("gdt_resize_plan_create", _c.c_int, [_c.c_int, _c.c_int, _c.c_double, _P]),
    ("gdt_resize_plan_destroy", None, [_P]),
    ("gdt_resize_plan_info", _c.c_int, [_P, _P, _P, _P, _P]),
    ("gdt_resize_workspace_bytes", _c.c_size_t, [_P]),
    ("gdt_resize_u8", _c.c_int, [_P, _P, _c.c_size_t, _P, _P, _c.c_size_t, _P]),
    ("gdt_resize_batch_workspace_bytes", _c.c_size_t, [_P, _c.c_int]),
    ("gdt_resize_u8_batch", _c.c_int, [_P, _P, _P, _c.c_int, _P, _P, _c.c_size_t, _P]),
    ("gdt_debug_k5_bytewise", _c.c_int, [_c.c_int]),
    ("gdt_debug_k5_planar", _c.c_int, [_c.c_int]),
    ("gdt_debug_resize_coeffs", _c.c_int, [_c.c_int, _c.c_float, _c.c_float, _c.c_int, _P, _P, _P, _c.c_size_t]),
]

_lib = None
_lock = threading.Lock()
_initialised_devices = set()
launch_count = 0  # number of library compute calls made by this process (bench.py reports kernel launches from it)

# kernels launched by each entry point (for bench.py's gpu_launches claim)
KERNELS_PER_CALL = {"clahe": 2, "meanstd_adapt": 1, "gem": 2, "gem_whiten": 4, "gem_pool": 1, "l2n_rows": 1, "desc_post": 1, "desc_post_whiten": 3, "db_prepare": 2, "score_topk_exact": 2,
                    "topk_merge": 1, "topk_pack": 1, "topk_unpack": 1, "probe_scores": 1, "rank_counts": 3, "map_eval": 1, "resize": 2}


class GdtError(RuntimeError):
    pass


def load():
    """dlopen the library (once). Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise GdtError("%s is missing: run `python -m gandtr_b200._build` (or __graft_entry__.build()); "
                               "gandtr_b200 has no CPU fallback" % LIB_PATH)
            lib = ctypes.CDLL(LIB_PATH)
            for name, restype, argtypes in SYMBOLS:
                fn = getattr(lib, name)
                fn.restype = restype
                fn.argtypes = argtypes
            _lib = lib
    return _lib


def check(rc, what):
    if rc != GDT_OK:
        lib = load()
        msg = lib.gdt_status_string(rc).decode()
        cuda = lib.gdt_last_cuda_error().decode()
        raise GdtError("%s failed: %s (%d)%s" % (what, msg, rc, (" -- " + cuda) if rc == -5 and cuda else ""))


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _require(t, dtype, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise GdtError("%s must be a CUDA tensor (gandtr_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise GdtError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise GdtError("%s must be contiguous" % name)
    return t


def _workspace(nbytes, device):
    # torch's caching allocator returns 512-byte aligned blocks
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def ensure_init(device):
    """gdt_init on `device` (uploads the CLAHE tables) once per process and device."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise GdtError("gandtr_b200 kernels need a CUDA device, got %s" % dev)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx in _initialised_devices:
        return
    lib = load()
    lut = np.fromfile(LUT_PATH, dtype="<i2")
    if lut.size != 33 * 33 * 33 * 3:
        raise GdtError("corrupt table file %s" % LUT_PATH)
    with torch.cuda.device(idx):
        torch.cuda.current_stream().synchronize()
        check(lib.gdt_init(lut.ctypes.data_as(ctypes.c_void_p)), "gdt_init")
    _initialised_devices.add(idx)


K1_DEFAULT_PACK = 0                     # pass B scalar (1 = packed f32x2); must match g_k1_pack
K1_DEFAULT_CONFIG = (0, 0, 0, -1, 4)    # texab, spltex, fytex, chroma_a (-1 = automatic), occ_a -- must match clahe_sm100.cu's statics


def k1_config_default():
    """Restore K1's built-in work split / pipe choice after an A/B run (debug hook)."""
    lib = load()
    check(lib.gdt_debug_k1_config(*K1_DEFAULT_CONFIG), "gdt_debug_k1_config")
    check(lib.gdt_debug_k1_rec32(1), "gdt_debug_k1_rec32")
    check(lib.gdt_debug_k1_persist(1), "gdt_debug_k1_persist")
    check(lib.gdt_debug_k1_div1(1), "gdt_debug_k1_div1")
    check(lib.gdt_debug_k1_chroma_f(0), "gdt_debug_k1_chroma_f")
    check(lib.gdt_debug_k1_pack(K1_DEFAULT_PACK), "gdt_debug_k1_pack")


def _f3(v):
    a = (ctypes.c_float * 3)(*[float(x) for x in v])
    return a


def _count(kind):
    global launch_count
    launch_count += KERNELS_PER_CALL[kind]


# ---- K1 ----------------------------------------------------------------------------------------------

def clahe_u8(rgb_hwc, mean, std, clip_limit=1.0, grid=8, out=None):
    """[n,h,w,3] uint8 CUDA -> [n,3,h,w] float32: fused pil2np | apply_clahe | totensor | normalize."""
    _require(rgb_hwc, torch.uint8, "rgb_hwc")
    if rgb_hwc.dim() != 4 or rgb_hwc.shape[3] != 3:
        raise GdtError("rgb_hwc must be [n, h, w, 3]")
    n, h, w, _ = rgb_hwc.shape
    ensure_init(rgb_hwc.device)
    lib = load()
    if out is None:
        out = torch.empty((n, 3, h, w), dtype=torch.float32, device=rgb_hwc.device)
    else:
        _require(out, torch.float32, "out")
    with torch.cuda.device(rgb_hwc.device):
        ws = _workspace(lib.gdt_clahe_workspace_bytes(n, h, w, grid), rgb_hwc.device)
        check(lib.gdt_clahe_u8(_ptr(rgb_hwc), n, h, w, float(clip_limit), int(grid), _f3(mean), _f3(std), _ptr(out),
                               _ptr(ws), ws.numel(), _stream()), "gdt_clahe_u8")
    _count("clahe")
    return out


def clahe_f32(x_chw, in_mean, in_std, out_mean, out_std, clip_limit=1.0, grid=8, out=None):
    """[n,3,h,w] float32 CUDA normalised with (in_mean, in_std) -> same shape, CLAHE applied, re-normalised."""
    _require(x_chw, torch.float32, "x_chw")
    if x_chw.dim() != 4 or x_chw.shape[1] != 3:
        raise GdtError("x_chw must be [n, 3, h, w]")
    n, _, h, w = x_chw.shape
    ensure_init(x_chw.device)
    lib = load()
    if out is None:
        out = torch.empty_like(x_chw)
    with torch.cuda.device(x_chw.device):
        ws = _workspace(lib.gdt_clahe_workspace_bytes(n, h, w, grid), x_chw.device)
        check(lib.gdt_clahe_f32(_ptr(x_chw), n, h, w, float(clip_limit), int(grid), _f3(in_mean), _f3(in_std),
                                _f3(out_mean), _f3(out_std), _ptr(out), _ptr(ws), ws.numel(), _stream()), "gdt_clahe_f32")
    _count("clahe")
    return out


def meanstd_adapt(x, in_mean, in_std, out_mean, out_std, out=None):
    """[n,3,h,w] or [3,h,w] float32 CUDA -> ((x * in_std + in_mean) - out_mean) / out_std, same shape (MeanStdPost)."""
    _require(x, torch.float32, "x")
    if x.dim() not in (3, 4) or x.shape[-3] != 3:
        raise GdtError("x must be [n, 3, h, w] or [3, h, w]")
    n = x.shape[0] if x.dim() == 4 else 1
    plane = x.shape[-1] * x.shape[-2]
    if out is None:
        out = torch.empty_like(x)
    else:
        _require(out, torch.float32, "out")
    with torch.cuda.device(x.device):
        check(load().gdt_meanstd_adapt(_ptr(x), n, plane, _f3(in_mean), _f3(in_std), _f3(out_mean), _f3(out_std), _ptr(out),
                                       _stream()), "gdt_meanstd_adapt")
    _count("meanstd_adapt")
    return out


# ---- K2 ----------------------------------------------------------------------------------------------

def whiten_prepare(P):
    """[dim, c] float32 CUDA projection -> [2, dim, c] TF32 hi / lo halves for the tcgen05 projection kernel (once per
    learned whitening)."""
    _require(P, torch.float32, "P")
    if P.dim() != 2:
        raise GdtError("P must be [dim, c]")
    dim, c = P.shape
    out = torch.empty((2, dim, c), dtype=torch.float32, device=P.device)
    with torch.cuda.device(P.device):
        check(load().gdt_whiten_prepare(_ptr(P), int(P.stride(0)), c, dim, _ptr(out), _stream()), "gdt_whiten_prepare")
    _count("l2n_rows")
    return out


def _check_split(P_split, P, dim):
    if P_split is None:
        return None
    _require(P_split, torch.float32, "P_split")
    if tuple(P_split.shape) != (2, dim, P.shape[1]):
        raise GdtError("P_split must be whiten_prepare(P[:dim]) with shape [2, dim, c]")
    return P_split


def gem_whiten(fmaps, p, eps=1e-6, aggregate=False, msp_is_p=False, P=None, m=None, dim=None, P_split=None):
    """fmaps: list (one per scale) of [n,c,h,w] float32 CUDA feature maps -> descriptors [n, dim] float32.
    p: 1-element float32 CUDA tensor (GeM exponent, read on the device). P_split: whiten_prepare(P[:dim]) selects the
    tcgen05 projection kernel."""
    if isinstance(fmaps, torch.Tensor):
        fmaps = [fmaps]
    scales = len(fmaps)
    if not 1 <= scales <= GDT_MAX_SCALES:
        raise GdtError("1..%d scales supported" % GDT_MAX_SCALES)
    fmaps = [_require(f, torch.float32, "fmap") for f in fmaps]
    n, c = fmaps[0].shape[0], fmaps[0].shape[1]
    for f in fmaps:
        if f.dim() != 4 or f.shape[0] != n or f.shape[1] != c:
            raise GdtError("all feature maps must be [n, c, h, w] with equal n and c")
    dev = fmaps[0].device
    _require(p, torch.float32, "p")
    lib = load()
    flags = (GDT_GEM_AGGREGATE if aggregate else 0) | (GDT_GEM_MSP_IS_P if msp_is_p else 0)
    if P is not None:
        _require(P, torch.float32, "P")
        _require(m, torch.float32, "m")
        dim = int(dim or P.shape[0])
        if P.dim() != 2 or P.shape[1] != c or dim > P.shape[0] or m.numel() != c:
            raise GdtError("whitening shapes do not match the feature maps")
        ldP = P.stride(0)
    else:
        dim, ldP = c, 0
    desc = torch.empty((n, dim), dtype=torch.float32, device=dev)
    ptrs = (ctypes.c_void_p * scales)(*[f.data_ptr() for f in fmaps])
    hs = (ctypes.c_int * scales)(*[int(f.shape[2]) for f in fmaps])
    ws_ = (ctypes.c_int * scales)(*[int(f.shape[3]) for f in fmaps])
    with torch.cuda.device(dev):
        ws = _workspace(lib.gdt_gem_whiten_workspace_bytes(n, c, scales, dim), dev)
        P_split = _check_split(P_split, P, dim) if P is not None else None
        check(lib.gdt_gem_whiten(ptrs, hs, ws_, n, c, scales, _ptr(p), float(eps), flags,
                                 _ptr(P) if P is not None else None, int(ldP),
                                 _ptr(P_split) if P_split is not None else None, _ptr(m) if m is not None else None,
                                 dim, _ptr(desc), _ptr(ws), ws.numel(), _stream()), "gdt_gem_whiten")
    _count("gem_whiten" if P is not None else "gem")
    return desc


def gem_pool(fmap, p, eps=1e-6):
    """GeM.forward: [n,c,h,w] float32 CUDA -> pooled [n,c] (not normalised)."""
    _require(fmap, torch.float32, "fmap")
    _require(p, torch.float32, "p")
    if fmap.dim() != 4:
        raise GdtError("fmap must be [n, c, h, w]")
    n, c, h, w = fmap.shape
    out = torch.empty((n, c), dtype=torch.float32, device=fmap.device)
    with torch.cuda.device(fmap.device):
        check(load().gdt_gem_pool(_ptr(fmap), n, c, h, w, _ptr(p), float(eps), _ptr(out), _stream()), "gdt_gem_pool")
    _count("gem_pool")
    return out


def l2n_rows(x, eps=1e-6):
    """L2N.forward on [n, dim] rows."""
    _require(x, torch.float32, "x")
    if x.dim() != 2:
        raise GdtError("x must be [n, dim]")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(load().gdt_l2n_rows(_ptr(x), x.shape[0], x.shape[1], float(eps), _ptr(out), _stream()), "gdt_l2n_rows")
    _count("l2n_rows")
    return out


def desc_post(descs, msp, P=None, m=None, dim=None, P_split=None):
    """Aggregation (msp: None = none, float, or 1-element CUDA tensor read on the device) and / or whitening of
    already L2-normalised per-scale descriptors, each [n, c] -> [n, dim]."""
    descs = [_require(d, torch.float32, "desc") for d in descs]
    scales = len(descs)
    n, c = descs[0].shape
    for d in descs:
        if tuple(d.shape) != (n, c):
            raise GdtError("all per-scale descriptors must be [n, c]")
    dev = descs[0].device
    flags, msp_dev, msp_host = 0, None, 1.0
    if msp is not None:
        flags |= GDT_GEM_AGGREGATE
        if isinstance(msp, torch.Tensor):
            _require(msp, torch.float32, "msp")
            flags |= GDT_GEM_MSP_IS_P
            msp_dev = msp
        else:
            msp_host = float(msp)
    elif scales != 1:
        raise GdtError("several scales need an aggregation exponent")
    if P is not None:
        _require(P, torch.float32, "P")
        _require(m, torch.float32, "m")
        dim = int(dim or P.shape[0])
        if P.dim() != 2 or P.shape[1] != c or dim > P.shape[0] or m.numel() != c:
            raise GdtError("whitening shapes do not match the descriptors")
        ldP = P.stride(0)
    else:
        dim, ldP = c, 0
    out = torch.empty((n, dim), dtype=torch.float32, device=dev)
    ptrs = (ctypes.c_void_p * scales)(*[d.data_ptr() for d in descs])
    lib = load()
    with torch.cuda.device(dev):
        ws = _workspace(lib.gdt_desc_post_workspace_bytes(n, c, dim), dev)
        P_split = _check_split(P_split, P, dim) if P is not None else None
        check(lib.gdt_desc_post(ptrs, n, c, scales, _ptr(msp_dev) if msp_dev is not None else None, msp_host, flags,
                                _ptr(P) if P is not None else None, int(ldP),
                                _ptr(P_split) if P_split is not None else None, _ptr(m) if m is not None else None, dim,
                                _ptr(out), _ptr(ws), ws.numel(), _stream()), "gdt_desc_post")
    _count("desc_post_whiten" if P is not None else "desc_post")
    return out


# ---- K3 ----------------------------------------------------------------------------------------------

def db_prepare(db):
    """[ndb, d] float32 CUDA database shard -> (fp16 shadow [ndb, d], stats [4] float32: max row norm, power-of-two
    scale, max rounding-residual norm, max shadow-row norm)."""
    _require(db, torch.float32, "db")
    ndb, d = db.shape
    lib = load()
    shadow = torch.empty((ndb, d), dtype=torch.float16, device=db.device)
    norm_max = torch.empty(4, dtype=torch.float32, device=db.device)
    with torch.cuda.device(db.device):
        ws = _workspace(lib.gdt_db_prepare_workspace_bytes(ndb, d), db.device)
        check(lib.gdt_db_prepare(_ptr(db), ndb, d, _ptr(shadow), _ptr(norm_max), _ptr(ws), ws.numel(), _stream()),
              "gdt_db_prepare")
    _count("db_prepare")
    return shadow, norm_max


def db_prepare_sharded(db, allreduce_max):
    """gdt_db_prepare for one shard of a row-sharded database: `allreduce_max(tensor)` must reduce (MAX, in place)
    over the ranks, so that every shard is converted with the same scale and filtered with the same error bound."""
    _require(db, torch.float32, "db")
    ndb, d = db.shape
    lib = load()
    shadow = torch.empty((ndb, d), dtype=torch.float16, device=db.device)
    stats = torch.zeros(4, dtype=torch.float32, device=db.device)
    with torch.cuda.device(db.device):
        if ndb:
            check(lib.gdt_db_prepare_norm(_ptr(db), ndb, d, _ptr(stats), _stream()), "gdt_db_prepare_norm")
        allreduce_max(stats[0:1])
        if ndb:
            check(lib.gdt_db_prepare_convert(_ptr(db), ndb, d, _ptr(shadow), _ptr(stats), _stream()), "gdt_db_prepare_convert")
        allreduce_max(stats[1:4])     # an empty shard learns scale and bounds from its peers
    _count("db_prepare")
    return shadow, stats


def _filter_launches(ndb, k):
    seed = max(32, (16 * k + 255) // 256)
    return 2 + (1 if (ndb + 255) // 256 > seed else 0)        # q_prepare, seed (+ main) filter pass


class FilterState:
    """What lives between gdt_score_topk_filter and gdt_score_topk_finalize of one shard: the workspace (candidate
    segments, thresholds), the status words and `hist`, the int32 [nq, 256] view of the per-query score histogram that
    the ranks sum (in place) before finalising."""

    def __init__(self, ws, status, hist, nq, ndb, d, k):
        self.ws, self.status, self.hist, self.nq, self.ndb, self.d, self.k = ws, status, hist, nq, ndb, d, k


def score_topk_filter(q, shadow, stats, k):
    """First phase of the row-sharded search on one shard: tcgen05 coarse pass, candidates + histogram in the workspace."""
    global launch_count
    _require(q, torch.float32, "q")
    _require(shadow, torch.float16, "shadow")
    _require(stats, torch.float32, "stats")
    nq, d = q.shape
    ndb = shadow.shape[0]
    if shadow.shape[1] != d:
        raise GdtError("q and shadow dimensions differ")
    lib = load()
    dev = q.device
    status = torch.empty(4, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws = _workspace(lib.gdt_score_topk_workspace_bytes(nq, ndb, d, k), dev)
        check(lib.gdt_score_topk_filter(_ptr(q), _ptr(shadow), _ptr(stats), nq, ndb, d, k, _ptr(status), _ptr(ws), ws.numel(),
                                        _stream()), "gdt_score_topk_filter")
        off, nbytes = ctypes.c_size_t(), ctypes.c_size_t()
        check(lib.gdt_score_topk_exchange_layout(nq, ndb, d, k, ctypes.byref(off), ctypes.byref(nbytes)),
              "gdt_score_topk_exchange_layout")
    hist = ws[off.value:off.value + nbytes.value].view(torch.int32).view(nq, 256)
    launch_count += _filter_launches(ndb, k)
    return FilterState(ws, status, hist, nq, ndb, d, k)


def score_topk_finalize(q, db, state, index_base=0):
    """Second phase: survivors above the threshold implied by `state.hist` (local, or summed over the shards) are
    re-scored exactly. Returns (scores [nq,k], idx [nq,k], status [4]) on the device."""
    global launch_count
    _require(q, torch.float32, "q")
    _require(db, torch.float32, "db")
    nq, d = q.shape
    k = state.k
    if (nq, db.shape[0], d) != (state.nq, state.ndb, state.d) or db.shape[1] != d:
        raise GdtError("finalize arguments do not match the filter pass")
    dev = q.device
    scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
    idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(load().gdt_score_topk_finalize(_ptr(q), _ptr(db), nq, state.ndb, d, k, int(index_base), _ptr(scores), _ptr(idx),
                                             _ptr(state.status), _ptr(state.ws), state.ws.numel(), _stream()),
              "gdt_score_topk_finalize")
    launch_count += 1
    return scores, idx, state.status


def score_topk_two_phase(q, db, shadow, stats, k, index_base=0, exchange=None):
    """Filter -> `exchange(hist)` (an in-place SUM all-reduce of the int32 [nq, 256] histogram view) -> finalize."""
    _require(db, torch.float32, "db")
    state = score_topk_filter(q, shadow, stats, k)
    if exchange is not None:
        exchange(state.hist)
    return score_topk_finalize(q, db, state, index_base=index_base)


def debug_k3_coarse_scores(q, shadow, stats, k):
    """Test hook: (coarse [nq, ndb] raw tensor-core scores, meta [nq, 4] = scale, 1/scale, margin = 2*E_q, sq)."""
    _require(q, torch.float32, "q")
    _require(shadow, torch.float16, "shadow")
    _require(stats, torch.float32, "stats")
    nq, d = q.shape
    ndb = shadow.shape[0]
    lib = load()
    dev = q.device
    coarse = torch.full((nq, ndb), float("nan"), dtype=torch.float32, device=dev)
    meta = torch.empty((nq, 4), dtype=torch.float32, device=dev)
    status = torch.empty(4, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        ws = _workspace(lib.gdt_score_topk_workspace_bytes(nq, ndb, d, k), dev)
        check(lib.gdt_debug_k3_coarse_scores(_ptr(q), _ptr(shadow), _ptr(stats), nq, ndb, d, k, _ptr(coarse), _ptr(meta),
                                             _ptr(status), _ptr(ws), ws.numel(), _stream()), "gdt_debug_k3_coarse_scores")
    return coarse, meta


def score_topk_workspace_bytes(nq, ndb, d, k):
    return load().gdt_score_topk_workspace_bytes(nq, ndb, d, k)


def score_topk(q, db, shadow, norm_max, k, index_base=0, ws=None, out=None):
    """tcgen05 path. Returns (scores [nq,k] f32, idx [nq,k] i64, status [4] i32), all on the device, async."""
    _require(q, torch.float32, "q")
    _require(db, torch.float32, "db")
    _require(shadow, torch.float16, "shadow")
    _require(norm_max, torch.float32, "norm_max")
    nq, d = q.shape
    ndb = db.shape[0]
    if db.shape[1] != d or tuple(shadow.shape) != (ndb, d):
        raise GdtError("q, db and shadow shapes do not match")
    lib = load()
    dev = q.device
    if out is None:
        scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
        idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
        status = torch.empty(4, dtype=torch.int32, device=dev)
    else:
        scores, idx, status = out
    with torch.cuda.device(dev):
        if ws is None:
            ws = _workspace(lib.gdt_score_topk_workspace_bytes(nq, ndb, d, k), dev)
        check(lib.gdt_score_topk(_ptr(q), _ptr(db), _ptr(shadow), _ptr(norm_max), nq, ndb, d, k, int(index_base),
                                 _ptr(scores), _ptr(idx), _ptr(status), _ptr(ws), ws.numel(), _stream()), "gdt_score_topk")
    global launch_count
    launch_count += _filter_launches(ndb, k) + 1                  # q_prepare, seed (+ main) filter pass, finalize
    return scores, idx, status


def score_topk_exact(q, db, k, index_base=0):
    """CUDA-core exact path (fp64-accumulated). Materialises [nq, ndb] scores in the workspace."""
    _require(q, torch.float32, "q")
    _require(db, torch.float32, "db")
    nq, d = q.shape
    ndb = db.shape[0]
    if db.shape[1] != d:
        raise GdtError("q and db dimensions differ")
    lib = load()
    dev = q.device
    scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
    idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        ws = _workspace(lib.gdt_score_topk_exact_workspace_bytes(nq, ndb, d, k), dev)
        check(lib.gdt_score_topk_exact(_ptr(q), _ptr(db), nq, ndb, d, k, int(index_base), _ptr(scores), _ptr(idx),
                                       _ptr(ws), ws.numel(), _stream()), "gdt_score_topk_exact")
    _count("score_topk_exact")
    return scores, idx


def topk_merge(scores, idx):
    """[g, nq, k] per-shard lists -> merged [nq, k]."""
    _require(scores, torch.float32, "scores")
    _require(idx, torch.int64, "idx")
    g, nq, k = scores.shape
    out_s = torch.empty((nq, k), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=scores.device)
    with torch.cuda.device(scores.device):
        check(load().gdt_topk_merge(_ptr(scores), _ptr(idx), g, nq, k, _ptr(out_s), _ptr(out_i), _stream()),
              "gdt_topk_merge")
    _count("topk_merge")
    return out_s, out_i


def topk_pack(scores, idx):
    """(scores, GLOBAL idx) lists of any shape -> uint64 rank keys (as an int64 tensor): the 8-byte exchange format."""
    _require(scores, torch.float32, "scores")
    _require(idx, torch.int64, "idx")
    keys = torch.empty(scores.shape, dtype=torch.int64, device=scores.device)
    with torch.cuda.device(scores.device):
        check(load().gdt_topk_pack(_ptr(scores), _ptr(idx), scores.numel(), _ptr(keys), _stream()), "gdt_topk_pack")
    _count("topk_pack")
    return keys


def topk_merge_packed(keys):
    """[g, nq, k] packed per-shard lists -> merged (scores [nq, k], idx [nq, k]); idx[q, 0] == -2 marks an overflow."""
    _require(keys, torch.int64, "keys")
    g, nq, k = keys.shape
    out_s = torch.empty((nq, k), dtype=torch.float32, device=keys.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=keys.device)
    with torch.cuda.device(keys.device):
        check(load().gdt_topk_merge_packed(_ptr(keys), g, nq, k, _ptr(out_s), _ptr(out_i), _stream()), "gdt_topk_merge_packed")
    _count("topk_merge")
    return out_s, out_i


def topk_merge_packed_keys(keys):
    """[g, nq, k] packed per-shard lists -> merged packed keys [nq, k] (the query-sharded merge's intermediate form)."""
    _require(keys, torch.int64, "keys")
    g, nq, k = keys.shape
    out = torch.empty((nq, k), dtype=torch.int64, device=keys.device)
    with torch.cuda.device(keys.device):
        check(load().gdt_topk_merge_packed_keys(_ptr(keys), g, nq, k, _ptr(out), _stream()), "gdt_topk_merge_packed_keys")
    _count("topk_merge")
    return out


def topk_unpack(keys):
    """packed keys of any shape -> (scores, idx): padding -> (-inf, -1), an overflow marker -> (-inf, -2)."""
    _require(keys, torch.int64, "keys")
    s = torch.empty(keys.shape, dtype=torch.float32, device=keys.device)
    i = torch.empty(keys.shape, dtype=torch.int64, device=keys.device)
    with torch.cuda.device(keys.device):
        check(load().gdt_topk_unpack(_ptr(keys), keys.numel(), _ptr(s), _ptr(i), _stream()), "gdt_topk_unpack")
    _count("topk_unpack")
    return s, i


# ---- N1: JPEG decode on the device (nvJPEG library) ---------------------------------------------------

def jpeg_available():
    return bool(load().gdt_jpeg_available())


def jpeg_decode_batch(streams, device):
    """streams: list of `bytes` (JPEG files) -> list of uint8 CUDA [h, w, 3] tensors, decoded by ONE batched nvJPEG call
    (hardware JPEG engines when available). Library decode: not bit-identical to libjpeg."""
    lib = load()
    dev = torch.device(device)
    n = len(streams)
    if n == 0:
        return []
    with torch.cuda.device(dev):
        bufs = [(ctypes.c_ubyte * len(b)).from_buffer_copy(b) for b in streams]
        ws, hs = (ctypes.c_int * n)(), (ctypes.c_int * n)()
        for i, b in enumerate(bufs):
            w, h = ctypes.c_int(), ctypes.c_int()
            check(lib.gdt_jpeg_dims(ctypes.cast(b, ctypes.c_void_p), len(streams[i]), ctypes.byref(w), ctypes.byref(h)), "gdt_jpeg_dims")
            ws[i], hs[i] = w.value, h.value
        outs = [torch.empty((hs[i], ws[i], 3), dtype=torch.uint8, device=dev) for i in range(n)]
        ptrs = (ctypes.c_void_p * n)(*[ctypes.cast(b, ctypes.c_void_p).value for b in bufs])
        lens = (ctypes.c_size_t * n)(*[len(b) for b in streams])
        dsts = (ctypes.c_void_p * n)(*[o.data_ptr() for o in outs])
        check(lib.gdt_jpeg_decode_batch(ptrs, lens, n, dsts, ws, hs, _stream()), "gdt_jpeg_decode_batch")
        torch.cuda.current_stream().synchronize()        # the host bit-stream buffers must outlive the decode
    return outs


# ---- K4 ----------------------------------------------------------------------------------------------

def probe_scores(q, db, probe_idx, index_base=0, out=None):
    _require(q, torch.float32, "q")
    _require(db, torch.float32, "db")
    _require(probe_idx, torch.int64, "probe_idx")
    nq, d = q.shape
    pmax = probe_idx.shape[1]
    if out is None:
        out = torch.zeros((nq, pmax), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        check(load().gdt_probe_scores(_ptr(q), _ptr(db), nq, db.shape[0], d, int(index_base), _ptr(probe_idx), pmax,
                                      _ptr(out), _stream()), "gdt_probe_scores")
    _count("probe_scores")
    return out


def rank_counts(q, db, probe_idx, probe_score, index_base=0, out=None):
    _require(q, torch.float32, "q")
    _require(db, torch.float32, "db")
    _require(probe_idx, torch.int64, "probe_idx")
    _require(probe_score, torch.float32, "probe_score")
    nq, d = q.shape
    pmax = probe_idx.shape[1]
    if out is None:
        out = torch.zeros((nq, pmax), dtype=torch.int64, device=q.device)
    lib = load()
    with torch.cuda.device(q.device):
        ws = _workspace(lib.gdt_rank_counts_workspace_bytes(nq, pmax), q.device)
        check(lib.gdt_rank_counts(_ptr(q), _ptr(db), nq, db.shape[0], d, int(index_base), _ptr(probe_idx),
                                  _ptr(probe_score), pmax, _ptr(out), _ptr(ws), ws.numel(), _stream()), "gdt_rank_counts")
    _count("rank_counts")
    return out


def map_eval(pos_rank, junk_rank, npos, njunk, kappas, nres=None):
    """-> (ap [nq] float64, prk [nq, nk] float64). kappas: list of ints or an int32 CUDA tensor; nres: optional int32 [nq]
    recall denominators (defaults to npos)."""
    _require(pos_rank, torch.int64, "pos_rank")
    _require(junk_rank, torch.int64, "junk_rank")
    _require(npos, torch.int32, "npos")
    _require(njunk, torch.int32, "njunk")
    if nres is not None:
        _require(nres, torch.int32, "nres")
    nq = pos_rank.shape[0]
    dev = pos_rank.device
    if isinstance(kappas, torch.Tensor):
        kap, nk = _require(kappas, torch.int32, "kappas"), kappas.numel()
    else:
        kap, nk = torch.tensor(list(kappas) or [1], dtype=torch.int32, device=dev), len(kappas)
    ap = torch.empty(nq, dtype=torch.float64, device=dev)
    prk = torch.empty((nq, max(nk, 1)), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(load().gdt_map_eval(_ptr(pos_rank), pos_rank.shape[1], _ptr(junk_rank), junk_rank.shape[1], _ptr(npos),
                                  _ptr(njunk), _ptr(nres) if nres is not None else None, nq, _ptr(kap), nk, _ptr(ap),
                                  _ptr(prk), _stream()), "gdt_map_eval")
    _count("map_eval")
    return ap, prk[:, :nk]


# ---- N3: mining ------------------------------------------------------------------------------------

def diverse_anchors(pool, ranks, first=0):
    """pool: [n, d] float32 CUDA rows; ranks: int32 CUDA [steps] ascending rank picked per round.
    -> (picked int32 [steps + 1], picked_score float32 [steps]) on the device, no host sync."""
    global launch_count
    _require(pool, torch.float32, "pool")
    _require(ranks, torch.int32, "ranks")
    n, d = pool.shape
    steps = ranks.numel()
    dev = pool.device
    picked = torch.empty(steps + 1, dtype=torch.int32, device=dev)
    score = torch.empty(max(steps, 1), dtype=torch.float32, device=dev)
    lib = load()
    with torch.cuda.device(dev):
        ws = _workspace(lib.gdt_diverse_anchors_workspace_bytes(n), dev)
        check(lib.gdt_diverse_anchors(_ptr(pool), n, d, _ptr(ranks), steps, int(first), _ptr(picked), _ptr(score), _ptr(ws),
                                      ws.numel(), _stream()), "gdt_diverse_anchors")
    launch_count += 2 * steps
    return picked, score[:steps]


# ---- N2: whitening learning ------------------------------------------------------------------------

def syrk_f64(A, alpha=1.0):
    """A: [d, n] float64 CUDA (rows contiguous) -> alpha * A @ A.T as a symmetric [d, d] float64 tensor."""
    global launch_count
    if not isinstance(A, torch.Tensor) or not A.is_cuda or A.dtype != torch.float64 or A.dim() != 2 or A.stride(1) != 1:
        raise GdtError("A must be a float64 CUDA matrix [d, n] with contiguous rows (gandtr_b200 has no CPU path)")
    d, n = A.shape
    C = torch.empty((d, d), dtype=torch.float64, device=A.device)
    if n == 0:
        return C.zero_()
    lib = load()
    with torch.cuda.device(A.device):
        ws = _workspace(lib.gdt_syrk_f64_workspace_bytes(d, n), A.device)
        check(lib.gdt_syrk_f64(_ptr(A), d, n, int(A.stride(0)), float(alpha), _ptr(C), _ptr(ws), ws.numel(), _stream()),
              "gdt_syrk_f64")
    launch_count += 2
    return C


# ---- K5: dataset image geometry ------------------------------------------------------------------

def thumbnail_geometry(w, h, imsize):
    """Host only: (out_w, out_h, fx, fy, resized) Pillow picks for `Image.thumbnail((imsize, imsize), LANCZOS)`."""
    ow, oh, fx, fy = (_c.c_int(), _c.c_int(), _c.c_int(), _c.c_int())
    rc = load().gdt_thumbnail_geometry(int(w), int(h), float(imsize), _c.byref(ow), _c.byref(oh), _c.byref(fx), _c.byref(fy))
    if rc < 0:
        check(rc, "gdt_thumbnail_geometry")
    return ow.value, oh.value, fx.value, fy.value, bool(rc)


class ResizePlan:
    """Filter coefficients of one (w, h, imsize) geometry on the current device (gdt_resize_plan_create)."""

    def __init__(self, w, h, imsize, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise GdtError("gandtr_b200 kernels need a CUDA device, got %s" % self.device)
        self._h = _c.c_void_p()
        with torch.cuda.device(self.device):
            check(load().gdt_resize_plan_create(int(w), int(h), float(imsize), _c.byref(self._h)), "gdt_resize_plan_create")
        ow, oh = _c.c_int(), _c.c_int()
        check(load().gdt_resize_plan_info(self._h, _c.byref(ow), _c.byref(oh), None, None), "gdt_resize_plan_info")
        self.in_size, self.out_size = (int(w), int(h)), (ow.value, oh.value)
        self.ws_bytes = load().gdt_resize_workspace_bytes(self._h)

    def __del__(self):
        try:
            if self._h:
                load().gdt_resize_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass


def resize_u8(plan, src, out=None):
    """src: uint8 CUDA tensor [h, w, 3] (any row stride: a crop view is fine) -> [out_h, out_w, 3] contiguous."""
    if not isinstance(src, torch.Tensor) or not src.is_cuda or src.dtype != torch.uint8 or src.dim() != 3 or src.shape[2] != 3:
        raise GdtError("resize_u8 needs a uint8 CUDA tensor [h, w, 3] (gandtr_b200 has no CPU path)")
    if src.stride(2) != 1 or src.stride(1) != 3:
        raise GdtError("resize_u8: pixels must be packed RGB (strides (*, 3, 1)), got %s" % (src.stride(),))
    if (src.shape[1], src.shape[0]) != plan.in_size:
        raise GdtError("resize_u8: plan is for %s, image is %s" % (plan.in_size, (src.shape[1], src.shape[0])))
    ow, oh = plan.out_size
    if out is None:
        out = torch.empty((oh, ow, 3), dtype=torch.uint8, device=src.device)
    with torch.cuda.device(src.device):
        ws = _workspace(plan.ws_bytes, src.device)
        check(load().gdt_resize_u8(plan._h, _ptr(src), src.stride(0), _ptr(out), _ptr(ws), ws.numel(), _stream()), "gdt_resize_u8")
    _count("resize")
    return out


def resize_u8_batch(plan, srcs, out=None):
    """srcs: list of uint8 CUDA tensors [h, w, 3] of the plan's geometry (any row stride each) -> [n, out_h, out_w, 3]
    contiguous, one launch per pass for the whole list."""
    n = len(srcs)
    if n == 0:
        raise GdtError("resize_u8_batch needs at least one image")
    for src in srcs:
        if not isinstance(src, torch.Tensor) or not src.is_cuda or src.dtype != torch.uint8 or src.dim() != 3 or src.shape[2] != 3:
            raise GdtError("resize_u8_batch needs uint8 CUDA tensors [h, w, 3] (gandtr_b200 has no CPU path)")
        if src.stride(2) != 1 or src.stride(1) != 3:
            raise GdtError("resize_u8_batch: pixels must be packed RGB (strides (*, 3, 1)), got %s" % (src.stride(),))
        if (src.shape[1], src.shape[0]) != plan.in_size:
            raise GdtError("resize_u8_batch: plan is for %s, image is %s" % (plan.in_size, (src.shape[1], src.shape[0])))
    ow, oh = plan.out_size
    dev = srcs[0].device
    if out is None:
        out = torch.empty((n, oh, ow, 3), dtype=torch.uint8, device=dev)
    ptrs = (ctypes.c_void_p * n)(*[s.data_ptr() for s in srcs])
    strides = (ctypes.c_size_t * n)(*[int(s.stride(0)) for s in srcs])
    lib = load()
    with torch.cuda.device(dev):
        ws = _workspace(lib.gdt_resize_batch_workspace_bytes(plan._h, n), dev)
        check(lib.gdt_resize_u8_batch(plan._h, ptrs, strides, n, _ptr(out), _ptr(ws), ws.numel(), _stream()), "gdt_resize_u8_batch")
    _count("resize")
    return out
