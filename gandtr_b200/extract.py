"""Batched descriptor extraction -- replaces `extract_vectors / extract_ss / extract_ms`
(mdir/external/cirtorch/networks/imageretrievalnet.py:312-359) and the dataset half of it
(`ImagesFromList.__getitem__`, genericdataset.py:66-102; `imresize`, datahelpers.py:75-82).

The reference runs batch size 1 with a `.cpu()` sync per image. Here: images are decoded / cropped / resized on the
host exactly as the reference does (PIL, LANCZOS thumbnail), consecutive images of equal size are stacked into one
pinned uint8 batch, and each batch is ONE K1 launch pair (CLAHE transform) + stock backbone + ONE K2 call. Descriptors
stay on the device as rows of the local database shard ([n_local, D]); `extract_vectors` is the reference-shaped
wrapper (D x n tensor on the host).

Data parallel: `rank` / `world_size` take a contiguous block of the image list (`retrieval.shard_bounds`), which is
exactly the row sharding `ShardedIndex` expects -- no collective on this path.
"""
import numpy as np
import torch

from . import _lib
from .retrieval import shard_bounds

__all__ = ["imresize", "load_image", "crop_box", "crop_like_pil", "extract_descriptors", "extract_vectors", "extract_ss", "extract_ms", "HostBatchUploader"]


class HostBatchUploader:
    """Double-buffered host -> device staging of uint8 image batches on a dedicated copy stream, so the PCIe transfer of
    batch i+1 overlaps the kernels of batch i. `upload(host_batch)` returns a device tensor that is valid on the
    caller's current stream; `release(device_batch)` (called after the last kernel that reads it was enqueued) lets
    the slot be overwritten two uploads later."""

    def __init__(self, device, slots=2):
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.slots = [None] * slots
        self.ready = [torch.cuda.Event() for _ in range(slots)]
        self.free = [None] * slots
        self.i = 0

    def upload(self, host_batch):
        j = self.i % len(self.slots)
        self.i += 1
        if self.slots[j] is None or self.slots[j].shape != host_batch.shape or self.slots[j].dtype != host_batch.dtype:
            self.slots[j] = torch.empty(host_batch.shape, dtype=host_batch.dtype, device=self.device)
            self.free[j] = None
            # The block comes from the compute stream's pool: kernels already queued there may still be using its previous
            # contents (a workspace, an older slot). The copy stream must not write into it before they are done.
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.copy_stream):
            if self.free[j] is not None:
                self.copy_stream.wait_event(self.free[j])        # the consumer of the previous occupant has finished
            self.slots[j].copy_(host_batch, non_blocking=True)
            self.ready[j].record(self.copy_stream)
        torch.cuda.current_stream(self.device).wait_event(self.ready[j])
        return self.slots[j]

    def release(self, device_batch):
        for j, s in enumerate(self.slots):
            if s is device_batch:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
                self.free[j] = ev



def imresize(img, imsize):
    """datahelpers.py:75-82: in-place LANCZOS thumbnail so that max(size) <= imsize; arrays pass through."""
    if isinstance(img, np.ndarray):
        return img
    from PIL import Image
    img.thumbnail((imsize, imsize), getattr(Image, "LANCZOS", Image.Resampling.LANCZOS))
    return img


def crop_box(bbx):
    """Image.crop's box arithmetic: every coordinate is rounded (half to even), not truncated."""
    return tuple(int(round(v)) for v in bbx)


def crop_like_pil(arr, bbx):
    """img.crop(bbx) on an [h, w, c] array / tensor: rounded box, pixels outside the image are black."""
    x0, y0, x1, y1 = crop_box(bbx)
    h, w = arr.shape[0], arr.shape[1]
    if 0 <= x0 <= x1 <= w and 0 <= y0 <= y1 <= h:
        return arr[y0:y1, x0:x1]
    out = (torch.zeros if isinstance(arr, torch.Tensor) else np.zeros)((max(y1 - y0, 0), max(x1 - x0, 0)) + tuple(arr.shape[2:]),
                                                                      **({"dtype": arr.dtype, "device": arr.device} if isinstance(arr, torch.Tensor) else {"dtype": arr.dtype}))
    sx0, sy0, sx1, sy1 = max(x0, 0), max(y0, 0), min(x1, w), min(y1, h)
    if sx1 > sx0 and sy1 > sy0:
        out[sy0 - y0:sy1 - y0, sx0 - x0:sx1 - x0] = arr[sy0:sy1, sx0:sx1]
    return out


def load_image(item, imsize=None, bbx=None):
    """One element of the image list -> uint8 HWC RGB array. `item` is a path, a PIL image or a uint8 array.
    Crop / resize order and the bounding-box scale rule follow genericdataset.py:88-97."""
    from PIL import Image
    if isinstance(item, np.ndarray):
        if bbx:
            item = crop_like_pil(item, bbx)
        return np.ascontiguousarray(item)
    if isinstance(item, Image.Image):
        img = item.convert("RGB")
    else:
        with open(item, "rb") as f:
            img = Image.open(f).convert("RGB")
    imfullsize = max(img.size)
    if bbx:
        img = img.crop(tuple(bbx))
    if imsize is not None:
        img = imresize(img, imsize * max(img.size) / imfullsize if bbx else imsize)
    return np.asarray(img)


class _Decode(torch.utils.data.Dataset):
    """Worker side: decode (+ crop / LANCZOS thumbnail on the host unless `device_resize`)."""

    def __init__(self, images, imsize, bbxs, device_resize=False):
        self.images, self.imsize, self.bbxs, self.device_resize = images, imsize, bbxs, device_resize

    def __len__(self):
        return len(self.images)

    def __getitem__(self, i):
        if self.device_resize:      # decode only: crop and thumbnail run on the GPU (K5, gandtr_b200/loader.py)
            # arrays are never resized by the reference (imresize passes them through): flag them
            return torch.from_numpy(load_image(self.images[i], None, None).copy()), isinstance(self.images[i], np.ndarray)
        return torch.from_numpy(load_image(self.images[i], self.imsize, self.bbxs[i] if self.bbxs is not None else None).copy())


def _net_rows(net, x):
    """`net(input)` of the reference loop (imageretrievalnet.py:341-353) on a whole batch -> [b, D'] rows.
    A `SingleNetwork` runs its stage wrappers (the hub models' eval stack {cirwhiten, cirmultiscale} is ONE fused K2
    call, network.py `_forward_fused`); a bare `ImageRetrievalNet` is features -> GeM -> L2N."""
    if not hasattr(net, "wrappers"):
        return net.descriptors([net.feature_map(x)])
    if net._fused_plan() is not None:
        out = net(x)
        return out.unsqueeze(0) if out.dim() == 1 else out.t()
    # arbitrary wrapper stacks follow the reference's calling convention: one image per call
    rows = [net(x[j:j + 1]).reshape(-1) for j in range(x.shape[0])]
    return torch.stack(rows)


def _descriptors_for_batch(net, x, ms, msp):
    """x: [b,3,h,w] normalised CUDA batch -> [b, D']. Mirrors extract_ss (ms == [1]) and extract_ms; `net` is called
    the way the reference calls it, so a SingleNetwork's whitening / multi-scale wrappers take part."""
    if len(ms) == 1 and ms[0] == 1:
        return _net_rows(net, x)
    per_scale = []
    for s in ms:
        xs = x if s == 1 else torch.nn.functional.interpolate(x, scale_factor=s, mode="bilinear", align_corners=False)
        per_scale.append(_net_rows(net, xs).contiguous())                  # net(input_t) per scale
    # v = sum_s d^msp ; v /= len(ms) ; v = v^(1/msp) ; v /= ||v||        (imageretrievalnet.py:344-357)
    return _lib.desc_post(per_scale, msp if isinstance(msp, torch.Tensor) else float(msp))


def extract_descriptors(net, images, image_size, transform, bbxs=None, ms=(1,), msp=1, batch_size=32, workers=4,
                        rank=0, world_size=1, print_freq=0, device_resize=None):
    """-> [n_local, D] float32 CUDA tensor: rows lo..hi of the descriptor matrix, (lo, hi) = shard_bounds(len(images)).
    `net` is a gandtr_b200 SingleNetwork or ImageRetrievalNet; `transform` a gandtr_b200.transforms.Compose.
    `device_resize`: the workers only decode; bounding-box crop and the LANCZOS thumbnail run on the GPU (K5,
    bit-identical to the host path), so the host cores are left to the JPEG decoder. Default (None): on whenever there
    is geometry work to do (an `image_size` or bounding boxes); without any, whole batches are uploaded as they are.
    False keeps Pillow's crop / thumbnail in the workers (the reference's arrangement)."""
    if device_resize is None:
        device_resize = image_size is not None or (bbxs is not None and any(b for b in bbxs))
    model = getattr(net, "model", net)
    net.eval()                                                  # SingleNetwork.eval() also selects the eval wrappers
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise _lib.GdtError("extract_descriptors needs the network on a CUDA device")
    lo, hi = shard_bounds(len(images), world_size, rank)
    local = list(images[lo:hi])
    local_bbxs = list(bbxs[lo:hi]) if bbxs is not None else None
    out = None            # allocated from the first batch: a whitening wrapper may reduce the dimensionality
    loader = torch.utils.data.DataLoader(_Decode(local, image_size, local_bbxs, device_resize), batch_size=None, shuffle=False,
                                         num_workers=min(workers, max(len(local), 1)) if len(local) > 8 else 0)
    pending, shape, start = [], None, 0
    uploader = HostBatchUploader(dev)
    geometry = None
    if device_resize:
        from .loader import DeviceImageLoader
        geometry = DeviceImageLoader(imsize=image_size, device=dev)

    def flush():
        nonlocal pending, start, out
        if not pending:
            return
        batch = torch.stack(pending)                           # device_resize: already on the device
        if not device_resize:
            batch = batch.pin_memory()                         # stays referenced until the descriptors are enqueued
        with torch.no_grad():
            if device_resize:
                x = transform.batch(batch)
            else:
                staged = uploader.upload(batch)
                x = transform.batch(staged)
                uploader.release(staged)                       # K1 was the only reader of the uint8 batch
            rows = _descriptors_for_batch(net, x, list(ms), msp)
            if out is None:
                out = torch.empty((len(local), rows.shape[1]), dtype=torch.float32, device=dev)
            out[start:start + len(pending)] = rows
        start += len(pending)
        pending = []

    for i, img in enumerate(loader):
        if device_resize:
            img, is_array = img
            bbx = local_bbxs[i] if local_bbxs is not None else None
            img = geometry.resize(img, bbx=bbx) if not is_array else geometry.crop_only(img, bbx)
        if shape is not None and (tuple(img.shape) != shape or len(pending) >= batch_size):
            flush()
        shape = tuple(img.shape)
        pending.append(img)
        if print_freq and ((i + 1) % print_freq == 0 or (i + 1) == len(local)):
            print("\r>>>> {}/{} done...".format(i + 1, len(local)), end="")
    flush()
    if print_freq:
        print("")
    if out is None:
        out = torch.empty((0, model.meta["out_channels"]), dtype=torch.float32, device=dev)
    return out


def extract_vectors(net, images, image_size, transform, bbxs=None, ms=[1], msp=1, print_freq=10, device=None, **kw):
    """Reference signature and return convention (imageretrievalnet.py:312-339): D x n float32 tensor on the host."""
    vecs = extract_descriptors(net, images, image_size, transform, bbxs=bbxs, ms=ms, msp=msp, print_freq=print_freq, **kw)
    return vecs.t().contiguous().cpu()


def extract_ss(net, input):
    """imageretrievalnet.py:341-342."""
    return net(input).cpu().data.squeeze()


def extract_ms(net, input, ms, msp):
    """imageretrievalnet.py:344-359."""
    model = getattr(net, "model", net)
    return _descriptors_for_batch(net, input.to(next(model.parameters()).device), list(ms), msp).squeeze(0).cpu()
